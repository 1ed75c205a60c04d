// Microbenchmark: does streaming data INTO shared memory (bulk async copies, as the TMA weight stream does) slow down
// SS-mode tcgen05.mma whose operands are read FROM shared memory?  One CTA per SM: warp 0 lane 0 issues MMAs
// (M=128, N in {64,128,256}), warp 1 lane 0 keeps `depth` 16 KB bulk copies global->smem in flight (or none).
#include "../../refid_b200/csrc/common.cuh"
#include "../../refid_b200/csrc/common.cu"
using namespace refid;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(320, 1) k(int N, int iters, int depth, const uint8_t* src, long long* out, long long* copied, int pollers, int commit_every) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, cbar[8], pbar, dbar[4];
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  if (threadIdx.x == 0) { mbar_init(&pbar, 1); for (int i = 0; i < 4; ++i) mbar_init(&dbar[i], 1); mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&cbar[i], 1); done = 0; fence_barrier_init(); }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 320) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 32 * 1024;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t aa = a0 + (uint32_t)((it & 1) * 16384), bb = b0 + (uint32_t)((it & 1) * 32768);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tm + (uint32_t)((it & 1) * N), make_smem_desc(aa + kk * 32, 16, 1024, 2), make_smem_desc(bb + kk * 32, 16, 1024, 2), idesc, 1);
      // stage-release commits, as a streamed-operand pipeline issues them (nobody waits on these barriers)
      if (commit_every > 0 && ((it + 1) & (commit_every - 1)) == 0) umma_commit(&dbar[(it >> 2) & 3]);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0, 1);
    out[blockIdx.x] = clock64() - t0;
    done = 1;
    mbar_arrive(&pbar);
  } else if (threadIdx.x == 32 && depth > 0) {
    // copy stream into the upper 96 KB of shared memory, `depth` chunks of 16 KB in flight
    uint8_t* dst0 = smem + 100 * 1024;
    const uint8_t* s = src + (size_t)blockIdx.x * (1 << 20);
    long long n = 0;
    uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < depth; ++i) { mbar_arrive_expect_tx(&cbar[i], 16384); bulk_g2s(dst0 + i * 16384, s + ((n++ * 16384) & ((1 << 20) - 1)), 16384, &cbar[i]); }
    while (!done) {
      for (int i = 0; i < depth; ++i) {
        mbar_wait(&cbar[i], ph[i], 2); ph[i] ^= 1;
        mbar_arrive_expect_tx(&cbar[i], 16384);
        bulk_g2s(dst0 + i * 16384, s + ((n++ * 16384) & ((1 << 20) - 1)), 16384, &cbar[i]);
      }
    }
    for (int i = 0; i < depth; ++i) mbar_wait(&cbar[i], ph[i], 3);
    copied[blockIdx.x] = n * 16384;
  }
  if (threadIdx.x >= 64 && (int)(threadIdx.x >> 5) - 2 < pollers) {
    // epilogue-warp stand-ins: whole warps spinning on an mbarrier that completes only when the MMA thread is done
    while (!mbar_try_wait(&pbar, 0)) {}
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long *d, *c; cudaMalloc(&d, 148 * 8); cudaMalloc(&c, 148 * 8);
  uint8_t* src; cudaMalloc(&src, 148ull << 20); cudaMemset(src, 0, 148ull << 20);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int N : {64, 128, 256}) for (int pollers : {0, 8}) for (int depth : {0, 4}) for (int ce : {0, 1, 2, 4, 16}) {
    if (ce > 0 && (pollers == 0 || depth == 0)) continue;
    const int iters = 4000;
    cudaMemset(c, 0, 148 * 8);
    k<<<148, 320, 200 * 1024 + 1024>>>(N, iters, depth, src, d, c, pollers, ce);
    cudaError_t e = cudaGetLastError(); cudaError_t e2 = cudaDeviceSynchronize(); if (e == cudaSuccess) e = e2;
    long long h[148], hc[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hc, c, sizeof(hc), cudaMemcpyDeviceToHost);
    long long mx = 0, cp = 0; for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; cp += hc[i]; }
    printf("N=%3d polling warps %d copies in flight %d commit every %d MMAs: %.1f cyc/MMA, copy stream %.1f B/clk/SM  %s\n", N, pollers, depth, ce * 4, (double)mx / (iters * 4),
           (double)cp / 148 / mx, cudaGetErrorString(e));
  }
  return 0;
}
