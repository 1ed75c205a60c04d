// Microbenchmark: tcgen05.mma.cta_group::2 (M = 256 over a CTA pair) -- semantics check + issue rate per N, SS mode bf16.
// Question it answers: a Cout = 64 conv is bound by the shared-memory operand fetch at 48 cycles per 128x64x16 MMA (67 % of
// the tensor peak, umma_rate.cu).  With cta_group::2 each SM supplies its own 128-row A tile but only HALF of B: does the
// pair run at max(32, (4096 + 1024) / 128 = 40) cycles per instruction?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate_cta2 umma_rate_cta2.cu ; run on a B200.
#include "../../refid_b200/csrc/common.cuh"
#include "../../refid_b200/csrc/common.cu"
using namespace refid;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// bounded wait: a protocol mistake must end the kernel, not hang the box
__device__ __forceinline__ bool wait_bounded(uint64_t* bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 24); ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_rate2(int N, int iters, long long* out, float* dout) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  // A = 1.0 everywhere; B = 1.0 in the leader's half, 2.0 in the peer's half (bf16 pairs)
  const uint32_t one = 0x3f803f80u, two = 0x40004000u;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = one;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem + 96 * 1024)[i] = rank ? two : one;
  fence_proxy_async();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  cluster_sync_all();
  const uint32_t tm = slot;
  const uint32_t idesc = make_idesc_bf16(256, N, 0, 0);
  const uint32_t a0 = smem_u32(smem), b0 = a0 + 96 * 1024;
  long long dt = 0;
  if (rank == 0 && threadIdx.x == 0) {
    uint64_t da[4], db[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      da[u] = make_smem_desc(a0 + (uint32_t)(u * 16384), 16, 1024, 2);
      db[u] = make_smem_desc(b0 + (uint32_t)(u * 16384), 16, 1024, 2);
    }
    // (1) semantics: ONE K=16 MMA, no accumulate
    umma2_bf16(tm, da[0], db[0], idesc, 0);
    umma2_commit_mc(&bar, 3);
  }
  const bool ok0 = wait_bounded(&bar, 0);
  tc_fence_after();
  {
    const int w = threadIdx.x >> 5;
    float v[16], u[16];
    tmem_ld16(tm + ((uint32_t)(w * 32) << 16), v);
    tmem_ld16(tm + ((uint32_t)(w * 32) << 16) + (uint32_t)(N / 2), u);
    tmem_ld_wait();
    if ((threadIdx.x & 31) == 0) {
      dout[(blockIdx.x * 4 + w) * 2] = ok0 ? v[0] : -1.f;
      dout[(blockIdx.x * 4 + w) * 2 + 1] = u[0];
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (rank == 0 && threadIdx.x == 0) {
    uint64_t da[4], db[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      da[u] = make_smem_desc(a0 + (uint32_t)(u * 16384), 16, 1024, 2);
      db[u] = make_smem_desc(b0 + (uint32_t)(u * 16384), 16, 1024, 2);
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; it += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma2_bf16(tm + (uint32_t)((u & 1) * N), da[u] + (uint64_t)(k * 2), db[u] + (uint64_t)(k * 2), idesc, 1);
      }
    }
    umma2_commit_mc(&bar, 3);
    const bool ok1 = wait_bounded(&bar, 1);
    dt = clock64() - t0;
    out[blockIdx.x / 2] = ok1 ? dt : -1;
  } else {
    wait_bounded(&bar, 1);
  }
  tc_fence_before(); __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  float* dd; cudaMalloc(&dd, 148 * 4 * 2 * 4);
  { cudaError_t e0 = cudaFuncSetAttribute(k_rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); printf("attr: %s\n", cudaGetErrorString(e0)); }
  for (int N : {32, 64, 128, 256}) {
    const int iters = 2000;
    cudaMemset(dd, 0, 148 * 4 * 2 * 4);
    k_rate2<<<148, 128, 201 * 1024 + 1024>>>(N, iters, d, dd);
    cudaError_t e = cudaGetLastError(); cudaError_t e2 = cudaDeviceSynchronize(); if (e == cudaSuccess) e = e2;
    long long h[74]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    float hd[148 * 8]; cudaMemcpy(hd, dd, sizeof(hd), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 74; ++i) mx = h[i] > mx ? h[i] : mx;
    double cyc = (double)mx / (iters * 4);
    printf("cta_group::2 M=256 N=%3d: %.1f cyc per instruction (= per 128xNx16 of EACH SM; math floor %.0f)  D[row0][0]=%g D[row0][N/2]=%g (CTA0) | %g %g (CTA1)  %s\n",
           N, cyc, 128.0 * N / 256.0, hd[0], hd[1], hd[8], hd[9], cudaGetErrorString(e));
  }
  printf("expected semantics: A = 1, B = 1 (leader's half of N) / 2 (peer's half): D = 16 in columns [0, N/2), 32 in [N/2, N), in BOTH CTAs' TMEM\n");
  return 0;
}
