// Microbenchmark: the exact MMA stream of the resident-weight 3x3 halo conv (M=128 pixels x N cout, K=64 slab, nine taps as
// shifted A descriptors into a 10-pixel-pitch halo patch), fully unrolled with constant-folded descriptors, against
// variants without the shifts / with a dense A tile.  Answers: what is the per-MMA floor of this access pattern?
#include "../../refid_b200/csrc/common.cuh"
#include "../../refid_b200/csrc/common.cu"
using namespace refid;

template <int N, int NM, int VAR>
__global__ void __launch_bounds__(320, 1) k(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 320) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = __shfl_sync(0xffffffffu, slot, 0);
  constexpr uint32_t PITCH = 10, PXB = 128, SBO = VAR == 2 ? 1024u : PITCH * PXB;
  constexpr uint32_t A_BYTES = (16 * NM + 2) * PITCH * PXB, B_TILE = N * 128;
  constexpr uint32_t IDESC = make_idesc_bf16(128, N, 0, 0);
  if (threadIdx.x < 32) {
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 2 * ((A_BYTES + 1023) & ~1023u);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t sa = it & 1;
      const uint64_t a_desc0 = make_smem_desc(a_base + sa * ((A_BYTES + 1023) & ~1023u), 16, SBO, 2);
      const uint64_t b_desc0 = make_smem_desc(b_base, 16, 1024, 2);
      if (elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t tap_off = VAR >= 1 ? 0u : (uint32_t)((tap / 3) * PITCH + tap % 3) * PXB;
#pragma unroll
          for (int j = 0; j < NM; ++j)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ad = a_desc0 + (uint64_t)((tap_off + (uint32_t)j * 16u * SBO + (uint32_t)kk * 32u) >> 4);
              const uint64_t bd = b_desc0 + (uint64_t)(((VAR == 3 ? 0u : (uint32_t)tap * B_TILE) + (uint32_t)kk * 32u) >> 4);
              umma_bf16(tm + (uint32_t)((it & 1) * NM * N + j * N), ad, bd, IDESC, 1);
            }
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0, 1);
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int N, int NM, int VAR>
void run(long long* d, const char* what) {
  cudaFuncSetAttribute(k<N, NM, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int iters = 400;
  k<N, NM, VAR><<<148, 320, 200 * 1024 + 1024>>>(iters, d);
  cudaError_t e = cudaGetLastError(); cudaError_t e2 = cudaDeviceSynchronize(); if (e == cudaSuccess) e = e2;
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d NM=%d %-46s: %.1f cyc/MMA  %s\n", N, NM, what, (double)mx / (iters * 36.0 * NM), cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
#define ALL(N, NM) \
  run<N, NM, 0>(d, "halo patch, shifted taps, SBO 1280 (kernel)"); \
  run<N, NM, 1>(d, "halo patch, no shifts, SBO 1280"); \
  run<N, NM, 2>(d, "dense tile, no shifts, SBO 1024"); \
  run<N, NM, 3>(d, "kernel pattern, one weight tile for all taps");
  ALL(32, 2) ALL(64, 1) ALL(64, 2)
  return 0;
}
