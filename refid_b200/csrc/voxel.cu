// Event -> voxel-grid rasterisation on the GPU (reference `events_to_voxel_grid`,
// basicsr/data/event_util.py:6-66, called per sliding two-event-chunk window at basicsr/data/image_npy_dataset.py:175-188).
// Same per-event arithmetic and types as the reference (float32 normalised time stamp, float64 bilinear weights, polarity
// 0 -> -1, truncation toward zero, the `ti < bins` / `ti + 1 < bins` validity rules, t[0] / t[-1] as the time span).  The
// reference adds event by event into a float32 grid (np.add.at); here the weights are accumulated as 2^-36 fixed point in
// int64 atomics -- exact, order-independent, bit-reproducible -- and rounded to float32 once, which differs from the
// sequential float32 sum by its rounding error only (<= 1e-6 relative in the tests).  Events outside the grid are dropped
// (the reference would raise or wrap).  Scatter-add: atomics / HBM-bound, 16 B read per event.
#include "common.cuh"

namespace refid {
namespace {

constexpr int kVoxThreads = 256;
constexpr double kVoxScale = 68719476736.0;  // 2^36

__global__ void __launch_bounds__(kVoxThreads) k_voxel_scatter(const float4* __restrict__ ev, long n, int bins, int W, int H,
                                                               long long* __restrict__ grid) {
  pdl_launch_dependents();
  pdl_wait();
  const float first = ev[0].x, last = ev[n - 1].x;
  float span = last - first;
  if (span == 0.f) span = 1.0f;
  const float nb1 = (float)(bins - 1);
  const long plane = (long)W * H;
  for (long i = (long)blockIdx.x * kVoxThreads + threadIdx.x; i < n; i += (long)gridDim.x * kVoxThreads) {
    const float4 e = ev[i];  // (timestamp, x, y, polarity)
    const float ts = __fdiv_rn(__fmul_rn(nb1, __fsub_rn(e.x, first)), span);  // float32, the reference's operation order
    const int x = (int)e.y, y = (int)e.z;  // astype(int): truncation
    if (x < 0 || x >= W || y < 0 || y >= H) continue;
    const double pol = e.w == 0.f ? -1.0 : (double)e.w;
    const long long ti = (long long)ts;  // truncation toward zero
    if (ti < 0) continue;
    const double dt = (double)ts - (double)ti;
    const long pix = (long)x + (long)y * W;
    if (ti < bins) atomicAdd(reinterpret_cast<unsigned long long*>(grid + ti * plane + pix),
                             (unsigned long long)(long long)llrint(pol * (1.0 - dt) * kVoxScale));
    if (ti + 1 < bins) atomicAdd(reinterpret_cast<unsigned long long*>(grid + (ti + 1) * plane + pix),
                                 (unsigned long long)(long long)llrint(pol * dt * kVoxScale));
  }
}

// fixed point -> float32, (bins,H,W) or (H,W,bins)
__global__ void __launch_bounds__(kVoxThreads) k_voxel_finish(const long long* __restrict__ grid, int bins, long plane, int hwc,
                                                              float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = plane * bins;
  for (long i = (long)blockIdx.x * kVoxThreads + threadIdx.x; i < total; i += (long)gridDim.x * kVoxThreads) {
    const float v = (float)((double)grid[i] / kVoxScale);
    if (hwc) out[(i % plane) * bins + i / plane] = v;
    else out[i] = v;
  }
}

}  // namespace
}  // namespace refid

extern "C" {
int refid_events_to_voxel(const float* events, long n, int num_bins, int width, int height, int hwc, void* scratch, float* voxel,
                          void* stream) {
  using namespace refid;
  REFID_REQUIRE(events && n > 0 && num_bins > 0 && width > 0 && height > 0 && scratch && voxel, "events_to_voxel: bad arguments");
  REFID_REQUIRE(reinterpret_cast<uintptr_t>(events) % 16 == 0, "events_to_voxel: the [n][4] event array must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long total = (long)num_bins * width * height;
  REFID_CUDA_CHECK(cudaMemsetAsync(scratch, 0, sizeof(long long) * total, s));
  long b1 = (n + kVoxThreads - 1) / kVoxThreads, b2 = (total + kVoxThreads - 1) / kVoxThreads;
  if (b1 > 148 * 16) b1 = 148 * 16;
  if (b2 > 148 * 16) b2 = 148 * 16;
  REFID_CUDA_CHECK(launch_k(k_voxel_scatter, dim3((unsigned)b1), dim3(kVoxThreads), 0, s, reinterpret_cast<const float4*>(events), n,
                            num_bins, width, height, static_cast<long long*>(scratch)));
  REFID_CUDA_CHECK(launch_k(k_voxel_finish, dim3((unsigned)b2), dim3(kVoxThreads), 0, s, (const long long*)scratch, num_bins,
                            (long)width * height, hwc, voxel));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}
}
