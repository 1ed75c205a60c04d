// Full-frame validation tiling on the GPU: the crop / merge pair of the reference's `grids` mode
// (basicsr/models/twoImage_event_recurrent_model.py:115-126 transposes, :201-243 crop placement, :245-266 overlap average).
//   refid_grids_crop : (planes,H,W) frame -> (ncrops,planes,cs,cs) crops, each in one of the 8 dihedral orientations
//   refid_grids_merge: (ncrops,planes,cs,cs) network outputs -> (planes,H,W): every output pixel averages the crops that
//                      cover it, un-oriented on the fly -- ONE pass, no atomics, no count image, no zero-fill.
// Both are pure gathers: 4 B read + 4 B written per element (merge: x the local cover count); HBM bound.
#include "common.cuh"

namespace refid {
namespace {

// part[y][x] = crop[a][b]: the reference's `transpose(t, k)` = rot90(flip_W(t) if k >= 4, k % 4) over (H, W).
__device__ __forceinline__ void orient_src(int k, int n, int y, int x, int& a, int& b) {
  switch (k & 3) {
    case 0: a = y; b = x; break;
    case 1: a = x; b = n - 1 - y; break;          // rot90 once: out[i][j] = in[j][n-1-i]
    case 2: a = n - 1 - y; b = n - 1 - x; break;
    default: a = n - 1 - x; b = y; break;
  }
  if (k >= 4) b = n - 1 - b;
}
// the part element that holds crop[a][b] (inverse of orient_src)
__device__ __forceinline__ void orient_dst(int k, int n, int a, int b, int& y, int& x) {
  if (k >= 4) b = n - 1 - b;
  switch (k & 3) {
    case 0: y = a; x = b; break;
    case 1: x = a; y = n - 1 - b; break;
    case 2: y = n - 1 - a; x = n - 1 - b; break;
    default: y = b; x = n - 1 - a; break;
  }
}

__global__ void __launch_bounds__(256) k_grids_crop(const float* __restrict__ src, int planes, int H, int W,
                                                    const int* __restrict__ idx, int ncrops, int cs, float* __restrict__ dst) {
  const long total = (long)ncrops * planes * cs * cs;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int x = (int)(e % cs);
    long r = e / cs;
    const int y = (int)(r % cs);
    r /= cs;
    const int p = (int)(r % planes), c = (int)(r / planes);
    const int i = idx[3 * c], j = idx[3 * c + 1], k = idx[3 * c + 2];
    int a, b;
    orient_src(k, cs, y, x, a, b);
    dst[e] = __ldg(src + ((size_t)p * H + (i + a)) * W + (j + b));
  }
}

__global__ void __launch_bounds__(256) k_grids_merge(const float* __restrict__ parts, int planes, int H, int W,
                                                     const int* __restrict__ idx, int ncrops, int cs, float* __restrict__ dst) {
  extern __shared__ int sidx[];
  for (int t = threadIdx.x; t < 3 * ncrops; t += blockDim.x) sidx[t] = idx[t];
  __syncthreads();
  const long total = (long)planes * H * W;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int X = (int)(e % W);
    const long r = e / W;
    const int Y = (int)(r % H), p = (int)(r / H);
    float sum = 0.f, cnt = 0.f;
    for (int c = 0; c < ncrops; ++c) {  // accumulation in crop order, like the reference's loop (bit-identical sums)
      const int i = sidx[3 * c], j = sidx[3 * c + 1], k = sidx[3 * c + 2];
      const int a = Y - i, b = X - j;
      if (a < 0 || a >= cs || b < 0 || b >= cs) continue;
      int y, x;
      orient_dst(k, cs, a, b, y, x);
      sum += __ldg(parts + (((size_t)c * planes + p) * cs + y) * cs + x);
      cnt += 1.f;
    }
    dst[e] = sum / cnt;
  }
}

}  // namespace
}  // namespace refid

extern "C" {

// src (planes,H,W) fp32; idx: ncrops x (i, j, trans_idx) int32 on the device; dst (ncrops,planes,cs,cs) fp32.
int refid_grids_crop(const float* src, int planes, int H, int W, const int* idx, int ncrops, int crop_size, float* dst,
                     void* stream) {
  using namespace refid;
  REFID_REQUIRE(src && idx && dst && planes > 0 && ncrops > 0, "grids_crop: bad argument");
  REFID_REQUIRE(crop_size > 0 && crop_size <= H && crop_size <= W, "grids_crop: crop_size %d does not fit a %dx%d frame",
                crop_size, H, W);
  const long total = (long)ncrops * planes * crop_size * crop_size;
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  k_grids_crop<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, planes, H, W, idx, ncrops, crop_size, dst);
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// parts (ncrops,planes,cs,cs) fp32 -> dst (planes,H,W): average over the covering crops (every pixel must be covered).
int refid_grids_merge(const float* parts, int planes, int H, int W, const int* idx, int ncrops, int crop_size, float* dst,
                      void* stream) {
  using namespace refid;
  REFID_REQUIRE(parts && idx && dst && planes > 0 && ncrops > 0, "grids_merge: bad argument");
  REFID_REQUIRE(ncrops <= 4096, "grids_merge: %d crops exceed the shared-memory index table", ncrops);
  const long total = (long)planes * H * W;
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  k_grids_merge<<<(unsigned)blocks, 256, (size_t)ncrops * 12, static_cast<cudaStream_t>(stream)>>>(parts, planes, H, W, idx,
                                                                                                   ncrops, crop_size, dst);
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"
