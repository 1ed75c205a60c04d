// Training-sample assembly on the GPU (second half): what `__getitem__` does between the voxel grid and
// the tensors the network sees -- basicsr/data/image_npy_dataset.py:189-232 with basicsr/data/transforms.py:88-129 (augment)
// and :163-238 (triple_random_crop): one crop window, horizontal / vertical flip, transpose, and the channel packing
//   lq    = [ blurry_0 (3) | voxel[1:m] | blurry_1 (3) | voxel[m+2+n:] ]         (:212-221, "deblur voxel")
//   voxel = sliding two-bin windows voxel[t:t+2], t = 0 .. num_bins-2            (:226-232)
// as ONE gather: out[p][y][x] = src(plane_table[p])[crop + flips + transpose of (y,x)].  4 B read + 4 B written per element.
#include "common.cuh"

namespace refid {
namespace {

__global__ void __launch_bounds__(256) k_crop_flip_gather(const float* __restrict__ src_a, const float* __restrict__ src_b,
                                                          int H, int W, const int* __restrict__ plane_table, int nplanes, int top,
                                                          int left, int ph, int pw, int hflip, int vflip, int rot90,
                                                          float* __restrict__ dst) {
  const int oh = rot90 ? pw : ph, ow = rot90 ? ph : pw;
  const long total = (long)nplanes * oh * ow;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int x = (int)(e % ow);
    const long r = e / ow;
    const int y = (int)(r % oh), p = (int)(r / oh);
    int a = rot90 ? x : y, b = rot90 ? y : x;  // transpose(1,0,2) is applied last (transforms.py:127-128)
    if (vflip) a = ph - 1 - a;                 // cv2.flip(img, 0)
    if (hflip) b = pw - 1 - b;                 // cv2.flip(img, 1)
    const int t = plane_table[p];
    const float* s = t >= 0 ? src_a + (size_t)t * H * W : src_b + (size_t)(-1 - t) * H * W;
    dst[e] = __ldg(s + (size_t)(top + a) * W + (left + b));
  }
}

}  // namespace
}  // namespace refid

extern "C" int refid_crop_flip_gather(const float* src_a, const float* src_b, int H, int W, const int* plane_table, int nplanes,
                                      int top, int left, int ph, int pw, int hflip, int vflip, int rot90, float* dst,
                                      void* stream) {
  using namespace refid;
  REFID_REQUIRE(src_a && plane_table && dst && nplanes > 0, "crop_flip_gather: bad argument");
  REFID_REQUIRE(top >= 0 && left >= 0 && ph > 0 && pw > 0 && top + ph <= H && left + pw <= W,
                "crop_flip_gather: window (%d,%d)+(%d,%d) outside a %dx%d frame", top, left, ph, pw, H, W);
  const long total = (long)nplanes * ph * pw;
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  k_crop_flip_gather<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src_a, src_b, H, W, plane_table, nplanes, top,
                                                                                     left, ph, pw, hflip, vflip, rot90, dst);
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}
