// tensor2img quantisation + PSNR on the GPU (reference basicsr/utils/img_util.py:59-121
// `tensor2img`: clamp to [0,1], x255, round half to even, uint8, CHW RGB -> HWC BGR; basicsr/metrics/psnr_ssim.py:9-61
// `calculate_psnr`: float64 mean of squared uint8 differences inside the crop border).  The squared differences are
// summed as integers (exact, order-independent), so the PSNR the host forms from them is bit-identical to the reference's.
// HBM-bound: 8 B read (+ 2 B written when the uint8 images are requested) per element.
#include "common.cuh"

namespace refid {
namespace {

constexpr int kMetThreads = 256;

__device__ __forceinline__ int quant_u8(float v) {
  v = fminf(fmaxf(v, 0.f), 1.f);
  return (int)rintf(v * 255.0f);  // round half to even, as numpy's .round()
}

// grid = (blocks per frame, frames); frame layout (C,H,W) fp32; images (H,W,C) uint8 with the channel order reversed
__global__ void __launch_bounds__(kMetThreads) k_quant_psnr(const float* __restrict__ pred, const float* __restrict__ gt, int C,
                                                            int H, int W, int crop, int reverse_channels,
                                                            unsigned long long* __restrict__ ssd, unsigned int* __restrict__ maxv,
                                                            unsigned char* __restrict__ img_pred, unsigned char* __restrict__ img_gt) {
  pdl_launch_dependents();
  pdl_wait();
  const int f = blockIdx.y;
  const long hw = (long)H * W;
  const float* pf = pred + (size_t)f * C * hw;
  const float* gf = gt ? gt + (size_t)f * C * hw : nullptr;
  unsigned int acc = 0, mx = 0;  // per thread at most 255^2 * C * (pixels / threads): the launcher keeps this below 2^32
  for (long p = (long)blockIdx.x * kMetThreads + threadIdx.x; p < hw; p += (long)gridDim.x * kMetThreads) {
    const int y = (int)(p / W), x = (int)(p % W);
    const bool inside = y >= crop && y < H - crop && x >= crop && x < W - crop;
    for (int c = 0; c < C; ++c) {
      const int a = quant_u8(pf[c * hw + p]);
      const int oc = reverse_channels ? C - 1 - c : c;
      if (img_pred) img_pred[((size_t)f * hw + p) * C + oc] = (unsigned char)a;
      if (gf) {
        const int b = quant_u8(gf[c * hw + p]);
        if (img_gt) img_gt[((size_t)f * hw + p) * C + oc] = (unsigned char)b;
        if (inside) {
          acc += (unsigned int)((a - b) * (a - b));
          mx = max(mx, (unsigned int)a);
        }
      }
    }
  }
  if (!gf) return;
  __shared__ unsigned long long ssum[kMetThreads / 32];
  __shared__ unsigned int smax[kMetThreads / 32];
  unsigned long long s = acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    ssum[threadIdx.x >> 5] = s;
    smax[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kMetThreads / 32; ++w) {
      s += ssum[w];
      mx = max(mx, smax[w]);
    }
    atomicAdd(&ssd[f], s);  // integer: exact and order-independent
    atomicMax(&maxv[f], mx);
  }
}

}  // namespace
}  // namespace refid

extern "C" {
int refid_quant_psnr(const float* pred, const float* gt, int frames, int C, int H, int W, int crop_border, int reverse_channels,
                     unsigned long long* ssd, unsigned int* max_pred, unsigned char* img_pred, unsigned char* img_gt, void* stream) {
  using namespace refid;
  REFID_REQUIRE(pred && frames > 0 && C > 0 && H > 0 && W > 0 && crop_border >= 0, "quant_psnr: bad arguments");
  REFID_REQUIRE(!gt || (ssd && max_pred), "quant_psnr: ssd / max_pred outputs are required with a ground truth");
  REFID_REQUIRE(2 * crop_border < H && 2 * crop_border < W, "quant_psnr: crop_border %d leaves no pixels", crop_border);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long hw = (long)H * W;
  long blocks = (hw + kMetThreads * 8 - 1) / (kMetThreads * 8);  // <= 8 pixels per thread: no 32-bit overflow of the partials
  if (blocks < 1) blocks = 1;
  REFID_REQUIRE(blocks <= 65535 * 16L, "quant_psnr: frame too large");
  if (gt) {
    REFID_CUDA_CHECK(cudaMemsetAsync(ssd, 0, sizeof(unsigned long long) * frames, s));
    REFID_CUDA_CHECK(cudaMemsetAsync(max_pred, 0, sizeof(unsigned int) * frames, s));
  }
  REFID_CUDA_CHECK(launch_k(k_quant_psnr, dim3((unsigned)blocks, (unsigned)frames), dim3(kMetThreads), 0, s, pred, gt, C, H, W,
                            crop_border, reverse_channels, ssd, max_pred, img_pred, img_gt));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}
}
