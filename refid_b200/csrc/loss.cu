// Charbonnier loss (reference basicsr/models/losses/losses.py:28-30 `sqrt((pred - target)^2 + eps)`,
// :143-173 CharbonnierLoss with reduction mean | sum and loss_weight): value AND gradient in one pass over pred / target.
// HBM-bound: 8 B read + 4 B written per element; per-block partial sums reduced in a fixed order (bit-reproducible).
#include "common.cuh"

namespace refid {
namespace {

constexpr int kLossThreads = 256;
constexpr int kLossBlocks = 148 * 8;

__global__ void __launch_bounds__(kLossThreads) k_charbonnier(const float* __restrict__ pred, const float* __restrict__ gt,
                                                              float* __restrict__ grad, double* __restrict__ partial, long n,
                                                              float eps, float gscale) {
  pdl_launch_dependents();
  pdl_wait();
  float acc = 0.f;
  const long n4 = n >> 2;
  for (long i = (long)blockIdx.x * kLossThreads + threadIdx.x; i < n4; i += (long)gridDim.x * kLossThreads) {
    const float4 p = reinterpret_cast<const float4*>(pred)[i];
    const float4 g = reinterpret_cast<const float4*>(gt)[i];
    float4 d = make_float4(p.x - g.x, p.y - g.y, p.z - g.z, p.w - g.w);
    float4 r = make_float4(sqrtf(d.x * d.x + eps), sqrtf(d.y * d.y + eps), sqrtf(d.z * d.z + eps), sqrtf(d.w * d.w + eps));
    acc += (r.x + r.y) + (r.z + r.w);
    if (grad) reinterpret_cast<float4*>(grad)[i] = make_float4(gscale * d.x / r.x, gscale * d.y / r.y, gscale * d.z / r.z, gscale * d.w / r.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {  // ragged tail
    const long i = (n4 << 2) + threadIdx.x;
    const float d = pred[i] - gt[i], r = sqrtf(d * d + eps);
    acc += r;
    if (grad) grad[i] = gscale * d / r;
  }
  __shared__ float swarp[kLossThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) swarp[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) s += (double)swarp[w];
    partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(32) k_charbonnier_finish(const double* __restrict__ partial, int nparts, double scale,
                                                           float* __restrict__ loss) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nparts; ++i) s += partial[i];  // fixed order
    *loss = (float)(s * scale);
  }
}

// x *= *s, skipped entirely (no memory traffic) when the device scalar is exactly 1 -- the upstream gradient of a training
// loss; the decision is taken on the device so the host never synchronises to look at it
__global__ void __launch_bounds__(kLossThreads) k_scale_by_scalar(float* __restrict__ x, const float* __restrict__ s, long n) {
  pdl_launch_dependents();
  pdl_wait();
  const float v = *s;
  if (v == 1.0f) return;
  for (long i = (long)blockIdx.x * kLossThreads + threadIdx.x; i < n; i += (long)gridDim.x * kLossThreads) x[i] *= v;
}

}  // namespace

int launch_scale_by_scalar(float* x, const float* scalar, long n, cudaStream_t s) {
  long want = (n + kLossThreads - 1) / kLossThreads;
  const int blocks = (int)(want < 1 ? 1 : (want > kLossBlocks ? kLossBlocks : want));
  REFID_CUDA_CHECK(launch_k(k_scale_by_scalar, dim3(blocks), dim3(kLossThreads), 0, s, x, scalar, n));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// reduction: 0 = sum, 1 = mean.  `partial` = kLossBlocks doubles of device scratch.
int launch_charbonnier(const float* pred, const float* gt, float* grad, double* partial, float* loss, long n, float eps,
                       float loss_weight, int reduction_mean, cudaStream_t s) {
  REFID_REQUIRE(n > 0, "charbonnier: empty input");
  REFID_REQUIRE((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(grad)) % 16 == 0,
                "charbonnier: pred / target / grad must be 16-byte aligned");
  const double scale = reduction_mean ? (double)loss_weight / (double)n : (double)loss_weight;
  long want = ((n >> 2) + kLossThreads - 1) / kLossThreads;
  const int blocks = (int)(want < 1 ? 1 : (want > kLossBlocks ? kLossBlocks : want));
  REFID_CUDA_CHECK(launch_k(k_charbonnier, dim3(blocks), dim3(kLossThreads), 0, s, pred, gt, grad, partial, n, eps, (float)scale));
  REFID_CUDA_CHECK(launch_k(k_charbonnier_finish, dim3(1), dim3(32), 0, s, (const double*)partial, blocks, scale, loss));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace refid

extern "C" {
int refid_charbonnier_scratch_bytes(void) { return (int)(refid::kLossBlocks * sizeof(double)); }
int refid_charbonnier(const float* pred, const float* target, float* grad, void* scratch, float* loss, long n, float eps,
                      float loss_weight, int reduction_mean, void* stream) {
  return refid::launch_charbonnier(pred, target, grad, static_cast<double*>(scratch), loss, n, eps, loss_weight, reduction_mean,
                                   static_cast<cudaStream_t>(stream));
}
int refid_scale_by_device_scalar(float* x, const float* scalar, long n, void* stream) {
  return refid::launch_scale_by_scalar(x, scalar, n, static_cast<cudaStream_t>(stream));
}
}
