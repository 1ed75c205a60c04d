// Global-norm gradient clip + Adam / AdamW step over all parameter tensors in two multi-tensor passes (SURVEY.md 8f rank 2;
// reference basicsr/models/twoImage_event_recurrent_model.py:304-307: `clip_grad_norm_(net_g.parameters(), 0.01)` then
// `optimizer_g.step()`, optimizer built at :67-95 as torch.optim.AdamW / Adam).  Arithmetic follows torch's single-tensor
// implementation: total_norm = || (||g_i||_2)_i ||_2, coef = min(1, max_norm / (total_norm + 1e-6)), g *= coef;
// AdamW: p *= 1 - lr*wd; Adam: g += wd*p; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).
// HBM-bound: pass 1 reads every gradient once (4 B/element), pass 2 reads g, p, m, v and writes p, m, v (28 B/element).
#include "common.cuh"

#include <vector>

namespace refid {
namespace {

constexpr int kOptThreads = 256;
constexpr int kOptBlocks = 148 * 4;
constexpr long kOptChunk = 16384;  // elements per work item

struct OptChunk {
  int tensor;
  int len;
  long off;
};

struct OptState {
  int ntensors = 0;
  int nchunks = 0;
  std::vector<long> numel;
  OptChunk* d_chunks = nullptr;
  // per-call table [5][ntensors] of 8-byte entries: param, grad, exp_avg, exp_avg_sq pointers and the tensor's two bias-
  // correction floats {1 - b1^t, 1 / sqrt(1 - b2^t)} -- staged through pinned memory, 4 slots
  static constexpr int kSlots = 4;
  void** h_ptrs[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  void** d_ptrs[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t done[kSlots];
  int slot = 0;
  double* d_partial = nullptr;  // [kOptBlocks]
  float* d_scalars = nullptr;   // [0] total_norm, [1] clip coefficient
};

__global__ void __launch_bounds__(kOptThreads) k_grad_sqnorm(const OptChunk* __restrict__ chunks, int nchunks,
                                                             void* const* __restrict__ ptrs, int ntensors,
                                                             double* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  float acc = 0.f;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const OptChunk ch = chunks[c];
    const float* g = static_cast<const float*>(ptrs[ntensors + ch.tensor]);
    if (!g) continue;  // parameter without a gradient this step
    g += ch.off;
    for (int i = threadIdx.x; i < ch.len; i += kOptThreads) {
      const float v = g[i];
      acc += v * v;
    }
  }
  __shared__ float swarp[kOptThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) swarp[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kOptThreads / 32; ++w) s += (double)swarp[w];
    partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(32) k_clip_coef(const double* __restrict__ partial, int nparts, float max_norm,
                                                  float* __restrict__ scalars) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nparts; ++i) s += partial[i];  // fixed order: bit-reproducible
    const float total = (float)sqrt(s);
    float coef = 1.0f;
    if (max_norm > 0.f) {
      coef = max_norm / (total + 1e-6f);
      if (coef > 1.0f) coef = 1.0f;
    }
    scalars[0] = total;
    scalars[1] = coef;
  }
}

__global__ void __launch_bounds__(kOptThreads) k_adam_step(const OptChunk* __restrict__ chunks, int nchunks,
                                                           void* const* __restrict__ ptrs, int ntensors,
                                                           const float* __restrict__ scalars, float lr, float beta1, float beta2,
                                                           float eps, float weight_decay, float decay, int decoupled) {
  pdl_launch_dependents();
  pdl_wait();
  const float coef = scalars[1];
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const OptChunk ch = chunks[c];
    const float* g = static_cast<const float*>(ptrs[ntensors + ch.tensor]);
    if (!g) continue;  // torch skips parameters whose .grad is None (no decay, no moment update)
    g += ch.off;
    float* p = static_cast<float*>(ptrs[ch.tensor]) + ch.off;
    float* m = static_cast<float*>(ptrs[2 * ntensors + ch.tensor]) + ch.off;
    float* v = static_cast<float*>(ptrs[3 * ntensors + ch.tensor]) + ch.off;
    const float2 bc = reinterpret_cast<const float2*>(ptrs + 4 * ntensors)[ch.tensor];  // per-tensor step count
    const float step_size = lr / bc.x, rsqrt_bias2 = bc.y;
    for (int i = threadIdx.x; i < ch.len; i += kOptThreads) {
      float gi = g[i] * coef, pi = p[i];
      if (decoupled) pi *= decay;  // 1 - lr*wd, formed in double on the host as torch does
      else gi += weight_decay * pi;
      const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
      const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      p[i] = pi - step_size * mi / (sqrtf(vi) * rsqrt_bias2 + eps);
    }
  }
}

}  // namespace
}  // namespace refid

extern "C" {

int refid_optim_create(int ntensors, const long* numel, void** handle) {
  using namespace refid;
  REFID_REQUIRE(ntensors > 0 && numel && handle, "optim_create: bad arguments");
  OptState* st = new OptState();
  st->ntensors = ntensors;
  st->numel.assign(numel, numel + ntensors);
  std::vector<OptChunk> chunks;
  for (int t = 0; t < ntensors; ++t)
    for (long off = 0; off < numel[t]; off += kOptChunk) {
      const long len = numel[t] - off < kOptChunk ? numel[t] - off : kOptChunk;
      chunks.push_back(OptChunk{t, (int)len, off});
    }
  st->nchunks = (int)chunks.size();
  REFID_CUDA_CHECK(cudaMalloc(&st->d_chunks, chunks.size() * sizeof(OptChunk)));
  REFID_CUDA_CHECK(cudaMemcpy(st->d_chunks, chunks.data(), chunks.size() * sizeof(OptChunk), cudaMemcpyHostToDevice));
  for (int s = 0; s < OptState::kSlots; ++s) {
    REFID_CUDA_CHECK(cudaMallocHost(&st->h_ptrs[s], 5 * ntensors * sizeof(void*)));
    REFID_CUDA_CHECK(cudaMalloc(&st->d_ptrs[s], 5 * ntensors * sizeof(void*)));
    REFID_CUDA_CHECK(cudaEventCreateWithFlags(&st->done[s], cudaEventDisableTiming));
  }
  REFID_CUDA_CHECK(cudaMalloc(&st->d_partial, kOptBlocks * sizeof(double)));
  REFID_CUDA_CHECK(cudaMalloc(&st->d_scalars, 2 * sizeof(float)));
  *handle = st;
  return 0;
}

int refid_optim_destroy(void* handle) {
  using namespace refid;
  OptState* st = static_cast<OptState*>(handle);
  if (!st) return 0;
  cudaFree(st->d_chunks);
  for (int s = 0; s < OptState::kSlots; ++s) {
    cudaFreeHost(st->h_ptrs[s]);
    cudaFree(st->d_ptrs[s]);
    cudaEventDestroy(st->done[s]);
  }
  cudaFree(st->d_partial);
  cudaFree(st->d_scalars);
  delete st;
  return 0;
}

int refid_optim_step(void* handle, float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                     float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay, const long* steps,
                     int decoupled, float* norm_out, void* stream) {
  using namespace refid;
  OptState* st = static_cast<OptState*>(handle);
  REFID_REQUIRE(st && params && grads && exp_avg && exp_avg_sq && steps, "optim_step: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n = st->ntensors, slot = st->slot;
  st->slot = (slot + 1) % OptState::kSlots;
  REFID_CUDA_CHECK(cudaEventSynchronize(st->done[slot]));  // the staging slot's previous upload has been consumed
  for (int t = 0; t < n; ++t) {
    st->h_ptrs[slot][t] = params[t];
    st->h_ptrs[slot][n + t] = const_cast<float*>(grads[t]);
    st->h_ptrs[slot][2 * n + t] = exp_avg[t];
    st->h_ptrs[slot][3 * n + t] = exp_avg_sq[t];
    const double tstep = (double)(steps[t] < 1 ? 1 : steps[t]);
    float bc[2] = {(float)(1.0 - pow((double)beta1, tstep)), (float)(1.0 / sqrt(1.0 - pow((double)beta2, tstep)))};
    memcpy(&st->h_ptrs[slot][4 * n + t], bc, sizeof(bc));
  }
  REFID_CUDA_CHECK(cudaMemcpyAsync(st->d_ptrs[slot], st->h_ptrs[slot], 5 * n * sizeof(void*), cudaMemcpyHostToDevice, s));
  REFID_CUDA_CHECK(cudaEventRecord(st->done[slot], s));
  const int blocks = st->nchunks < kOptBlocks ? st->nchunks : kOptBlocks;
  float* scalars = norm_out ? norm_out : st->d_scalars;
  REFID_CUDA_CHECK(launch_k(k_grad_sqnorm, dim3(blocks), dim3(kOptThreads), 0, s, (const OptChunk*)st->d_chunks, st->nchunks,
                            (void* const*)st->d_ptrs[slot], n, st->d_partial));
  REFID_CUDA_CHECK(launch_k(k_clip_coef, dim3(1), dim3(32), 0, s, (const double*)st->d_partial, blocks, max_norm, scalars));
  REFID_CUDA_CHECK(launch_k(k_adam_step, dim3(blocks), dim3(kOptThreads), 0, s, (const OptChunk*)st->d_chunks, st->nchunks,
                            (void* const*)st->d_ptrs[slot], n, (const float*)scalars, lr, beta1, beta2, eps, weight_decay,
                            (float)(1.0 - (double)lr * (double)weight_decay), decoupled));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"
