// Memory-bound kernels (see elementwise.cuh).  Every thread moves 16-byte vectors (8 bf16 channels); reductions are
// done in registers -> warp shuffles -> shared memory -> one atomicAdd per block and channel.
#include "elementwise.cuh"

namespace refid {

namespace {

constexpr int kEwThreads = 256;

inline unsigned blocks_for(long n, int per_block) {
  long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  return (unsigned)b;
}

__device__ __forceinline__ float act_mask(float sv, int act, float slope) {
  if (act == ACT_LRELU) return sv > 0.f ? 1.f : slope;
  if (act == ACT_MULT) return sv;  // saved derivative (GELU layers)
  return 1.f;
}

// ---------------------------------------------------------------------------------------------
// input / output-gradient layout conversion
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads) k_unroll5(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int B,
                                                        int T, int Cin, int H, int W, int Kp, int f16, int t0, int Tn) {
  pdl_launch_dependents();
  pdl_wait();
  // (kx, c) of every unrolled channel k = kx*Cin + c, once per block (the per-element integer divisions made the r1 kernel
  // instruction-bound: 0.8 ms for the event head's input); channel group fastest, so consecutive threads write consecutive
  // 16-byte chunks of one pixel.  One block walks whole pixel rows: no 64-bit division in the loop.
  __shared__ short2 kmap[320];
  for (int k = threadIdx.x; k < Kp; k += blockDim.x) kmap[k] = make_short2((short)(k / Cin), (short)(k % Cin));
  __syncthreads();
  const int G = Kp / 8;
  const int per_row = W * G;
  const long rows = (long)B * Tn * H;
  const size_t plane = (size_t)H * W;
  for (long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int y = (int)(row % H);
    const int n_out = (int)(row / H);
    const int t = t0 + n_out / B, b = n_out % B;  // time steps [t0, t0+Tn) of the T in the input; output image (t-t0)*B + b
    const float* src = in + ((size_t)(b * T + t) * Cin) * plane + (size_t)y * W;
    __nv_bfloat16* dst = out + (size_t)row * W * Kp;
    for (int e = threadIdx.x; e < per_row; e += blockDim.x) {
      const int g = e % G, x = e / G;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const short2 kc = kmap[g * 8 + i];
        const int xx = x + kc.x - 2;
        v[i] = (kc.x < 5 && xx >= 0 && xx < W) ? __ldg(src + (size_t)kc.y * plane + xx) : 0.f;
      }
      store8_rt(dst + (size_t)x * Kp + g * 8, v, f16);
    }
  }
}

__global__ void __launch_bounds__(kEwThreads) k_gout_pack(const float* __restrict__ gout, __nv_bfloat16* __restrict__ out, int B,
                                                          int T, int Cv, int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = (long)B * T * H * W;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long hw = (long)H * W;
    const long pix = idx % hw;
    const int n_out = (int)(idx / hw);
    const int t = n_out / B, b = n_out % B;
    const float* src = gout + ((size_t)(b * T + t) * Cv) * hw + pix;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < Cv && c < 8; ++c) v[c] = __ldg(src + (size_t)c * hw);
    __nv_bfloat16* o = out + (size_t)idx * 32;
    store8(o, v);
    const uint4 z = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4*>(o)[1] = z;
    reinterpret_cast<uint4*>(o)[2] = z;
    reinterpret_cast<uint4*>(o)[3] = z;
  }
}

// ---------------------------------------------------------------------------------------------
// column sum (bias gradients)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads) k_colsum(const __nv_bfloat16* __restrict__ in, long rows, int C,
                                                       float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[kEwThreads][9];
  const int G = C / 8;                 // vector groups per row (4..32)
  const int lanes = kEwThreads / G;    // row lanes per block
  const int g = threadIdx.x % G, rl = threadIdx.x / G;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (rl < lanes) {
    for (long r = (long)blockIdx.x * lanes + rl; r < rows; r += (long)gridDim.x * lanes) {
      float v[8];
      load8(in + (size_t)r * C + g * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = acc[i];
  __syncthreads();
  if (threadIdx.x < C) {
    const int c = threadIdx.x, gg = c / 8, i = c % 8;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[l * G + gg][i];
    atomicAdd(out + c, s);
  }
}

// ---------------------------------------------------------------------------------------------
// masked accumulate
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads) k_addmask(const AddMaskArgs p) {
  pdl_launch_dependents();
  pdl_wait();
  const long n8 = p.n / 8;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float t[8];
    if (p.a) {
      load8(p.a + i * 8, t);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += t[k];
    }
    if (p.b) {
      load8(p.b + i * 8, t);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += t[k];
    }
    if (p.f) {
      const float4 f0 = reinterpret_cast<const float4*>(p.f + i * 8)[0], f1 = reinterpret_cast<const float4*>(p.f + i * 8)[1];
      v[0] += f0.x; v[1] += f0.y; v[2] += f0.z; v[3] += f0.w;
      v[4] += f1.x; v[5] += f1.y; v[6] += f1.z; v[7] += f1.w;
    }
    if (p.dstf) {
      float4* o = reinterpret_cast<float4*>(p.dstf + i * 8);
      float4 f0 = o[0], f1 = o[1];
      f0.x += v[0]; f0.y += v[1]; f0.z += v[2]; f0.w += v[3];
      f1.x += v[4]; f1.y += v[5]; f1.z += v[6]; f1.w += v[7];
      o[0] = f0;
      o[1] = f1;
    }
    if (p.dst) {
      if (p.dst_acc) {
        load8(p.dst + i * 8, t);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += t[k];
      }
      if (p.sv) {
        load8(p.sv + i * 8, t);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] *= act_mask(t[k], p.act, p.slope);
      }
      store8(p.dst + i * 8, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over 64 channels: 8 lanes per pixel, 8 channels per lane
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sum8lanes(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

__device__ __forceinline__ void ln_stats(const float* x, float& mu, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
  mu = sum8lanes(s) * (1.f / 64.f);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) q += (x[k] - mu) * (x[k] - mu);
  rstd = rsqrtf(sum8lanes(q) * (1.f / 64.f) + 1e-6f);
}

__global__ void __launch_bounds__(kEwThreads) k_ln_fwd(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                       long npix, int f16) {
  pdl_launch_dependents();
  pdl_wait();
  const long nvec = npix * 8;
  const long iters = (nvec + (long)gridDim.x * blockDim.x - 1) / ((long)gridDim.x * blockDim.x);
  for (long it = 0; it < iters; ++it) {  // uniform trip count: shuffles need whole warps
    const long i = it * gridDim.x * blockDim.x + (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = i < nvec;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (ok) load8_rt(x + i * 8, v, f16);
    float mu, rstd;
    ln_stats(v, mu, rstd);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (v[k] - mu) * rstd;
    if (ok) store8_rt(y + i * 8, v, f16);
  }
}

__global__ void __launch_bounds__(kEwThreads) k_ln_bwd(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gy,
                                                       const __nv_bfloat16* __restrict__ add, __nv_bfloat16* dst, int dst_acc,
                                                       float* dstf, long npix) {
  pdl_launch_dependents();
  pdl_wait();
  const long nvec = npix * 8;
  const long iters = (nvec + (long)gridDim.x * blockDim.x - 1) / ((long)gridDim.x * blockDim.x);
  for (long it = 0; it < iters; ++it) {
    const long i = it * gridDim.x * blockDim.x + (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = i < nvec;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0}, g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (ok) {
      load8(x + i * 8, v);
      load8(gy + i * 8, g);
    }
    float mu, rstd;
    ln_stats(v, mu, rstd);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = (v[k] - mu) * rstd;  // xhat
      sg += g[k];
      sgx += g[k] * v[k];
    }
    sg = sum8lanes(sg) * (1.f / 64.f);
    sgx = sum8lanes(sgx) * (1.f / 64.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = rstd * (g[k] - sg - v[k] * sgx);
    if (!ok) continue;
    float t[8];
    if (add) {
      load8(add + i * 8, t);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += t[k];
    }
    if (dstf) {
      float4* o = reinterpret_cast<float4*>(dstf + i * 8);
      float4 f0 = o[0], f1 = o[1];
      f0.x += v[0]; f0.y += v[1]; f0.z += v[2]; f0.w += v[3];
      f1.x += v[4]; f1.y += v[5]; f1.z += v[6]; f1.w += v[7];
      o[0] = f0;
      o[1] = f1;
    }
    if (dst) {
      if (dst_acc) {
        load8(dst + i * 8, t);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += t[k];
      }
      store8(dst + i * 8, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// depthwise 3x3 (C = 64), forward and backward.  A block stages the halo patch of an 8 x 32 pixel tile (10 x 34 pixels x
// 128 B, zero-filled outside the image) in shared memory with cp.async and every thread walks one (column, 8-channel group)
// down the tile's rows, reading its nine neighbours from shared memory: the nine-fold neighbourhood re-reads cost shared-
// memory bandwidth instead of L2 round trips (the per-pixel version was latency-bound at 5-9x its HBM time).
// ---------------------------------------------------------------------------------------------
constexpr int kDwPW = kDwTW + 2, kDwPH = kDwTH + 2;
constexpr int kDwPatchBytes = kDwPH * kDwPW * 128;  // 43520

__device__ __forceinline__ void dw_stage_patch(uint8_t* patch, const __nv_bfloat16* __restrict__ src, int n, int y0, int x0, int H,
                                               int W) {
  const uint32_t base = smem_u32(patch);
  for (int i = threadIdx.x; i < kDwPH * kDwPW * 8; i += kEwThreads) {
    const int px = i >> 3, chunk = i & 7;
    const int y = y0 - 1 + px / kDwPW, x = x0 - 1 + px % kDwPW;
    const bool ok = y >= 0 && y < H && x >= 0 && x < W;
    const __nv_bfloat16* g = src + (((size_t)n * H + (ok ? y : 0)) * W + (ok ? x : 0)) * 64 + chunk * 8;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(base + (uint32_t)i * 16u), "l"(g), "r"(ok ? 16 : 0) : "memory");
  }
}
__device__ __forceinline__ void dw_stage_wait() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();
}
__device__ __forceinline__ void lds8_rt(const uint8_t* p, float* f, int f16) {
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  if (f16) {
    f[0] = cvt_lo<true>(a.x); f[1] = cvt_hi<true>(a.x); f[2] = cvt_lo<true>(a.y); f[3] = cvt_hi<true>(a.y);
    f[4] = cvt_lo<true>(a.z); f[5] = cvt_hi<true>(a.z); f[6] = cvt_lo<true>(a.w); f[7] = cvt_hi<true>(a.w);
    return;
  }
  f[0] = bf16_lo(a.x); f[1] = bf16_hi(a.x); f[2] = bf16_lo(a.y); f[3] = bf16_hi(a.y);
  f[4] = bf16_lo(a.z); f[5] = bf16_hi(a.z); f[6] = bf16_lo(a.w); f[7] = bf16_hi(a.w);
}
__device__ __forceinline__ void lds8(const uint8_t* p, float* f) {
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  f[0] = bf16_lo(a.x); f[1] = bf16_hi(a.x); f[2] = bf16_lo(a.y); f[3] = bf16_hi(a.y);
  f[4] = bf16_lo(a.z); f[5] = bf16_hi(a.z); f[6] = bf16_lo(a.w); f[7] = bf16_hi(a.w);
}

// grid = (tiles per image, N): a block never straddles two samples, so the pooled sums are written as per-block partials
// [N][dw_pool_parts(H, W)][64] and reduced in a fixed order by k_se_fwd (bit-reproducible forward pass)
__global__ void __launch_bounds__(kEwThreads, 2) k_dw_fwd(const __nv_bfloat16* __restrict__ a, const float* __restrict__ w,
                                                       const float* __restrict__ bias, __nv_bfloat16* __restrict__ d,
                                                       __nv_bfloat16* __restrict__ g, float* pool, int N, int H, int W, int f16) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) uint8_t patch[kDwPatchBytes];
  __shared__ float spool[8][64];
  const int tiles_x = (W + kDwTW - 1) / kDwTW;
  const int n = blockIdx.y, y0 = (blockIdx.x / tiles_x) * kDwTH, x0 = (blockIdx.x % tiles_x) * kDwTW;
  dw_stage_patch(patch, a, n, y0, x0, H, W);
  const int grp = threadIdx.x & 7, col = threadIdx.x >> 3;
  float wr[8][9], br[8];  // this thread's 8 channels: weights and bias in registers
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    br[k] = bias[grp * 8 + k];
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[k][t] = w[(grp * 8 + k) * 9 + t];
  }
  dw_stage_wait();
  const int x = x0 + col;
  float psum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
  for (int r = 0; r < kDwTH; ++r) {
    const int y = y0 + r;
    if (y >= H || x >= W) break;  // (x is loop-invariant: a column outside the image does nothing)
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = br[k];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float v[8];
        lds8_rt(patch + ((r + ky) * kDwPW + col + kx) * 128 + grp * 16, v, f16);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v[k] * wr[k][ky * 3 + kx];
      }
    }
    float gv[8], dg[8];  // the "pre-activation" buffer holds gelu'(z): the backward multiplies (ACT_MULT)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      gelu_both_f(acc[k], &gv[k], &dg[k]);
      psum[k] += gv[k];
    }
    const size_t off = (((size_t)n * H + y) * W + x) * 64 + grp * 8;
    if (d) store8_rt(d + off, dg, f16);  // forward-only plans keep no derivative
    store8_rt(g + off, gv, f16);
  }
  if (pool) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      psum[k] += __shfl_xor_sync(0xffffffffu, psum[k], 8);
      psum[k] += __shfl_xor_sync(0xffffffffu, psum[k], 16);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) spool[warp][lane * 8 + k] = psum[k];
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      float s = 0.f;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) s += spool[wq][threadIdx.x];
      pool[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 64 + threadIdx.x] = s;
    }
  }
}

// Backward: ga[p] = sum_t gd[p - off_t] w[t];  gw[c][t] = sum_p a[p + off_t] gd[p];  gb[c] = sum_p gd[p].
// Persistent blocks, one per SM, patches double-buffered (the next tile is staged while this one is computed): the weight-
// gradient partials stay in registers over all of a block's tiles, so every output address receives one atomic per block.
constexpr int kDwBwdSmem = 4 * kDwPatchBytes;

__global__ void __launch_bounds__(kEwThreads, 1) k_dw_bwd(const __nv_bfloat16* __restrict__ gd, const __nv_bfloat16* __restrict__ a,
                                                          const float* __restrict__ w, __nv_bfloat16* __restrict__ ga, float* gw,
                                                          float* gb, int N, int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t dw_smem[];
  __shared__ __align__(16) float sw[9 * 64];  // [tap][channel]
  __shared__ float sred[64 * 10];
  for (int i = threadIdx.x; i < 64 * 9; i += blockDim.x) sw[(i % 9) * 64 + i / 9] = w[i];
  for (int i = threadIdx.x; i < 64 * 10; i += blockDim.x) sred[i] = 0.f;
  const int grp = threadIdx.x & 7, col = threadIdx.x >> 3;
  const int tiles_x = (W + kDwTW - 1) / kDwTW, tiles_y = (H + kDwTH - 1) / kDwTH;
  const int total = N * tiles_y * tiles_x;
  float aw[8][9];
  float ab[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    ab[k] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) aw[k][t] = 0.f;
  }
  auto stage = [&](int tile, int buf) {
    const int n = tile / (tiles_y * tiles_x), y0 = ((tile / tiles_x) % tiles_y) * kDwTH, x0 = (tile % tiles_x) * kDwTW;
    dw_stage_patch(dw_smem + (size_t)(2 * buf) * kDwPatchBytes, a, n, y0, x0, H, W);
    dw_stage_patch(dw_smem + (size_t)(2 * buf + 1) * kDwPatchBytes, gd, n, y0, x0, H, W);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if ((int)blockIdx.x < total) stage(blockIdx.x, 0);
  int buf = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, buf ^= 1) {
    const int n = tile / (tiles_y * tiles_x), y0 = ((tile / tiles_x) % tiles_y) * kDwTH, x0 = (tile % tiles_x) * kDwTW;
    const bool more = tile + (int)gridDim.x < total;
    __syncthreads();  // the readers of the other buffer (previous tile) are done; sw / sred are initialised
    if (more) stage(tile + gridDim.x, buf ^ 1);
    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const uint8_t* pa = dw_smem + (size_t)(2 * buf) * kDwPatchBytes;
    const uint8_t* pg = pa + kDwPatchBytes;
    const int x = x0 + col;
#pragma unroll 1
    for (int r = 0; r < kDwTH; ++r) {
      const int y = y0 + r;
      float gc[8];
      lds8(pg + ((r + 1) * kDwPW + col + 1) * 128 + grp * 16, gc);  // gd[p] (zero outside the image)
#pragma unroll
      for (int k = 0; k < 8; ++k) ab[k] += gc[k];
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int t = ky * 3 + kx;
          float v[8];
          lds8(pa + ((r + ky) * kDwPW + col + kx) * 128 + grp * 16, v);  // a[p + off_t]
#pragma unroll
          for (int k = 0; k < 8; ++k) aw[k][t] += v[k] * gc[k];
          lds8(pg + ((r + 2 - ky) * kDwPW + col + 2 - kx) * 128 + grp * 16, v);  // gd[p - off_t]
          const float4 w0 = *reinterpret_cast<const float4*>(sw + t * 64 + grp * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(sw + t * 64 + grp * 8 + 4);
          acc[0] += v[0] * w0.x; acc[1] += v[1] * w0.y; acc[2] += v[2] * w0.z; acc[3] += v[3] * w0.w;
          acc[4] += v[4] * w1.x; acc[5] += v[5] * w1.y; acc[6] += v[6] * w1.z; acc[7] += v[7] * w1.w;
        }
      }
      if (y < H && x < W) store8(ga + (((size_t)n * H + y) * W + x) * 64 + grp * 8, acc);
    }
  }
  // reduce the 32 columns that share a channel group: lanes grp, grp+8, grp+16, grp+24 of each warp, then the 8 warps
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float v = aw[k][t];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      aw[k][t] = v;
    }
    float v = ab[k];
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    ab[k] = v;
  }
  __syncthreads();
  if ((threadIdx.x & 31) < 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(&sred[(grp * 8 + k) * 9 + t], aw[k][t]);
      atomicAdd(&sred[64 * 9 + grp * 8 + k], ab[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 9; i += blockDim.x) atomicAdd(gw + i, sred[i]);
  if (threadIdx.x < 64) atomicAdd(gb + threadIdx.x, sred[64 * 9 + threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------
// squeeze-excite MLP: one block of 64 threads per sample
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_se_fwd(const float* __restrict__ pool_part, int parts, float inv_hw, SeParams p, float* s,
                                                float* save_mean, float* save_z) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float m[64], z[32], red[4][64];
  const int n = blockIdx.x, c = threadIdx.x & 63, part = threadIdx.x >> 6;  // 4 threads per channel
  {
    // fixed-order reduction of the depthwise kernel's per-block partial sums (bit-reproducible): each of the 4 threads of
    // a channel sums every 4th partial in order, the four results are added in order
    float sum = 0.f;
    for (int b = part; b < parts; b += 4) sum += pool_part[((size_t)n * parts + b) * 64 + c];
    red[part][c] = sum;
  }
  __syncthreads();
  if (part == 0) {
    m[c] = (((red[0][c] + red[1][c]) + red[2][c]) + red[3][c]) * inv_hw;
    save_mean[n * 64 + c] = m[c];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int j = threadIdx.x;
    float acc = p.b1[j];
    for (int i = 0; i < 64; ++i) acc += p.w1[j * 64 + i] * m[i];
    z[j] = fmaxf(acc, 0.f);
    save_z[n * 32 + j] = z[j];
  }
  __syncthreads();
  if (part == 0) {
    float acc = p.b2[c];
    for (int j = 0; j < 32; ++j) acc += p.w2[c * 32 + j] * z[j];
    const float sig = 1.f / (1.f + __expf(-acc));
    s[n * 64 + c] = sig;
    red[0][c] = sig;
  }
  // The gate folded into the consuming 1x1 conv (fusion_modules.py:312-317: x*se, x_e*se, then conv3 on their concat):
  // conv3(cat(g_i*s, g_e*s)) = (W3 . diag(s,s)) cat(g_i, g_e), so this sample's conv3 weights are scaled on their K side
  // here and the gated 128-channel tensor is never written.  w3: fp32 [k = 128][co = 64] (the flat gradient layout);
  // wf: 16-bit [co][k] (forward operand), wd: [k][co] (data-gradient operand), both per sample.
  if (p.w3) {
    __syncthreads();
    const float* gate = red[0];
    __nv_bfloat16* wf = p.wf + (size_t)n * 8192;
    // forward layout [co][k]: thread -> (co, 8 consecutive k): one 16-byte store; reads of w3 are strided by 64 floats (L2)
    for (int i = threadIdx.x; i < 1024; i += 256) {
      const int co = i >> 4, k0 = (i & 15) * 8;
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = p.w3[(k0 + q) * 64 + co] * gate[(k0 + q) & 63];
      store8_rt(wf + co * 128 + k0, v, p.f16);
    }
    if (p.wd) {  // data-gradient layout [k][co]: same order as w3 -- coalesced both ways
      __nv_bfloat16* wd = p.wd + (size_t)n * 8192;
      for (int i = threadIdx.x; i < 1024; i += 256) {
        const int k = i >> 3, c0 = (i & 7) * 8;
        const float g = gate[k & 63];
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = p.w3[k * 64 + c0 + q] * g;
        store8_rt(wd + k * 64 + c0, v, p.f16);
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_se_bwd(const float* __restrict__ gs, const float* __restrict__ s,
                                                const float* __restrict__ save_mean, const float* __restrict__ save_z, float inv_hw,
                                                SeParams p, float* gpool) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float gq[64], gz[32], m[64], z[32], red[4][64];
  const int n = blockIdx.x, c = threadIdx.x & 63, part = threadIdx.x >> 6;
  if (p.mwg) {
    // gate folded into conv3: M = per-sample weight gradient of the scaled conv, [k = 128][co = 64]:
    //   dL/ds[c] = sum_{k in {c, c+64}} sum_co W3[k][co] M[k][co]      (dL/dW3 is formed once per direction: k_gate_wgrad)
    // 4 threads per gate channel, 16 output channels each, rotated so that neighbouring threads hit different L2 sectors
    const float* M = p.mwg + (size_t)n * 8192;
    float a = 0.f;
#pragma unroll 4
    for (int q = 0; q < 16; ++q) {
      const int co = part * 16 + ((q + c) & 15);
      a += p.w3[c * 64 + co] * M[c * 64 + co] + p.w3[(c + 64) * 64 + co] * M[(c + 64) * 64 + co];
    }
    red[part][c] = a;
  }
  __syncthreads();
  if (threadIdx.x >= 64) return;
  const float sv = s[n * 64 + c];
  const float gsv = p.mwg ? ((red[0][c] + red[1][c]) + red[2][c]) + red[3][c] : gs[n * 64 + c];
  gq[c] = gsv * sv * (1.f - sv);  // gradient at the pre-sigmoid logits
  m[c] = save_mean[n * 64 + c];
  if (c < 32) z[c] = save_z[n * 32 + c];
  asm volatile("bar.sync 1, 64;" ::: "memory");
  atomicAdd(p.gb2 + c, gq[c]);
  for (int j = 0; j < 32; ++j) atomicAdd(p.gw2 + c * 32 + j, gq[c] * z[j]);
  if (c < 32) {
    float acc = 0.f;
    for (int o = 0; o < 64; ++o) acc += p.w2[o * 32 + c] * gq[o];
    gz[c] = z[c] > 0.f ? acc : 0.f;
    atomicAdd(p.gb1 + c, gz[c]);
  }
  asm volatile("bar.sync 1, 64;" ::: "memory");
  float acc = 0.f;
  for (int j = 0; j < 32; ++j) {
    acc += p.w1[j * 64 + c] * gz[j];
    atomicAdd(p.gw1 + j * 64 + c, gz[j] * m[c]);
  }
  gpool[n * 64 + c] = acc * inv_hw;
}

// dL/dW3[k][co] += sum_j M_j[k][co] * s_j[k % 64] over all (step, sample) pairs j of one direction: M [count][128][64],
// s [count][64].  One launch per direction after the sweep's back-propagation.
__global__ void __launch_bounds__(256) k_gate_wgrad(const float* __restrict__ M, const float* __restrict__ s, int count,
                                                    float* __restrict__ gw3) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * 256 + threadIdx.x;  // 0 .. 8191
  const int k = (i >> 6) & 63;
  const int j0 = blockIdx.y, dj = gridDim.y;
  float acc = 0.f;
  for (int j = j0; j < count; j += dj) acc += M[(size_t)j * 8192 + i] * __ldg(s + (size_t)j * 64 + k);
  atomicAdd(gw3 + i, acc);
}

// ---------------------------------------------------------------------------------------------
// weight repacking
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads) k_pack(const float* __restrict__ flat, __nv_bfloat16* __restrict__ wpack,
                                                     const PackDesc* __restrict__ descs, int f16) {
  pdl_launch_dependents();
  pdl_wait();
  const PackDesc d = descs[blockIdx.y];
  const long per_tap = (long)d.R * d.Cc;
  const long total = per_tap * d.ntaps;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int t = (int)(i / per_tap);
    const long r = i % per_tap;
    long dst = d.dst_off + (d.dst_tap_stride ? (long)t * d.dst_tap_stride + r : i);
    if (d.transpose && d.dst_pitch) dst = d.dst_off + (long)t * (d.dst_tap_stride ? d.dst_tap_stride : per_tap) + (r / d.R) * d.dst_pitch + r % d.R;
    if (d.tapmap[t] < 0) {
      wpack[dst] = __float2bfloat16(0.f);  // +0 in either format
      continue;
    }
    long src;
    if (d.transpose) {  // dst [t][Cc][R]
      const int j = (int)(r / d.R), ii = (int)(r % d.R);
      src = (long)d.tapmap[t] * per_tap + (long)ii * d.Cc + j;
    } else {
      src = (long)d.tapmap[t] * per_tap + r;
    }
    const float wv = flat[d.src_off + src];
    if (f16) reinterpret_cast<__half*>(wpack)[dst] = __float2half_rn(wv);
    else wpack[dst] = __float2bfloat16(wv);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
int launch_unroll5(const float* in, __nv_bfloat16* out, int B, int T, int Cin, int H, int W, int Kp, cudaStream_t s, int f16,
                   int t0, int Tn) {
  REFID_REQUIRE(Kp % 32 == 0 && Kp >= 5 * Cin, "unroll5: Kp=%d too small for Cin=%d", Kp, Cin);
  if (Tn < 0) Tn = T;
  REFID_REQUIRE(t0 >= 0 && t0 + Tn <= T, "unroll5: steps [%d,%d) outside T=%d", t0, t0 + Tn, T);
  REFID_REQUIRE(Kp <= 320, "unroll5: Kp=%d exceeds the channel table", Kp);
  long rows = (long)B * Tn * H;
  unsigned blocks = rows > 148L * 16 ? 148u * 16u : (unsigned)rows;
  REFID_CUDA_CHECK(launch_k(k_unroll5, dim3(blocks), dim3(kEwThreads), 0, s, in, out, B, T, Cin, H, W, Kp, f16, t0, Tn));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_gout_pack(const float* gout, __nv_bfloat16* out, int B, int T, int Cv, int H, int W, cudaStream_t s) {
  REFID_REQUIRE(Cv <= 8, "gout_pack: out_chn %d > 8 unsupported", Cv);
  const long total = (long)B * T * H * W;
  unsigned blocks = blocks_for(total, kEwThreads);
  if (blocks > 148u * 32u) blocks = 148u * 32u;
  REFID_CUDA_CHECK(launch_k(k_gout_pack, dim3(blocks), dim3(kEwThreads), 0, s, gout, out, B, T, Cv, H, W));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_colsum(const __nv_bfloat16* in, long rows, int C, float* out, cudaStream_t s) {
  REFID_REQUIRE(C % 8 == 0 && C <= 256 && (kEwThreads % (C / 8)) == 0, "colsum: unsupported C=%d", C);
  const int lanes = kEwThreads / (C / 8);
  unsigned blocks = blocks_for(rows, lanes * 16);
  if (blocks > 148u * 4u) blocks = 148u * 4u;
  REFID_CUDA_CHECK(launch_k(k_colsum, dim3(blocks), dim3(kEwThreads), 0, s, in, rows, C, out));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(kEwThreads) k_sum_series(const __nv_bfloat16* __restrict__ base, long slot_elems, int T,
                                                           __nv_bfloat16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long n8 = slot_elems / 8;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 4
    for (int t = 0; t < T; ++t) {
      float u[8];
      load8(base + (size_t)t * slot_elems + i * 8, u);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += u[k];
    }
    store8(out + i * 8, v);
  }
}

// out[t * slot_elems + i] = in[i] for t < k: one read, k writes (image-branch features entering every step of a chunk)
__global__ void __launch_bounds__(kEwThreads) k_repeat(const uint4* __restrict__ in, long n16, int k, uint4* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long)gridDim.x * blockDim.x) {
    const uint4 v = in[i];
    for (int t = 0; t < k; ++t) out[(size_t)t * n16 + i] = v;
  }
}

int launch_repeat(const __nv_bfloat16* in, long slot_elems, int k, __nv_bfloat16* out, cudaStream_t s) {
  REFID_REQUIRE(slot_elems % 8 == 0, "repeat: slot_elems=%ld not a multiple of 8", slot_elems);
  unsigned blocks = blocks_for(slot_elems / 8, kEwThreads);
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  REFID_CUDA_CHECK(launch_k(k_repeat, dim3(blocks), dim3(kEwThreads), 0, s, reinterpret_cast<const uint4*>(in), slot_elems / 8, k,
                            reinterpret_cast<uint4*>(out)));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_sum_series(const __nv_bfloat16* base, long slot_elems, int T, __nv_bfloat16* out, cudaStream_t s) {
  REFID_REQUIRE(slot_elems % 8 == 0, "sum_series: slot_elems=%ld not a multiple of 8", slot_elems);
  unsigned blocks = blocks_for(slot_elems / 8, kEwThreads);
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  REFID_CUDA_CHECK(launch_k(k_sum_series, dim3(blocks), dim3(kEwThreads), 0, s, base, slot_elems, T, out));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_addmask(const AddMaskArgs& a, cudaStream_t s) {
  REFID_REQUIRE(a.n % 8 == 0, "addmask: n=%ld not a multiple of 8", a.n);
  unsigned blocks = blocks_for(a.n / 8, kEwThreads);
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  REFID_CUDA_CHECK(launch_k(k_addmask, dim3(blocks), dim3(kEwThreads), 0, s, a));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_ln_fwd(const __nv_bfloat16* x, __nv_bfloat16* y, long npix, cudaStream_t s, int f16) {
  unsigned blocks = blocks_for(npix * 8, kEwThreads);
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  REFID_CUDA_CHECK(launch_k(k_ln_fwd, dim3(blocks), dim3(kEwThreads), 0, s, x, y, npix, f16));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_ln_bwd(const __nv_bfloat16* x, const __nv_bfloat16* gy, const __nv_bfloat16* add, __nv_bfloat16* dst, int dst_acc,
                  float* dstf, long npix, cudaStream_t s) {
  unsigned blocks = blocks_for(npix * 8, kEwThreads);
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  REFID_CUDA_CHECK(launch_k(k_ln_bwd, dim3(blocks), dim3(kEwThreads), 0, s, x, gy, add, dst, dst_acc, dstf, npix));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_dw_fwd(const __nv_bfloat16* a, const float* w, const float* bias, __nv_bfloat16* d, __nv_bfloat16* g, float* pool,
                  int N, int H, int W, cudaStream_t s, int f16) {
  dim3 grid(dw_pool_parts(H, W), N);
  REFID_CUDA_CHECK(launch_k(k_dw_fwd, dim3(grid), dim3(kEwThreads), 0, s, a, w, bias, d, g, pool, N, H, W, f16));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_dw_bwd(const __nv_bfloat16* gd, const __nv_bfloat16* a, const float* w, __nv_bfloat16* ga, float* gw, float* gb,
                  int N, int H, int W, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    REFID_CUDA_CHECK(cudaFuncSetAttribute(k_dw_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwBwdSmem));
    attr_set = true;
  }
  long blocks = (long)N * dw_pool_parts(H, W);
  if (blocks > 148) blocks = 148;
  REFID_CUDA_CHECK(launch_k(k_dw_bwd, dim3((unsigned)blocks), dim3(kEwThreads), (size_t)kDwBwdSmem, s, gd, a, w, ga, gw, gb, N, H, W));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_se_fwd(const float* pool_part, int parts, float inv_hw, SeParams p, float* s, float* save_mean, float* save_z, int N,
                  cudaStream_t st) {
  REFID_CUDA_CHECK(launch_k(k_se_fwd, dim3(N), dim3(256), 0, st, pool_part, parts, inv_hw, p, s, save_mean, save_z));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_se_bwd(const float* gs, const float* s, const float* save_mean, const float* save_z, float inv_hw, SeParams p,
                  float* gpool, int N, cudaStream_t st) {
  REFID_CUDA_CHECK(launch_k(k_se_bwd, dim3(N), dim3(256), 0, st, gs, s, save_mean, save_z, inv_hw, p, gpool));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_gate_wgrad(const float* M, const float* s, int count, float* gw3, cudaStream_t st) {
  int gy = count < 8 ? count : 8;
  if (gy < 1) gy = 1;
  REFID_CUDA_CHECK(launch_k(k_gate_wgrad, dim3(32, gy), dim3(256), 0, st, M, s, count, gw3));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int launch_pack(const float* flat, __nv_bfloat16* wpack, const PackDesc* descs_dev, int ndesc, long max_elems, cudaStream_t st,
                int f16) {
  unsigned bx = blocks_for(max_elems, kEwThreads * 4);
  if (bx > 1024u) bx = 1024u;
  dim3 grid(bx, ndesc);
  REFID_CUDA_CHECK(launch_k(k_pack, dim3(grid), dim3(kEwThreads), 0, st, flat, wpack, descs_dev, f16));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace refid
