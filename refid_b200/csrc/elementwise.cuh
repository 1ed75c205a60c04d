// Memory-bound kernels around the tap-GEMM engine: layout conversion of the network inputs / output gradient,
// bias-gradient column sums, masked gradient accumulation, and the pieces of EGACA (event-guided adaptive channel
// attention, reference basicsr/models/archs/fusion_modules.py:237-333) that are not 1x1 GEMMs: per-pixel LayerNorm,
// depthwise 3x3 + GELU (+ global pooling) and the squeeze-excite MLP, which also folds the channel gate into the weights of
// the 1x1 conv that consumes the gated features.
// The forward kernels take `f16`: the 16-bit storage is fp16 instead of bf16 (forward-only plans).
// All activations are NHWC bf16, 16-byte vectorised (8 channels per thread); statistics and parameters are fp32.
#pragma once
#include "tapgemm.cuh"

namespace refid {

// fp32 NCHW (frame n_in = b*T + t) -> bf16 NHWC [(t-t0)*B + b][H][W][Kp], channel k = kx*Cin + c holds in[c][y][x+kx-2];
// steps t0 .. t0+Tn-1 (Tn < 0: all T).
int launch_unroll5(const float* in, __nv_bfloat16* out, int B, int T, int Cin, int H, int W, int Kp, cudaStream_t s, int f16 = 0,
                   int t0 = 0, int Tn = -1);
// fp32 (B,T,Cv,H,W) -> bf16 [t*B + b][H][W][32] (channels >= Cv zero).
int launch_gout_pack(const float* gout, __nv_bfloat16* out, int B, int T, int Cv, int H, int W, cudaStream_t s);
// out[c] += sum_rows in[row][c]   (bf16 [rows][C] contiguous, C % 8 == 0, C <= 256)
int launch_colsum(const __nv_bfloat16* in, long rows, int C, float* out, cudaStream_t s);

// dst = mask(sv) * ((dst_acc ? dst : 0) + a + b + f);   dstf += a + b + f (unmasked).  n = element count (% 8 == 0).
struct AddMaskArgs {
  const __nv_bfloat16* a;
  const __nv_bfloat16* b;
  const float* f;
  const __nv_bfloat16* sv;
  int act;
  float slope;
  __nv_bfloat16* dst;
  int dst_acc;
  float* dstf;
  long n;
};
int launch_addmask(const AddMaskArgs& a, cudaStream_t s);

// out[i] = sum_{t < T} base[t * slot_elems + i]   (fp32 accumulation, bf16 result); slot_elems % 8 == 0.
int launch_sum_series(const __nv_bfloat16* base, long slot_elems, int T, __nv_bfloat16* out, cudaStream_t s);

// out[t * slot_elems + i] = in[i], t < k (k copies along the image axis); slot_elems % 8 == 0.
int launch_repeat(const __nv_bfloat16* in, long slot_elems, int k, __nv_bfloat16* out, cudaStream_t s);

// Per-pixel LayerNorm over C = 64 channels without affine (the affine is folded into the following 1x1 conv).
int launch_ln_fwd(const __nv_bfloat16* x, __nv_bfloat16* y, long npix, cudaStream_t s, int f16 = 0);
// val = d(LN)/dx applied to gy;  dst = (dst_acc ? dst : 0) + add + val   or   dstf += val.
int launch_ln_bwd(const __nv_bfloat16* x, const __nv_bfloat16* gy, const __nv_bfloat16* add, __nv_bfloat16* dst, int dst_acc,
                  float* dstf, long npix, cudaStream_t s);

// Depthwise 3x3 (pad 1) over C = 64: z = dw(a) + bias, g = GELU(z), d = GELU'(z) (saved for the backward); pool[n][c] += sum_pix g.
// pool (optional): per-block partial sums [N][dw_pool_parts(H, W)][64] fp32, reduced in fixed order by launch_se_fwd.
constexpr int kDwTH = 8, kDwTW = 32;  // depthwise-conv pixel tile
inline int dw_pool_parts(int H, int W) { return ((H + kDwTH - 1) / kDwTH) * ((W + kDwTW - 1) / kDwTW); }
int launch_dw_fwd(const __nv_bfloat16* a, const float* w, const float* bias, __nv_bfloat16* d, __nv_bfloat16* g, float* pool,
                  int N, int H, int W, cudaStream_t s, int f16 = 0);
// ga = dw^T(gd); gw[c][tap] += sum a[p+tap] gd[p]; gb[c] += sum gd[p].
int launch_dw_bwd(const __nv_bfloat16* gd, const __nv_bfloat16* a, const float* w, __nv_bfloat16* ga, float* gw, float* gb,
                  int N, int H, int W, cudaStream_t s);

// Squeeze-excite MLP on the pooled event statistics: s = sigmoid(W2 relu(W1 mean + b1) + b2), C = 64, hidden 32.
struct SeParams {
  const float* w1;  // (32,64)
  const float* b1;
  const float* w2;  // (64,32)
  const float* b2;
  float* gw1;
  float* gb1;
  float* gw2;
  float* gb2;
  // gate folded into conv3 (null: off): fp32 master weights [128][64]; per-sample scaled 16-bit copies written by the
  // forward ([n][64][128] and, on training plans, [n][128][64]); backward: per-sample weight gradients M [n][128][64] in,
  // conv3's weight gradient out
  const float* w3;
  __nv_bfloat16* wf;
  __nv_bfloat16* wd;
  int f16;
  const float* mwg;
};
int launch_se_fwd(const float* pool_part, int parts, float inv_hw, SeParams p, float* s, float* save_mean, float* save_z, int N,
                  cudaStream_t st);
// gpool[n][c] = inv_hw * dL/dmean[n][c]
int launch_se_bwd(const float* gs, const float* s, const float* save_mean, const float* save_z, float inv_hw, SeParams p,
                  float* gpool, int N, cudaStream_t st);

// conv3's weight gradient with the gate folded in: gw3[k][co] += sum_j M[j][k][co] * s[j][k % 64]
int launch_gate_wgrad(const float* M, const float* s, int count, float* gw3, cudaStream_t st);
// Weight repacking: fp32 [ntaps][R][Cc] (gradient layout) -> bf16, per-tap copy or transpose, taps gathered by tapmap.
struct PackDesc {
  long src_off;  // floats
  long dst_off;  // bf16 elements
  int ntaps, R, Cc, transpose;
  long dst_tap_stride;  // elements between destination taps (0: contiguous, R*Cc)
  int dst_pitch;        // transpose mode: elements between destination rows (0: R)
  signed char tapmap[16];  // source tap of destination tap i; -1: zero block
};
int launch_pack(const float* flat, __nv_bfloat16* wpack, const PackDesc* descs_dev, int ndesc, long max_elems, cudaStream_t st,
                int f16 = 0);

}  // namespace refid
