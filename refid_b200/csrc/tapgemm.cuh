// Tap-GEMM: the im2col-free implicit-GEMM convolution engine (forward conv, data-gradient, transposed conv).
//
//   out[p, n] = epilogue( sum_{src} sum_{tap} sum_{c} A_src[ map_tap(p) ][c] * Wp[tap][n][k(src,c)] )
//
// p runs over a (N,H,W) grid of output-grid pixels in tiles of 128 (TN x TH x TW); map_tap is a spatial offset
// (dy,dx) in the coordinate space of a TMA tensor map (plain NHWC view, or one of four stride-2 parity views), so
// the 128 x BK activation tile of every tap is fetched by ONE cp.async.bulk.tensor with hardware zero padding.
// tcgen05.mma (M=128, N=BN, K=16) accumulates in TMEM; 4 epilogue warps apply bias / residual / activation or
// activation-derivative masks and write bf16 NHWC (optionally strided, e.g. stride-2 scatter of a transposed conv).
#pragma once
#include "common.cuh"

namespace refid {

constexpr int kMaxTaps = 16;
constexpr int kMaxNBlocks = 8;

// ACT_MULT: backward only -- the saved tensor already holds the activation's derivative (GELU layers store gelu'(z) next
// to gelu(z) in the forward pass, so the backward multiplies instead of evaluating erf + exp per element again)
enum ActKind : int { ACT_NONE = 0, ACT_LRELU = 1, ACT_GELU = 2, ACT_MULT = 3 };

// Epilogue of one N block (all pointers share pixel mapping, channel pitch C and channel offset coff).
//   a = acc + bias[n*bias_nstride + c] + pre + pre2
//   if sv_bits (halo engine): a *= bit ? 1 : slope
//   if sv:  a *= act'(sv)      (LRELU: sv>0 ? 1 : slope, evaluated on the saved OUTPUT; MULT: a *= sv, sv = saved derivative)
//   else :  if out_pre: out_pre = (act == GELU ? gelu'(a) : a);   a = act(a)
//   out = a;  out_f32 += a;  out2 = a + post;  out_nchw[c < nchw_C] = a
struct EpiDesc {
  __nv_bfloat16* out;
  __nv_bfloat16* out2;
  __nv_bfloat16* out_pre;
  float* out_f32;
  float* out_nchw;      // optional fp32 NCHW store of the first nchw_C channels (pred): [n*nchw_nstride + c*OH*OW + y*OW + x]
  long nchw_nstride;
  int nchw_C;
  int nchw_B;           // > 0: the launch's image n is (step n / nchw_B, sample n % nchw_B) of a (B,T,C,H,W) tensor:
  long nchw_tstride;    //      [(n % nchw_B)*nchw_nstride + (n / nchw_B)*nchw_tstride + ...] (a chunk of time steps in one launch)
  const __nv_bfloat16* post;
  const __nv_bfloat16* pre;
  const __nv_bfloat16* pre2;
  const __nv_bfloat16* sv;
  const float* bias;
  int bias_nstride;
  int C, coff;
  int osy, osx, ooy, oox, OH, OW;
  int act;
  float slope;
  // Activation sign bits (halo-conv engine, training plans): one 32-bit word per pixel and 32-channel group, bit i = the
  // pre-activation of channel 32*g + i is > 0.  A LeakyReLU / ReLU forward epilogue writes them (out_bits); the backward
  // epilogue that targets the tensor reads them (sv_bits) INSTEAD of the 16-bit tensor itself: 1/16 of the bytes for the
  // derivative mask, the most common epilogue operand of the data-gradient launches.
  uint32_t* out_bits;
  const uint32_t* sv_bits;
  int out_bits_pitch, sv_bits_pitch;  // words per pixel
};

struct TapGemmParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  EpiDesc epi[kMaxNBlocks];
  int num_taps;
  signed char tap_dy[kMaxTaps], tap_dx[kMaxTaps], tap_map[kMaxTaps];
  int wrows_per_tap;  // weight rows per tap (all N blocks)
  int w_row0;         // first weight row of this launch
  int nsrc;           // K-concatenated sources (1 or 2); src s uses tmA[s] unless parity_mode
  int src_slabs[2];   // BK-slabs per source
  int parity_mode;    // tap_map selects tmA (stride-2 views); single source
  int TW, TH, TN, tiles_x, tiles_y;
  int N, H, W;  // output-grid extent
  int num_stages;
  int f16;  // fp16 instead of bf16 storage (forward-only plans)
  int w_img_rows;  // > 0: image n uses the weight rows [n * w_img_rows, ...) (per-sample weights: EGACA's gate folded into
                   // conv3, fusion_modules.py:312-317); needs TN == 1
};

// Host: geometry helper -- picks TW x TH x TN = 128 for an (N,H,W) grid.
void pick_tile(int N, int H, int W, int* TW, int* TH, int* TN);
// Host: launch. grid = (tiles, n_blocks).
int launch_tapgemm(TapGemmParams& p, int BN, int BK, int n_blocks, cudaStream_t stream);

}  // namespace refid
