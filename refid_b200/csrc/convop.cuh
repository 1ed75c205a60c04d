// Host-side description of one convolution-shaped op and its lowering onto the tap-GEMM / wgrad kernels.
#pragma once
#include "haloconv.cuh"
#include "halowgrad.cuh"
#include "tapgemm.cuh"
#include "wgrad.cuh"

namespace refid {

enum ConvKind : int {
  CK_3X3 = 0,          // 3x3 stride 1 pad 1 (also its data-gradient, with flipped/transposed packed weights)
  CK_1X1 = 1,          // 1x1
  CK_DOWN4 = 2,        // 4x4 stride 2 pad 1 forward: 16 taps over four stride-2 parity views; output grid H/2 x W/2
  CK_UP2 = 3,          // ConvTranspose 2x2 stride 2 forward: one tap, N blocks = the 4 output parities (stride-2 scatter)
  CK_DOWN4_DGRAD = 4,  // data-gradient of CK_DOWN4 for ONE output parity (4 taps on dY, stride-2 scatter); 4 launches
  CK_UP2_DGRAD = 5,    // data-gradient of CK_UP2: 2x2 stride-2 conv over dOut (4 taps, parity views); grid H/2 x W/2
  CK_ROWS5 = 6,        // 5x5 head conv on an x-unrolled input (channel = kx*Cin + c): 5 vertical taps dy=-2..2
  CK_DOWN4_HALO = 8,       // CK_DOWN4 forward on the halo-conv engine: a 3x3 conv over the four stride-2 parity views of the
                           // input (K = 4*Cin), each view using 4 of the 9 taps (weights [tap9][Cout][parity*Cin])
  CK_DOWN4_DGRAD_HALO = 7  // data-gradient of CK_DOWN4, all four output parities in ONE halo-conv launch: the four parities
                           // read the same 3x3 neighbourhood of dY, each uses 4 of its 9 taps (weights [tap9][parity][Cin][Cout])
};

struct ActSrc {
  const __nv_bfloat16* ptr;
  int C;      // channels read (K extent of this source)
  int pitch;  // elements per pixel in memory
  int nmod;   // this source has nmod images, image n of the launch reads image n % nmod (0: off); halo-conv engine / wgrad
};

struct OutGroup {
  int channels;  // output channels of this group (multiple of the chosen BN)
  EpiDesc epi;   // out/out2/pre/.../bias/act/slope/C/coff set by caller; geometry fields filled by the lowering
};

struct ConvDesc {
  float* bias_grad;  // build_wgrad only: fp32 [Cout] bias-gradient accumulator the launch may fill (see WgradLaunch::bias_done)
  int kind;
  int parity;  // CK_DOWN4_DGRAD only: output parity py*2+px
  ActSrc src[2];
  int nsrc;
  int N, H, W;  // spatial extent of the INPUT (src) tensors
  const __nv_bfloat16* w;  // packed weights [w_rows][w_cols] bf16 (see weights.cuh for the layouts)
  long w_rows;
  int w_cols;
  int wrows_per_tap;
  int w_row0;
  int f16;  // operands and outputs are fp16 (forward-only plans); 0: bf16
  int w_img_rows;        // build_conv: image n uses weight rows n * w_img_rows + ... (tap-GEMM engine, one image per tile)
  long out_img_stride;   // build_wgrad: per-image outputs, floats between consecutive images' matrices (0: one summed matrix)
};
// one image per 128-pixel tile for this grid?  (per-image weights / per-image weight gradients need it)
bool one_image_per_tile(int N, int H, int W);

// Build params once (tensor maps are encoded here), launch many times.
struct TapGemmLaunch {
  TapGemmParams p;
  int BN, BK, n_blocks;
  int use_halo;  // stride-1 3x3 / 1x1 with 64-channel-multiple sources run on the halo-conv engine
  int NM;
  HaloConvParams hp;
  EpiDesc* epi(int i) { return use_halo ? &hp.epi[i] : &p.epi[i]; }
  int num_epi() const { return use_halo ? hp.n_blocks * (BN / hp.epi_seg) : n_blocks; }
};
int build_conv(const ConvDesc& d, const OutGroup* groups, int ngroups, TapGemmLaunch* out);
inline int run_conv(TapGemmLaunch& l, cudaStream_t s) {
  return l.use_halo ? launch_haloconv(l.hp, l.BN, l.NM, s) : launch_tapgemm(l.p, l.BN, l.BK, l.n_blocks, s);
}

struct WgradLaunch {
  WgradParams p;
  int pixel_chunks;
  int use_halo;  // stride-1 3x3 with 64-channel-multiple operands runs on the halo-wgrad kernel
  int bias_done; // the launch also accumulates the bias gradient (ConvDesc::bias_grad)
  HaloWgradParams hp;
};
// d describes the tapped operand P (kind, sources, taps); q is the un-shifted operand on the op's output grid.
// out: fp32 [taps * sum(src C)][q.C], accumulated atomically.
int build_wgrad(const ConvDesc& d, ActSrc q, float* out, WgradLaunch* l);
inline int run_wgrad(WgradLaunch& l, cudaStream_t s) {
  return l.use_halo ? launch_halowgrad(l.hp, s) : launch_wgrad(l.p, l.pixel_chunks, s);
}

}  // namespace refid
