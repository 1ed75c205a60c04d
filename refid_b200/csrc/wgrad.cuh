// Weight-gradient GEMM:  D[(tap, cP)][cQ] += sum_pixels P[map_tap(p)][cP] * Q[p][cQ]      (fp32, atomically accumulated)
//
// P is the "tapped" operand (the conv input for ordinary convs: same TMA views / tap offsets as the forward
// A operand), Q the un-shifted operand (the output gradient).  Both tiles arrive as [128 pixels][channels] boxes,
// i.e. MN-major UMMA operands with K = pixels; one M tile = 128 rows = (128/CB) channel blocks of (tap, channel-slab).
#pragma once
#include "tapgemm.cuh"

namespace refid {

struct WgradParams {
  CUtensorMap tmP[4];
  CUtensorMap tmQ;
  float* out;  // [taps * CP_total][CQ] fp32, atomically accumulated
  int num_taps;
  signed char tap_dy[kMaxTaps], tap_dx[kMaxTaps], tap_map[kMaxTaps];
  int nsrc, src_blocks[2];  // CB-channel blocks per P source
  int src_nmod[2];          // P source s holds src_nmod[s] images that repeat along the launch's image axis (0: off)
  int parity_mode;
  int CB;           // channel block of P maps (64 or 32)
  int CBq;          // channel block of the Q map (64 or 32)
  int CQ;           // total Q channels (row pitch of out)
  int BNq;          // Q channels per CTA (<=256)
  int num_mtiles;   // ceil(taps*blocks_per_tap / (128/CB))
  int mt_per_cta;   // M tiles accumulated per CTA (mt_per_cta*BNq <= 512)
  int total_rows;   // taps * CP_total
  int TW, TH, TN, tiles_x, tiles_y, num_tiles;
  int N, H, W;
  int num_stages;
  int img_chunks;        // > 0: per-image outputs -- CTA x works on image x / img_chunks only (its tiles x % img_chunks,
  long out_img_stride;   //      + img_chunks, ...) and accumulates into out + image * out_img_stride; needs TN == 1
};

int launch_wgrad(WgradParams& p, int pixel_chunks, cudaStream_t stream);

}  // namespace refid
