// Halo-wgrad: halo-reusing weight-gradient kernel for stride-1 3x3 convolutions (see halowgrad.cu).
#pragma once
#include "tapgemm.cuh"

namespace refid {

struct HaloWgradParams {
  CUtensorMap tmP[4];  // conv input per source (down: the four stride-2 parity views of the one source): dims (C, W, H, N), box (64, 10, 18 | 16, 1), 128B swizzle
  CUtensorMap tmQ;     // output gradient: dims (CQ, W, H, N), box (64, 8, 16, 1)
  float* out;          // [9 * cp_total][CQ] fp32, accumulated with reductions
  float* bias_out;     // optional [CQ] fp32: += column sums of the output gradient (bias gradient), computed from the G
                       // tiles in shared memory by the otherwise idle epilogue warps of one job per Cout block
  int down;            // 4x4 stride-2 conv: sources = parity views, 2 x 2 taps per view (see halowgrad.cu)
  int mode;            // 64: Cout == 64; 128: Cout % 128 == 0 and every source a multiple of 128 channels
  int CQ, cp_total;
  int nsrc, src_slabs[2], total_slabs;
  int tiles_x, tiles_y, N, H, W, num_tiles;
  int jobs, chunks;
  int num_stages;
};

int launch_halowgrad(HaloWgradParams& p, cudaStream_t stream);

}  // namespace refid
