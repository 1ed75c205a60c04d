// Halo-conv: persistent, halo-reusing implicit-GEMM engine for the stride-1 3x3 / 1x1 convolutions (forward and
// data-gradient) that carry ~90 % of the network's FLOPs.
//
// Why a second engine: the tap-GEMM kernel (tapgemm.cuh) fetches one 128 x BK activation tile PER TAP from L2, i.e. 9x
// the input for a 3x3 conv.  At N = Cout = 64 a 128x64x16 MMA takes 32 cycles, so the tensor pipe wants 128 B/clk/SM of
// A operand while the L2 can deliver ~42 B/clk/SM (6300 B/clk chip-wide): the kernel is L2-bound at ~1/3 of peak before
// any fixed cost.  Here ONE (TH+2) x (TW+2) halo patch per 64-channel K slab is staged in shared memory by a single TMA
// box load (zero fill = padding) and all nine taps are issued from it with shifted UMMA descriptors (start address moved
// by whole 128-byte pixel rows; the swizzle phase is carried in the descriptor's base-offset field).  The kernel is
// persistent (one CTA per SM loops over tiles), accumulators are double-buffered in TMEM so the epilogue of tile i
// overlaps the MMAs of tile i+1, weights are either resident in shared memory for the whole kernel (C = 64 layers) or
// streamed tap by tap through their own ring, and one CTA iteration can process two vertically adjacent pixel tiles
// (M = 256) against the same weight tiles.
#pragma once
#include "tapgemm.cuh"

namespace refid {

struct HaloConvParams {
  CUtensorMap tmA[4];  // per source: dims (C, W, H, N), box (64, pitch_px, patch_rows, 1), 128B swizzle (the stride-2 `down`
                       // forward conv has four sources: the stride-2 parity views of its input)
  CUtensorMap tmB;     // packed weights [rows][K], box (64, BN)
  EpiDesc epi[kMaxNBlocks];
  int epi_seg;      // output channels per EpiDesc (power of two, divides BN)
  int epi_shift;    // log2(epi_seg)
  unsigned short tap_mask[kMaxNBlocks];  // per N block: taps to execute (0 = all); streamed-weight path only
  int epi_inputs;   // some EpiDesc reads a global operand (pre, pre2, sv, post)
  int kc;           // channels per K slab: 64 (128B-swizzled pixel rows) or 32 (64B)
  int num_taps;     // 9 or 1
  int halo;         // 1 (3x3) or 0 (1x1)
  int pitch_px;     // pixels per patch row in shared memory (multiple of 8)
  int patch_rows;   // 16*NM + 2*halo
  int wrows_per_tap, w_row0;
  int nsrc, src_slabs[4];
  int src_nmod[4];  // source s holds src_nmod[s] images that repeat along the launch's image axis (0: off)
  unsigned short slab_mask[16];  // per K slab: taps to execute (0 = all); with `masked`, resident weight tiles are compact
  unsigned char slab_b0[16];     // masked + resident: index of the slab's first weight tile
  int masked, resident_tiles;
  int n_blocks;
  int tiles_x, tiles_y, N, H, W;  // tiles are 8 wide x 16*NM high
  int num_items;                  // tiles * n_blocks
  int resident_b;                 // weights stay in shared memory for the whole kernel (n_blocks == 1)
  int stages_a, stages_b;
  int f16;  // activations / packed weights are fp16 (forward-only plans) instead of bf16
  int w_img_rows;   // > 0: image n reads weight rows n * w_img_rows + ... (per-sample weights; streamed-weight path)
  int bias_images;  // > 0: the bias table holds one row per image (EpiDesc::bias_nstride), bias_images = N
  int epi_l2pf;     // epilogue warps L2-prefetch the next item's global operands
  int f32_rmw;      // diagnostic: fp32 accumulation targets by read-modify-write instead of vector reductions
  // Tail splitting (two-block tiles, NM == 2): the persistent CTAs process `full_items` whole items (a multiple of the grid)
  // and then the remaining items as HALF items (one 16-row block each) -- when the remainder is at most half a round, the
  // last round costs half: 512 tiles on 148 SMs take 3.5 instead of 4 rounds.  0: off.
  int tail_split, full_items;
};

// Largest per-sample bias table (one row of n_blocks*BN floats per image) a launch may keep in shared memory.  REFID_BIAS_KB
// (diagnostic) overrides the default of 96 KB (192 images x 128 channels: EGACA's data-gradient over a chunk of time steps;
// all 23 steps of the benchmark batch of 8 -- measured 1 % faster than chunks of 8 steps with a 32 KB table).
size_t halo_max_bias_table();
#define kHaloMaxBiasTable (::refid::halo_max_bias_table())
int launch_haloconv(const HaloConvParams& p, int BN, int NM, cudaStream_t stream);
// Shared-memory plan; returns 0 when the configuration does not fit.  mode 0: resident weights if they fit, else
// streamed; 1: resident only; 2: streamed only.
int haloconv_plan(HaloConvParams* p, int BN, int NM, int mode = 0);

}  // namespace refid
