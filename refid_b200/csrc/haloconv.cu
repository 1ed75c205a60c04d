// Halo-conv host side: shared-memory plan and the dispatch onto the two storage flavours (bf16: this unit; fp16:
// haloconv_f16.cu).  The kernel template is in haloconv_kernel.cuh.
#include "haloconv_kernel.cuh"

namespace refid {
int launch_haloconv_f16(const HaloConvParams& p, int BN, int NM, cudaStream_t stream);  // haloconv_f16.cu


size_t halo_max_bias_table() {
  static const size_t v = getenv("REFID_BIAS_KB") ? (size_t)atoi(getenv("REFID_BIAS_KB")) * 1024 : 96 * 1024;
  return v;
}

int haloconv_plan(HaloConvParams* p, int BN, int NM, int mode) {
  if (2 * NM * BN > 512) return 0;
  p->patch_rows = 16 * NM + 2 * p->halo;
  int total_slabs = 0;
  for (int i = 0; i < p->nsrc; ++i) total_slabs += p->src_slabs[i];
  // weights resident for the whole kernel when they fit next to >= 2 activation stages
  if (p->n_blocks == 1 && mode != 2 && !p->w_img_rows) {  // (per-image weights are streamed per item)
    p->resident_b = 1;
    p->stages_b = 0;
    // activation stages: two for convs with fewer than three K slabs (one tile of look-ahead for the single-slab 64 -> 64
    // layers), else three.  REFID_HALO_SA = 3 / 4 (diagnostic) tries more: measured neutral without epilogue operands and 3 %
    // slower with them (tools/conv_bench.py, round 2) -- the activation ring is not what those layers wait for.
    static const int sa_env = getenv("REFID_HALO_SA") ? atoi(getenv("REFID_HALO_SA")) : 2;
    const int sa_hi = sa_env <= 2 ? (total_slabs >= 3 ? 3 : 2) : (sa_env > kHaloMaxStages ? kHaloMaxStages : sa_env);
    for (int sa = sa_hi; sa >= 2; --sa) {
      p->stages_a = sa;
      if (halo_smem_bytes(*p, BN) <= kHaloSmemMax) return 1;
    }
  }
  if (mode == 1) return 0;
  p->resident_b = 0;
  p->stages_a = 2;
  for (int sb = kHaloMaxStages; sb >= 3; --sb) {
    p->stages_b = sb;
    if (halo_smem_bytes(*p, BN) <= kHaloSmemMax) {
      p->stages_a = 3;
      if (halo_smem_bytes(*p, BN) > kHaloSmemMax) p->stages_a = 2;
      return 1;
    }
  }
  return 0;
}

int launch_haloconv(const HaloConvParams& p, int BN, int NM, cudaStream_t stream) {
  return p.f16 ? launch_haloconv_f16(p, BN, NM, stream) : launch_haloconv_flavour<false>(p, BN, NM, stream);
}

}  // namespace refid


#ifdef REFID_HALO_TIMING
// diagnostic build only: per-CTA cycle counters of the halo-conv MMA warp (total, acc wait, A wait, B wait, issue, items)
extern "C" int refid_debug_halo_timing(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, refid::g_halo_t, sizeof(long long) * 148 * 8) == cudaSuccess ? 0 : 1;
}
#endif
