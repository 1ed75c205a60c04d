#include "convop.cuh"

#include <stdlib.h>
#include <string.h>

namespace refid {

namespace {

struct TapTable {
  int n;
  int dy[kMaxTaps], dx[kMaxTaps], map[kMaxTaps];
  int parity_mode;
  int grid_div;  // output grid = input / grid_div
};

// For the stride-2 4x4 conv (pad 1): input row = 2*y + k - 1.
//   forward tap k -> parity view (k+1)&1, offset floor((k-1)/2) in that view.
//   data-gradient for output parity q (row 2*i+q): the two contributing kernel rows and the dY offsets:
//     q=0: k=1 (dY row i), k=3 (dY row i-1);  q=1: k=0 (dY row i+1), k=2 (dY row i).
inline int down_fwd_off(int k) { return k == 0 ? -1 : (k == 3 ? 1 : 0); }
inline int down_dgrad_off(int q, int a) { return q == 0 ? (a == 0 ? 0 : -1) : (a == 0 ? 1 : 0); }

int fill_taps(int kind, int parity, TapTable* t) {
  memset(t, 0, sizeof(*t));
  t->grid_div = 1;
  switch (kind) {
    case CK_3X3:
      t->n = 9;
      for (int i = 0; i < 9; ++i) {
        t->dy[i] = i / 3 - 1;
        t->dx[i] = i % 3 - 1;
      }
      return 0;
    case CK_ROWS5:
      t->n = 5;
      for (int i = 0; i < 5; ++i) t->dy[i] = i - 2;
      return 0;
    case CK_1X1:
    case CK_UP2:
      t->n = 1;
      return 0;
    case CK_DOWN4:
      t->n = 16;
      t->parity_mode = 1;
      t->grid_div = 2;
      for (int ky = 0; ky < 4; ++ky)
        for (int kx = 0; kx < 4; ++kx) {
          const int i = ky * 4 + kx;
          t->dy[i] = down_fwd_off(ky);
          t->dx[i] = down_fwd_off(kx);
          t->map[i] = ((ky + 1) & 1) * 2 + ((kx + 1) & 1);
        }
      return 0;
    case CK_DOWN4_DGRAD: {
      t->n = 4;
      const int py = parity >> 1, px = parity & 1;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          t->dy[a * 2 + b] = down_dgrad_off(py, a);
          t->dx[a * 2 + b] = down_dgrad_off(px, b);
        }
      return 0;
    }
    case CK_DOWN4_DGRAD_HALO:
    case CK_DOWN4_HALO:
      set_error("halo-only conv kind %d needs 64-channel-multiple operands and even H, W", kind);
      return 1;
    case CK_UP2_DGRAD:
      t->n = 4;
      t->parity_mode = 1;
      t->grid_div = 2;
      for (int i = 0; i < 4; ++i) t->map[i] = i;
      return 0;
  }
  set_error("fill_taps: unknown conv kind %d", kind);
  return 1;
}

}  // namespace

// 4-D NHWC map whose box is the halo patch of one 8 x (16*NM) pixel tile: (64 channels, pitch_px, patch_rows, 1 image).
static int make_patch_map(CUtensorMap* m, const ActSrc& s, int N, int H, int W, int pitch_px, int patch_rows, int kc) {
  return make_act_map(m, s.ptr, s.nmod ? s.nmod : N, H, W, s.pitch, s.C, 0, 0, 1, kc, pitch_px, patch_rows, 1);
}

// Returns 1 when the op was lowered onto the halo-conv engine, 0 when it does not qualify, -1 on error.
static int try_build_halo(const ConvDesc& d, const OutGroup* groups, int ngroups, TapGemmLaunch* out) {
  const bool down_dgrad = d.kind == CK_DOWN4_DGRAD_HALO;
  const bool down_fwd = d.kind == CK_DOWN4_HALO;  // 3x3 over the four stride-2 parity views of the single source
  const bool up2 = d.kind == CK_UP2;              // ConvTranspose 2x2 s2: a 1x1 conv whose four N blocks are the output parities
  const bool scatter4 = down_dgrad || up2;        // four replicas of the output groups, scattered with stride 2
  if (d.kind != CK_3X3 && d.kind != CK_1X1 && !down_dgrad && !down_fwd && !up2) return 0;
  if (d.w_img_rows && d.kind != CK_1X1) return 0;  // per-image weights: 1x1 only here (else the tap-GEMM engine)
  if (scatter4 && ngroups != 1) return 0;
  if (down_fwd && (d.nsrc != 1 || d.src[0].C % 64 || (d.H & 1) || (d.W & 1) || 4 * (d.src[0].C / 64) > 16)) return 0;
  const int GH = down_fwd ? d.H / 2 : d.H, GW = down_fwd ? d.W / 2 : d.W;  // grid the pixel tiles run over
  for (int g = 0; g < ngroups; ++g)  // GELU epilogues exist on the halo engine for 1x1 convs only (EGACA); else tap-GEMM
    if (groups[g].epi.act == ACT_GELU && d.kind != CK_1X1) return 0;
  static const int disabled = getenv("REFID_NO_HALO") ? 1 : 0;
  if (disabled) return 0;
  int ktot = 0, kc = 64;
  for (int s = 0; s < d.nsrc; ++s) {
    if (d.src[s].C % 32) return 0;
    if (d.src[s].C % 64) kc = 32;  // 32-channel slabs: 64-byte pixel rows, 64B swizzle
    ktot += d.src[s].C;
  }
  if (down_fwd) ktot *= 4;
  if (ktot != d.w_cols) return 0;
  int total = 0, seg = 1 << 30;
  for (int g = 0; g < ngroups; ++g) {
    if (groups[g].channels % 32) return 0;
    total += groups[g].channels;
    int c = groups[g].channels & -groups[g].channels;  // largest power of two dividing the group
    if (c < seg) seg = c;
  }
  // widest N tile that divides the output channels: measured on B200, 256 beats 128 even for the 64-tile bottleneck convs
  // (the weight stream per pixel halves), and 128/64 splits of C = 256 layers were 20-80 % slower
  int BN = kc == 64 ? 256 : 128;
  while (BN > 32 && total % BN) BN >>= 1;
  if (total % BN) return 0;
  {
    // small grids: a narrower N tile while the launch would otherwise occupy at most half of the SMs (same total weight
    // stream, twice the busy SMs); never below 64 output channels per tile (N = 32 MMAs run at a third of the rate)
    static const int no_small = getenv("REFID_NO_SMALL_TILES") ? 1 : 0;
    static int sms = 0;
    if (!sms && (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0) != cudaSuccess || sms <= 0)) sms = 148;
    auto items_for = [&](int bn) {
      const int nm = (2 * 2 * bn <= 512 && GH > 16) ? 2 : 1;
      return (long)((GW + 7) / 8) * ((GH + 16 * nm - 1) / (16 * nm)) * d.N * (total / bn) * (scatter4 ? 4 : 1);
    };
    while (!no_small && kc == 64 && BN > 64 && items_for(BN) * 2 <= sms && (scatter4 ? 4 : 1) * (total / (BN / 2)) <= kMaxNBlocks) BN >>= 1;
  }
  if (seg > BN) seg = BN;
  if (BN % seg || (total / seg) > kMaxNBlocks) return 0;
  HaloConvParams& h = out->hp;
  memset(&h, 0, sizeof(h));
  h.f16 = d.f16;
  h.w_img_rows = d.w_img_rows;
  for (int g = 0; g < ngroups; ++g)
    if (groups[g].epi.bias && groups[g].epi.bias_nstride) h.bias_images = d.N;
  if (h.bias_images && (size_t)h.bias_images * total * 4 > kHaloMaxBiasTable) return 0;  // per-sample bias table must fit shared memory
  h.num_taps = (d.kind == CK_1X1 || up2) ? 1 : 9;
  h.halo = (d.kind == CK_1X1 || up2) ? 0 : 1;
  h.pitch_px = h.halo ? 10 : 8;
  h.wrows_per_tap = d.wrows_per_tap;
  h.w_row0 = d.w_row0;
  h.nsrc = down_fwd ? 4 : d.nsrc;
  h.kc = kc;
  for (int s = 0; s < h.nsrc; ++s) h.src_slabs[s] = d.src[down_fwd ? 0 : s].C / kc;
  for (int s = 0; s < d.nsrc && !down_fwd; ++s) h.src_nmod[s] = d.src[s].nmod;
  if (down_fwd) {
    // parity view q = (py,px) holds in[2i+py][2j+px]; input row 2Y+ky-1 = 2(Y+dy)+py  =>  py = 0 uses dy in {0,+1},
    // py = 1 uses dy in {-1,0} (same in x): 4 of the 9 taps per view, 16 (view, tap) pairs = the 16 kernel taps
    const int per_q = d.src[0].C / kc;
    h.masked = 1;
    int tiles = 0;
    for (int ks = 0; ks < 4 * per_q; ++ks) {
      const int py = (ks / per_q) >> 1, px = (ks / per_q) & 1;
      unsigned m = 0;
      for (int t9 = 0; t9 < 9; ++t9) {
        const int dy = t9 / 3 - 1, dx = t9 % 3 - 1;
        if ((py ? dy <= 0 : dy >= 0) && (px ? dx <= 0 : dx >= 0)) m |= 1u << t9;
      }
      h.slab_mask[ks] = (unsigned short)m;
      h.slab_b0[ks] = (unsigned char)tiles;
      tiles += 4;
    }
    h.resident_tiles = tiles;
  }
  h.n_blocks = (scatter4 ? 4 : 1) * (total / BN);
  if (h.n_blocks > kMaxNBlocks || (scatter4 && 4 * (total / seg) > kMaxNBlocks)) return 0;
  if (down_dgrad) {
    // N blocks enumerate (output parity q = qy*2+qx, channel block); parity q of dX[2i+qy][2j+qx] uses the taps (dy,dx) of
    // the dY neighbourhood with ky = qy + 1 - 2*dy and kx = qx + 1 - 2*dx inside the 4x4 kernel
    const int per_q = total / BN;
    for (int nb = 0; nb < h.n_blocks; ++nb) {
      const int qy = (nb / per_q) >> 1, qx = (nb / per_q) & 1;
      unsigned m = 0;
      for (int t9 = 0; t9 < 9; ++t9) {
        const int ky = qy + 1 - 2 * (t9 / 3 - 1), kx = qx + 1 - 2 * (t9 % 3 - 1);
        if (ky >= 0 && ky < 4 && kx >= 0 && kx < 4) m |= 1u << t9;
      }
      h.tap_mask[nb] = (unsigned short)m;
    }
  }
  h.epi_seg = seg;
  h.epi_shift = 0;
  while ((1 << h.epi_shift) < seg) ++h.epi_shift;
  h.N = d.N;
  h.H = GH;
  h.W = GW;
  int n_ein = 0;
  for (int g = 0; g < ngroups; ++g) {
    const EpiDesc& e = groups[g].epi;
    n_ein += (e.pre ? 1 : 0) + (e.pre2 ? 1 : 0) + ((e.sv || e.post || e.sv_bits) ? 1 : 0);
  }
  h.epi_inputs = n_ein > 0;
  static const int l2pf = getenv("REFID_EPI_L2PF") ? atoi(getenv("REFID_EPI_L2PF")) : 0;
  h.epi_l2pf = l2pf;
  static const int rmw = getenv("REFID_F32_RMW") ? atoi(getenv("REFID_F32_RMW")) : 0;
  h.f32_rmw = rmw;
  static const int nobits = getenv("REFID_NO_SIGNBITS") ? 1 : 0;  // diagnostic: 16-bit mask operands as in round 1
  if (nobits)
    for (int g = 0; g < ngroups; ++g) const_cast<OutGroup*>(groups)[g].epi.sv_bits = nullptr;
  // preference: resident weights (two pixel tiles per item, else one) before streamed weights -- re-streaming the
  // weights of a C=64 dual-source conv per 256-pixel item costs more than the smaller M (measured 111 vs 83 us MMA-side)
  int nm_pref = (2 * 2 * BN <= 512 && GH > 16) ? 2 : 1;
  {
    // small grids (the reference's own batch of 1 per GPU): when two-block tiles would occupy at most half of the SMs,
    // one-block tiles double the number of busy SMs for the same work
    static const int no_small = getenv("REFID_NO_SMALL_TILES") ? 1 : 0;
    static int sms = 0;
    if (!sms && (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0) != cudaSuccess || sms <= 0)) sms = 148;
    const long items2 = (long)((GW + 7) / 8) * ((GH + 31) / 32) * d.N * h.n_blocks;
    if (!no_small && nm_pref == 2 && items2 * 2 <= sms) nm_pref = 1;
  }
  int NM = 0;
  if (haloconv_plan(&h, BN, nm_pref, 1)) NM = nm_pref;
  else if (haloconv_plan(&h, BN, 1, 1)) NM = 1;
  else if (haloconv_plan(&h, BN, nm_pref, 2)) NM = nm_pref;
  else if (haloconv_plan(&h, BN, 1, 2)) NM = 1;
  if (!NM) return 0;
  h.tiles_x = (GW + 7) / 8;
  h.tiles_y = (GH + 16 * NM - 1) / (16 * NM);
  h.num_items = h.tiles_x * h.tiles_y * d.N * h.n_blocks;
  if (down_fwd) {
    for (int q = 0; q < 4; ++q)
      if (make_act_map(&h.tmA[q], d.src[0].ptr, d.N, d.H, d.W, d.src[0].pitch, d.src[0].C, q >> 1, q & 1, 2, kc, h.pitch_px,
                       h.patch_rows, 1))
        return -1;
  } else {
    for (int s = 0; s < d.nsrc; ++s)
      if (make_patch_map(&h.tmA[s], d.src[s], d.N, d.H, d.W, h.pitch_px, h.patch_rows, kc)) return -1;
  }
  if (make_mat_map(&h.tmB, d.w, d.w_rows, d.w_cols, kc, BN)) return -1;
  int nd = 0;
  for (int q = 0; q < (scatter4 ? 4 : 1); ++q)
    for (int g = 0; g < ngroups; ++g)
      for (int c = 0; c < groups[g].channels; c += seg) {
        EpiDesc e = groups[g].epi;
        // the two addends are interchangeable; the halo epilogue prefetches `pre` two groups ahead but loads `pre2` at use
        // (~800 exposed cycles per group): a lone pending addend (the residual passthrough of every trunk conv1 /
        // main.0 data-gradient) must ride in the prefetched slot
        if (!e.pre && e.pre2) {
          e.pre = e.pre2;
          e.pre2 = nullptr;
        }
        if (e.sv_bits) e.sv = nullptr;  // the sign bits replace the 16-bit mask operand on this engine
        e.coff += c;
        e.osy = e.osx = scatter4 ? 2 : 1;
        e.ooy = scatter4 ? (q >> 1) : 0;
        e.oox = scatter4 ? (q & 1) : 0;
        e.OH = scatter4 ? 2 * d.H : GH;
        e.OW = scatter4 ? 2 * d.W : GW;
        h.epi[nd++] = e;
      }
  out->use_halo = 1;
  out->BN = BN;
  out->BK = 64;
  out->NM = NM;
  out->n_blocks = h.n_blocks;
  return 1;
}

int build_conv(const ConvDesc& d, const OutGroup* groups, int ngroups, TapGemmLaunch* out) {
  out->use_halo = 0;
  out->NM = 1;
  {
    const int r = try_build_halo(d, groups, ngroups, out);
    if (r < 0) return 1;
    if (r > 0) return 0;
  }
  TapTable tt;
  if (fill_taps(d.kind, d.parity, &tt)) return 1;
  TapGemmParams& p = out->p;
  memset(&p, 0, sizeof(p));
  p.f16 = d.f16;
  REFID_REQUIRE(d.nsrc == 1 || (d.nsrc == 2 && !tt.parity_mode), "build_conv: dual source not allowed with parity views");
  REFID_REQUIRE(!d.src[0].nmod && !d.src[1].nmod, "build_conv: a repeated source needs the halo-conv engine");
  REFID_REQUIRE(d.H % tt.grid_div == 0 && d.W % tt.grid_div == 0, "build_conv: H,W must be even for stride-2 ops");

  int BK = 64;
  for (int s = 0; s < d.nsrc; ++s) {
    REFID_REQUIRE(d.src[s].C % 32 == 0, "build_conv: source channels %d not a multiple of 32", d.src[s].C);
    if (d.src[s].C % 64) BK = 32;
  }
  int BN = 128;
  for (int g = 0; g < ngroups; ++g) {
    REFID_REQUIRE(groups[g].channels % 32 == 0, "build_conv: group channels %d not a multiple of 32", groups[g].channels);
    while (groups[g].channels % BN) BN >>= 1;
  }
  const int gh = d.H / tt.grid_div, gw = d.W / tt.grid_div;
  p.N = d.N;
  p.H = gh;
  p.W = gw;
  pick_tile(d.N, gh, gw, &p.TW, &p.TH, &p.TN);
  p.tiles_x = (gw + p.TW - 1) / p.TW;
  p.tiles_y = (gh + p.TH - 1) / p.TH;
  p.num_taps = tt.n;
  p.parity_mode = tt.parity_mode;
  for (int i = 0; i < tt.n; ++i) {
    p.tap_dy[i] = (signed char)tt.dy[i];
    p.tap_dx[i] = (signed char)tt.dx[i];
    p.tap_map[i] = (signed char)tt.map[i];
  }
  p.nsrc = d.nsrc;
  int ktot = 0;
  for (int s = 0; s < d.nsrc; ++s) {
    p.src_slabs[s] = d.src[s].C / BK;
    ktot += d.src[s].C;
  }
  REFID_REQUIRE(ktot == d.w_cols, "build_conv: K mismatch: sources %d vs weight cols %d", ktot, d.w_cols);
  if (tt.parity_mode) {
    for (int par = 0; par < 4; ++par)
      if (make_act_map(&p.tmA[par], d.src[0].ptr, d.N, d.H, d.W, d.src[0].pitch, d.src[0].C, par >> 1, par & 1, 2, BK, p.TW,
                       p.TH, p.TN))
        return 1;
  } else {
    for (int s = 0; s < d.nsrc; ++s)
      if (make_act_map(&p.tmA[s], d.src[s].ptr, d.N, d.H, d.W, d.src[s].pitch, d.src[s].C, 0, 0, 1, BK, p.TW, p.TH, p.TN))
        return 1;
  }
  if (make_mat_map(&p.tmB, d.w, d.w_rows, d.w_cols, BK, BN)) return 1;
  p.wrows_per_tap = d.wrows_per_tap;
  p.w_row0 = d.w_row0;
  p.w_img_rows = d.w_img_rows;
  REFID_REQUIRE(!d.w_img_rows || p.TN == 1, "build_conv: per-image weights need one image per tile (%dx%d grid)", gh, gw);

  // N blocks
  int nb = 0;
  const int reps = d.kind == CK_UP2 ? 4 : 1;
  for (int r = 0; r < reps; ++r) {
    for (int g = 0; g < ngroups; ++g) {
      for (int c = 0; c < groups[g].channels; c += BN) {
        REFID_REQUIRE(nb < kMaxNBlocks, "build_conv: too many N blocks");
        EpiDesc e = groups[g].epi;
        e.coff += c;
        e.osy = e.osx = 1;
        e.ooy = e.oox = 0;
        e.OH = gh;
        e.OW = gw;
        if (d.kind == CK_UP2) {
          e.osy = e.osx = 2;
          e.ooy = r >> 1;
          e.oox = r & 1;
          e.OH = 2 * gh;
          e.OW = 2 * gw;
        } else if (d.kind == CK_DOWN4_DGRAD) {
          e.osy = e.osx = 2;
          e.ooy = d.parity >> 1;
          e.oox = d.parity & 1;
          e.OH = 2 * gh;
          e.OW = 2 * gw;
        }
        p.epi[nb++] = e;
      }
    }
  }
  out->BN = BN;
  out->BK = BK;
  out->n_blocks = nb;
  return 0;
}

// Returns 1 when lowered onto the halo-wgrad kernel, 0 when the op does not qualify, -1 on error.
static int try_build_halowgrad(const ConvDesc& d, ActSrc q, float* outp, WgradLaunch* l) {
  const bool down = d.kind == CK_DOWN4;
  if (d.kind != CK_3X3 && !down) return 0;
  if (d.out_img_stride) return 0;
  if (down && (d.nsrc != 1 || d.src[0].C % 64 || (d.H & 1) || (d.W & 1) || q.C == 32)) return 0;
  for (int s = 0; s < d.nsrc; ++s)
    if (d.src[s].nmod) return 0;
  static const int disabled = getenv("REFID_NO_HALO_WGRAD") ? 1 : 0;
  if (disabled) return 0;
  int mode = 0;
  if (q.C == 64) mode = 64;
  else if (q.C % 128 == 0) mode = 128;
  else if (q.C == 32) mode = 32;
  if (!mode) return 0;
  int cp_total = 0;
  for (int s = 0; s < d.nsrc; ++s) {
    if (d.src[s].C % mode) return 0;
    cp_total += d.src[s].C;
  }
  HaloWgradParams& h = l->hp;
  memset(&h, 0, sizeof(h));
  h.mode = mode;
  h.out = outp;
  h.bias_out = d.bias_grad;
  h.CQ = q.C;
  h.cp_total = cp_total;
  h.nsrc = d.nsrc;
  const int slab_c = mode == 32 ? 32 : 64;
  for (int s = 0; s < d.nsrc; ++s) h.src_slabs[s] = d.src[s].C / slab_c;
  h.total_slabs = (down ? 4 : 1) * (cp_total / slab_c);  // down: every parity view carries all Cin channels
  h.down = down ? 1 : 0;
  const int GH = down ? d.H / 2 : d.H, GW = down ? d.W / 2 : d.W;  // grid of the output gradient
  h.N = d.N;
  h.H = GH;
  h.W = GW;
  h.tiles_x = (GW + 7) / 8;
  h.tiles_y = (GH + 15) / 16;
  h.num_tiles = h.tiles_x * h.tiles_y * d.N;
  h.jobs = mode == 128 ? (h.total_slabs / 2) * (down ? 2 : 3) * (q.C / 128) : h.total_slabs;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      num_sms = 148;
  }
  h.chunks = num_sms / h.jobs;
  if (h.chunks < 1) h.chunks = 1;
  if (h.chunks > h.num_tiles) h.chunks = h.num_tiles;
  const int rows = mode == 128 ? 16 : 18;
  if (down) {
    for (int v = 0; v < 4; ++v)
      if (make_act_map(&h.tmP[v], d.src[0].ptr, d.N, d.H, d.W, d.src[0].pitch, d.src[0].C, v >> 1, v & 1, 2, slab_c, 10, rows, 1))
        return -1;
  } else {
    for (int s = 0; s < d.nsrc; ++s)
      if (make_act_map(&h.tmP[s], d.src[s].ptr, d.N, d.H, d.W, d.src[s].pitch, d.src[s].C, 0, 0, 1, slab_c, 10, rows, 1)) return -1;
  }
  if (make_act_map(&h.tmQ, q.ptr, d.N, GH, GW, q.pitch, q.C, 0, 0, 1, slab_c, 8, 16, 1)) return -1;
  l->use_halo = 1;
  l->bias_done = d.bias_grad != nullptr;
  return 1;
}

int build_wgrad(const ConvDesc& d, ActSrc q, float* outp, WgradLaunch* l) {
  TapTable tt;
  if (fill_taps(d.kind, d.parity, &tt)) return 1;
  l->use_halo = 0;
  l->bias_done = 0;
  {
    const int r = try_build_halowgrad(d, q, outp, l);
    if (r < 0) return 1;
    if (r > 0) return 0;
  }
  WgradParams& p = l->p;
  memset(&p, 0, sizeof(p));
  REFID_REQUIRE(d.nsrc == 1 || (d.nsrc == 2 && !tt.parity_mode), "build_wgrad: dual source not allowed with parity views");
  int CB = 64;
  int cp_total = 0;
  for (int s = 0; s < d.nsrc; ++s) {
    REFID_REQUIRE(d.src[s].C % 32 == 0, "build_wgrad: source channels %d", d.src[s].C);
    if (d.src[s].C % 64) CB = 32;
    cp_total += d.src[s].C;
  }
  REFID_REQUIRE(q.C % 32 == 0, "build_wgrad: q channels %d", q.C);
  const int CBq = (q.C % 64) ? 32 : 64;
  const int gh = d.H / tt.grid_div, gw = d.W / tt.grid_div;
  p.N = d.N;
  p.H = gh;
  p.W = gw;
  pick_tile(d.N, gh, gw, &p.TW, &p.TH, &p.TN);
  p.tiles_x = (gw + p.TW - 1) / p.TW;
  p.tiles_y = (gh + p.TH - 1) / p.TH;
  p.num_tiles = p.tiles_x * p.tiles_y * ((d.N + p.TN - 1) / p.TN);
  p.num_taps = tt.n;
  p.parity_mode = tt.parity_mode;
  for (int i = 0; i < tt.n; ++i) {
    p.tap_dy[i] = (signed char)tt.dy[i];
    p.tap_dx[i] = (signed char)tt.dx[i];
    p.tap_map[i] = (signed char)tt.map[i];
  }
  p.nsrc = d.nsrc;
  for (int s = 0; s < d.nsrc; ++s) {
    p.src_blocks[s] = d.src[s].C / CB;
    p.src_nmod[s] = d.src[s].nmod;
    REFID_REQUIRE(!d.src[s].nmod || (!tt.parity_mode && d.src[s].nmod % p.TN == 0),
                  "build_wgrad: repeated source needs image-aligned tiles (nmod %d, TN %d)", d.src[s].nmod, p.TN);
  }
  p.CB = CB;
  p.CBq = CBq;
  p.CQ = q.C;
  p.BNq = q.C > 256 ? 256 : q.C;
  REFID_REQUIRE(q.C % p.BNq == 0, "build_wgrad: q.C=%d", q.C);
  const int bpt = 128 / CB;
  const int total_blocks = tt.n * (cp_total / CB);
  p.num_mtiles = (total_blocks + bpt - 1) / bpt;
  p.mt_per_cta = 512 / p.BNq;
  if (p.mt_per_cta > p.num_mtiles) p.mt_per_cta = p.num_mtiles;
  p.total_rows = tt.n * cp_total;
  p.out = outp;
  if (tt.parity_mode) {
    for (int par = 0; par < 4; ++par)
      if (make_act_map(&p.tmP[par], d.src[0].ptr, d.N, d.H, d.W, d.src[0].pitch, d.src[0].C, par >> 1, par & 1, 2, CB, p.TW,
                       p.TH, p.TN))
        return 1;
  } else {
    for (int s = 0; s < d.nsrc; ++s)
      if (make_act_map(&p.tmP[s], d.src[s].ptr, d.src[s].nmod ? d.src[s].nmod : d.N, d.H, d.W, d.src[s].pitch, d.src[s].C, 0,
                       0, 1, CB, p.TW, p.TH, p.TN))
        return 1;
  }
  if (make_act_map(&p.tmQ, q.ptr, d.N, gh, gw, q.pitch, q.C, 0, 0, 1, CBq, p.TW, p.TH, p.TN)) return 1;
  const int mt_groups = (p.num_mtiles + p.mt_per_cta - 1) / p.mt_per_cta;
  const int ctas_per_chunk = mt_groups * (q.C / p.BNq);
  int chunks = (296 + ctas_per_chunk - 1) / ctas_per_chunk;
  if (chunks > p.num_tiles) chunks = p.num_tiles;
  if (chunks < 1) chunks = 1;
  l->pixel_chunks = chunks;
  if (d.out_img_stride) {
    REFID_REQUIRE(p.TN == 1, "build_wgrad: per-image outputs need one image per tile (%dx%d grid)", gh, gw);
    p.out_img_stride = d.out_img_stride;
    // ~one CTA per SM over all images: every CTA ends with one atomic per output element, so fewer, longer CTAs win
    p.img_chunks = (148 + d.N * ctas_per_chunk - 1) / (d.N * ctas_per_chunk);
    if (p.img_chunks < 1) p.img_chunks = 1;
  }
  return 0;
}

bool one_image_per_tile(int N, int H, int W) {
  int tw, th, tn;
  pick_tile(N, H, W, &tw, &th, &tn);
  return tn == 1;
}

}  // namespace refid
