// Halo-conv kernel template (see haloconv.cuh).  sm_100a: TMA halo patch -> shifted UMMA descriptors -> TMEM (double
// buffered) -> epilogue warps; persistent over tiles.  Included by haloconv.cu (bf16 storage, training and reference
// precision) and haloconv_f16.cu (fp16 storage, forward-only plans); F16 selects the operand format of the instruction
// descriptor and the 16-bit <-> fp32 conversions of the epilogue.
#pragma once
#include "haloconv.cuh"

#include <stdlib.h>

namespace refid {
#ifdef REFID_HALO_TIMING
// diagnostic build: REFID_F32_RMW=2 makes the MMA warp skip the tcgen05.mma issue (commits only), so the kernel time becomes
// max(TMA, epilogue) -- is the epilogue slow by itself or because it shares the SM with running MMAs?
#define HALO_UMMA(...) do { if (p.f32_rmw != 2) umma_bf16(__VA_ARGS__); } while (0)
static __device__ long long g_halo_t[148 * 8];  // one copy per translation unit; haloconv.cu (bf16) reports its own
#define HT_DECL long long ht_acc = 0, ht_a = 0, ht_b = 0, ht_mma = 0, ht_t0 = clock64(), ht_x
#define HT_BEGIN ht_x = clock64()
#define HT_END(v) v += clock64() - ht_x
#else
#define HALO_UMMA(...) umma_bf16(__VA_ARGS__)
#define HT_DECL
#define HT_BEGIN
#define HT_END(v)
#endif

namespace {

#ifndef REFID_HALO_EPI_WARPS
#define REFID_HALO_EPI_WARPS 8
#endif
constexpr int kHaloEpiWarps = REFID_HALO_EPI_WARPS;  // 8 or 16: two or four epilogue warps per TMEM lane quarter
constexpr int kHaloEpiSplit = kHaloEpiWarps / 4;     // 32-channel groups of a quarter are dealt round-robin to its warps
constexpr int kHaloThreads = 64 + 32 * kHaloEpiWarps;  // warp0: TMA producer, warp1: MMA issuer + TMEM owner, then the epilogue warps
constexpr int kHaloMaxStages = 8;

// ---- epilogue with register prefetch --------------------------------------------------------------------------------
// The epilogue's global reads (residuals, pending gradient addends, activation masks, skip-sum operands) do not depend
// on the accumulator, so they are issued one 32-channel group AHEAD of use -- the first group of a tile before the
// accumulator-ready wait -- instead of load -> ~800-cycle stall -> use in every 16-channel step.
struct EpiPF {
  uint4 a[4], c[4];  // pre, (sv | post): 32 channels of this thread's pixel each (pre2 is rare and loaded at use)
  uint32_t m;        // activation sign bits of the 32 channels (sv_bits)
};

// LEAN instantiations serve launches whose epilogues only use {bias, pre, sign-bit masks, LeakyReLU / none, out}: the forms of
// most trunk convs.  Everything else (second addend, 16-bit masks, GELU, skip-sum second output, fp32 / NCHW targets) is
// compiled out, which shortens the hot loop body to what the instruction cache holds and frees the second operand set's
// registers (ncu: 0.6-1.1 stall cycles per issued instruction were instruction fetch in the general kernel).
template <bool LEAN>
__device__ __forceinline__ void epi_prefetch32(const EpiDesc& e, size_t off, size_t pix, int cword, bool valid, EpiPF& f) {
  if (!valid) return;
  if (e.sv_bits) f.m = __ldg(e.sv_bits + pix * (size_t)e.sv_bits_pitch + cword);
  if (e.pre) {
    ldg256(e.pre + off, f.a[0], f.a[1]);
    ldg256(e.pre + off + 16, f.a[2], f.a[3]);
  }
  if (!LEAN) {
    const __nv_bfloat16* third = e.sv ? e.sv : e.post;
    if (third) {
      ldg256(third + off, f.c[0], f.c[1]);
      ldg256(third + off + 16, f.c[2], f.c[3]);
    }
  }
}

template <bool F16>
__device__ __forceinline__ void unpack16(const uint4* q, float* f) {
  const uint32_t w[8] = {q[0].x, q[0].y, q[0].z, q[0].w, q[1].x, q[1].y, q[1].z, q[1].w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f[2 * i] = cvt_lo<F16>(w[i]);
    f[2 * i + 1] = cvt_hi<F16>(w[i]);
  }
}

template <bool F16>
__device__ __forceinline__ void unpack32(const uint4* q, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[8 * i + 0] = cvt_lo<F16>(q[i].x); f[8 * i + 1] = cvt_hi<F16>(q[i].x);
    f[8 * i + 2] = cvt_lo<F16>(q[i].y); f[8 * i + 3] = cvt_hi<F16>(q[i].y);
    f[8 * i + 4] = cvt_lo<F16>(q[i].z); f[8 * i + 5] = cvt_hi<F16>(q[i].z);
    f[8 * i + 6] = cvt_lo<F16>(q[i].w); f[8 * i + 7] = cvt_hi<F16>(q[i].w);
  }
}
// v[8k..8k+7] (+)= the 8 bf16 values of q (one 16-byte chunk): keeps temporaries at 8 registers instead of 32
template <bool F16>
__device__ __forceinline__ void add_chunk8(float* v, const uint4& q) {
  v[0] += cvt_lo<F16>(q.x); v[1] += cvt_hi<F16>(q.x); v[2] += cvt_lo<F16>(q.y); v[3] += cvt_hi<F16>(q.y);
  v[4] += cvt_lo<F16>(q.z); v[5] += cvt_hi<F16>(q.z); v[6] += cvt_lo<F16>(q.w); v[7] += cvt_hi<F16>(q.w);
}
// out2 = v + post, packed and stored chunk by chunk (no second 32-float array)
template <bool F16>
__device__ __forceinline__ void store32_sum(__nv_bfloat16* ptr, const float* v, const uint4* post) {
  uint4 t[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint4 q = post[k];
    t[k].x = cvt_pack<F16>(v[8 * k + 0] + cvt_lo<F16>(q.x), v[8 * k + 1] + cvt_hi<F16>(q.x));
    t[k].y = cvt_pack<F16>(v[8 * k + 2] + cvt_lo<F16>(q.y), v[8 * k + 3] + cvt_hi<F16>(q.y));
    t[k].z = cvt_pack<F16>(v[8 * k + 4] + cvt_lo<F16>(q.z), v[8 * k + 5] + cvt_hi<F16>(q.z));
    t[k].w = cvt_pack<F16>(v[8 * k + 6] + cvt_lo<F16>(q.w), v[8 * k + 7] + cvt_hi<F16>(q.w));
  }
  stg256(ptr, t[0], t[1]);
  stg256(ptr + 16, t[2], t[3]);
}
template <bool F16>
__device__ __forceinline__ void store32(__nv_bfloat16* ptr, const float* v) {
  uint4 t[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    t[i].x = cvt_pack<F16>(v[8 * i + 0], v[8 * i + 1]);
    t[i].y = cvt_pack<F16>(v[8 * i + 2], v[8 * i + 3]);
    t[i].z = cvt_pack<F16>(v[8 * i + 4], v[8 * i + 5]);
    t[i].w = cvt_pack<F16>(v[8 * i + 6], v[8 * i + 7]);
  }
  stg256(ptr, t[0], t[1]);
  stg256(ptr + 16, t[2], t[3]);
}

// Epilogue arithmetic of 32 consecutive output channels of one pixel (same as epi_apply16 in tapgemm.cu; see EpiDesc),
// global operands taken from registers (prefetched from global memory or read from the TMA-staged tiles).  On return
// v = the `out` values and, when e.out2, v2 = out + post.  The rare fp32 / NCHW / pre-activation outputs are written here.
// GELU (exact erf: ~60 instructions per element, twice) is compiled only into the GELU instantiations.
template <bool GELU, bool INPUTS, bool F16, bool LEAN>
__device__ __forceinline__ void epi_math32(const EpiDesc& e, float* v, float* v2, size_t off, int cseg, int n, int y, int x,
                                           const EpiPF& f, const float* sbias, uint32_t* bits_dst, bool f32_rmw) {
  if (e.bias) {  // bias table of the launch in shared memory (with ~227 KB of smem in use the L1 is too small to cache it)
    const uint32_t b4 = smem_u32(sbias);  // explicit shared-space loads (the pointer would otherwise be treated as generic)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 t;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(b4 + 16 * i));
      v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
    }
  }
  if (INPUTS && e.pre) {
#pragma unroll
    for (int k = 0; k < 4; ++k) add_chunk8<F16>(v + 8 * k, f.a[k]);
  }
  if (INPUTS && !LEAN && e.pre2) {
    uint4 r[4];
    ldg256(e.pre2 + off, r[0], r[1]);
    ldg256(e.pre2 + off + 16, r[2], r[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) add_chunk8<F16>(v + 8 * k, r[k]);
  }
  if (INPUTS && e.sv_bits) {
    const float sl = e.slope;
    const uint32_t m = f.m;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= ((m >> i) & 1u) ? 1.f : sl;
  } else if (INPUTS && !LEAN && e.sv) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      add_chunk8<F16>(t, f.c[k]);
      if (e.act == ACT_MULT) {  // saved derivative (GELU layers)
#pragma unroll
        for (int i = 0; i < 8; ++i) v[8 * k + i] *= t[i];
      } else {
        const float sl = e.slope;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[8 * k + i] *= (t[i] > 0.f ? 1.f : sl);
      }
    }
  } else {
    if (GELU && !LEAN && e.act == ACT_GELU) {
      // out = gelu(z); out_pre = gelu'(z) for the backward pass (chunk-wise: 8 temporaries)
#pragma unroll
      if (e.out_pre) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float dg[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) gelu_both_f(v[8 * k + i], &v[8 * k + i], &dg[i]);
          store8_rt(e.out_pre + off + 8 * k, dg, F16 ? 1 : 0);
        }
      } else {  // forward-only plans keep no derivative
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_f(v[i]);
      }
    } else {
      if (!LEAN && e.out_pre) store32<F16>(e.out_pre + off, v);
      if (e.act == ACT_LRELU) {
        const float sl = e.slope;
        if (bits_dst) {  // training plans: the derivative mask as sign bits (see EpiDesc); one compare serves mask and value
          uint32_t m = 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const bool pos = v[i] > 0.f;
            m |= (pos ? 1u : 0u) << i;
            v[i] = pos ? v[i] : v[i] * sl;
          }
          *bits_dst = m;
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * sl;
        }
      }
    }
  }
  if (!LEAN && e.out_nchw && cseg == 0) {
    const size_t nimg = e.nchw_B ? (size_t)(n % e.nchw_B) * e.nchw_nstride + (size_t)(n / e.nchw_B) * e.nchw_tstride
                                 : (size_t)n * e.nchw_nstride;
    float* o = e.out_nchw + nimg + (size_t)y * e.OW + (size_t)x;
    const size_t plane = (size_t)e.OH * e.OW;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < e.nchw_C) o[i * plane] = v[i];
  }
  if (!LEAN && e.out_f32) {
    // fp32 accumulation target: fire-and-forget vector reductions (the add happens in the L2; no load, no exposed latency --
    // the read-modify-write version stalled ~1 us per 32-channel group).  Every element receives exactly ONE add per launch
    // and launches are stream-ordered, so the result does not depend on any ordering.
    float* o = e.out_f32 + off;
    if (f32_rmw) {  // diagnostic (REFID_F32_RMW=1): the round-1 read-modify-write
      float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 t = o4[i];
        t.x += v[4 * i]; t.y += v[4 * i + 1]; t.z += v[4 * i + 2]; t.w += v[4 * i + 3];
        o4[i] = t;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * i), "f"(v[4 * i]), "f"(v[4 * i + 1]),
                     "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                     : "memory");
    }
  }
  (void)v2;
}

// Shifted taps: a tap (dy,dx) only moves the START ADDRESS of the A descriptor by whole 128-byte pixel rows inside the halo
// patch.  The 128B swizzle XOR is a function of the absolute shared-memory address bits on both the TMA write and the UMMA
// read side, so no descriptor base-offset is needed (verified on B200 with tools/halo_tap_probe.py).
//
// Issue-rate notes (measured with tools/ubench/umma_rate.cu): an SS-mode 128xNx16 MMA never takes less than ~62.5 cycles,
// so the single issuing thread has <= 62 cycles per MMA at N <= 128.  The producer and MMA warps therefore run their
// loops warp-uniformly (warp index via shuffle, elect.sync only around the issue) so that descriptors live in uniform
// registers, and every per-MMA descriptor is `base + compile-time constant` (TAPS / pitch / NM are template parameters).
// SINGLE (LEAN kernels only): the launch has ONE EpiDesc, so every descriptor field is an immediate constant-bank operand and
// the per-pixel address arithmetic is loop-invariant; a dynamically indexed descriptor costs a constant load (LDC, tens of
// cycles) in front of every field test of every 32-channel group.
// Work unit w of a CTA's round-robin sequence -> (item, first 16-row block, number of blocks).  Whole items first; with tail
// splitting the remainder comes as half items (see HaloConvParams::tail_split).
template <int NM>
__device__ __forceinline__ void halo_work(const HaloConvParams& p, int w, int& item, int& j0, int& nm_eff) {
  if (NM == 1 || !p.tail_split || w < p.full_items) {
    item = w;
    j0 = 0;
    nm_eff = NM;
  } else {
    const int sub = w - p.full_items;
    item = p.full_items + sub / NM;
    j0 = sub % NM;
    nm_eff = 1;
  }
}
__device__ __forceinline__ int halo_work_count(const HaloConvParams& p, int NM) {
  return (NM > 1 && p.tail_split) ? p.full_items + (p.num_items - p.full_items) * NM : p.num_items;
}

template <int BN, int NM, int TAPS, int KC, bool GELU, bool INPUTS, bool F16, bool LEAN, bool SINGLE = false>
__global__ void __launch_bounds__(kHaloThreads, 1) haloconv_kernel(const __grid_constant__ HaloConvParams p) {
  constexpr int HALO = TAPS == 9 ? 1 : 0;
  constexpr int PITCH = TAPS == 9 ? 10 : 8;       // pixels per patch row in shared memory
  constexpr int PATCH_ROWS = 16 * NM + 2 * HALO;
  constexpr uint32_t IDESC = make_idesc_16(128, BN, 0, 0, F16);
  constexpr uint32_t PXB = KC * 2;             // bytes of one pixel row of a K slab (KC = 64 or 32 channels)
  constexpr uint32_t SWZ = KC == 64 ? 2u : 4u;  // UMMA layout type: 128B / 64B swizzle
  constexpr int KSTEPS = KC / 16;
  constexpr uint32_t B_TILE = BN * PXB;        // bytes of one (tap, slab) weight tile
  constexpr uint32_t SBO_B = 8u * PXB;
  constexpr uint32_t ACC_COLS = NM * BN;       // TMEM columns of one accumulator buffer
  constexpr uint32_t TMEM_COLS = (2 * ACC_COLS) <= 32 ? 32 : ((2 * ACC_COLS) <= 64 ? 64 : ((2 * ACC_COLS) <= 128 ? 128 : ((2 * ACC_COLS) <= 256 ? 256 : 512)));
  static_assert(2 * ACC_COLS <= 512, "accumulators exceed TMEM");
  constexpr uint32_t A_TX = (uint32_t)PATCH_ROWS * PITCH * PXB;
  constexpr uint32_t A_BYTES = (A_TX + 1023u) & ~1023u;
  constexpr uint32_t SBO = (uint32_t)PITCH * PXB;  // 8-pixel group (one tile row) to the next tile row

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int SA = p.stages_a, SB = p.stages_b;
  int total_slabs = 0;
  for (int i = 0; i < p.nsrc; ++i) total_slabs += p.src_slabs[i];
  uint8_t* a_base = smem;
  uint8_t* b_base = smem + (size_t)SA * A_BYTES;
  const size_t b_total = p.resident_b ? (size_t)(p.masked ? p.resident_tiles : total_slabs * TAPS) * B_TILE : (size_t)SB * B_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + b_total);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kHaloMaxStages;
  uint64_t* b_full = a_empty + kHaloMaxStages;
  uint64_t* b_empty = b_full + kHaloMaxStages;
  uint64_t* acc_full = b_empty + kHaloMaxStages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* wres_bar = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wres_bar + 1);
  // bias table of the launch: [n_blocks * BN] floats, 16-byte aligned (read as float4)
  float* sbias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kHaloMaxStages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kHaloEpiWarps);
    }
    mbar_init(wres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  // set-up above overlaps the tail of the kernel before; from here on global memory is read
  pdl_wait();
  const int bias_row = p.n_blocks * BN;  // floats per table row; one row per image when the bias is per sample
  for (int i = threadIdx.x; i < bias_row * (p.bias_images ? p.bias_images : 1); i += kHaloThreads) {
    const int img = i / bias_row, ch = i - img * bias_row;
    const EpiDesc& e = p.epi[ch >> p.epi_shift];
    sbias[i] = e.bias ? __ldg(e.bias + (size_t)img * e.bias_nstride + e.coff + (ch & (p.epi_seg - 1))) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  // dependents may be scheduled from here on: this CTA holds its TMEM columns, so a co-resident CTA of the next kernel
  // cannot take them first and then block in its own pdl_wait()
  pdl_launch_dependents();

  if (warp == 0) {
    // ---------------- TMA producer (whole warp runs the loop; one elected lane issues) ----------------
    if (elect_one()) {
      tma_prefetch_desc(&p.tmA[0]);
      tma_prefetch_desc(&p.tmB);
      if (p.resident_b) {
        mbar_arrive_expect_tx(wres_bar, (uint32_t)b_total);
        int idx = 0;
        for (int ks = 0; ks < total_slabs; ++ks)
          for (int tap = 0; tap < TAPS; ++tap) {
            if (p.masked && !((p.slab_mask[ks] >> tap) & 1u)) continue;  // masked: only the used (slab, tap) tiles, compact
            tma_load_2d(b_base + (size_t)(p.masked ? idx : ks * TAPS + tap) * B_TILE, &p.tmB, wres_bar, ks * KC,
                        p.w_row0 + tap * p.wrows_per_tap);
            ++idx;
          }
      }
    }
    RingPos ra, rb;
    const int nwork = halo_work_count(p, NM);
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
      int item, j0, nm_eff;
      halo_work<NM>(p, w, item, j0, nm_eff);
      const int nblk = item % p.n_blocks, tile = item / p.n_blocks;
      const int x0 = (tile % p.tiles_x) * 8;
      const int y0 = ((tile / p.tiles_x) % p.tiles_y) * (16 * NM) + j0 * 16;  // (a half item loads the whole patch box from its own first row)
      const int n = tile / tiles_per_img;
      int ks = 0;
      for (int src = 0; src < p.nsrc; ++src) {
        for (int slab = 0; slab < p.src_slabs[src]; ++slab, ++ks) {
          const int sa = ra.s;
          mbar_wait(&a_empty[sa], ra.ph ^ 1u, 0x700 + sa);
          if (elect_one()) {
            mbar_arrive_expect_tx(&a_full[sa], A_TX);
            tma_load_4d(a_base + (size_t)sa * A_BYTES, &p.tmA[src], &a_full[sa], slab * KC, x0 - HALO, y0 - HALO,
                        p.src_nmod[src] ? n % p.src_nmod[src] : n);
          }
          ra.advance(SA);
          if (!p.resident_b) {
            const unsigned tmask = (p.tap_mask[nblk] ? p.tap_mask[nblk] : 0xFFFFu) & (p.slab_mask[ks] ? p.slab_mask[ks] : 0xFFFFu);
            for (int tap = 0; tap < TAPS; ++tap) {
              if (!((tmask >> tap) & 1u)) continue;
              const int sb = rb.s;
              mbar_wait(&b_empty[sb], rb.ph ^ 1u, 0x710 + sb);
              if (elect_one()) {
                mbar_arrive_expect_tx(&b_full[sb], B_TILE);
                tma_load_2d(b_base + (size_t)sb * B_TILE, &p.tmB, &b_full[sb], ks * KC,
                            p.w_row0 + tap * p.wrows_per_tap + nblk * BN + n * p.w_img_rows);
              }
              rb.advance(SB);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (whole warp runs the loop; one elected lane issues) ----------------
    // The tcgen05 queue is about one MMA deep: every instruction this warp spends between MMAs is tensor-pipe idle time,
    // so the loop carries no integer divisions and the descriptors are `stage base + constant`.
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    if (p.resident_b) {
      mbar_wait(wres_bar, 0, 0x720);
      tc_fence_after();
    }
    const uint32_t a_base_u = smem_u32(a_base), b_base_u = smem_u32(b_base);
    RingPos ra, rb;
    uint32_t it = 0;
    HT_DECL;
    const int nwork = halo_work_count(p, NM);
    for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
      int item, j0, nm_eff;
      halo_work<NM>(p, w, item, j0, nm_eff);
      const uint32_t buf = it & 1;
      const int nblk_i = p.n_blocks == 1 ? 0 : item % p.n_blocks;
      const unsigned nmask = p.tap_mask[nblk_i] ? p.tap_mask[nblk_i] : 0xFFFFu;
      bool first = true;  // the first MMA of the item overwrites the accumulator
      HT_BEGIN;
      mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1, 0x730 + buf);
      HT_END(ht_acc);
      tc_fence_after();
      const uint32_t acc = tm + buf * ACC_COLS;
      for (int ks = 0; ks < total_slabs; ++ks) {
        const int sa = ra.s;
        HT_BEGIN;
        mbar_wait(&a_full[sa], ra.ph, 0x740 + sa);
        HT_END(ht_a);
        tc_fence_after();
        const uint64_t a_desc0 = make_smem_desc(a_base_u + (uint32_t)sa * A_BYTES, 16, SBO, SWZ);
        const unsigned tmask = nmask & (p.slab_mask[ks] ? p.slab_mask[ks] : 0xFFFFu);
        if (p.resident_b && p.masked) {
          // resident, compact weight tiles: only the taps this K slab uses (runtime tap loop, no weight barriers)
          int bt = p.slab_b0[ks];
#pragma unroll 1
          for (int tap = 0; tap < TAPS; ++tap) {
            if (!((tmask >> tap) & 1u)) continue;
            const uint32_t tap_off = (uint32_t)((TAPS == 9 ? tap / 3 : 0) * PITCH + (TAPS == 9 ? tap % 3 : 0)) * PXB;
            const uint64_t a_desc = a_desc0 + (uint64_t)(tap_off >> 4);
            const uint64_t b_desc = make_smem_desc(b_base_u + (uint32_t)bt * B_TILE, 16, SBO_B, SWZ);
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < NM; ++j) {
                if (j >= nm_eff) continue;
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  const uint64_t ad = a_desc + (uint64_t)(((uint32_t)j * 16u * SBO + (uint32_t)k * 32u) >> 4);
                  const uint64_t bd = b_desc + (uint64_t)(((uint32_t)k * 32u) >> 4);
                  HALO_UMMA(acc + j * BN, ad, bd, IDESC, (!first || k > 0) ? 1u : 0u);
                }
              }
            }
            __syncwarp();
            first = false;
            ++bt;
          }
          if (elect_one()) umma_commit(&a_empty[sa]);
          __syncwarp();
        } else if (p.resident_b) {
          const uint64_t b_desc0 = make_smem_desc(b_base_u + (uint32_t)(ks * TAPS) * B_TILE, 16, SBO_B, SWZ);
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
              const uint32_t tap_off = (uint32_t)(((TAPS == 9 ? tap / 3 : 0)) * PITCH + (TAPS == 9 ? tap % 3 : 0)) * PXB;
#pragma unroll
              for (int j = 0; j < NM; ++j) {
                if (j >= nm_eff) continue;
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  const uint64_t ad = a_desc0 + (uint64_t)((tap_off + (uint32_t)j * 16u * SBO + (uint32_t)k * 32u) >> 4);
                  const uint64_t bd = b_desc0 + (uint64_t)(((uint32_t)tap * B_TILE + (uint32_t)k * 32u) >> 4);
                  HALO_UMMA(acc + j * BN, ad, bd, IDESC, (ks > 0 || tap > 0 || k > 0) ? 1u : 0u);
                }
              }
            }
            umma_commit(&a_empty[sa]);
          }
          __syncwarp();
        } else if (TAPS == 9 && BN < 256 && (tmask & 0x1FFu) == 0x1FFu) {
          // streamed weights, all nine taps: one issue block per kernel row (three weight stages, 3*NM*KSTEPS MMAs).  The
          // tcgen05 queue is about one MMA deep, so every instruction the issuing thread spends between MMAs is tensor-pipe
          // idle time: the three barrier probes are issued together (their latencies overlap), all descriptors are
          // `stage base + constant`, and the per-block overhead (probe, elect, commit, warp sync) is paid once per 3 taps.
#pragma unroll
          for (int row = 0; row < 3; ++row) {
            RingPos r0 = rb, r1 = rb;
            r1.advance(SB);
            RingPos r2 = r1;
            r2.advance(SB);
            HT_BEGIN;
            const bool ok0 = mbar_try_wait(&b_full[r0.s], r0.ph);
            const bool ok1 = mbar_try_wait(&b_full[r1.s], r1.ph);
            const bool ok2 = mbar_try_wait(&b_full[r2.s], r2.ph);
            if (!ok0) mbar_wait(&b_full[r0.s], r0.ph, 0x750 + r0.s);
            if (!ok1) mbar_wait(&b_full[r1.s], r1.ph, 0x750 + r1.s);
            if (!ok2) mbar_wait(&b_full[r2.s], r2.ph, 0x750 + r2.s);
            HT_END(ht_b);
            const uint64_t bdq[3] = {make_smem_desc(b_base_u + r0.s * B_TILE, 16, SBO_B, SWZ),
                                     make_smem_desc(b_base_u + r1.s * B_TILE, 16, SBO_B, SWZ),
                                     make_smem_desc(b_base_u + r2.s * B_TILE, 16, SBO_B, SWZ)};
            uint64_t* const eq[3] = {&b_empty[r0.s], &b_empty[r1.s], &b_empty[r2.s]};
            HT_BEGIN;
            if (elect_one()) {
#pragma unroll
              for (int q = 0; q < 3; ++q) {
                const uint32_t tap_off = (uint32_t)(row * PITCH + q) * PXB;
#pragma unroll
                for (int j = 0; j < NM; ++j) {
                  if (j >= nm_eff) continue;
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k) {
                    const uint64_t ad = a_desc0 + (uint64_t)((tap_off + (uint32_t)j * 16u * SBO + (uint32_t)k * 32u) >> 4);
                    const uint64_t bd = bdq[q] + (uint64_t)(((uint32_t)k * 32u) >> 4);
                    HALO_UMMA(acc + j * BN, ad, bd, IDESC, (!first || row > 0 || q > 0 || k > 0) ? 1u : 0u);
                  }
                }
                umma_commit(eq[q]);
              }
              if (row == 2) umma_commit(&a_empty[sa]);
            }
            __syncwarp();
            HT_END(ht_mma);
            rb = r2;
            rb.advance(SB);
          }
          first = false;
        } else {
#pragma unroll 1
          for (int tap = 0; tap < TAPS; ++tap) {
            if (!((tmask >> tap) & 1u)) continue;
            const int sb = rb.s;
            HT_BEGIN;
            mbar_wait(&b_full[sb], rb.ph, 0x750 + sb);
            HT_END(ht_b);
            const uint32_t tap_off = (uint32_t)((TAPS == 9 ? tap / 3 : 0) * PITCH + (TAPS == 9 ? tap % 3 : 0)) * PXB;
            const uint64_t a_desc = a_desc0 + (uint64_t)(tap_off >> 4);
            const uint64_t b_desc = make_smem_desc(b_base_u + (uint32_t)sb * B_TILE, 16, SBO_B, SWZ);
            HT_BEGIN;
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < NM; ++j) {
                if (j >= nm_eff) continue;
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  const uint64_t ad = a_desc + (uint64_t)(((uint32_t)j * 16u * SBO + (uint32_t)k * 32u) >> 4);
                  const uint64_t bd = b_desc + (uint64_t)(((uint32_t)k * 32u) >> 4);
                  HALO_UMMA(acc + j * BN, ad, bd, IDESC, (!first || k > 0) ? 1u : 0u);
                }
              }
              umma_commit(&b_empty[sb]);
            }
            __syncwarp();
            HT_END(ht_mma);
            first = false;
            rb.advance(SB);
          }
          if (elect_one()) umma_commit(&a_empty[sa]);
          __syncwarp();
        }
        ra.advance(SA);
      }
      if (elect_one()) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
#ifdef REFID_HALO_TIMING
    if (lane == 0 && blockIdx.x < 148) {
      long long* o = g_halo_t + blockIdx.x * 8;
      o[0] = clock64() - ht_t0; o[1] = ht_acc; o[2] = ht_a; o[3] = ht_b; o[4] = ht_mma; o[5] = it;
    }
#endif
  } else {
    // ---------------- epilogue: TMEM -> registers -> global, overlapped with the next item's MMAs ----------------
    const int q = warp & 3;  // TMEM lane quarter this warp may access (hardware rule: lanes 32*(warp%4)..+31)
    const int m = q * 32 + lane;
    const int ty = m >> 3, tx = m & 7;
    constexpr int GPT = BN / 32;  // 32-channel groups per pixel tile
    constexpr int G = NM * GPT;
    const int epi_mask = p.epi_seg - 1;
    {
      // ---- direct mode: eight warps, two per TMEM lane quarter splitting the groups (even / odd); global operands
      //      prefetched one group ahead into registers, 256-bit global accesses ----
      const int hsel = (warp - 2) >> 2;
      uint32_t it = 0;
      const int nwork = halo_work_count(p, NM);
      for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
        int item, j0, nm_eff;
        halo_work<NM>(p, w, item, j0, nm_eff);
        const int nblk = item % p.n_blocks, tile = item / p.n_blocks;
        const int x = (tile % p.tiles_x) * 8 + tx;
        const int y0 = ((tile / p.tiles_x) % p.tiles_y) * (16 * NM) + j0 * 16;
        const int n = tile / tiles_per_img;
        const uint32_t buf = it & 1;
        auto group_ctx = [&](int g, const EpiDesc*& e, size_t& off, size_t& pix, int& cword, int& cseg, int& y, bool& valid) {
          const int j = g / GPT, c0 = (g % GPT) * 32;
          y = y0 + j * 16 + ty;
          valid = (j < nm_eff) && (y < p.H) && (x < p.W);
          const int ch = nblk * BN + c0;
          if constexpr (SINGLE) e = &p.epi[0];
          else e = &p.epi[ch >> p.epi_shift];
          cseg = ch & epi_mask;
          pix = ((size_t)n * e->OH + (size_t)(y * e->osy + e->ooy)) * e->OW + (size_t)(x * e->osx + e->oox);
          off = pix * (size_t)e->C + e->coff + cseg;
          cword = (e->coff + cseg) >> 5;
        };
        // 32-channel groups of this thread's pixel: TMEM -> registers -> arithmetic -> 256-bit global stores.  INPUTS
        // instantiations (residuals, masks, skip-sum operands) load group g+2's global operands while group g is
        // computed; the others carry neither the prefetch registers nor the second output.
        auto prefetch = [&](int g, EpiPF& f) {
          const EpiDesc* e;
          size_t off, pix;
          int cseg, y, cword;
          bool valid;
          group_ctx(g, e, off, pix, cword, cseg, y, valid);
          epi_prefetch32<LEAN>(*e, off, pix, cword, valid, f);
        };
        auto process = [&](int g, const EpiPF& f) {
          const EpiDesc* e;
          size_t off, pix;
          int cseg, y, cword;
          bool valid;
          group_ctx(g, e, off, pix, cword, cseg, y, valid);
          const int j = g / GPT, c0 = (g % GPT) * 32;
          float v[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_COLS + (uint32_t)(j * BN + c0);
          tmem_ld16(taddr, v);
          tmem_ld16(taddr + 16, v + 16);
          tmem_ld_wait();
          if (valid) {
            epi_math32<GELU, INPUTS, F16, LEAN>(*e, v, nullptr, off, cseg, n, y, x, f,
                                          sbias + (p.bias_images ? n * bias_row : 0) + nblk * BN + c0,
                                          e->out_bits ? e->out_bits + pix * (size_t)e->out_bits_pitch + cword : nullptr,
                                          p.f32_rmw != 0);
            if (e->out) store32<F16>(e->out + off, v);
            if (INPUTS && !LEAN && e->out2) store32_sum<F16>(e->out2 + off, v, f.c);
          }
        };
        // L2 prefetch of the NEXT item's epilogue operands (residuals, masks, skip-sum addends): they are whole activation
        // tensors streamed from HBM; the register prefetch below covers two groups (~1-2 k cycles), which under a loaded
        // memory system is about one DRAM round trip -- with the lines already in L2 it covers the rest comfortably
        if (INPUTS && !LEAN && p.epi_l2pf) {
          const int item2 = item + (int)gridDim.x;
          if (item2 < p.num_items) {
            const int nblk2 = item2 % p.n_blocks, tile2 = item2 / p.n_blocks;
            const int x2 = (tile2 % p.tiles_x) * 8 + tx;
            const int y02 = ((tile2 / p.tiles_x) % p.tiles_y) * (16 * NM);
            const int n2 = tile2 / tiles_per_img;
#pragma unroll 1
            for (int g = hsel; g < G; g += kHaloEpiSplit) {
              const int j = g / GPT, c0 = (g % GPT) * 32;
              const int y2 = y02 + j * 16 + ty;
              if (y2 >= p.H || x2 >= p.W) continue;
              const int ch = nblk2 * BN + c0;
              const EpiDesc& e2 = p.epi[ch >> p.epi_shift];
              const size_t pix = ((size_t)n2 * e2.OH + (size_t)(y2 * e2.osy + e2.ooy)) * e2.OW + (size_t)(x2 * e2.osx + e2.oox);
              const size_t off2 = pix * (size_t)e2.C + e2.coff + (ch & epi_mask);
              if (e2.pre) asm volatile("prefetch.global.L2 [%0];" ::"l"(e2.pre + off2));
              const __nv_bfloat16* third = e2.sv ? e2.sv : e2.post;
              if (third) asm volatile("prefetch.global.L2 [%0];" ::"l"(third + off2));
              if (e2.pre2) asm volatile("prefetch.global.L2 [%0];" ::"l"(e2.pre2 + off2));
            }
          }
        }
        EpiPF fa, fb;  // two register sets, alternating: no copies
        // the global operands do not depend on the accumulator: the first TWO groups are requested before the
        // accumulator-ready wait, every later group two groups ahead of its use
        if (INPUTS && hsel < G) prefetch(hsel, fa);
        if (INPUTS && hsel + kHaloEpiSplit < G) prefetch(hsel + kHaloEpiSplit, fb);
        mbar_wait(&acc_full[buf], (it >> 1) & 1, 0x760 + buf);
        tc_fence_after();
        if (INPUTS) {
#pragma unroll 1
          for (int g = hsel; g < G; g += 2 * kHaloEpiSplit) {
            process(g, fa);
            if (g + 2 * kHaloEpiSplit < G) prefetch(g + 2 * kHaloEpiSplit, fa);
            if (g + kHaloEpiSplit < G) {
              process(g + kHaloEpiSplit, fb);
              if (g + 3 * kHaloEpiSplit < G) prefetch(g + 3 * kHaloEpiSplit, fb);
            }
          }
        } else {
#pragma unroll 1
          for (int g = hsel; g < G; g += kHaloEpiSplit) process(g, fa);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

constexpr size_t kHaloSmemMax = 227 * 1024;
constexpr size_t kHaloBarBytes = (4 * kHaloMaxStages + 5) * sizeof(uint64_t) + 48;

inline size_t halo_smem_bytes(const HaloConvParams& p, int BN) {
  const size_t a_bytes = ((size_t)p.patch_rows * p.pitch_px * p.kc * 2 + 1023) & ~(size_t)1023;
  int total_slabs = 0;
  for (int i = 0; i < p.nsrc; ++i) total_slabs += p.src_slabs[i];
  const size_t b_total = p.resident_b ? (size_t)(p.masked ? p.resident_tiles : total_slabs * p.num_taps) * BN * p.kc * 2
                                      : (size_t)p.stages_b * BN * p.kc * 2;
  return (size_t)p.stages_a * a_bytes + b_total + kHaloBarBytes +
         (size_t)p.n_blocks * BN * 4 * (p.bias_images ? p.bias_images : 1) + 1024;
}


template <int BN, int NM, int TAPS, int KC, bool GELU, bool INPUTS, bool F16, bool LEAN, bool SINGLE = false>
int launch_halo_inst(const HaloConvParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    REFID_CUDA_CHECK(cudaFuncSetAttribute(haloconv_kernel<BN, NM, TAPS, KC, GELU, INPUTS, F16, LEAN, SINGLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHaloSmemMax));
    configured = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    REFID_CUDA_CHECK(cudaGetDevice(&dev));
    REFID_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = p.num_items < num_sms ? p.num_items : num_sms;
  static const int no_split = getenv("REFID_NO_TAIL_SPLIT") ? 1 : 0;
  const int full = (p.num_items / grid) * grid, rem = p.num_items - full;
  if (NM == 2 && !no_split && rem > 0 && rem * NM <= grid && !p.epi_l2pf) {
    HaloConvParams q = p;
    q.tail_split = 1;
    q.full_items = full;
    REFID_CUDA_CHECK(launch_k(haloconv_kernel<BN, NM, TAPS, KC, GELU, INPUTS, F16, LEAN, SINGLE>, dim3(grid), dim3(kHaloThreads), halo_smem_bytes(p, BN), stream, q));
    REFID_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  REFID_CUDA_CHECK(launch_k(haloconv_kernel<BN, NM, TAPS, KC, GELU, INPUTS, F16, LEAN, SINGLE>, dim3(grid), dim3(kHaloThreads), halo_smem_bytes(p, BN), stream, p));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}


// Instantiation table of one storage flavour; each flavour lives in its own translation unit.
template <bool F16>
int launch_haloconv_flavour(const HaloConvParams& p, int BN, int NM, cudaStream_t stream) {
  REFID_REQUIRE((p.num_taps == 9 && p.halo == 1 && p.pitch_px == 10) || (p.num_taps == 1 && p.halo == 0 && p.pitch_px == 8),
                "haloconv: taps/halo/pitch %d/%d/%d unsupported", p.num_taps, p.halo, p.pitch_px);
  bool gelu = false;
  for (int i = 0; i < kMaxNBlocks; ++i) gelu = gelu || p.epi[i].act == ACT_GELU;
  REFID_REQUIRE(!gelu || (p.num_taps == 1 && p.kc == 64), "haloconv: GELU epilogues are instantiated for 1x1 / 64-channel slabs only");
  const bool in = p.epi_inputs != 0;
  // lean epilogue: only {bias, pre, sign-bit masks, LeakyReLU / none, out} anywhere in the launch
  bool lean = !p.epi_l2pf && !p.f32_rmw && !gelu;
  static const int no_lean = getenv("REFID_NO_LEAN") ? 1 : 0;
  if (no_lean) lean = false;
  for (int i = 0; i < kMaxNBlocks && lean; ++i) {
    const EpiDesc& e = p.epi[i];
    if (e.out_nchw || e.out_f32 || e.pre2 || e.sv || e.post || e.out2 || e.out_pre || (e.act != ACT_NONE && e.act != ACT_LRELU)) lean = false;
  }
  static const int no_single = getenv("REFID_NO_SINGLE") ? 1 : 0;
  const bool single = !no_single && p.n_blocks * (BN >> p.epi_shift) == 1;
  static const int census = getenv("REFID_EPI_CENSUS") ? 1 : 0;  // diagnostic: which epilogue features launches use
  if (census) {
    unsigned f = 0;
    for (int i = 0; i < kMaxNBlocks; ++i) {
      const EpiDesc& e = p.epi[i];
      f |= (e.pre ? 1u : 0) | (e.pre2 ? 2u : 0) | (e.sv ? 4u : 0) | (e.sv_bits ? 8u : 0) | (e.post || e.out2 ? 16u : 0) |
           (e.out_pre ? 32u : 0) | (e.out_f32 ? 64u : 0) | (e.out_nchw ? 128u : 0) | (e.act == ACT_GELU ? 256u : 0) |
           (e.act == ACT_MULT ? 512u : 0) | (e.out_bits ? 1024u : 0) | (e.osx > 1 ? 2048u : 0);
    }
    fprintf(stderr, "EPI_CENSUS taps=%d kc=%d BN=%d NM=%d items=%d feat=0x%x lean=%d\n", p.num_taps, p.kc, BN, NM, p.num_items, f, (int)lean);
  }
#define HPICK(bn, nm, taps, kc, gl) \
  return in ? launch_halo_inst<bn, nm, taps, kc, gl, true, F16, false>(p, stream) : launch_halo_inst<bn, nm, taps, kc, gl, false, F16, false>(p, stream)
#define HPICK9(bn, nm, taps, kc)                                                                                             \
  {                                                                                                                          \
    if constexpr (taps == 9) {                                                                                               \
      if (lean && single)                                                                                                    \
        return in ? launch_halo_inst<bn, nm, taps, kc, false, true, F16, true, true>(p, stream)                               \
                  : launch_halo_inst<bn, nm, taps, kc, false, false, F16, true, true>(p, stream);                             \
    }                                                                                                                        \
    return lean ? (in ? launch_halo_inst<bn, nm, taps, kc, false, true, F16, true>(p, stream)                                 \
                      : launch_halo_inst<bn, nm, taps, kc, false, false, F16, true>(p, stream))                               \
                : (in ? launch_halo_inst<bn, nm, taps, kc, false, true, F16, false>(p, stream)                                \
                      : launch_halo_inst<bn, nm, taps, kc, false, false, F16, false>(p, stream));                             \
  }
#define HINST(bn, nm)                                \
  if (BN == bn && NM == nm && p.kc == 64) {         \
    if (p.num_taps == 9) HPICK9(bn, nm, 9, 64);      \
    if (gelu) HPICK(bn, nm, 1, 64, true);            \
    HPICK9(bn, nm, 1, 64);                           \
  }
#define HINST32(bn, nm)                              \
  if (BN == bn && NM == nm && p.kc == 32) {         \
    if (p.num_taps == 9) HPICK9(bn, nm, 9, 32);      \
    HPICK9(bn, nm, 1, 32);                           \
  }
  HINST(32, 1) HINST(64, 1) HINST(128, 1) HINST(256, 1)
  HINST(32, 2) HINST(64, 2) HINST(128, 2)
  HINST32(32, 1) HINST32(64, 1) HINST32(128, 1)
  HINST32(32, 2) HINST32(64, 2) HINST32(128, 2)
#undef HPICK
#undef HPICK9
#undef HINST32
#undef HINST
  set_error("haloconv: unsupported BN=%d NM=%d", BN, NM);
  return 1;
}


}  // namespace
}  // namespace refid
