// Halo-wgrad: weight gradient of the stride-1 3x3 convolutions with ONE activation-patch load per pixel tile.
//
//   D[(tap, ci)][co] += sum_pixels X[p + tap][ci] * G[p][co]
//
// The generic wgrad kernel (wgrad.cu) fetches a separate 128-pixel box of X per tap (9x the input from L2) and re-reads G
// for every M tile.  Here a CTA owns a JOB -- a fixed set of accumulator tiles that fills TMEM -- and streams pixel
// tiles (8 wide x 16 tall) through it: per tile one TMA box with the halo patch of X and one with G; every tap is a
// shifted UMMA descriptor into the patch (both operands MN-major, K = pixels, two 8-pixel rows per K=16 step).
//
//   MODE 64  (Cout = 64):   job = one 64-channel slab of X, all 9 taps.  An MMA has M = 128 = TWO TAPS: the descriptor's
//            leading-dimension byte offset (the stride between the two 64-row halves of A) is simply the address
//            difference of the two taps inside the patch.  5 MMA groups (4 pairs + the last tap), N = 64, 320 TMEM columns.
//   MODE 128 (Cout % 128 == 0): job = (pair of X slabs, one kernel row dy, 128-channel block of Cout).  M = 128 = the two
//            slabs (LBO = slab stride in shared memory), one MMA group per dx, N = 128, 384 TMEM columns.
//   MODE 32  (Cout = 32, 32-channel slabs, 64-byte pixel rows / 64B swizzle): job = one 32-channel slab, all 9 taps.  M = 128 =
//            FOUR 32-row blocks one pixel (64 B) apart = the three dx taps of one kernel row + a don't-care block; one MMA
//            group per dy, N = 32, 96 TMEM columns.
//
// The 4x4 stride-2 `down` conv (p.down) is the same kernel on the four stride-2 parity views of its input: view (py,px)
// holds in[2i+py][2j+px] and contributes the taps dy in (py ? {-1,0} : {0,+1}), dx likewise -- 2 x 2 taps per view = the 16
// kernel taps (ky = 2*dy + py + 1, kx = 2*dx + px + 1).  A slab's view index selects the tensor map and the tap set.
//
// Accumulation stays in TMEM over the CTA's whole pixel range (all T*B images of the batched launch); the fp32 result is
// added to global memory once per CTA with vectorised reductions.
#include "halowgrad.cuh"

namespace refid {

namespace {

constexpr int kHWThreads = 192;  // warp0: TMA producer, warp1: MMA issuer + TMEM owner, warps2-5: epilogue
constexpr int kHWMaxStages = 6;
constexpr uint32_t PITCH = 10;   // patch pixels per row (8 + 2 halo columns)

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int MODE>
struct HWCfg;
template <>
struct HWCfg<64> {
  static constexpr int ROWS = 18;                           // patch rows (16 + 2 halo rows)
  static constexpr uint32_t P_BYTES = 18 * PITCH * 128;     // one slab patch (23040 B)
  static constexpr uint32_t P_ALLOC = 23552;                // 1024-aligned
  static constexpr uint32_t Q_BYTES = 128 * 64 * 2;         // 16 KB
  static constexpr uint32_t STAGE = P_ALLOC + Q_BYTES;      // 39936
  static constexpr uint32_t TX = P_BYTES + Q_BYTES;
  static constexpr int GROUPS = 5, N = 64, COLS = 512;      // 320 used
};
template <>
struct HWCfg<128> {
  static constexpr int ROWS = 16;
  static constexpr uint32_t P_BYTES = 16 * PITCH * 128;     // 20480 B per slab (1024-aligned)
  static constexpr uint32_t P_ALLOC = 2 * P_BYTES;          // slab pair
  static constexpr uint32_t Q_BYTES = 2 * 128 * 64 * 2;     // two 64-channel halves of the 128-channel block
  static constexpr uint32_t STAGE = P_ALLOC + Q_BYTES;      // 73728
  static constexpr uint32_t TX = P_ALLOC + Q_BYTES;
  static constexpr int GROUPS = 3, N = 128, COLS = 512;     // 384 used
};

template <>
struct HWCfg<32> {
  static constexpr int ROWS = 18;
  static constexpr uint32_t P_BYTES = 18 * PITCH * 64;      // one 32-channel slab patch (11520 B)
  static constexpr uint32_t P_ALLOC = 12288;                // 1024-aligned (the don't-care block may read past the patch)
  static constexpr uint32_t Q_BYTES = 128 * 32 * 2;         // 8 KB
  static constexpr uint32_t STAGE = P_ALLOC + Q_BYTES;      // 20480
  static constexpr uint32_t TX = P_BYTES + Q_BYTES;
  static constexpr int GROUPS = 3, N = 32, COLS = 128;      // 96 used
};

template <int MODE>
__global__ void __launch_bounds__(kHWThreads, 1) halowgrad_kernel(const __grid_constant__ HaloWgradParams p) {
  using Cfg = HWCfg<MODE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.num_stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)S * Cfg::STAGE);
  uint64_t* empty_bar = full_bar + kHWMaxStages;
  uint64_t* acc_bar = empty_bar + kHWMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  // job / pixel-range decode (consecutive CTAs = same pixel range, different jobs: they share the tile loads in L2)
  const int job = blockIdx.x % p.jobs;
  const int chunk = blockIdx.x / p.jobs;
  const int t_begin = (int)(((long)p.num_tiles * chunk) / p.chunks);
  const int t_end = (int)(((long)p.num_tiles * (chunk + 1)) / p.chunks);
  int slab, dy = 0, coblk = 0;  // slab: global 64-channel slab index (MODE 64) or first slab of the pair (MODE 128)
  const int ndy = p.down ? 2 : 3;  // kernel rows a MODE 128 job family covers
  int dyi = 0;
  if (MODE == 64 || MODE == 32) {
    slab = job;
  } else {
    const int pairs = p.total_slabs / 2;
    slab = (job % pairs) * 2;
    dyi = (job / pairs) % ndy;
    dy = dyi - 1;
    coblk = job / (pairs * ndy);
  }
  int src = slab >= p.src_slabs[0] ? 1 : 0;
  int slab_in_src = slab - (src ? p.src_slabs[0] : 0);
  // `down`: source = parity view of the slab; first tap offsets of the view in y / x (the second is +1)
  int py = 0, px = 0;
  if (p.down) {
    src = slab / p.src_slabs[0];
    slab_in_src = slab % p.src_slabs[0];
    py = src >> 1;
    px = src & 1;
    dy = (py ? -1 : 0) + dyi;
  }
  const int dy0 = py ? -1 : 0, dx0 = px ? -1 : 0;

  // bias gradient: one job per Cout block also column-sums the G tiles it streams (4 extra arrivals free a stage)
  const bool colsum = p.bias_out != nullptr && (MODE == 128 ? (slab == 0 && dyi == 0) : job == 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kHWMaxStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], colsum ? 5 : 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::COLS);
    tmem_relinquish();
  }
  pdl_wait();  // set-up above overlaps the tail of the kernel before
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  pdl_launch_dependents();  // after the TMEM allocation (see haloconv.cu)

  if (t_end > t_begin) {
    if (warp == 0) {
      // ---------------- TMA producer ----------------
      if (elect_one()) {
        tma_prefetch_desc(&p.tmP[src]);
        tma_prefetch_desc(&p.tmQ);
      }
      int it = 0;
      RingPos ring;
      for (int t = t_begin; t < t_end; ++t, ++it, ring.advance(S)) {
        const int x0 = (t % p.tiles_x) * 8;
        const int y0 = ((t / p.tiles_x) % p.tiles_y) * 16;
        const int n = t / tiles_per_img;
        const int s = (int)ring.s;
        mbar_wait(&empty_bar[s], ring.ph ^ 1u, 0x800 + s);
        if (elect_one()) {
          uint8_t* st = smem + (size_t)s * Cfg::STAGE;
          mbar_arrive_expect_tx(&full_bar[s], Cfg::TX);
          if (MODE == 64 || MODE == 32) {
            tma_load_4d(st, &p.tmP[src], &full_bar[s], slab_in_src * MODE, x0 - 1, y0 - 1, n);
            tma_load_4d(st + Cfg::P_ALLOC, &p.tmQ, &full_bar[s], 0, x0, y0, n);
          } else {
            tma_load_4d(st, &p.tmP[src], &full_bar[s], slab_in_src * 64, x0 - 1, y0 + dy, n);
            tma_load_4d(st + Cfg::P_BYTES, &p.tmP[src], &full_bar[s], slab_in_src * 64 + 64, x0 - 1, y0 + dy, n);
            tma_load_4d(st + Cfg::P_ALLOC, &p.tmQ, &full_bar[s], coblk * 128, x0, y0, n);
            tma_load_4d(st + Cfg::P_ALLOC + 16384, &p.tmQ, &full_bar[s], coblk * 128 + 64, x0, y0, n);
          }
        }
      }
    } else if (warp == 1) {
      // ---------------- MMA issuer ----------------
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_u = smem_u32(smem);
      constexpr uint32_t IDESC = make_idesc_bf16(128, Cfg::N, 1, 1);
      constexpr uint32_t PXB = MODE == 32 ? 64u : 128u;  // bytes per pixel row of a slab
      constexpr uint32_t SBO_A = PITCH * PXB;          // next 8-pixel K group = next tile row of the patch
      int it = 0;
      RingPos ring;
      for (int t = t_begin; t < t_end; ++t, ++it, ring.advance(S)) {
        const int s = (int)ring.s;
        mbar_wait(&full_bar[s], ring.ph, 0x810 + s);
        tc_fence_after();
        const uint32_t st = smem_u + (uint32_t)s * Cfg::STAGE;
        // B: G tile [128 px][64 co] per half; K step = 16 pixel rows of 128 B = 2048 B; halves 16 KB apart
        const uint64_t bd0 = make_smem_desc(st + Cfg::P_ALLOC, 16384, 8u * PXB, MODE == 32 ? 4u : 2u);
        if (elect_one()) {
          if (MODE == 32) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
              for (int g = 0; g < 3; ++g) {  // g = kernel row; the four M blocks are the pixels dx = -1, 0, +1, (+2: discarded)
                const uint32_t offa = (uint32_t)((2 * ks + g) * (int)PITCH) * 64u;
                const uint64_t ad = make_smem_desc(st + offa, 64u, SBO_A, 4u);
                umma_bf16(tm + (uint32_t)(g * 32), ad, bd0 + (uint64_t)(ks * 64), IDESC, (it > 0 || ks > 0) ? 1u : 0u);
              }
            }
          } else if (MODE == 64 && p.down) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
              for (int g = 0; g < 2; ++g) {  // g = the view's kernel row; the two M halves are its two dx taps (one pixel apart)
                const uint32_t offa = (uint32_t)((2 * ks + dy0 + g + 1) * (int)PITCH + dx0 + 1) * 128u;
                const uint64_t ad = make_smem_desc(st + offa, 128u, SBO_A, 2u);
                umma_bf16(tm + (uint32_t)(g * 64), ad, bd0 + (uint64_t)(ks * 128), IDESC, (it > 0 || ks > 0) ? 1u : 0u);
              }
            }
          } else if (MODE == 64) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
              for (int g = 0; g < 5; ++g) {
                const int ta = 2 * g, tb = g < 4 ? 2 * g + 1 : 8;
                const uint32_t offa = (uint32_t)((2 * ks + ta / 3) * (int)PITCH + ta % 3) * 128u;
                const uint32_t offb = (uint32_t)((2 * ks + tb / 3) * (int)PITCH + tb % 3) * 128u;
                const uint32_t lbo = g < 4 ? offb - offa : 128u;  // group 4: second half is a don't-care copy
                const uint64_t ad = make_smem_desc(st + offa, lbo, SBO_A, 2u);
                umma_bf16(tm + (uint32_t)(g * 64), ad, bd0 + (uint64_t)(ks * 128), IDESC, (it > 0 || ks > 0) ? 1u : 0u);
              }
            }
          } else if (p.down) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
              for (int g = 0; g < 2; ++g) {  // the view's two dx taps
                const uint32_t offa = (uint32_t)((2 * ks) * (int)PITCH + dx0 + g + 1) * 128u;
                const uint64_t ad = make_smem_desc(st + offa, Cfg::P_BYTES, SBO_A, 2u);
                umma_bf16(tm + (uint32_t)(g * 128), ad, bd0 + (uint64_t)(ks * 128), IDESC, (it > 0 || ks > 0) ? 1u : 0u);
              }
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
              for (int g = 0; g < 3; ++g) {
                const uint32_t offa = (uint32_t)((2 * ks) * (int)PITCH + g) * 128u;
                const uint64_t ad = make_smem_desc(st + offa, Cfg::P_BYTES, SBO_A, 2u);
                umma_bf16(tm + (uint32_t)(g * 128), ad, bd0 + (uint64_t)(ks * 128), IDESC, (it > 0 || ks > 0) ? 1u : 0u);
              }
            }
          }
          umma_commit(&empty_bar[s]);
          if (t == t_end - 1) umma_commit(acc_bar);
        }
        __syncwarp();
      }
    } else {
      // ---------------- epilogue: TMEM -> fp32 global reductions, once per CTA ----------------
      const int q = warp & 3;
      const int m = q * 32 + lane;  // accumulator row
      if (colsum && MODE == 32) {
        // thread m sums pixel row m of the [128 px][32 co] G tile (64-byte rows, 64B swizzle)
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        const uint32_t smem_u = smem_u32(smem);
        RingPos ring;
        for (int t = t_begin; t < t_end; ++t, ring.advance(S)) {
          const int s = (int)ring.s;
          mbar_wait(&full_bar[s], ring.ph, 0x830 + s);
          const uint32_t gt = smem_u + (uint32_t)s * Cfg::STAGE + Cfg::P_ALLOC;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 u;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                         : "r"(gt + (uint32_t)m * 64u + (uint32_t)((k ^ ((m >> 1) & 3)) << 4)));
            acc[8 * k + 0] += bf16_lo(u.x); acc[8 * k + 1] += bf16_hi(u.x);
            acc[8 * k + 2] += bf16_lo(u.y); acc[8 * k + 3] += bf16_hi(u.y);
            acc[8 * k + 4] += bf16_lo(u.z); acc[8 * k + 5] += bf16_hi(u.z);
            acc[8 * k + 6] += bf16_lo(u.w); acc[8 * k + 7] += bf16_hi(u.w);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float v = acc[i];
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          acc[i] = v;
        }
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(p.bias_out + i, acc[i]);
        }
      } else if (colsum) {
        // thread m sums pixel row m (MODE 64) or rows (m & 63), (m & 63) + 64 of its 64-channel half (MODE 128)
        float acc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = 0.f;
        const uint32_t smem_u = smem_u32(smem);
        RingPos ring;
        for (int t = t_begin; t < t_end; ++t, ring.advance(S)) {
          const int s = (int)ring.s;
          mbar_wait(&full_bar[s], ring.ph, 0x830 + s);
          const uint32_t gt = smem_u + (uint32_t)s * Cfg::STAGE + Cfg::P_ALLOC + (MODE == 128 ? (uint32_t)(m >> 6) * 16384u : 0u);
#pragma unroll
          for (int rr = 0; rr < (MODE == 128 ? 2 : 1); ++rr) {
            const int row = MODE == 128 ? (m & 63) + rr * 64 : m;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              uint4 u;
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                           : "r"(gt + (uint32_t)row * 128u + (uint32_t)((k ^ (row & 7)) << 4)));
              acc[8 * k + 0] += bf16_lo(u.x); acc[8 * k + 1] += bf16_hi(u.x);
              acc[8 * k + 2] += bf16_lo(u.y); acc[8 * k + 3] += bf16_hi(u.y);
              acc[8 * k + 4] += bf16_lo(u.z); acc[8 * k + 5] += bf16_hi(u.z);
              acc[8 * k + 6] += bf16_lo(u.w); acc[8 * k + 7] += bf16_hi(u.w);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
        // all lanes of a warp hold the same channels (and in MODE 128 the same half): reduce over lanes, one atomic per channel
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          float v = acc[i];
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          acc[i] = v;
        }
        if (lane == 0) {
          float* bo = p.bias_out + (MODE == 128 ? coblk * 128 + (m >> 6) * 64 : 0);
#pragma unroll
          for (int i = 0; i < 64; ++i) atomicAdd(bo + i, acc[i]);
        }
      }
      mbar_wait(acc_bar, 0, 0x820);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < (p.down ? 2 : Cfg::GROUPS); ++g) {
        long row;
        bool valid = true;
        if (p.down) {
          // kernel tap of this accumulator row: MODE 64: row g of the view, dx half m >> 6; MODE 128: row dy, dx = g
          const int ddy = MODE == 64 ? dy0 + g : dy, ddx = MODE == 64 ? dx0 + (m >> 6) : dx0 + g;
          const int tap = (2 * ddy + py + 1) * 4 + (2 * ddx + px + 1);
          row = (long)tap * p.cp_total + slab_in_src * 64 + (MODE == 64 ? (m & 63) : m);
        } else if (MODE == 32) {
          valid = m < 96;  // the fourth 32-row block is the don't-care pixel
          row = (long)(g * 3 + (m >> 5)) * p.cp_total + slab * 32 + (m & 31);
        } else if (MODE == 64) {
          const int tap = m < 64 ? 2 * g : 2 * g + 1;
          valid = tap < 9;
          row = (long)tap * p.cp_total + slab * 64 + (m & 63);
        } else {
          const int tap = (dy + 1) * 3 + g;
          row = (long)tap * p.cp_total + slab * 64 + m;
        }
        float* orow = p.out + row * p.CQ + coblk * 128;
#pragma unroll 1
        for (int c0 = 0; c0 < Cfg::N; c0 += 16) {
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * Cfg::N + c0), v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) red_add_v4(orow + c0 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, Cfg::COLS);
  }
}

template <int MODE>
int launch_hw_inst(HaloWgradParams& p, cudaStream_t stream) {
  using Cfg = HWCfg<MODE>;
  static bool configured = false;
  if (!configured) {
    REFID_CUDA_CHECK(cudaFuncSetAttribute(halowgrad_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  int stages = (int)((220 * 1024) / Cfg::STAGE);
  if (stages > kHWMaxStages) stages = kHWMaxStages;
  p.num_stages = stages;
  const size_t smem = (size_t)stages * Cfg::STAGE + (2 * kHWMaxStages + 1) * sizeof(uint64_t) + 16 + 1024;
  REFID_CUDA_CHECK(launch_k(halowgrad_kernel<MODE>, dim3(p.jobs * p.chunks), dim3(kHWThreads), smem, stream, p));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace

int launch_halowgrad(HaloWgradParams& p, cudaStream_t stream) {
  return p.mode == 64 ? launch_hw_inst<64>(p, stream) : (p.mode == 32 ? launch_hw_inst<32>(p, stream) : launch_hw_inst<128>(p, stream));
}

}  // namespace refid
