// Tap-GEMM kernel (see tapgemm.cuh).  sm_100a: TMA -> smem ring -> tcgen05.mma -> TMEM -> epilogue warps.
#include "tapgemm.cuh"

namespace refid {

namespace {

constexpr int kThreads = 192;  // warp0: TMA producer, warp1: MMA issuer + TMEM owner, warps2-5: epilogue

template <bool F16>
__device__ __forceinline__ void load16(const __nv_bfloat16* p, float* f) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f[2 * i] = cvt_lo<F16>(w[i]);
    f[2 * i + 1] = cvt_hi<F16>(w[i]);
  }
}
template <bool F16>
__device__ __forceinline__ void store16(__nv_bfloat16* p, const float* f) {
  uint4 a, b;
  a.x = cvt_pack<F16>(f[0], f[1]);
  a.y = cvt_pack<F16>(f[2], f[3]);
  a.z = cvt_pack<F16>(f[4], f[5]);
  a.w = cvt_pack<F16>(f[6], f[7]);
  b.x = cvt_pack<F16>(f[8], f[9]);
  b.y = cvt_pack<F16>(f[10], f[11]);
  b.z = cvt_pack<F16>(f[12], f[13]);
  b.w = cvt_pack<F16>(f[14], f[15]);
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = a;
  q[1] = b;
}

// Epilogue of 16 consecutive output channels [c0, c0+16) of one pixel (see EpiDesc in tapgemm.cuh).
template <bool F16>
__device__ __forceinline__ void epi_apply16(const EpiDesc& e, float* v, const float* bias, size_t base, int c0, int n, int y,
                                            int x) {
  if (bias) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += __ldg(bias + e.coff + c0 + i);
  }
  if (e.pre) {
    float t[16];
    load16<F16>(e.pre + base + c0, t);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += t[i];
  }
  if (e.pre2) {
    float t[16];
    load16<F16>(e.pre2 + base + c0, t);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += t[i];
  }
  if (e.sv) {
    float t[16];
    load16<F16>(e.sv + base + c0, t);
    if (e.act == ACT_MULT) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] *= t[i];
    } else {
      const float sl = e.slope;
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] *= (t[i] > 0.f ? 1.f : sl);
    }
  } else {
    if (e.act == ACT_GELU) {
      float dg[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) gelu_both_f(v[i], &v[i], &dg[i]);
      if (e.out_pre) store16<F16>(e.out_pre + base + c0, dg);
    } else {
      if (e.out_pre) store16<F16>(e.out_pre + base + c0, v);
      if (e.act == ACT_LRELU) {
        const float sl = e.slope;
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * sl;
      }
    }
  }
  if (e.out) store16<F16>(e.out + base + c0, v);
  if (e.out_nchw && c0 == 0) {
    const size_t nimg = e.nchw_B ? (size_t)(n % e.nchw_B) * e.nchw_nstride + (size_t)(n / e.nchw_B) * e.nchw_tstride
                                 : (size_t)n * e.nchw_nstride;
    float* o = e.out_nchw + nimg + (size_t)(y * e.osy + e.ooy) * e.OW + (size_t)(x * e.osx + e.oox);
    const size_t plane = (size_t)e.OH * e.OW;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < e.nchw_C) o[i * plane] = v[i];
  }
  if (e.out_f32) {
    float4* o = reinterpret_cast<float4*>(e.out_f32 + base + c0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 t = o[i];
      t.x += v[4 * i];
      t.y += v[4 * i + 1];
      t.z += v[4 * i + 2];
      t.w += v[4 * i + 3];
      o[i] = t;
    }
  }
  if (e.out2) {
    float t[16];
    load16<F16>(e.post + base + c0, t);
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] += v[i];
    store16<F16>(e.out2 + base + c0, t);
  }
}

template <int BN, int BK, bool F16>
__global__ void __launch_bounds__(kThreads, 1) tapgemm_kernel(const __grid_constant__ TapGemmParams p) {
  constexpr int A_BYTES = 128 * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t SWZ = (BK == 64) ? 2u : 4u;    // UMMA layout type: 128B / 64B swizzle
  constexpr uint32_t SBO = 8u * BK * 2u;            // 8 rows of one swizzle span
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr uint32_t IDESC = make_idesc_16(128, BN, 0, 0, F16);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.num_stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)S * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* acc_bar = empty_bar + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  const int tile = blockIdx.x;
  const int tx_i = tile % p.tiles_x;
  const int ty_i = (tile / p.tiles_x) % p.tiles_y;
  const int tn_i = tile / (p.tiles_x * p.tiles_y);
  const int x0 = tx_i * p.TW, y0 = ty_i * p.TH, n0 = tn_i * p.TN;
  const int nblk = blockIdx.y;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  pdl_wait();  // set-up above overlaps the tail of the kernel before
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // after the TMEM allocation (see haloconv.cu)

  const int total_slabs = p.src_slabs[0] + (p.nsrc > 1 ? p.src_slabs[1] : 0);
  const int total_iters = total_slabs * p.num_taps;

  if (warp == 0) {
    // TMA producer: the whole warp runs the loop (uniform control flow), one elected lane issues.
    if (elect_one()) tma_prefetch_desc(&p.tmB);
    RingPos ring;
    int kglob = 0;
    for (int src = 0; src < p.nsrc; ++src) {
      for (int slab = 0; slab < p.src_slabs[src]; ++slab, kglob += BK) {
        for (int tap = 0; tap < p.num_taps; ++tap, ring.advance(S)) {
          const int s = (int)ring.s;
          const uint32_t ph = ring.ph;
          mbar_wait(&empty_bar[s], ph ^ 1, 0x100 + s);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
            uint8_t* a_dst = smem + (size_t)s * STAGE_BYTES;
            uint8_t* b_dst = a_dst + A_BYTES;
            const CUtensorMap* am = &p.tmA[p.parity_mode ? p.tap_map[tap] : src];
            tma_load_4d(a_dst, am, &full_bar[s], slab * BK, x0 + p.tap_dx[tap], y0 + p.tap_dy[tap], n0);
            tma_load_2d(b_dst, &p.tmB, &full_bar[s], kglob, p.w_row0 + tap * p.wrows_per_tap + nblk * BN + n0 * p.w_img_rows);
          }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: uniform loop, elected lane issues; descriptors are `stage base + constant`.
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t smem_u = smem_u32(smem);
    RingPos ring;
    for (int it = 0; it < total_iters; ++it, ring.advance(S)) {
      const int s = (int)ring.s;
      const uint32_t ph = ring.ph;
      mbar_wait(&full_bar[s], ph, 0x200 + s);
      tc_fence_after();
      const uint32_t a_addr = smem_u + (uint32_t)s * STAGE_BYTES;
      const uint64_t ad0 = make_smem_desc(a_addr, 16, SBO, SWZ);
      const uint64_t bd0 = make_smem_desc(a_addr + A_BYTES, 16, SBO, SWZ);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(tm, ad0 + (uint64_t)(k * 2), bd0 + (uint64_t)(k * 2), IDESC, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
        if (it == total_iters - 1) umma_commit(acc_bar);
      }
      __syncwarp();
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> global ----------------
    const EpiDesc& e = p.epi[nblk];
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;
    const int tn = m / (p.TH * p.TW);
    const int ty = (m / p.TW) % p.TH;
    const int tx = m % p.TW;
    const int n = n0 + tn, y = y0 + ty, x = x0 + tx;
    const bool valid = (n < p.N) && (y < p.H) && (x < p.W);
    const size_t pix = ((size_t)n * e.OH + (size_t)(y * e.osy + e.ooy)) * e.OW + (size_t)(x * e.osx + e.oox);
    const size_t base = pix * (size_t)e.C + e.coff;
    const float* bias = e.bias ? e.bias + (size_t)n * e.bias_nstride : nullptr;

    mbar_wait(acc_bar, 0, 0x300);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (valid) epi_apply16<F16>(e, v, bias, base, c0, n, y, x);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN, int BK, bool F16>
int launch_inst(TapGemmParams& p, int n_blocks, cudaStream_t stream) {
  constexpr int STAGE_BYTES = 128 * BK * 2 + BN * BK * 2;
  const int total_slabs = p.src_slabs[0] + (p.nsrc > 1 ? p.src_slabs[1] : 0);
  const int total_iters = total_slabs * p.num_taps;
  int stages = (200 * 1024) / STAGE_BYTES;
  if (stages > 8) stages = 8;
  if (stages > total_iters) stages = total_iters;
  if (stages < 1) stages = 1;
  p.num_stages = stages;
  const size_t smem = (size_t)stages * STAGE_BYTES + (2 * stages + 1) * sizeof(uint64_t) + 16 + 1024;
  static int configured = 0;  // per instantiation
  if (configured < (int)smem) {
    REFID_CUDA_CHECK(cudaFuncSetAttribute(tapgemm_kernel<BN, BK, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 227 * 1024;
  }
  const int tiles = p.tiles_x * p.tiles_y * ((p.N + p.TN - 1) / p.TN);
  dim3 grid(tiles, n_blocks);
  REFID_CUDA_CHECK(launch_k(tapgemm_kernel<BN, BK, F16>, dim3(grid), dim3(kThreads), smem, stream, p));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int pow2ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

}  // namespace

void pick_tile(int N, int H, int W, int* TW, int* TH, int* TN) {
  int tw = pow2ceil(W < 16 ? W : 16);
  int th = 128 / tw;
  int hp = pow2ceil(H);
  if (th > hp) th = hp;
  *TW = tw;
  *TH = th;
  *TN = 128 / (tw * th);
  (void)N;
}

int launch_tapgemm(TapGemmParams& p, int BN, int BK, int n_blocks, cudaStream_t stream) {
  REFID_REQUIRE(n_blocks >= 1 && n_blocks <= kMaxNBlocks, "tapgemm: bad n_blocks %d", n_blocks);
  REFID_REQUIRE(p.num_taps >= 1 && p.num_taps <= kMaxTaps, "tapgemm: bad num_taps %d", p.num_taps);
  REFID_REQUIRE(p.TW * p.TH * p.TN == 128, "tapgemm: tile %dx%dx%d != 128", p.TW, p.TH, p.TN);
  REFID_REQUIRE(!p.w_img_rows || p.TN == 1, "tapgemm: per-image weights need one image per tile (TN=%d)", p.TN);
#define INST(bn, bk) \
  if (BN == bn && BK == bk) return p.f16 ? launch_inst<bn, bk, true>(p, n_blocks, stream) : launch_inst<bn, bk, false>(p, n_blocks, stream);
  INST(32, 32) INST(64, 32) INST(128, 32) INST(256, 32)
  INST(32, 64) INST(64, 64) INST(128, 64) INST(256, 64)
#undef INST
  set_error("tapgemm: unsupported BN=%d BK=%d", BN, BK);
  return 1;
}

}  // namespace refid
