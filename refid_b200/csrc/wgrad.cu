// Weight-gradient kernel (see wgrad.cuh). MN-major UMMA operands straight from NHWC pixel tiles; K = pixels.
#include "wgrad.cuh"

namespace refid {

namespace {

constexpr int kWThreads = 192;
constexpr int P_TILE_BYTES = 128 * 128 * 2;  // 128 pixels x 128 M-rows (bf16)

__global__ void __launch_bounds__(kWThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.num_stages;
  const int Q_TILE_BYTES = 128 * p.BNq * 2;
  const int STAGE_BYTES = P_TILE_BYTES + Q_TILE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)S * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* acc_bar = empty_bar + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int bpt = 128 / p.CB;  // channel blocks per M tile
  const int blocks_per_tap = p.src_blocks[0] + (p.nsrc > 1 ? p.src_blocks[1] : 0);
  const int total_blocks = blocks_per_tap * p.num_taps;
  const int mt0 = blockIdx.y * p.mt_per_cta;
  int my_mt = p.num_mtiles - mt0;
  if (my_mt > p.mt_per_cta) my_mt = p.mt_per_cta;
  const int nb = blockIdx.z;
  // tile walk of this CTA: every gridDim.x-th tile of the launch, or (per-image outputs) every img_chunks-th tile of ONE image
  int t_first = blockIdx.x, t_stride = gridDim.x, t_limit = p.num_tiles, t_base = 0, img = 0;
  if (p.img_chunks) {
    img = blockIdx.x / p.img_chunks;
    t_first = blockIdx.x % p.img_chunks;
    t_stride = p.img_chunks;
    t_limit = p.tiles_x * p.tiles_y;
    t_base = img * t_limit;
  }
  int my_tiles = 0;
  for (int k = t_first; k < t_limit; k += t_stride) ++my_tiles;

  uint32_t ncols = 32;
  while ((int)ncols < p.mt_per_cta * p.BNq) ncols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, ncols);
    tmem_relinquish();
  }
  pdl_wait();  // set-up above overlaps the tail of the kernel before
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // after the TMEM allocation (see haloconv.cu)

  if (my_tiles > 0 && my_mt > 0) {
    if (warp == 0) {
      // TMA producer: uniform loop, one elected lane issues
      RingPos ring;
      for (int k = t_first; k < t_limit; k += t_stride) {
        const int pt = t_base + k;
        const int tx_i = pt % p.tiles_x;
        const int ty_i = (pt / p.tiles_x) % p.tiles_y;
        const int tn_i = pt / (p.tiles_x * p.tiles_y);
        const int x0 = tx_i * p.TW, y0 = ty_i * p.TH, n0 = tn_i * p.TN;
        for (int mt = 0; mt < my_mt; ++mt, ring.advance(S)) {
          const int s = (int)ring.s;
          const uint32_t ph = ring.ph;
          mbar_wait(&empty_bar[s], ph ^ 1, 0x400 + s);
          const int b0 = (mt0 + mt) * bpt;
          int nvalid = total_blocks - b0;
          if (nvalid > bpt) nvalid = bpt;
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(nvalid * 128 * p.CB * 2 + Q_TILE_BYTES));
            uint8_t* p_dst = smem + (size_t)s * STAGE_BYTES;
            uint8_t* q_dst = p_dst + P_TILE_BYTES;
            for (int j = 0; j < nvalid; ++j) {
              const int b = b0 + j;
              const int tap = b / blocks_per_tap;
              int cb = b % blocks_per_tap;
              int src = 0;
              if (cb >= p.src_blocks[0]) {
                src = 1;
                cb -= p.src_blocks[0];
              }
              const CUtensorMap* pm = &p.tmP[p.parity_mode ? p.tap_map[tap] : src];
              const int nmod = p.parity_mode ? 0 : p.src_nmod[src];
              tma_load_4d(p_dst + (size_t)j * 128 * p.CB * 2, pm, &full_bar[s], cb * p.CB, x0 + p.tap_dx[tap],
                          y0 + p.tap_dy[tap], nmod ? n0 % nmod : n0);
            }
            const int nq = p.BNq / p.CBq;
            for (int j = 0; j < nq; ++j)
              tma_load_4d(q_dst + (size_t)j * 128 * p.CBq * 2, &p.tmQ, &full_bar[s], nb * p.BNq + j * p.CBq, x0, y0, n0);
          }
        }
      }
    } else if (warp == 1) {
      // MMA issuer: uniform loop, elected lane issues; per-MMA descriptors are `stage base + ks * step`
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_u = smem_u32(smem);
      const uint32_t idesc = make_idesc_bf16(128, p.BNq, 1, 1);
      const uint32_t swzP = p.CB == 64 ? 2u : 4u, swzQ = p.CBq == 64 ? 2u : 4u;
      const uint32_t lboP = 128u * p.CB * 2u, sboP = 8u * p.CB * 2u;
      const uint32_t lboQ = 128u * p.CBq * 2u, sboQ = 8u * p.CBq * 2u;
      const uint64_t kstepP = (uint64_t)((16u * p.CB * 2u) >> 4), kstepQ = (uint64_t)((16u * p.CBq * 2u) >> 4);
      RingPos ring;
      for (int t = 0; t < my_tiles; ++t) {
        for (int mt = 0; mt < my_mt; ++mt, ring.advance(S)) {
          const int s = (int)ring.s;
          const uint32_t ph = ring.ph;
          mbar_wait(&full_bar[s], ph, 0x500 + s);
          tc_fence_after();
          const uint32_t p_addr = smem_u + (uint32_t)s * STAGE_BYTES;
          const uint64_t ad0 = make_smem_desc(p_addr, lboP, sboP, swzP);
          const uint64_t bd0 = make_smem_desc(p_addr + P_TILE_BYTES, lboQ, sboQ, swzQ);
          const uint32_t d_addr = tm + (uint32_t)(mt * p.BNq);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_bf16(d_addr, ad0 + ks * kstepP, bd0 + ks * kstepQ, idesc, (t > 0 || ks > 0) ? 1u : 0u);
            umma_commit(&empty_bar[s]);
          }
          __syncwarp();
        }
      }
      if (elect_one()) umma_commit(acc_bar);
      __syncwarp();
    } else {
      const int q = warp & 3;
      const int m = q * 32 + lane;
      mbar_wait(acc_bar, 0, 0x600);
      tc_fence_after();
      for (int mt = 0; mt < my_mt; ++mt) {
        const int row = (mt0 + mt) * 128 + m;
        const bool valid = row < p.total_rows;
        float* orow = p.out + (size_t)img * p.out_img_stride + (size_t)row * p.CQ + (size_t)nb * p.BNq;
#pragma unroll 1
        for (int c0 = 0; c0 < p.BNq; c0 += 16) {
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * p.BNq + c0), v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicAdd(orow + c0 + i, v[i]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, ncols);
  }
}

}  // namespace

int launch_wgrad(WgradParams& p, int pixel_chunks, cudaStream_t stream) {
  REFID_REQUIRE(p.CB == 64 || p.CB == 32, "wgrad: CB must be 32/64");
  REFID_REQUIRE(p.CBq == 64 || p.CBq == 32, "wgrad: CBq must be 32/64");
  REFID_REQUIRE(p.BNq % 16 == 0 && p.BNq >= 16 && p.BNq <= 256 && p.CQ % p.BNq == 0 && p.BNq % p.CBq == 0,
                "wgrad: bad BNq=%d CQ=%d", p.BNq, p.CQ);
  REFID_REQUIRE(p.mt_per_cta >= 1 && p.mt_per_cta * p.BNq <= 512, "wgrad: TMEM overflow mt=%d BNq=%d", p.mt_per_cta, p.BNq);
  const int stage_bytes = P_TILE_BYTES + 128 * p.BNq * 2;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 1) stages = 1;
  p.num_stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + (2 * stages + 1) * sizeof(uint64_t) + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    REFID_CUDA_CHECK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  if (pixel_chunks > p.num_tiles) pixel_chunks = p.num_tiles;
  if (pixel_chunks < 1) pixel_chunks = 1;
  if (p.img_chunks) {
    REFID_REQUIRE(p.TN == 1 && p.src_nmod[0] == 0 && p.src_nmod[1] == 0, "wgrad: per-image outputs need one image per tile");
    const int tpi = p.tiles_x * p.tiles_y;
    if (p.img_chunks > tpi) p.img_chunks = tpi;
    pixel_chunks = p.N * p.img_chunks;
  }
  const int mt_groups = (p.num_mtiles + p.mt_per_cta - 1) / p.mt_per_cta;
  dim3 grid(pixel_chunks, mt_groups, p.CQ / p.BNq);
  REFID_CUDA_CHECK(launch_k(wgrad_kernel, dim3(grid), dim3(kWThreads), smem, stream, p));
  REFID_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace refid
