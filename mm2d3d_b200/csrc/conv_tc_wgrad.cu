// tcgen05 / TMEM weight gradient of the rule-table convolution for sm_100a.
//
//   dW[k][ci][co] = sum_j in[tbl(j,k)][ci] * dout[j][co]
//
// With the forward kernel's "virtual K" vk = k*C_in + ci (conv_tc.cu) this is, per tile of 128
// output rows,   dWflat[vk][co] += A_tile^T [vk x 128] . G_tile [128 x co]
// where A_tile is the gathered, zero-filled stage the forward kernel builds ([128 rows][32 vk] per
// K-block), here written with the 32-byte-granule swizzle that MN-major tf32 operands require
// (SWIZZLE_128B_BASE32B).  Read as an MN-major operand a stage is a 32-wide slice of M, so four
// consecutive K-blocks (stages 16 KB apart = the descriptor's LBO) form one M=128 operand and
// every 8 rows (two 4-row swizzle atoms, SBO = 512 B) are one tf32 K-step: 16 tcgen05.mma per
// (tile, group of 4 K-blocks), accumulating the group's [128 vk x C_out] slab of dW in TMEM over all
// tiles the CTA owns.  dout tiles are copied (no gather) into the same row-major swizzled form and
// serve as the MN-major B operand.  Grid = (tile splits) x (group splits): a CTA keeps its TMEM
// accumulators for its whole tile range and adds them to dW once (atomics: tile_splits * |dW|).
//
// Warp roles (14 warps): 0-7 gather producers (warp w owns ring stage w, tc_gather.cuh), 8-11
// epilogue (TMEM -> atomics), 12 MMA issuer + TMEM allocator, 13 dout-tile loader.
#include "tc_gather.cuh"

namespace {

using namespace tc;

constexpr int kTileM = 128;
constexpr int kStageBytes = kTileM * 128;
constexpr int kStages = 8;  // two groups of four K-blocks
constexpr int kThreads = (kStages + 6) * 32;
constexpr int kMaxK = 27;

struct WgParams {
  GatherArgs ga;
  const float* dout;
  float* dw;
  int c_out;
  int n_pad, nblk;        // padded N, 32-wide N blocks of the dout tile
  int num_tiles, tile_splits, groups_total, gpc;  // gpc = groups per CTA
  int g_bufs, tmem_cols;
  int* err;
};

template <bool ONEHOT>
__global__ void __launch_bounds__(kThreads, 1)
k_wgrad_tc(const WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_base = smem_base;
  const uint32_t g_bytes = (uint32_t)p.nblk * kStageBytes;
  const uint32_t g_base = a_base + kStages * kStageBytes;
  const uint32_t l_base = g_base + (uint32_t)p.g_bufs * g_bytes;
  const uint32_t bar_base = l_base + kStages * kListBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (bar_base - smem_base));
  auto a_full = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (uint32_t)(kStages + s); };
  auto g_full = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kStages + s); };
  auto g_empty = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kStages + 2 + s); };
  const uint32_t acc_full = bar_base + 8u * (uint32_t)(2 * kStages + 4);
  constexpr int kNumBars = 2 * kStages + 5;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + kNumBars);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(bars + kNumBars + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x;                    // which tiles
  const int g0 = blockIdx.y * p.gpc;               // first group of this CTA
  const int ng = min(p.gpc, p.groups_total - g0);  // groups of this CTA
  const int kb_lo = g0 * 4, kb_n = ng * 4;         // K-blocks [kb_lo, kb_lo + kb_n) per tile

  for (uint32_t i = threadIdx.x; i < (kStages * kStageBytes + (uint32_t)p.g_bufs * g_bytes) / 16; i += kThreads)
    reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(a_full(s), 32);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(g_full(s), 32);
      mbar_init(g_empty(s), 1);
    }
    mbar_init(acc_full, 1);
    *abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc(smem_u32((const void*)tmem_slot), (uint32_t)p.tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kStages) {
    // =================================================================== gather producer: owns stage `warp`
    const uint32_t stage = a_base + (uint32_t)warp * kStageBytes;
    const uint32_t list = l_base + (uint32_t)warp * kListBytes;
    uint32_t filled = 0, round = 0;
    int lt = 0, kbi = warp;  // item = (lt-th tile of this CTA, local K-block kbi); every 8th item is ours
    while (kbi >= kb_n) { kbi -= kb_n; ++lt; }
    int tile = split + lt * p.tile_splits;
    int nbv[4 * kMaxSegs];
    load_entries<ONEHOT>(p.ga, tile < p.num_tiles, (int64_t)tile * kTileM, kb_lo + kbi, lane, nbv);
    while (tile < p.num_tiles) {
      int n_kbi = kbi + kStages, n_lt = lt;
      while (n_kbi >= kb_n) { n_kbi -= kb_n; ++n_lt; }
      const int n_tile = split + n_lt * p.tile_splits;
      int nbn[4 * kMaxSegs];
      load_entries<ONEHOT>(p.ga, n_tile < p.num_tiles, (int64_t)n_tile * kTileM, kb_lo + n_kbi, lane, nbn);
      const int cnt = build_list(p.ga, kb_lo + kbi, lane, nbv, filled, list);  // off the stage's critical path
      if (!mbar_wait(a_empty(warp), (round & 1u) ^ 1u, abort_flag)) goto done;
      issue_copies<true>(p.ga, stage, lane, cnt, list, a_full(warp));
#pragma unroll
      for (int i = 0; i < 4 * kMaxSegs; ++i) nbv[i] = nbn[i];
      kbi = n_kbi; lt = n_lt; tile = n_tile;
      ++round;
    }
  } else if (warp < 12) {
    // =================================================================== epilogue: TMEM -> dW (+=)
    const int ew = warp & 3;
    if (!mbar_wait(acc_full, 0u, abort_flag)) goto done;
    tc_fence_after();
    const int m = ew * 32 + lane;  // row of the group's [128 x N] slab
    for (int gi = 0; gi < ng; ++gi) {
      const int vk = (g0 + gi) * 128 + m;  // virtual-K row = k*c_in + ci  ->  dW row
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(gi * p.n_pad);
      for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
        float acc[16];
        tmem_ld16(taddr + (uint32_t)c0, acc);
        if (vk < p.ga.K * p.ga.c_in) {
          float* dst = p.dw + (int64_t)vk * p.c_out + c0;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < p.c_out && acc[j] != 0.f) atomicAdd(dst + j, acc[j]);
        }
      }
    }
    tc_fence_before();
  } else if (warp == 12) {
    // =================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(128, p.n_pad, 1, 1);
      uint32_t it = 0, tile_iter = 0;
      bool ok = true;
      for (int tile = split; tile < p.num_tiles && ok; tile += p.tile_splits, ++tile_iter) {
        const int buf = (int)(tile_iter % (uint32_t)p.g_bufs);
        const uint32_t use = tile_iter / (uint32_t)p.g_bufs;
        if (!mbar_wait(g_full(buf), use & 1u, abort_flag)) break;
        const uint32_t gb = g_base + (uint32_t)buf * g_bytes;
        for (int gi = 0; gi < ng && ok; ++gi) {
          const int s0 = (int)(it % (uint32_t)kStages);  // 0 or 4: the group's first stage
          for (int j = 0; j < 4; ++j)
            if (!mbar_wait(a_full(s0 + j), (it / (uint32_t)kStages) & 1u, abort_flag)) { ok = false; break; }
          if (!ok) break;
          tc_fence_after();
          const uint32_t ab = a_base + (uint32_t)s0 * kStageBytes;
          const uint32_t d_tmem = tmem_base + (uint32_t)(gi * p.n_pad);
          // per K-step of 8 rows: two 4-row swizzle atoms (SBO = 512 B), four 32-wide M / nblk N blocks (LBO)
          for (int r8 = 0; r8 < 16; ++r8)
            umma_tf32(d_tmem, make_desc_sw128_base32(ab + r8 * 1024, kStageBytes, 512),
                      make_desc_sw128_base32(gb + r8 * 1024, kStageBytes, 512), idesc, (tile_iter | (uint32_t)r8) != 0);
          for (int j = 0; j < 4; ++j) umma_commit(a_empty(s0 + j));
          it += 4;
        }
        if (ok) umma_commit(g_empty(buf));
      }
      if (ok) umma_commit(acc_full);
    }
  } else {
    // =================================================================== dout-tile loader
    const int nch = p.c_out >> 2;
    uint32_t tile_iter = 0;
    for (int tile = split; tile < p.num_tiles; tile += p.tile_splits, ++tile_iter) {
      const int buf = (int)(tile_iter % (uint32_t)p.g_bufs);
      const uint32_t use = tile_iter / (uint32_t)p.g_bufs;
      if (!mbar_wait(g_empty(buf), (use & 1u) ^ 1u, abort_flag)) goto done;
      const uint32_t gb = g_base + (uint32_t)buf * g_bytes;
      for (int r0 = 0; r0 < kTileM; r0 += 4) {
        const int r = r0 + (lane >> 3);
        const int64_t row = (int64_t)tile * kTileM + r;
        const bool ok = row < p.ga.n_out;
        const float* src = p.dout + (ok ? row : 0) * p.c_out;
        for (int ch = lane & 7; ch < nch; ch += 8)
          cp_async16(gb + (uint32_t)(ch >> 3) * kStageBytes + (uint32_t)r * 128u + swz_base32(ch & 7, r), src + ch * 4,
                     ok ? 16u : 0u);
      }
      cp_async_arrive(g_full(buf));
    }
  }
done:
  asm volatile("cp.async.wait_all;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && *abort_flag && p.err) atomicExch(p.err, 1);
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

template <bool ONEHOT>
int launch_wgrad_tc(const WgParams& p, dim3 grid, size_t smem, cudaStream_t stream) {
  static bool once = false;
  if (!once) {
    MM3D_CUDA(cudaFuncSetAttribute(k_wgrad_tc<ONEHOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MM3D_CUDA(cudaFuncSetAttribute(k_wgrad_tc<ONEHOT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared));
    once = true;
  }
  k_wgrad_tc<ONEHOT><<<grid, kThreads, smem, stream>>>(p);
  return MM3D_OK;
}

}  // namespace

int* mm3d_device_err_flag();  // conv_tc.cu

int mm3d_conv_wgrad_tc_supported(int c_in, int c_out, int K) {
  return (c_in % 4) == 0 && c_in >= 16 && (c_out % 4) == 0 && c_out >= 4 && c_out <= 128 && K <= kMaxK;
}

int mm3d_conv_wgrad_tc(const float* in, int64_t n_in, int c_in, const float* d_out, int64_t n_out, int c_out,
                       float* d_weight, int K, const int32_t* tbl, int64_t tbl_stride,
                       const uint8_t* onehot_off, int accumulate, cudaStream_t stream) {
  MM3D_REQUIRE(mm3d_conv_wgrad_tc_supported(c_in, c_out, K), MM3D_ERR_UNSUPPORTED,
               "tcgen05 wgrad: unsupported shape c_in %d c_out %d K %d", c_in, c_out, K);
  MM3D_REQUIRE(n_out < (1ll << 31) && n_in * (int64_t)c_in < (1ll << 32), MM3D_ERR_UNSUPPORTED,
               "tcgen05 wgrad: tensor too large for 32-bit element offsets");
  MM3D_REQUIRE((((uintptr_t)in | (uintptr_t)d_out | (uintptr_t)d_weight) & 15) == 0, MM3D_ERR_INVALID,
               "tcgen05 wgrad: pointers must be 16-byte aligned");
  if (!accumulate) MM3D_CUDA(cudaMemsetAsync(d_weight, 0, sizeof(float) * (size_t)K * c_in * c_out, stream));
  if (n_out == 0) return MM3D_OK;
  WgParams p;
  p.ga = GatherArgs{in, tbl, tbl_stride, onehot_off, (int)n_out, c_in, K, c_in / 4, K * (c_in / 4)};
  p.dout = d_out; p.dw = d_weight; p.c_out = c_out;
  p.n_pad = (c_out + 15) / 16 * 16;
  p.nblk = (p.n_pad + 31) / 32;
  p.num_tiles = (int)mm3d_cdiv(n_out, kTileM);
  const int kbt = (p.ga.nq + 7) / 8;
  p.groups_total = (kbt + 3) / 4;
  int gpc = 512 / p.n_pad;  // TMEM: one [128 x n_pad] accumulator per group
  if (gpc > p.groups_total) gpc = p.groups_total;
  p.gpc = gpc;
  const int group_splits = (p.groups_total + gpc - 1) / gpc;
  int tile_splits = MM3D_NUM_SMS / group_splits;
  if (tile_splits < 1) tile_splits = 1;
  if (tile_splits > p.num_tiles) tile_splits = p.num_tiles;
  p.tile_splits = tile_splits;
  int cols = 32;
  while (cols < gpc * p.n_pad) cols <<= 1;
  p.tmem_cols = cols;
  p.g_bufs = p.nblk <= 2 ? 2 : 1;
  p.err = mm3d_device_err_flag();
  const size_t smem = 1024 + (size_t)kStages * kStageBytes + (size_t)p.g_bufs * p.nblk * kStageBytes +
                      (size_t)kStages * kListBytes + 8 * (2 * kStages + 5) + 64;
  dim3 grid((unsigned)tile_splits, (unsigned)group_splits);
  int rc = onehot_off ? launch_wgrad_tc<true>(p, grid, smem, stream) : launch_wgrad_tc<false>(p, grid, smem, stream);
  if (rc) return rc;
  mm3d_count_launches(1);
  MM3D_CHECK_LAUNCH("mm3d_conv_wgrad_tc");
  return MM3D_OK;
}
