// tcgen05 / TMEM rule-table convolution (forward and dgrad) for sm_100a.
//
//   out[j,:] = sum_k in[tbl(j,k),:] . W'[k]          j in a tile of 128 output rows
//
// Output-stationary implicit GEMM.  A CTA owns 128 output rows; the reduction dimension is the
// "virtual K" = (kernel offset k, input channel ci) flattened, cut into K-blocks of 32 tf32
// (= one 128-byte shared-memory row).  For every K-block the 128 gathered rows (zero where the
// neighbour is absent) form a K-major SWIZZLE_128B tile that tcgen05.mma (M=128, N=C_out,
// kind::tf32) multiplies with the matching block of the pre-swizzled weight image, accumulating
// ALL offsets in one TMEM accumulator.  No atomics, no read-modify-write of `out`, each output row
// stored once: deterministic.  Packing offsets back to back along K means a 16-channel layer needs
// 14 K-blocks instead of 27 and no layer pays per-offset padding.
//
// Warp roles (448 threads):   warps 0-7   gather producers (cp.async, 8 lanes per 128-byte row piece)
//                             warps 8-11  epilogue (TMEM lanes 32(w-8).. -> registers -> global)
//                             warp  12    MMA issuer (one lane) + TMEM allocator
//                             warp  13    weight loader (cp.async.bulk of the weight image)
// Pipelines (mbarriers):      A ring   a_full[S] (256 cp.async arrivals) / a_empty[S] (tcgen05.commit)
//                             B ring   b_full[SB] (expect_tx + bulk copy) / b_empty[SB] (tcgen05.commit)
//                             accum    acc_full[2] (tcgen05.commit)       / acc_empty[2] (128 arrivals)
// The accumulator is double buffered in TMEM, so the epilogue of tile i overlaps the MMAs of i+1.
//
// The producer loop is the critical instruction stream (one 16-byte slot decision per lane and
// pass), so it is branch-free, has the ring stage as a compile-time index (unrolled by S), advances
// (offset, channel-chunk) incrementally instead of dividing, and fetches the rule-table entries of
// the NEXT K-block into registers while the current one is being copied (no table staging in
// shared memory: the smem saved is what lets 2-3 CTAs share an SM and overlap their streams).
//
// Absent neighbours cost NO shared-memory traffic: stages are zeroed once and a thread re-zeroes
// (cp.async with src-size 0) only a slot it filled the previous time the stage was used -- the
// 3^3 tables are 10-30 % dense, so the fill stays proportional to the real pairs.
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kTileM = 128;
constexpr int kKBlock = 32;                 // tf32 elements per 128-byte row of a K-block
constexpr int kStageBytes = kTileM * 128;   // one A stage = one K-block of 128 rows = 16 KB
constexpr int kProducers = 256;             // 8 warps: 32 rows x 8 chunk lanes per pass, 4 passes
constexpr int kEpilogue = 128;
constexpr int kThreads = kProducers + kEpilogue + 64;
constexpr int kMaxStages = 6;
constexpr int kMaxBStages = 4;
constexpr int kMaxK = 27;

struct TcParams {
  const float* in;
  float* out;
  const float* wimg;  // [kbt][n_pad][32] tf32, rows 128-byte swizzled
  const int32_t* tbl;
  int64_t tbl_stride;
  const uint8_t* onehot_off;
  int n_out, c_in, c_out, K;
  int cq;         // 16-byte chunks per input row (c_in / 4)
  int nq;         // K * cq   chunks of the virtual K
  int kbt;        // K-blocks per tile = ceil(nq / 8)
  int d8, m8;     // 8 / cq and 8 % cq: how (offset, chunk) advance from one K-block to the next
  int n_pad, num_tiles, b_stages, tmem_cols;
  int* err;
};

// Weight image: one [n_pad][32] block per K-block; element (n, e) of block kb is Wsel(k)[ci][n]
// with (k, ci) = divmod(kb*32 + e, c_in) (0 beyond the virtual K / c_out), tf32-rounded, and the
// 16-byte chunks of every 128-byte row XOR-swizzled with (n & 7): exactly the bytes a SWIZZLE_128B
// K-major B tile has in shared memory, so the kernel bulk-copies it verbatim.
__global__ void k_weight_image(const float* __restrict__ w, float* __restrict__ img, int K, int c_in, int c_out,
                               int kbt, int n_pad, int transposed, int mirror) {
  const int64_t total = (int64_t)kbt * n_pad * kKBlock;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int e_sw = (int)(i % kKBlock);
    const int n = (int)((i / kKBlock) % n_pad);
    const int kb = (int)(i / ((int64_t)kKBlock * n_pad));
    const int chunk = (e_sw >> 2) ^ (n & 7);  // un-swizzle: which logical chunk lives here
    const int vk = kb * kKBlock + chunk * 4 + (e_sw & 3);
    const int k = vk / c_in, ci = vk - k * c_in;
    float v = 0.f;
    if (k < K && n < c_out) {
      const int ks = mirror ? K - 1 - k : k;
      v = transposed ? __ldg(w + ((int64_t)ks * c_out + n) * c_in + ci)   // forward weight is [K][c_out][c_in]
                     : __ldg(w + ((int64_t)ks * c_in + ci) * c_out + n);  // [K][c_in][c_out]
    }
    img[i] = to_tf32(v);
  }
}

template <int S, bool ONEHOT>
__global__ void __launch_bounds__(kThreads, 1)
k_conv_tc(const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A stages][B stages][barriers][tmem ptr][abort]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const int SB = p.b_stages;
  const uint32_t b_bytes = (uint32_t)p.n_pad * 128u;
  const uint32_t b_stride = (b_bytes + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + (uint32_t)S * kStageBytes;
  const uint32_t bar_base = b_base + (uint32_t)SB * b_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (bar_base - smem_base));
  auto a_full = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (uint32_t)(kMaxStages + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kMaxStages + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kMaxStages + kMaxBStages + s); };
  auto acc_full = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kMaxStages + 2 * kMaxBStages + s); };
  auto acc_empty = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kMaxStages + 2 * kMaxBStages + 2 + s); };
  constexpr int kNumBars = 2 * kMaxStages + 2 * kMaxBStages + 4;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + kNumBars);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(bars + kNumBars + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup
  for (uint32_t i = threadIdx.x; i < (uint32_t)S * kStageBytes / 16; i += kThreads)
    reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(a_full(s), kProducers);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), kEpilogue);
    }
    *abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc(smem_u32((const void*)tmem_slot), (uint32_t)p.tmem_cols);
  fence_proxy_async();  // the zero fill must be visible to the tensor core's (async-proxy) reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // =================================================================== gather producers
    const int g = threadIdx.x >> 3;  // row within a 32-row pass
    const int c = threadIdx.x & 7;   // 16-byte chunk within the 128-byte K-block row
    // this thread's slot in a stage: row g (+32 per pass), swizzled chunk
    const uint32_t slot0 = a_base + (uint32_t)g * 128u + (uint32_t)((c ^ (g & 7)) << 4);
    const int k0 = c / p.cq, cc0 = c - k0 * p.cq;  // (offset, chunk) of virtual-K chunk q = c
    uint32_t filled[S];                             // bit ps: the slot of pass ps holds data, not zeros
#pragma unroll
    for (int s = 0; s < S; ++s) filled[s] = 0;

    // rule-table entries of one K-block for this lane's 4 rows (-1 = absent / out of range)
    int par_c[4], off_c[4], par_n[4], off_n[4];  // ONEHOT: parent / offset of the current and next tile
    auto load_onehot = [&](int tile, int (&par)[4], int (&off)[4]) {
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const int64_t row = (int64_t)tile * kTileM + ps * 32 + g;
        const bool ok = tile < p.num_tiles && row < p.n_out;
        par[ps] = ok ? __ldg(p.tbl + row) : -1;
        off[ps] = ok ? (int)__ldg(p.onehot_off + row) : -1;
      }
    };
    auto entries = [&](int tile, int kb, int k, const int (&par)[4], const int (&off)[4], int (&nb)[4]) {
      const bool q_ok = tile < p.num_tiles && kb * 8 + c < p.nq;
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        if (ONEHOT) {
          nb[ps] = (q_ok && off[ps] == k) ? par[ps] : -1;
        } else {
          const int64_t row = (int64_t)tile * kTileM + ps * 32 + g;
          nb[ps] = (q_ok && row < p.n_out) ? __ldg(p.tbl + (int64_t)k * p.tbl_stride + row) : -1;
        }
      }
    };

    int tile = blockIdx.x, kb = 0, k = k0, cc = cc0;
    uint32_t round = 0;
    int nb[4];
    if (ONEHOT) {
      load_onehot(tile, par_c, off_c);
      load_onehot(tile + gridDim.x, par_n, off_n);
    }
    entries(tile, kb, k, par_c, off_c, nb);
    while (tile < p.num_tiles) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        if (tile >= p.num_tiles) break;
        // coordinates of the next K-block, and its table entries (in flight while this one is copied)
        int n_tile = tile, n_kb = kb + 1, n_k = k + p.d8, n_cc = cc + p.m8;
        if (n_cc >= p.cq) { n_cc -= p.cq; ++n_k; }
        const bool new_tile = n_kb == p.kbt;
        if (new_tile) { n_kb = 0; n_tile += gridDim.x; n_k = k0; n_cc = cc0; }
        int nb_next[4];
        if (ONEHOT && new_tile) entries(n_tile, n_kb, n_k, par_n, off_n, nb_next);
        else entries(n_tile, n_kb, n_k, par_c, off_c, nb_next);

        if (!mbar_wait(a_empty(s), (round & 1u) ^ 1u, abort_flag)) goto done;
        const uint32_t src_off = (uint32_t)cc * 4u;
        const uint32_t f = filled[s];
        uint32_t nf = 0;
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          const bool have = nb[ps] >= 0;
          const float* src = have ? p.in + ((uint32_t)nb[ps] * (uint32_t)p.c_in + src_off) : p.in;
          if (have || ((f >> ps) & 1u))
            cp_async16(slot0 + (uint32_t)(s * kStageBytes + ps * 32 * 128), src, have ? 16u : 0u);
          nf |= (have ? 1u : 0u) << ps;
        }
        filled[s] = nf;
        cp_async_arrive(a_full(s));

        if (ONEHOT && new_tile) {
#pragma unroll
          for (int ps = 0; ps < 4; ++ps) { par_c[ps] = par_n[ps]; off_c[ps] = off_n[ps]; }
          load_onehot(n_tile + gridDim.x, par_n, off_n);
        }
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) nb[ps] = nb_next[ps];
        tile = n_tile; kb = n_kb; k = n_k; cc = n_cc;
      }
      ++round;
    }
  } else if (warp < 12) {
    // =================================================================== epilogue
    const int ew = warp - 8;
    uint32_t tile_iter = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tile_iter) {
      const int ab = (int)(tile_iter & 1u);
      if (!mbar_wait(acc_full(ab), (tile_iter >> 1) & 1u, abort_flag)) goto done;
      tc_fence_after();
      const int64_t row = (int64_t)tile * kTileM + ew * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(ab * p.n_pad);
      for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
        float acc[16];
        tmem_ld16(taddr + (uint32_t)c0, acc);
        if (row < p.n_out) {
          float* dst = p.out + row * p.c_out + c0;
          if ((p.c_out & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              if (c0 + j < p.c_out) *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.c_out) dst[j] = acc[j];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty(ab));
    }
  } else if (warp == 12) {
    // =================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(kTileM, p.n_pad);
      uint32_t it = 0, tile_iter = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x, ++tile_iter) {
        const int ab = (int)(tile_iter & 1u);
        if (!mbar_wait(acc_empty(ab), ((tile_iter >> 1) & 1u) ^ 1u, abort_flag)) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * p.n_pad);
        for (int kb = 0; kb < p.kbt; ++kb, ++it) {
          const int s = (int)(it % (uint32_t)S);
          const int bs = (int)(it % (uint32_t)SB);
          if (!mbar_wait(b_full(bs), (it / (uint32_t)SB) & 1u, abort_flag)) { ok = false; break; }
          if (!mbar_wait(a_full(s), (it / (uint32_t)S) & 1u, abort_flag)) { ok = false; break; }
          tc_fence_after();
          const int rem = p.nq * 4 - kb * kKBlock;  // virtual-K elements left
          const int ksteps = rem >= kKBlock ? kKBlock / 8 : (rem + 7) >> 3;
          const uint32_t a_addr = a_base + (uint32_t)s * kStageBytes;
          const uint32_t b_addr = b_base + (uint32_t)bs * b_stride;
          for (int ks = 0; ks < ksteps; ++ks)
            umma_tf32(d_tmem, make_desc_sw128(a_addr + ks * 32), make_desc_sw128(b_addr + ks * 32), idesc,
                      (kb | ks) != 0);
          umma_commit(a_empty(s));
          umma_commit(b_empty(bs));
        }
        if (ok) umma_commit(acc_full(ab));
      }
    }
  } else {
    // =================================================================== weight loader
    if (lane == 0) {
      uint32_t it = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x) {
        for (int kb = 0; kb < p.kbt; ++kb, ++it) {
          const int bs = (int)(it % (uint32_t)SB);
          if (!mbar_wait(b_empty(bs), ((it / (uint32_t)SB) & 1u) ^ 1u, abort_flag)) { ok = false; break; }
          mbar_arrive_expect_tx(b_full(bs), b_bytes);
          bulk_g2s(b_base + (uint32_t)bs * b_stride, p.wimg + (size_t)kb * p.n_pad * kKBlock, b_bytes, b_full(bs));
        }
      }
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

int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

template <int S, bool ONEHOT>
int launch_conv_tc(const TcParams& p, size_t smem, cudaStream_t stream) {
  static int regs = 0;
  if (!regs) {
    MM3D_CUDA(cudaFuncSetAttribute(k_conv_tc<S, ONEHOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    MM3D_CUDA(cudaFuncSetAttribute(k_conv_tc<S, ONEHOT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared));
    cudaFuncAttributes fa;
    MM3D_CUDA(cudaFuncGetAttributes(&fa, k_conv_tc<S, ONEHOT>));
    regs = fa.numRegs > 0 ? fa.numRegs : 64;
  }
  // persistent CTAs: as many as fit (registers, shared memory, threads, TMEM columns); tiles round-robin
  const int regs_alloc = (regs + 7) / 8 * 8;
  int per_sm = 65536 / (regs_alloc * kThreads);
  const int by_smem = (int)((227 * 1024) / (smem + 1024));
  const int by_tmem = 512 / p.tmem_cols;
  if (per_sm > by_smem) per_sm = by_smem;
  if (per_sm > 2048 / kThreads) per_sm = 2048 / kThreads;
  if (per_sm > by_tmem) per_sm = by_tmem;
  if (per_sm < 1) per_sm = 1;
  int grid = MM3D_NUM_SMS * per_sm;
  if (grid > p.num_tiles) grid = p.num_tiles;
  k_conv_tc<S, ONEHOT><<<grid, kThreads, smem, stream>>>(p);
  return MM3D_OK;
}

}  // namespace

// ---- host entry points used by capi.cu -------------------------------------------------------

static int tc_geometry(int c_in, int c_out, int K, int* kbt, int* n_pad) {
  *kbt = (K * (c_in / 4) + 7) / 8;
  *n_pad = (c_out + 15) / 16 * 16;
  return (*n_pad <= 256 && (c_in % 4) == 0 && c_in >= 4 && K <= kMaxK) ? 0 : 1;
}

// 1 when the tcgen05 kernel handles this shape (input rows must be whole 16-byte chunks)
int mm3d_conv_tc_supported(int c_in, int c_out, int K) {
  int kbt, n_pad;
  return tc_geometry(c_in, c_out, K, &kbt, &n_pad) == 0;
}

size_t mm3d_conv_tc_workspace_bytes(int c_in, int c_out, int K) {
  int kbt, n_pad;
  if (tc_geometry(c_in, c_out, K, &kbt, &n_pad)) return 0;
  return mm3d_align((size_t)kbt * n_pad * 128);  // weight image
}

// Sticky per-device error flag set by a kernel whose mbarrier pipeline timed out (never in a
// correct build; it turns a would-be GPU hang into a reportable error).
static int* g_err_flag[64] = {nullptr};
int* mm3d_device_err_flag() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!g_err_flag[dev]) {
    int* p = nullptr;
    if (cudaMalloc(&p, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, sizeof(int));
    g_err_flag[dev] = p;
  }
  return g_err_flag[dev];
}

extern "C" int mm3d_take_device_error(void) {
  int* p = mm3d_device_err_flag();
  if (!p) return -1;
  int v = 0;
  if (cudaMemcpy(&v, p, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (v) cudaMemset(p, 0, sizeof(int));
  return v;
}

int mm3d_conv_fwd_tc(const float* in, int64_t n_in, int c_in, float* out, int64_t n_out, int c_out,
                     const float* weight, int K, const int32_t* tbl, int64_t tbl_stride,
                     const uint8_t* onehot_off, int flags, void* ws, size_t ws_bytes, cudaStream_t stream) {
  int kbt, n_pad;
  MM3D_REQUIRE(tc_geometry(c_in, c_out, K, &kbt, &n_pad) == 0, MM3D_ERR_UNSUPPORTED,
               "tcgen05 conv: unsupported shape c_in %d c_out %d K %d", c_in, c_out, K);
  MM3D_REQUIRE(n_out < (1ll << 31) && n_in * (int64_t)c_in < (1ll << 32), MM3D_ERR_UNSUPPORTED,
               "tcgen05 conv: tensor too large for 32-bit element offsets");
  MM3D_REQUIRE(ws && ws_bytes >= mm3d_conv_tc_workspace_bytes(c_in, c_out, K), MM3D_ERR_WORKSPACE,
               "tcgen05 conv: workspace too small");
  MM3D_REQUIRE((((uintptr_t)in | (uintptr_t)out | (uintptr_t)ws) & 15) == 0, MM3D_ERR_INVALID,
               "tcgen05 conv: pointers must be 16-byte aligned");
  const bool tr = (flags & MM3D_CONV_TRANSPOSE_W) != 0, mir = (flags & MM3D_CONV_MIRROR_K) != 0;
  MM3D_REQUIRE(tr || !mir, MM3D_ERR_UNSUPPORTED, "MIRROR_K without TRANSPOSE_W not implemented");
  if (n_out == 0) return MM3D_OK;

  float* wimg = (float*)ws;
  k_weight_image<<<mm3d_grid((int64_t)kbt * n_pad * kKBlock, 256), 256, 0, stream>>>(weight, wimg, K, c_in, c_out, kbt,
                                                                                 n_pad, tr ? 1 : 0, mir ? 1 : 0);
  TcParams p;
  p.in = in; p.out = out; p.wimg = wimg; p.tbl = tbl; p.tbl_stride = tbl_stride; p.onehot_off = onehot_off;
  p.n_out = (int)n_out; p.c_in = c_in; p.c_out = c_out; p.K = K;
  p.cq = c_in / 4; p.nq = K * p.cq; p.kbt = kbt; p.d8 = 8 / p.cq; p.m8 = 8 % p.cq; p.n_pad = n_pad;
  p.num_tiles = (int)mm3d_cdiv(n_out, kTileM);
  p.tmem_cols = pow2_cols(2 * n_pad);
  p.err = mm3d_device_err_flag();
  const uint32_t b_stride = ((uint32_t)n_pad * 128u + 1023u) & ~1023u;
  // narrow N: 4 A + 4 B stages, several CTAs per SM; wide N: one CTA per SM with a 6-deep A ring
  const bool wide = n_pad > 96;
  const int S = wide ? 6 : 4;
  p.b_stages = wide ? (n_pad > 128 ? 2 : 3) : 4;
  const size_t smem = 1024 + (size_t)S * kStageBytes + (size_t)p.b_stages * b_stride +
                      8 * (2 * kMaxStages + 2 * kMaxBStages + 4) + 64;
  int rc;
  if (wide) rc = onehot_off ? launch_conv_tc<6, true>(p, smem, stream) : launch_conv_tc<6, false>(p, smem, stream);
  else      rc = onehot_off ? launch_conv_tc<4, true>(p, smem, stream) : launch_conv_tc<4, false>(p, smem, stream);
  if (rc) return rc;
  mm3d_count_launches(2);
  MM3D_CHECK_LAUNCH("mm3d_conv_fwd_tc");
  return MM3D_OK;
}
