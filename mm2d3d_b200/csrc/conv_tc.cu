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
// Warp roles ((S+5) warps):   warps 0..S-1    gather producers, warp w OWNS ring stage w (tc_gather.cuh)
//                                             and bulk-copies that K-block's weight block (cp.async.bulk)
//                             warps S..S+3    epilogue (TMEM lanes 32(w&3).. -> registers -> global)
//                             warp  S+4       MMA issuer (one lane) + TMEM allocator
// Pipelines (mbarriers):      ring     a_full[S] (32 cp.async arrivals + expect_tx of the weight block)
//                                      / a_empty[S] (tcgen05.commit): ONE wait and ONE commit per K-block,
//                                      because the single MMA-issuing thread is the serial resource
//                             accum    acc_full[2] (tcgen05.commit)       / acc_empty[2] (128 arrivals)
// The accumulator is double buffered in TMEM, so the epilogue of tile i overlaps the MMAs of i+1.
//
// Absent neighbours cost neither instructions nor shared-memory traffic beyond one coalesced table
// read per 32 rows: stages are zeroed once, the producer compacts the rows that need an action and
// re-zeroes (cp.async with src-size 0) only slots that held data the last time the stage was used.
#include <stdlib.h>

#include "tc_gather.cuh"

namespace {

using namespace tc;

constexpr int kTileM = 128;
constexpr int kKBlock = 32;                 // tf32 elements per 128-byte row of a K-block
constexpr int kStageBytes = kTileM * 128;   // one A stage = one K-block of 128 rows = 16 KB
constexpr int kEpilogue = 128;
constexpr int kMaxStages = 6;
constexpr int kMaxK = 27;

struct TcParams {
  GatherArgs ga;
  float* out;
  const float* wimg;  // [kbt][n_pad][32] tf32, rows 128-byte swizzled
  int c_out;
  int kbt;            // K-blocks per tile = ceil(K * cq / 8)
  int n_pad, num_tiles, b_stages, tmem_cols;
  int* err;
  long long* trace;  // debug: MMA-thread timestamps of CTA 0 (4 per item), or NULL
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
__global__ void __launch_bounds__((S + 5) * 32, 1)
k_conv_tc(const TcParams p) {
  constexpr int kThreads = (S + 5) * 32;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A stages][B stages][producer lists][barriers][tmem ptr][abort]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t b_bytes = (uint32_t)p.n_pad * 128u;
  const uint32_t b_stride = (b_bytes + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + (uint32_t)S * kStageBytes;
  const uint32_t l_base = b_base + (uint32_t)S * b_stride;  // one weight block per A stage
  const uint32_t bar_base = l_base + (uint32_t)S * kListBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (bar_base - smem_base));
  auto a_full = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (uint32_t)(kMaxStages + s); };
  auto acc_full = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kMaxStages + s); };
  auto acc_empty = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kMaxStages + 2 + s); };
  constexpr int kNumBars = 2 * kMaxStages + 4;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + kNumBars);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(bars + kNumBars + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup
  for (uint32_t i = threadIdx.x; i < (uint32_t)S * kStageBytes / 16; i += kThreads)
    reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(a_full(s), 33);  // 32 cp.async arrivals (gathered rows) + 1 expect_tx arrival (weight block)
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), kEpilogue);
    }
    *abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == S + 4) tmem_alloc(smem_u32((const void*)tmem_slot), (uint32_t)p.tmem_cols);
  fence_proxy_async();  // the zero fill must be visible to the tensor core's (async-proxy) reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < S) {
    // =================================================================== gather producer: owns stage `warp`
    const uint32_t stage = a_base + (uint32_t)warp * kStageBytes;
    const uint32_t list = l_base + (uint32_t)warp * kListBytes;
    uint32_t filled = 0, round = 0;
    int lt = 0, kb = warp;  // item = (lt-th tile of this CTA, K-block kb); this warp takes every S-th item
    while (kb >= p.kbt) { kb -= p.kbt; ++lt; }
    int tile = blockIdx.x + lt * gridDim.x;
    int nbv[4 * kMaxSegs];
    load_entries<ONEHOT>(p.ga, tile < p.num_tiles, (int64_t)tile * kTileM, kb, lane, nbv);
    while (tile < p.num_tiles) {
      int n_kb = kb + S, n_lt = lt;
      while (n_kb >= p.kbt) { n_kb -= p.kbt; ++n_lt; }
      const int n_tile = blockIdx.x + n_lt * gridDim.x;
      int nbn[4 * kMaxSegs];  // the next item's table entries are in flight while this one is copied
      load_entries<ONEHOT>(p.ga, n_tile < p.num_tiles, (int64_t)n_tile * kTileM, n_kb, lane, nbn);
      const int cnt = build_list(p.ga, kb, lane, nbv, filled, list);  // off the stage's critical path
      if (!mbar_wait(a_empty(warp), (round & 1u) ^ 1u, abort_flag)) goto done;
      if (lane == 0) {  // this K-block's weight block rides on the same barrier as the gathered rows
        mbar_arrive_expect_tx(a_full(warp), b_bytes);
        bulk_g2s(b_base + (uint32_t)warp * b_stride, p.wimg + (size_t)kb * p.n_pad * kKBlock, b_bytes, a_full(warp));
      }
      issue_copies<false>(p.ga, stage, lane, cnt, list, a_full(warp));
#pragma unroll
      for (int i = 0; i < 4 * kMaxSegs; ++i) nbv[i] = nbn[i];
      kb = n_kb; lt = n_lt; tile = n_tile;
      ++round;
    }
  } else if (warp < S + 4) {
    // =================================================================== epilogue
    const int ew = warp & 3;  // a warp may only touch TMEM lanes 32*(warp%4) .. +31
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
        if (row < p.ga.n_out) {
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
  } else if (warp == S + 4) {
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
          const bool tr = p.trace && blockIdx.x == 0 && it < 512;
          if (tr) p.trace[4 * it + 0] = p.trace[4 * it + 1] = clock64();
          if (!mbar_wait(a_full(s), (it / (uint32_t)S) & 1u, abort_flag)) { ok = false; break; }
          if (tr) p.trace[4 * it + 2] = clock64();
          tc_fence_after();
          const int rem = p.ga.nq * 4 - kb * kKBlock;  // virtual-K elements left
          const int ksteps = rem >= kKBlock ? kKBlock / 8 : (rem + 7) >> 3;
          const uint32_t a_addr = a_base + (uint32_t)s * kStageBytes;
          const uint32_t b_addr = b_base + (uint32_t)s * b_stride;
          for (int ks = 0; ks < ksteps; ++ks)
            umma_tf32(d_tmem, make_desc_sw128(a_addr + ks * 32), make_desc_sw128(b_addr + ks * 32), idesc,
                      (kb | ks) != 0);
          umma_commit(a_empty(s));  // frees the stage and its weight block
          if (tr) p.trace[4 * it + 3] = clock64();
        }
        if (ok) umma_commit(acc_full(ab));
      }
    }
  }
done:
  asm volatile("cp.async.wait_all;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && *abort_flag && p.err) atomicExch(p.err, 1);
  if (warp == S + 4) {
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
  constexpr int kThreads = (S + 5) * 32;
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
  // input rows must be whole 16-byte chunks, at least 4 of them (at most 3 segments per K-block)
  return (*n_pad <= 256 && (c_in % 4) == 0 && c_in >= 16 && K <= kMaxK) ? 0 : 1;
}

// 1 when the tcgen05 kernel handles this shape
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
  p.ga = GatherArgs{in, tbl, tbl_stride, onehot_off, (int)n_out, c_in, K, c_in / 4, K * (c_in / 4)};
  p.out = out; p.wimg = wimg; p.c_out = c_out; p.kbt = kbt; p.n_pad = n_pad;
  p.num_tiles = (int)mm3d_cdiv(n_out, kTileM);
  p.tmem_cols = pow2_cols(2 * n_pad);
  p.err = mm3d_device_err_flag();
  {
    const char* t = getenv("MM3D_TC_TRACE");  // debug: device pointer (decimal) of a 2048-entry int64 buffer
    p.trace = t ? (long long*)strtoull(t, nullptr, 10) : nullptr;
  }
  const uint32_t b_stride = ((uint32_t)n_pad * 128u + 1023u) & ~1023u;
  // narrow N: 3 stages (48 KB + small weight blocks) so that three CTAs share an SM; mid: 4; wide: 6
  const bool wide = n_pad > 96 && n_pad <= 128, narrow = n_pad <= 32;  // (N > 128: 24-32 KB weight blocks, 4 stages)
  const int S = wide ? 6 : narrow ? 3 : 4;
  p.b_stages = S;
  const size_t smem = 1024 + (size_t)S * kStageBytes + (size_t)S * b_stride + (size_t)S * kListBytes +
                      8 * (2 * kMaxStages + 4) + 64;
  int rc;
  if (wide)        rc = onehot_off ? launch_conv_tc<6, true>(p, smem, stream) : launch_conv_tc<6, false>(p, smem, stream);
  else if (narrow) rc = onehot_off ? launch_conv_tc<3, true>(p, smem, stream) : launch_conv_tc<3, false>(p, smem, stream);
  else             rc = onehot_off ? launch_conv_tc<4, true>(p, smem, stream) : launch_conv_tc<4, false>(p, smem, stream);
  if (rc) return rc;
  mm3d_count_launches(2);
  MM3D_CHECK_LAUNCH("mm3d_conv_fwd_tc");
  return MM3D_OK;
}
