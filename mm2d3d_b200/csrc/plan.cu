// Row plans (plan.cuh): order the output rows of a rule table by neighbour mask, permute the table
// accordingly and record, per tile of 128 rows, which offsets are present at all.
//
// One CTA owns a chunk of 8192 consecutive rows (64 tiles): it builds a sort key per row from the
// table (bit per offset; for the 3^3 table the rarest offsets -- corners, then edges, then faces --
// are the most significant bits, which measured best), sorts the chunk with a stable LSD radix sort
// in shared memory (8-bit digits, per-warp histograms, match_any ranking: no atomics, so the order
// is the same on every run), then writes the permutation, the permuted table (reads stay inside the
// chunk's 32 KB window of each table plane) and the tile masks.  Sorting per chunk instead of
// globally keeps it to one launch per table; it costs ~15 % more non-empty blocks than a global sort.
#include "plan.cuh"

namespace {

constexpr int kChunk = 8192;
constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kSeg = kChunk / kWarps;  // 256 consecutive elements per warp
constexpr int kIters = kSeg / 32;      // 8
constexpr int kPerThread = kChunk / kThreads;

struct BitPos {
  uint8_t p[32];  // sort-key bit of offset k (255 = not part of the key)
};

struct PlanSmem {
  uint32_t keys[2][kChunk];
  uint16_t idx[2][kChunk];
  uint16_t hist[kWarps * 256];
  uint32_t tmask[kChunk / 128];
  int warp_sums[kWarps];
};

template <bool ONEHOT>
__global__ void __launch_bounds__(kThreads, 1)
k_build_plan(const int32_t* __restrict__ tbl, int64_t tbl_stride, const uint8_t* __restrict__ onehot_off,
             const int32_t* __restrict__ n_dev, int K, int nbits, BitPos bp, int32_t* __restrict__ perm,
             uint32_t* __restrict__ tile_mask, int32_t* __restrict__ ptbl, int64_t pstride) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  PlanSmem& s = *reinterpret_cast<PlanSmem*>(smem_raw);
  const int64_t n = *n_dev;
  const int64_t base = (int64_t)blockIdx.x * kChunk;
  if (base >= n) return;
  const int cnt = (int)min((int64_t)kChunk, n - base);
  const int cnt_pad = (cnt + 127) & ~127;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t sentinel = nbits >= 32 ? 0xFFFFFFFFu : (1u << nbits) - 1u;  // sorts last (stable: after equal valid keys)

  // ---- keys
  for (int i = tid; i < kChunk; i += kThreads) {
    uint32_t key = sentinel;
    if (i < cnt) {
      const int64_t row = base + i;
      if (ONEHOT) {
        key = (uint32_t)__ldg(onehot_off + row);
      } else {
        key = 0;
        for (int k = 0; k < K; ++k)
          if (bp.p[k] != 255 && __ldg(tbl + (int64_t)k * tbl_stride + row) >= 0) key |= 1u << bp.p[k];
      }
    }
    s.keys[0][i] = key;
    s.idx[0][i] = (uint16_t)i;
  }
  if (tid < kChunk / 128) s.tmask[tid] = 0;
  __syncthreads();

  // ---- stable LSD radix sort of the chunk, 8 bits per pass
  int cur = 0;
  const uint32_t lt = (1u << lane) - 1u;
  for (int shift = 0; shift < nbits; shift += 8) {
    for (int i = tid; i < kWarps * 256 / 2; i += kThreads) reinterpret_cast<uint32_t*>(s.hist)[i] = 0;
    __syncthreads();
    uint16_t* h = s.hist + warp * 256;
    const uint32_t* kin = s.keys[cur];
    const uint16_t* iin = s.idx[cur];
    // per-warp digit histogram of the warp's segment
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
      const uint32_t d = (kin[warp * kSeg + it * 32 + lane] >> shift) & 255u;
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      if ((peers & lt) == 0) h[d] += (uint16_t)__popc(peers);
      __syncwarp();
    }
    __syncthreads();
    // exclusive scan over (digit major, warp minor)
    {
      const int d = tid >> 2, w0 = (tid & 3) * 8;
      int v[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = s.hist[(w0 + j) * 256 + d];
        sum += v[j];
      }
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) s.warp_sums[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        int ws = s.warp_sums[lane], wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, wi, o);
          if (lane >= o) wi += t;
        }
        s.warp_sums[lane] = wi - ws;
      }
      __syncthreads();
      int run = s.warp_sums[warp] + incl - sum;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s.hist[(w0 + j) * 256 + d] = (uint16_t)run;
        run += v[j];
      }
    }
    __syncthreads();
    // stable scatter
    uint32_t* kout = s.keys[cur ^ 1];
    uint16_t* iout = s.idx[cur ^ 1];
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
      const int i = warp * kSeg + it * 32 + lane;
      const uint32_t key = kin[i];
      const uint16_t id = iin[i];
      const uint32_t d = (key >> shift) & 255u;
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      const int rank = __popc(peers & lt);
      const int pos = (int)h[d] + rank;
      __syncwarp();
      if (rank == 0) h[d] = (uint16_t)(pos + __popc(peers));
      __syncwarp();
      kout[pos] = key;
      iout[pos] = id;
    }
    __syncthreads();
    cur ^= 1;
  }

  // ---- permutation, permuted table, tile masks
  const uint16_t* sidx = s.idx[cur];
#pragma unroll 1
  for (int j = 0; j < kPerThread; ++j) {
    const int i = tid + j * kThreads;  // warp-uniform tile: i >> 7
    if (i >= cnt_pad) break;
    const int64_t r = i < cnt ? base + (int64_t)sidx[i] : -1;
    perm[base + i] = (int32_t)r;
    int par = -1, off = -1;
    if (ONEHOT && r >= 0) {
      par = __ldg(tbl + r);
      off = (int)__ldg(onehot_off + r);
    }
    for (int k = 0; k < K; ++k) {
      int e = -1;
      if (r >= 0) e = ONEHOT ? (off == k ? par : -1) : __ldg(tbl + (int64_t)k * tbl_stride + r);
      ptbl[(int64_t)k * pstride + base + i] = e;
      const unsigned bal = __ballot_sync(0xffffffffu, e >= 0);
      if (lane == 0 && bal) atomicOr(&s.tmask[i >> 7], 1u << k);
    }
  }
  __syncthreads();
  if (tid < cnt_pad / 128) {
    const uint32_t m = s.tmask[tid];
    tile_mask[base / 128 + tid] = m ? m : 1u;  // a tile without any input still runs one (all-zero) block
  }
}

// Tiles by descending number of non-empty offsets, stable.  One CTA, warp w owns the bin popcount == 32 - w:
// it counts its tiles, the bins are scanned, then it compacts its tiles in order (ballot ranks).
__global__ void __launch_bounds__(1024, 1)
k_plan_order(const uint32_t* __restrict__ tile_mask, const int32_t* __restrict__ n_dev, int32_t* __restrict__ order) {
  __shared__ int bin_base[32];
  const int T = (int)(((int64_t)*n_dev + 127) / 128);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int want = 32 - warp;
  int total = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const bool mine = t < T && __popc(__ldg(tile_mask + t)) == want;
    total += __popc(__ballot_sync(0xffffffffu, mine));
  }
  if (lane == 0) bin_base[warp] = total;
  __syncthreads();
  if (warp == 0) {
    const int v = bin_base[lane];
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int x = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += x;
    }
    bin_base[lane] = incl - v;
  }
  __syncthreads();
  int pos = bin_base[warp];
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const bool mine = t < T && __popc(__ldg(tile_mask + t)) == want;
    const unsigned bal = __ballot_sync(0xffffffffu, mine);
    if (mine) order[pos + __popc(bal & ((1u << lane) - 1u))] = t;
    pos += __popc(bal);
  }
}

}  // namespace

extern "C" size_t mm3d_plan_bytes(int64_t n_cap, int K) {
  if (n_cap < 0 || K <= 0 || K > 32) return 0;
  return mm3d_plan_size(n_cap, K);
}

extern "C" int mm3d_build_plan(const int32_t* tbl, int64_t tbl_stride, const uint8_t* onehot_off,
                               const int32_t* n_dev, int64_t n_cap, int K, void* plan, size_t plan_bytes,
                               mm3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MM3D_REQUIRE(K > 0 && K <= 32, MM3D_ERR_INVALID, "plan: K must be in (0, 32]");
  MM3D_REQUIRE(n_cap >= 0 && n_cap < (1ll << 31) - 128, MM3D_ERR_UNSUPPORTED, "plan: too many rows");
  if (n_cap == 0) return MM3D_OK;
  MM3D_REQUIRE(tbl && n_dev && plan, MM3D_ERR_INVALID, "plan: null pointer");
  MM3D_REQUIRE(onehot_off || tbl_stride >= n_cap, MM3D_ERR_INVALID, "plan: tbl_stride < n_cap");
  MM3D_REQUIRE(plan_bytes >= mm3d_plan_size(n_cap, K), MM3D_ERR_WORKSPACE, "plan: buffer too small");
  MM3D_REQUIRE(((uintptr_t)plan & 255) == 0, MM3D_ERR_INVALID, "plan: buffer must be 256-byte aligned");
  char* b = (char*)plan;
  int32_t* perm = (int32_t*)b;
  uint32_t* tmask = (uint32_t*)(b + mm3d_plan_off_mask(n_cap));
  int32_t* order = (int32_t*)(b + mm3d_plan_off_order(n_cap));
  int32_t* ptbl = (int32_t*)(b + mm3d_plan_off_tbl(n_cap));
  const int64_t pstride = mm3d_plan_tiles(n_cap) * 128;

  BitPos bp;
  int nbits;
  for (int k = 0; k < 32; ++k) bp.p[k] = 255;
  if (onehot_off) {
    nbits = 1;
    while ((1 << nbits) < K) ++nbits;
  } else if (K == 27) {
    // 3^3: centre (always present) left out; corners most significant, then edges, then faces
    int bit = 25;
    for (int cls = 3; cls >= 1; --cls)
      for (int k = 0; k < 27; ++k) {
        const int dx = k / 9 - 1, dy = (k / 3) % 3 - 1, dz = k % 3 - 1;
        if (abs(dx) + abs(dy) + abs(dz) == cls) bp.p[k] = (uint8_t)bit--;
      }
    nbits = 26;
  } else {
    for (int k = 0; k < K; ++k) bp.p[k] = (uint8_t)k;
    nbits = K;
  }
  const unsigned grid = (unsigned)mm3d_cdiv(n_cap, kChunk);
  const size_t smem = sizeof(PlanSmem);
  if (onehot_off) {
    static bool once = false;
    if (!once) {
      MM3D_CUDA(cudaFuncSetAttribute(k_build_plan<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      once = true;
    }
    k_build_plan<true><<<grid, kThreads, smem, stream>>>(tbl, tbl_stride, onehot_off, n_dev, K, nbits, bp, perm, tmask,
                                                         ptbl, pstride);
  } else {
    static bool once = false;
    if (!once) {
      MM3D_CUDA(cudaFuncSetAttribute(k_build_plan<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      once = true;
    }
    k_build_plan<false><<<grid, kThreads, smem, stream>>>(tbl, tbl_stride, onehot_off, n_dev, K, nbits, bp, perm, tmask,
                                                          ptbl, pstride);
  }
  k_plan_order<<<1, 1024, 0, stream>>>(tmask, n_dev, order);
  mm3d_count_launches(2);
  MM3D_CHECK_LAUNCH("mm3d_build_plan");
  return MM3D_OK;
}
