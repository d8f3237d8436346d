// Row plans (plan.cuh): order the output rows of a rule table by neighbour mask, permute the table
// accordingly, record per tile of 128 rows which offsets are present at all, and list the tiles by
// cost.  One launch per table.
//
// One CTA owns a chunk of 8192 consecutive rows (64 tiles): it builds a sort key per row from the
// table, sorts the chunk with a stable LSD radix sort in shared memory (digits of up to 10 bits,
// per-warp histograms, ballot ranking: no atomics, so the order is the same on every run), then
// writes the permutation, the permuted table (reads stay inside the chunk's 32 KB window of each
// table plane) and the tile masks.  Sorting per chunk instead of globally keeps it to one launch;
// it costs ~15 % more non-empty blocks than a global sort.  The last CTA to finish orders the tiles.
//
// Sort keys.  3^3 table: bit 18 = "has any corner neighbour", bits 6..17 = the 12 edge offsets,
// bits 0..5 = the 6 face offsets (the centre is always present): rarest first measured best, and
// folding the 8 corner bits into one loses nothing measurable (0.194 vs 0.191 non-empty blocks)
// while saving a radix pass.  Other dense tables: bit k = offset k present.  (parent, offset)
// tables: the offset.
#include <stdlib.h>

#include "plan.cuh"

namespace {

constexpr int kChunk = 8192;
constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxIters = kChunk / kThreads;  // 8 rounds of 1024 elements
constexpr int kDigitBits = 10;                // widest radix digit

struct BitPos {
  uint8_t p[32];   // sort-key bit of offset k
  uint8_t en[32];  // 1 = offset k is part of the key
};

struct PlanSmem {
  uint32_t keys[2][kChunk];
  uint16_t idx[2][kChunk];
  uint16_t hist[kWarps << kDigitBits];
  uint32_t tmask[kChunk / 128];
  uint16_t dstart[1 << kDigitBits];  // pre-ordering: first sorted position of every coarse digit in the chunk
  int warp_sums[kWarps];
  int bin_base[32];
  int off_cnt[32];
  int is_last;
};

// One table of a batched build: the launch covers `blocks` CTAs for it starting at CTA `block0`.
struct PlanItem {
  const int32_t* tbl;
  int64_t tbl_stride;
  const uint8_t* onehot_off;
  const int32_t* n_dev;
  int32_t* perm;
  uint32_t* tile_mask;
  int32_t* order;
  uint32_t* off_tiles;
  int32_t* ptbl;
  int64_t pstride;
  // global pre-ordering (3^3 tables with more than one chunk; NULL = rows enter the chunks in natural order):
  int32_t* pre;      // rows stably sorted by the top kDigitBits bits of their key
  uint32_t* keys_g;  // key of every row
  uint32_t* chist;   // [chunk][digit] counts, then global start positions
  int K, nbits, kind, block0;  // kind: 0 = 3^3 table, 1 = other dense table with K <= 8, 2 = (parent, offset) K <= 8,
                               //       3 / 4 = generic dense / (parent, offset) with K <= 32
};
constexpr int kMaxBatch = 32;
struct PlanBatch {
  int n;
  PlanItem item[kMaxBatch];
  BitPos bp[2];  // [0] the 3^3 key layout, [1] identity (bit k = offset k)
};

// One stable radix pass over elements [0, 1024 * iters) of s.keys[cur] / s.idx[cur] on the digit (key >> shift) &
// (2^dbits - 1), into the other buffer.  Warp w owns the consecutive segment [w * seg, (w+1) * seg), seg = 32 * iters.
// Ranking inside a warp uses one ballot per digit bit (lanes with my digit = AND over the bits of ballot or ~ballot)
// -- match_any is much slower here -- and the peer masks of the histogram phase are reused by the scatter.
__device__ __forceinline__ void radix_pass(PlanSmem& s, int cur, int shift, int dbits, int iters) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const int nbins = 1 << dbits;
  const int seg = 32 * iters;
  for (int i = tid; i < kWarps * nbins / 2; i += kThreads) reinterpret_cast<uint32_t*>(s.hist)[i] = 0;
  __syncthreads();
  uint16_t* h = s.hist + warp * nbins;
  const uint32_t* kin = s.keys[cur];
  const uint16_t* iin = s.idx[cur];
  uint32_t peers[kMaxIters];
  // per-warp digit histogram of the warp's segment
#pragma unroll
  for (int it = 0; it < kMaxIters; ++it) {
    if (it < iters) {
      const uint32_t d = (kin[warp * seg + it * 32 + lane] >> shift) & (uint32_t)(nbins - 1);
      uint32_t pm = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < kDigitBits; ++b) {
        if (b < dbits) {
          const uint32_t bal = __ballot_sync(0xffffffffu, (d >> b) & 1u);
          pm &= ((d >> b) & 1u) ? bal : ~bal;
        }
      }
      peers[it] = pm;
      if ((pm & lt) == 0) h[d] += (uint16_t)__popc(pm);
      __syncwarp();
    }
  }
  __syncthreads();
  // exclusive scan over (digit major, warp minor): nbins * 32 counters, consecutive ones per thread
  {
    const int per = nbins * kWarps / kThreads;  // 32 (10 bits), 16, 8, ...; 0: only threads < nbins * 32 work
    int sum = 0;
    if (per >= 1) {
      for (int j = 0; j < per; ++j) {
        const int f = tid * per + j;  // flat index = digit * 32 + warp
        sum += s.hist[(f & 31) * nbins + (f >> 5)];
      }
    } else if (tid < nbins * kWarps) {
      sum = s.hist[(tid & 31) * nbins + (tid >> 5)];
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
    if (per >= 1) {
      for (int j = 0; j < per; ++j) {
        const int f = tid * per + j;
        const int a = (f & 31) * nbins + (f >> 5);
        const int v = s.hist[a];
        s.hist[a] = (uint16_t)run;
        run += v;
      }
    } else if (tid < nbins * kWarps) {
      s.hist[(tid & 31) * nbins + (tid >> 5)] = (uint16_t)run;
    }
  }
  __syncthreads();
  // stable scatter
  uint32_t* kout = s.keys[cur ^ 1];
  uint16_t* iout = s.idx[cur ^ 1];
#pragma unroll
  for (int it = 0; it < kMaxIters; ++it) {
    if (it < iters) {
      const int i = warp * seg + it * 32 + lane;
      const uint32_t key = kin[i];
      const uint16_t id = iin[i];
      const uint32_t d = (key >> shift) & (uint32_t)(nbins - 1);
      const uint32_t pm = peers[it];
      const int rank = __popc(pm & lt);
      const int pos = (int)h[d] + rank;
      __syncwarp();
      if (rank == 0) h[d] = (uint16_t)(pos + __popc(pm));
      __syncwarp();
      kout[pos] = key;
      iout[pos] = id;
    }
  }
  __syncthreads();
}

// KT = compile-time bound of the offset loops (27, 8 or 32 = generic); K <= KT is the real count.
template <int KT, bool ONEHOT>
__device__ __forceinline__ void build_plan_chunk(PlanSmem& s, const PlanItem& it_, const BitPos& bp, int chunk,
                                                 unsigned int* __restrict__ done_counter) {
  const int32_t* __restrict__ tbl = it_.tbl;
  const int64_t tbl_stride = it_.tbl_stride;
  const uint8_t* __restrict__ onehot_off = it_.onehot_off;
  const int K = it_.K, nbits = it_.nbits;
  int32_t* __restrict__ perm = it_.perm;
  uint32_t* __restrict__ tile_mask = it_.tile_mask;
  int32_t* __restrict__ order = it_.order;
  uint32_t* __restrict__ off_tiles = it_.off_tiles;
  int32_t* __restrict__ ptbl = it_.ptbl;
  const int64_t pstride = it_.pstride;
  const int64_t n = *it_.n_dev;
  const int64_t base = (int64_t)chunk * kChunk;
  if (base >= n) return;
  const int cnt = (int)min((int64_t)kChunk, n - base);
  const int cnt_pad = (cnt + 127) & ~127;
  const int iters = (cnt + kThreads - 1) / kThreads;  // rounds of 1024 elements that contain rows
  const int n_sort = iters * kThreads;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t sentinel = nbits >= 32 ? 0xFFFFFFFFu : (1u << nbits) - 1u;  // sorts last (stable: after equal valid keys)

  // ---- keys
  for (int i = tid; i < n_sort; i += kThreads) {
    uint32_t key = sentinel;
    if (i < cnt && it_.pre) {
      key = __ldcg(it_.keys_g + __ldcg(it_.pre + base + i));  // (pre-ordered: the keys were computed by k_plan_keys)
    } else if (i < cnt) {
      const int64_t row = base + i;
      if (ONEHOT) {
        key = (uint32_t)__ldg(onehot_off + row);
      } else {
        int v[KT];  // all K loads of the row in flight at once
#pragma unroll
        for (int k = 0; k < KT; ++k) v[k] = k < K ? __ldg(tbl + (int64_t)k * tbl_stride + row) : -1;
        asm volatile("" ::: "memory");  // keep the loads together (the compiler would sink each to its use)
        key = 0;
#pragma unroll
        for (int k = 0; k < KT; ++k) key |= ((uint32_t)(v[k] >= 0) & bp.en[k]) << bp.p[k];
      }
    }
    s.keys[0][i] = key;
    s.idx[0][i] = (uint16_t)i;
  }
  if (tid < kChunk / 128) s.tmask[tid] = 0;
  __syncthreads();

  // ---- stable LSD radix sort of elements [0, n_sort): ceil(nbits / 10) passes
  int cur = 0;
  const int passes = (nbits + kDigitBits - 1) / kDigitBits;
  const int dbits = (nbits + passes - 1) / passes;
  for (int shift = 0; shift < nbits; shift += dbits) {
    radix_pass(s, cur, shift, dbits, iters);
    cur ^= 1;
  }

  // ---- permutation, permuted table, tile masks
  const uint16_t* sidx = s.idx[cur];
#pragma unroll 1
  for (int i = tid; i < cnt_pad; i += kThreads) {  // warp-uniform bound (multiple of 128) and tile (i >> 7)
    int64_t r = i < cnt ? base + (int64_t)sidx[i] : -1;
    if (r >= 0 && it_.pre) r = __ldcg(it_.pre + r);  // local element -> row of the pre-ordered sequence
    perm[base + i] = (int32_t)r;
    int par = -1, off = -1;
    if (ONEHOT && r >= 0) {
      par = __ldg(tbl + r);
      off = (int)__ldg(onehot_off + r);
    }
    int v[KT];  // all K entries of the row in flight at once
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      v[k] = -1;
      if (k < K && r >= 0) v[k] = ONEHOT ? (off == k ? par : -1) : __ldg(tbl + (int64_t)k * tbl_stride + r);
    }
    asm volatile("" ::: "memory");
    uint32_t wmask = 0;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      if (k < K) ptbl[(int64_t)k * pstride + base + i] = v[k];
      if (__ballot_sync(0xffffffffu, v[k] >= 0)) wmask |= 1u << k;
    }
    if (lane == 0 && wmask) atomicOr(&s.tmask[i >> 7], wmask);
  }
  __syncthreads();
  if (tid < cnt_pad / 128) {
    const uint32_t m = s.tmask[tid];
    tile_mask[base / 128 + tid] = m ? m : 1u;  // a tile without any input still runs one (all-zero) block
  }

  // ---- the last CTA to get here lists the tiles by descending number of non-empty offsets (stable):
  // warp w owns the bin popcount == 32 - w, the bins are scanned, every warp compacts its tiles in order.
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int active = (unsigned int)((n + kChunk - 1) / kChunk);  // CTAs that got past the row-count test
    const unsigned int prev = atomicAdd(done_counter, 1u);
    s.is_last = prev == active - 1;
    if (s.is_last) *done_counter = 0;  // ready for the next build on this stream
  }
  __syncthreads();
  if (!s.is_last) return;
  __threadfence();
  const int T = (int)((n + 127) / 128);
  const int want = 32 - warp;
  if (tid < 32) s.bin_base[tid] = 0;
  if (tid < 32) s.off_cnt[tid] = 0;
  __syncthreads();
  for (int t = tid; t < T; t += kThreads) {
    const uint32_t m = __ldcg(tile_mask + t);
    atomicAdd(&s.bin_base[32 - __popc(m)], 1);
    for (uint32_t r = m; r; r &= r - 1) atomicAdd(&s.off_cnt[__ffs(r) - 1], 1);
  }
  __syncthreads();
  if (tid < 32) off_tiles[tid] = (uint32_t)s.off_cnt[tid];
  if (warp == 0) {
    const int v = s.bin_base[lane];
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int x = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += x;
    }
    s.bin_base[lane] = incl - v;
  }
  __syncthreads();
  int pos = s.bin_base[warp];
  uint32_t* sm = &s.keys[0][0];  // the sort buffers are free now: stage 16384 masks at a time
  for (int b0 = 0; b0 < T; b0 += 2 * kChunk) {
    const int nb = min(2 * kChunk, T - b0);
    __syncthreads();
    for (int t = tid; t < nb; t += kThreads) sm[t] = __ldcg(tile_mask + b0 + t);
    __syncthreads();
    for (int t0 = 0; t0 < nb; t0 += 32) {
      const int t = t0 + lane;
      const bool mine = t < nb && __popc(sm[t]) == want;
      const unsigned bal = __ballot_sync(0xffffffffu, mine);
      if (mine) order[pos + __popc(bal & ((1u << lane) - 1u))] = b0 + t;
      pos += __popc(bal);
    }
  }
}

// ---- global pre-ordering of the rows of 3^3 tables ------------------------------------------------------------------
// Sorting inside chunks of 8192 consecutive rows leaves 20-45 % more non-empty (tile, offset) blocks than a global sort
// (measured on nuScenes-shaped scans, levels 0-2: tools/plan_fill_stats.py).  Three small kernels in front of the
// chunk builder close most of that gap: rows are first ordered GLOBALLY, stably, by the top 10 bits of their key (a
// counting sort: per-chunk digit counts, one scan, per-chunk stable scatter with the radix pass above), and the chunk
// builder then sorts chunks of that sequence by the full key.  No atomics decide a position: the order is the same on
// every run.
template <int KT>
__device__ __forceinline__ uint32_t row_key(const PlanItem& it, const BitPos& bp, int64_t row) {
  int v[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) v[k] = k < it.K ? __ldg(it.tbl + (int64_t)k * it.tbl_stride + row) : -1;
  asm volatile("" ::: "memory");
  uint32_t key = 0;
#pragma unroll
  for (int k = 0; k < KT; ++k) key |= ((uint32_t)(v[k] >= 0) & bp.en[k]) << bp.p[k];
  return key;
}

// keys of a chunk's rows + its digit counts
__global__ void __launch_bounds__(kThreads, 1)
k_plan_keys(const __grid_constant__ PlanBatch batch) {
  __shared__ uint32_t cnt_s[1 << kDigitBits];
  mm3d_griddep_wait();
  int i = 0;
  while (i + 1 < batch.n && (int)blockIdx.x >= batch.item[i + 1].block0) ++i;
  const PlanItem& it = batch.item[i];
  const int chunk = (int)blockIdx.x - it.block0;
  if (!it.pre) return;
  const int64_t n = *it.n_dev, base = (int64_t)chunk * kChunk;
  if (base >= n) return;
  const int cnt = (int)min((int64_t)kChunk, n - base);
  for (int d = threadIdx.x; d < (1 << kDigitBits); d += kThreads) cnt_s[d] = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < cnt; j += kThreads) {
    const uint32_t key = row_key<27>(it, batch.bp[0], base + j);
    it.keys_g[base + j] = key;
    atomicAdd(&cnt_s[key >> (it.nbits - kDigitBits)], 1u);  // (counts do not depend on the order of the adds)
  }
  __syncthreads();
  for (int d = threadIdx.x; d < (1 << kDigitBits); d += kThreads) it.chist[(size_t)chunk * (1 << kDigitBits) + d] = cnt_s[d];
}

// counts -> global start position of every (digit, chunk) segment: digit major, chunk minor.  One CTA per table,
// thread d owns digit d.
__global__ void __launch_bounds__(1 << kDigitBits, 1)
k_plan_scan(const __grid_constant__ PlanBatch batch) {
  __shared__ uint32_t wsum[32];
  mm3d_griddep_wait();
  const PlanItem& it = batch.item[blockIdx.x];
  if (!it.pre) return;
  const int64_t n = *it.n_dev;
  const int chunks = (int)((n + kChunk - 1) / kChunk);
  const int d = threadIdx.x, lane = d & 31, warp = d >> 5;
  uint32_t total = 0;
  for (int c = 0; c < chunks; ++c) {
    uint32_t* p = it.chist + (size_t)c * (1 << kDigitBits) + d;
    const uint32_t v = *p;
    *p = total;
    total += v;
  }
  uint32_t incl = total;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t ws = wsum[lane];
    uint32_t wi = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    wsum[lane] = wi - ws;
  }
  __syncthreads();
  const uint32_t start = wsum[warp] + incl - total;
  for (int c = 0; c < chunks; ++c) it.chist[(size_t)c * (1 << kDigitBits) + d] += start;
}

// a chunk's rows to their global positions: one stable radix pass on the coarse digit inside the chunk, then every
// element goes to (start of its (digit, chunk) segment) + (its position among the chunk's rows of that digit)
__global__ void __launch_bounds__(kThreads, 1)
k_plan_preorder(const __grid_constant__ PlanBatch batch) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  PlanSmem& s = *reinterpret_cast<PlanSmem*>(smem_raw);
  mm3d_griddep_wait();
  int i = 0;
  while (i + 1 < batch.n && (int)blockIdx.x >= batch.item[i + 1].block0) ++i;
  const PlanItem& it = batch.item[i];
  const int chunk = (int)blockIdx.x - it.block0;
  if (!it.pre) return;
  const int64_t n = *it.n_dev, base = (int64_t)chunk * kChunk;
  if (base >= n) return;
  const int cnt = (int)min((int64_t)kChunk, n - base);
  const int iters = (cnt + kThreads - 1) / kThreads;
  const int n_sort = iters * kThreads;
  const int tid = threadIdx.x;
  const int shift = it.nbits - kDigitBits;
  const uint32_t sentinel = (1u << it.nbits) - 1u;  // (padding sorts behind the rows of the last digit)
  for (int j = tid; j < n_sort; j += kThreads) {
    s.keys[0][j] = j < cnt ? __ldcg(it.keys_g + base + j) : sentinel;
    s.idx[0][j] = (uint16_t)j;
  }
  __syncthreads();
  radix_pass(s, 0, shift, kDigitBits, iters);
  const uint32_t* ks = s.keys[1];
  const uint16_t* is = s.idx[1];
  for (int j = tid; j < n_sort; j += kThreads) {
    const uint32_t d = ks[j] >> shift;
    if (j == 0 || (ks[j - 1] >> shift) != d) s.dstart[d] = (uint16_t)j;
  }
  __syncthreads();
  const uint32_t* starts = it.chist + (size_t)chunk * (1 << kDigitBits);
  for (int j = tid; j < n_sort; j += kThreads) {
    const int id = (int)is[j];
    if (id >= cnt) continue;
    const uint32_t d = ks[j] >> shift;
    it.pre[__ldcg(starts + d) + (uint32_t)(j - (int)s.dstart[d])] = (int32_t)(base + id);
  }
}

// per-device completion counters of the plan builder, one per table of a batch (zero between builds)
unsigned int* g_done[64] = {nullptr};
unsigned int* done_counters() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!g_done[dev]) {
    unsigned int* p = nullptr;
    if (cudaMalloc(&p, kMaxBatch * sizeof(unsigned int)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, kMaxBatch * sizeof(unsigned int));
    g_done[dev] = p;
  }
  return g_done[dev];
}

__global__ void __launch_bounds__(kThreads, 1)
k_build_plans(const __grid_constant__ PlanBatch batch, unsigned int* __restrict__ done_counters) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  PlanSmem& s = *reinterpret_cast<PlanSmem*>(smem_raw);
  mm3d_griddep_wait();
  int i = 0;
  while (i + 1 < batch.n && (int)blockIdx.x >= batch.item[i + 1].block0) ++i;
  const PlanItem& it = batch.item[i];
  const int chunk = (int)blockIdx.x - it.block0;
  switch (it.kind) {
    case 0: build_plan_chunk<27, false>(s, it, batch.bp[0], chunk, done_counters + i); break;
    case 1: build_plan_chunk<8, false>(s, it, batch.bp[1], chunk, done_counters + i); break;
    case 2: build_plan_chunk<8, true>(s, it, batch.bp[1], chunk, done_counters + i); break;
    case 3: build_plan_chunk<32, false>(s, it, batch.bp[1], chunk, done_counters + i); break;
    default: build_plan_chunk<32, true>(s, it, batch.bp[1], chunk, done_counters + i); break;
  }
}

}  // namespace

extern "C" size_t mm3d_plan_bytes(int64_t n_cap, int K) {
  if (n_cap < 0 || K <= 0 || K > 32) return 0;
  return mm3d_plan_size(n_cap, K);
}

extern "C" int mm3d_build_plans(const mm3d_plan_desc* descs, int n_plans, mm3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MM3D_REQUIRE(n_plans >= 0 && (n_plans == 0 || descs), MM3D_ERR_INVALID, "plans: bad descriptor array");
  unsigned int* counters = done_counters();
  MM3D_REQUIRE(counters, MM3D_ERR_CUDA, "plans: could not allocate the completion counters");
  static bool once_dev[64] = {false};
  bool& once = once_dev[mm3d_device_slot()];
  if (!once) {
    MM3D_CUDA(cudaFuncSetAttribute(k_build_plans, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PlanSmem)));
    MM3D_CUDA(cudaFuncSetAttribute(k_plan_preorder, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PlanSmem)));
    once = true;
  }
  // MM3D_PLAN_NO_PREORDER=1: rows enter the chunks in natural order (the round-1 plans), for A/B measurements
  const bool no_preorder = getenv("MM3D_PLAN_NO_PREORDER") != nullptr;
  for (int first = 0; first < n_plans; first += kMaxBatch) {
    PlanBatch batch;
    batch.n = 0;
    int blocks = 0;
    bool any_pre = false;
    // key layouts: 3^3 -> bit 18 = any corner, 17..6 = edges, 5..0 = faces (centre left out); else bit k = offset k
    for (int k = 0; k < 32; ++k) {
      batch.bp[0].p[k] = 0; batch.bp[0].en[k] = 0;
      batch.bp[1].p[k] = (uint8_t)k; batch.bp[1].en[k] = 1;
    }
    int edge = 17, face = 5;
    for (int k = 0; k < 27; ++k) {
      const int dx = k / 9 - 1, dy = (k / 3) % 3 - 1, dz = k % 3 - 1;
      const int cls = abs(dx) + abs(dy) + abs(dz);
      if (cls == 0) continue;
      batch.bp[0].en[k] = 1;
      batch.bp[0].p[k] = (uint8_t)(cls == 3 ? 18 : cls == 2 ? edge-- : face--);
    }
    for (int i = first; i < n_plans && i < first + kMaxBatch; ++i) {
      const mm3d_plan_desc& d = descs[i];
      MM3D_REQUIRE(d.K > 0 && d.K <= 32, MM3D_ERR_INVALID, "plan: K must be in (0, 32]");
      MM3D_REQUIRE(d.n_cap >= 0 && d.n_cap < (1ll << 31) - 128, MM3D_ERR_UNSUPPORTED, "plan: too many rows");
      if (d.n_cap == 0) continue;
      MM3D_REQUIRE(d.tbl && d.n_dev && d.plan, MM3D_ERR_INVALID, "plan: null pointer");
      MM3D_REQUIRE(d.onehot_off || d.tbl_stride >= d.n_cap, MM3D_ERR_INVALID, "plan: tbl_stride < n_cap");
      MM3D_REQUIRE(d.plan_bytes >= mm3d_plan_size(d.n_cap, d.K), MM3D_ERR_WORKSPACE, "plan: buffer too small");
      MM3D_REQUIRE(((uintptr_t)d.plan & 255) == 0, MM3D_ERR_INVALID, "plan: buffer must be 256-byte aligned");
      const int64_t rows = d.n_rows_hint > 0 && d.n_rows_hint < d.n_cap ? d.n_rows_hint : d.n_cap;
      char* b = (char*)d.plan;
      PlanItem& it = batch.item[batch.n++];
      it.tbl = d.tbl; it.tbl_stride = d.tbl_stride; it.onehot_off = d.onehot_off; it.n_dev = d.n_dev;
      it.perm = (int32_t*)b;
      it.tile_mask = (uint32_t*)(b + mm3d_plan_off_mask(d.n_cap));
      it.order = (int32_t*)(b + mm3d_plan_off_order(d.n_cap));
      it.off_tiles = (uint32_t*)(b + mm3d_plan_off_cnt(d.n_cap));
      it.ptbl = (int32_t*)(b + mm3d_plan_off_tbl(d.n_cap));
      it.pstride = mm3d_plan_tiles(d.n_cap) * 128;
      it.K = d.K;
      if (d.onehot_off) {
        it.nbits = 1;
        while ((1 << it.nbits) < d.K) ++it.nbits;
        it.kind = d.K <= 8 ? 2 : 4;
      } else if (d.K == 27) {
        it.nbits = 19;
        it.kind = 0;
      } else {
        it.nbits = d.K;
        it.kind = d.K <= 8 ? 1 : 3;
      }
      it.pre = nullptr; it.keys_g = nullptr; it.chist = nullptr;
      if (it.kind == 0 && rows > kChunk && !no_preorder) {  // (a single chunk is sorted globally anyway)
        char* sc = b + mm3d_plan_off_scratch(d.n_cap, d.K);
        const size_t plane = ((size_t)mm3d_plan_tiles(d.n_cap) * 128 * 4 + 255) / 256 * 256;
        it.pre = (int32_t*)sc;
        it.keys_g = (uint32_t*)(sc + plane);
        it.chist = (uint32_t*)(sc + 2 * plane);
        any_pre = true;
      }
      it.block0 = blocks;
      blocks += (int)mm3d_cdiv(rows, kChunk);
    }
    if (blocks == 0) continue;
    if (any_pre) {
      MM3D_CUDA(mm3d_launch_pdl(k_plan_keys, dim3((unsigned)blocks), dim3(kThreads), 0, stream, batch));
      MM3D_CUDA(mm3d_launch_pdl(k_plan_scan, dim3((unsigned)batch.n), dim3(1 << kDigitBits), 0, stream, batch));
      MM3D_CUDA(mm3d_launch_pdl(k_plan_preorder, dim3((unsigned)blocks), dim3(kThreads), sizeof(PlanSmem), stream, batch));
      mm3d_count_launches(3);
    }
    MM3D_CUDA(mm3d_launch_pdl(k_build_plans, dim3((unsigned)blocks), dim3(kThreads), sizeof(PlanSmem), stream, batch, counters));
    mm3d_count_launches(1);
    MM3D_CHECK_LAUNCH("mm3d_build_plans");
  }
  return MM3D_OK;
}

extern "C" int mm3d_build_plan(const int32_t* tbl, int64_t tbl_stride, const uint8_t* onehot_off,
                               const int32_t* n_dev, int64_t n_cap, int K, void* plan, size_t plan_bytes,
                               mm3d_stream_t stream) {
  mm3d_plan_desc d;
  d.tbl = tbl; d.tbl_stride = tbl_stride; d.onehot_off = onehot_off; d.n_dev = n_dev; d.n_cap = n_cap;
  d.n_rows_hint = 0; d.K = K; d.plan = plan; d.plan_bytes = plan_bytes;
  return mm3d_build_plans(&d, 1, stream);
}
