// Gather producer shared by the tcgen05 forward/dgrad and wgrad kernels.
//
// One producer WARP owns one pipeline stage (a K-block = 8 sixteen-byte chunk columns x 128 rows).
// The rule tables of this network are 3-30 % dense, so a slot-per-lane scheme spends almost all of
// its instructions deciding that there is nothing to copy.  Here a lane owns a ROW (4 row groups of
// 32), reads that row's table entry once per (segment, row group) with a coalesced 128-byte load,
// and the rows that need any action are compacted (ballot + popc) into a small per-warp list in
// shared memory.  The warp then walks the list four rows at a time, 8 lanes per row, each lane
// issuing one 16-byte cp.async (real copy) or a zero-fill for a slot that held data the last time
// the stage was used.  Work is proportional to the pairs of the rule table, not to 27 x 128.
//
// A K-block covers chunk columns q = 8*kb .. 8*kb+7 of the virtual K; with cq chunks per input row
// these belong to up to 3 "segments" (offset k, first chunk cc, columns [col, col+ncols)).
#pragma once

#include "tc_common.cuh"

namespace tc {

constexpr int kGatherTile = 128;
constexpr int kMaxSegs = 3;                       // cq >= 4  =>  at most 3 segments per K-block
constexpr int kListEntries = kGatherTile * kMaxSegs;
constexpr int kListBytes = kListEntries * 8;      // per producer warp

struct GatherArgs {
  const float* in;
  const int32_t* tbl;
  int64_t tbl_stride;
  const uint8_t* onehot_off;
  int n_out, c_in, K, cq, nq;
};

template <bool BASE32>
__device__ __forceinline__ uint32_t chunk_off(int col, int row) {
  return BASE32 ? swz_base32(col, row) : (uint32_t)((col ^ (row & 7)) << 4);
}

// Table entries of K-block kb of the tile starting at row0 for this lane's 4 rows x up to 3 segments
// (-1 = absent, out of range, or segment not present).  Issued one item ahead of their use.
template <bool ONEHOT>
__device__ __forceinline__ void load_entries(const GatherArgs& a, bool tile_ok, int64_t row0, int kb, int lane,
                                             int (&nbv)[4 * kMaxSegs]) {
  const int q0 = kb * 8;
  const int k_first = q0 / a.cq;
  int cc = q0 - k_first * a.cq, col = 0;
  int par[4], off[4];
  if (ONEHOT) {
#pragma unroll
    for (int rg = 0; rg < 4; ++rg) {
      const int64_t row = row0 + rg * 32 + lane;
      const bool ok = tile_ok && row < a.n_out;
      par[rg] = ok ? __ldg(a.tbl + row) : -1;
      off[rg] = ok ? (int)__ldg(a.onehot_off + row) : -1;
    }
  }
#pragma unroll
  for (int sg = 0; sg < kMaxSegs; ++sg) {
    const int k = k_first + sg;
    const bool seg_ok = tile_ok && col < 8 && k < a.K;
#pragma unroll
    for (int rg = 0; rg < 4; ++rg) {
      int nb = -1;
      if (seg_ok) {
        if (ONEHOT) {
          nb = off[rg] == k ? par[rg] : -1;
        } else {
          const int64_t row = row0 + rg * 32 + lane;
          if (row < a.n_out) nb = __ldg(a.tbl + (int64_t)k * a.tbl_stride + row);
        }
      }
      nbv[sg * 4 + rg] = nb;
    }
    col += a.cq - cc;
    cc = 0;
  }
}

// Step 1 (needs no access to the stage, so it runs BEFORE the wait for the stage to be free):
// decide, for K-block kb, which (row, columns) of the stage must be written or re-zeroed and compact
// them into the warp's list.  `filled` holds, per lane, 8 column bits for each of its 4 rows
// (bit rg*8 + col = that slot holds data) and is advanced to the state after this K-block.
// All ballots are independent and the stores come last, so the 12 (segment, row group) tasks overlap.
__device__ __forceinline__ int build_list(const GatherArgs& a, int kb, int lane, const int (&nbv)[4 * kMaxSegs],
                                          uint32_t& filled, uint32_t list) {
  const int q0 = kb * 8;
  const int k_first = q0 / a.cq;
  int cc = q0 - k_first * a.cq, col = 0;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t meta[4 * (kMaxSegs + 1)], bal[4 * (kMaxSegs + 1)];
  int src[4 * (kMaxSegs + 1)];
#pragma unroll
  for (int sg = 0; sg <= kMaxSegs; ++sg) {
    // real segments first; one trailing pseudo segment covers the columns past the end of the virtual
    // K so that stale data from an earlier K-block on this stage gets cleared
    const int k = k_first + sg;
    const bool real = sg < kMaxSegs && col < 8 && k < a.K;
    const int ncols = real ? min(8 - col, a.cq - cc) : 8 - col;
    const uint32_t segmask = ncols > 0 ? ((1u << ncols) - 1u) << col : 0u;
    const int delta = cc - col + 8;  // source chunk = column + delta - 8
#pragma unroll
    for (int rg = 0; rg < 4; ++rg) {
      const int nb = real ? nbv[(sg < kMaxSegs ? sg : 0) * 4 + rg] : -1;
      const uint32_t old = (filled >> (rg * 8)) & 0xFFu;
      const uint32_t fillm = nb >= 0 ? segmask : 0u;
      const uint32_t clearm = old & segmask & ~fillm;
      filled = (filled & ~(segmask << (rg * 8))) | (fillm << (rg * 8));
      const bool act = (fillm | clearm) != 0u;
      bal[sg * 4 + rg] = __ballot_sync(0xffffffffu, act);
      meta[sg * 4 + rg] = act ? ((uint32_t)(rg * 32 + lane) | (fillm << 8) | (clearm << 16) | ((uint32_t)delta << 24)) : 0u;
      src[sg * 4 + rg] = nb;
    }
    col += ncols;
    cc = 0;
  }
  int cnt = 0;
#pragma unroll
  for (int t = 0; t < 4 * (kMaxSegs + 1); ++t) {
    if (meta[t]) {
      const int pos = cnt + __popc(bal[t] & lt);
      asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(list + (uint32_t)pos * 8u), "r"((uint32_t)src[t]), "r"(meta[t]) : "memory");
    }
    cnt += __popc(bal[t]);
  }
  __syncwarp();
  return cnt;
}

// Step 2 (after the stage is free): walk the list, four rows per pass, 8 lanes per row, one 16-byte
// cp.async (copy or zero fill) per lane; ends with the lanes' cp.async arrivals on `full_bar`.
template <bool BASE32>
__device__ __forceinline__ void issue_copies(const GatherArgs& a, uint32_t stage, int lane, int cnt, uint32_t list,
                                             uint32_t full_bar) {
  const int c = lane & 7;
  for (int e0 = lane >> 3; e0 < cnt; e0 += 8) {
    uint32_t enb[2], meta[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      meta[u] = 0;
      if (e0 + 4 * u < cnt)
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(enb[u]), "=r"(meta[u]) : "r"(list + (uint32_t)(e0 + 4 * u) * 8u) : "memory");
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = (int)(meta[u] & 0xFFu);
      const uint32_t dst = stage + (uint32_t)r * 128u + chunk_off<BASE32>(c, r);
      if ((meta[u] >> (8 + c)) & 1u) {
        const int src_chunk = c + (int)(meta[u] >> 24) - 8;
        cp_async16(dst, a.in + ((uint32_t)enb[u] * (uint32_t)a.c_in + (uint32_t)src_chunk * 4u), 16u);
      } else if ((meta[u] >> (16 + c)) & 1u) {
        cp_async16(dst, a.in, 0u);
      }
    }
  }
  cp_async_arrive(full_bar);
  __syncwarp();  // the list is rewritten by the next K-block
}

}  // namespace tc
