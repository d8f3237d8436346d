// Row plan of a rule table (shared by plan.cu and the tcgen05 kernels).
//
// The tensor-core kernels process OUTPUT rows in tiles of 128.  Which rows share a tile is free
// (every output row is produced by exactly one accumulator lane and stored once), so the plan
// orders rows by their neighbour mask: rows with the same set of present offsets sit together and
// most (tile, offset) blocks become EMPTY -- the kernels skip them -- while the non-empty ones are
// 60-80 % dense instead of 10-30 %.  A plan is one device buffer:
//
//   perm      int32 [T*128]      tile t holds output rows perm[128 t .. 128 t + 127] (-1 = padding)
//   tile_mask uint32[T]          bit k set  <=>  some row of tile t has an input at offset k
//   order     int32 [T]          the tiles [0, ceil(n/128)) by descending popcount(mask) (stable): kernels deal
//                                tiles to their CTAs from this list in serpentine order, which balances the
//                                number of non-empty blocks per CTA (a tile has 1..K of them)
//   off_tiles uint32[32]         number of tiles that contain offset k (the weight-gradient kernel shares its CTAs
//                                among offset groups in proportion)
//   tbl       int32 [K][T*128]   tbl[k][128 t + r] = input row of perm[128 t + r] at offset k, or -1
//   (3^3 tables: builder scratch behind it, mm3d_plan_off_scratch)
//
// with T = ceil(n_cap / 128).  Features stay in SparseConvNet row order everywhere; only the order
// in which rows are computed changes, and results do not depend on it (an all-zero K-block adds 0).
#pragma once

#include "common.cuh"

struct Mm3dPlanView {
  const int32_t* perm;
  const uint32_t* tile_mask;
  const int32_t* order;
  const uint32_t* off_tiles;
  const int32_t* tbl;
  int64_t stride;  // T * 128
};

__host__ __device__ inline int64_t mm3d_plan_tiles(int64_t n_cap) { return (n_cap + 127) / 128; }

__host__ __device__ inline size_t mm3d_plan_off_mask(int64_t n_cap) {
  return ((size_t)mm3d_plan_tiles(n_cap) * 128 * 4 + 255) / 256 * 256;
}
__host__ __device__ inline size_t mm3d_plan_off_order(int64_t n_cap) {
  return mm3d_plan_off_mask(n_cap) + ((size_t)mm3d_plan_tiles(n_cap) * 4 + 255) / 256 * 256;
}
__host__ __device__ inline size_t mm3d_plan_off_cnt(int64_t n_cap) {
  return mm3d_plan_off_order(n_cap) + ((size_t)mm3d_plan_tiles(n_cap) * 4 + 255) / 256 * 256;
}
__host__ __device__ inline size_t mm3d_plan_off_tbl(int64_t n_cap) { return mm3d_plan_off_cnt(n_cap) + 256; }
// behind the permuted table: scratch of the builder for 3^3 tables (the global pre-ordering of plan.cu): pre int32[T*128],
// keys uint32[T*128], per-chunk digit counters uint32[ceil(n_cap / 8192)][1024]
__host__ __device__ inline size_t mm3d_plan_off_scratch(int64_t n_cap, int K) {
  return mm3d_plan_off_tbl(n_cap) + ((size_t)K * mm3d_plan_tiles(n_cap) * 128 * 4 + 255) / 256 * 256;
}
__host__ __device__ inline size_t mm3d_plan_size(int64_t n_cap, int K) {
  size_t s = mm3d_plan_off_scratch(n_cap, K);
  if (K == 27) s += 2 * (((size_t)mm3d_plan_tiles(n_cap) * 128 * 4 + 255) / 256 * 256) + (size_t)((n_cap + 8191) / 8192) * 1024 * 4 + 256;
  return s;
}

inline Mm3dPlanView mm3d_plan_view(const void* plan, int64_t n_cap) {
  const char* b = (const char*)plan;
  Mm3dPlanView v;
  v.perm = (const int32_t*)b;
  v.tile_mask = (const uint32_t*)(b + mm3d_plan_off_mask(n_cap));
  v.order = (const int32_t*)(b + mm3d_plan_off_order(n_cap));
  v.off_tiles = (const uint32_t*)(b + mm3d_plan_off_cnt(n_cap));
  v.tbl = (const int32_t*)(b + mm3d_plan_off_tbl(n_cap));
  v.stride = mm3d_plan_tiles(n_cap) * 128;
  return v;
}

// The i-th tile of CTA `cta` out of `ctas`: position i*ctas + (cta, or mirrored on odd rounds) of the
// cost-ordered tile list; -1 past the end.  Mirroring makes the per-CTA cost sums telescope to ~equal.
__device__ __forceinline__ int mm3d_plan_local_tile(const int32_t* __restrict__ order, int num_tiles, int ctas, int cta, int i) {
  const int64_t pos = (int64_t)i * ctas + ((i & 1) ? ctas - 1 - cta : cta);
  return pos < num_tiles ? __ldg(order + pos) : -1;
}
