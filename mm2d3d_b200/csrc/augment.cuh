// Device arithmetic of "points (metres) -> voxel coordinates", shared by mm3d_scale_points (augment.cu: the
// coordinates are written out as the reference's int64 [N, 4]) and mm3d_voxelize_points (hash_meta.cu: the
// coordinates go straight into the voxel hash).  Both follow the reference's augment_and_scale_3d
// (lib/utils/augmentation_3d.py:83-158) + cast + filter (lib/dataset/nuscenes_dataloader.py:323-327) through the
// SAME functions, so the two entry points cannot drift apart.  Every step is an explicitly rounded single operation
// (no contraction): numpy evaluates them one by one.
#pragma once

#include "common.cuh"

// order-preserving map float -> uint32 so that atomicMin / atomicMax work on floats
__device__ __forceinline__ uint32_t mm3d_f2o(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float mm3d_o2f(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// c = (p . R) * scale in float32; the dot product as fma(p2, R2j, fma(p1, R1j, p0 * R0j))
__device__ __forceinline__ void mm3d_rotate_scale(const float* __restrict__ p, const float* __restrict__ R, float scale,
                                                  float (&c)[3]) {
  const float p0 = __ldg(p), p1 = __ldg(p + 1), p2 = __ldg(p + 2);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float r = __fmaf_rn(p2, R[6 + j], __fmaf_rn(p1, R[3 + j], __fmul_rn(p0, R[j])));
    c[j] = __fmul_rn(r, scale);
  }
}

// last sample whose first point is <= i
__device__ __forceinline__ int mm3d_sample_of(const int64_t* __restrict__ offs, int B, int64_t i) {
  int lo = 0, hi = B - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(offs + mid) <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// Voxel coordinates of point i of sample b once the sample's min / max (mm = [min x y z | max x y z] of sample b in
// ordered-uint form) are known.  Returns "inside [0, full_scale)^3"; the sample's first point records the sample's
// min_value / offset (what the reference returns beside the coordinates).
__device__ __forceinline__ bool mm3d_point_voxel(const float* __restrict__ pts, const int64_t* __restrict__ offs, int b,
                                                 int64_t i, const float* __restrict__ rot, float scale, int full_scale,
                                                 const double* __restrict__ u, const uint32_t* __restrict__ mm,
                                                 float* __restrict__ min_value, double* __restrict__ offset,
                                                 long long (&out)[3]) {
  float R[9], c[3];
#pragma unroll
  for (int q = 0; q < 9; ++q) R[q] = __ldg(rot + 9 * b + q);
  mm3d_rotate_scale(pts + 3 * i, R, scale, c);
  const bool first = i == __ldg(offs + b);
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float mn = mm3d_o2f(mm[6 * b + j]), mx = mm3d_o2f(mm[6 * b + 3 + j]);
    float v = __fsub_rn(c[j], mn);
    double off = 0.0;
    if (u) {
      // numpy: full_scale - coords.max(0) - 0.001 stays float32 (python scalars are weak), the product with
      // the float64 draws and the in-place += are float64, stored back as float32
      float room = __fsub_rn(__fsub_rn((float)full_scale, __fsub_rn(mx, mn)), 0.001f);
      room = room > 0.f ? room : 0.f;
      off = __dmul_rn((double)room, __ldg(u + 3 * b + j));
      v = (float)__dadd_rn((double)v, off);
    }
    if (first) {
      min_value[3 * b + j] = mn;
      offset[3 * b + j] = off;
    }
    const long long q = (long long)v;  // astype(int64): truncation
    out[j] = q;
    ok = ok && q >= 0 && q < full_scale;
  }
  return ok;
}

// per-sample min / max of the rotated, scaled points into mm[6 * B] (two launches on `stream`); augment.cu
int mm3d_launch_minmax(const float* points, const int64_t* sample_offsets, int B, int64_t n, const float* rot, float scale,
                       uint32_t* mm, cudaStream_t stream);
