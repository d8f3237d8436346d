// Points (metres) -> voxel coordinates on the GPU: the reference's augment_and_scale_3d
// (lib/utils/augmentation_3d.py:83-158) followed by its integer cast and receptive-field filter
// (lib/dataset/nuscenes_dataloader.py:323-327), per sample of a collated batch:
//
//     r      = p . R                      (row vector times the 3x3 augmentation matrix, float32)
//     c      = r * scale                  (float32)
//     c     -= min over the sample        (float32)
//     offset = clip(full_scale - max(c) - 0.001, 0) * u      u = the sample's three uniform draws (float64)
//     c      = float32(double(c) + offset)
//     coords = int64(c)  (x, y, z, sample);   keep = all(0 <= coords < full_scale)
//
// The random draws (matrix R, u) are made on the host exactly as the reference makes them; the data-dependent
// parts (min / max reductions, scaling, cast, filter) run here.  Arithmetic follows numpy's float32 / float64
// promotion step by step; the only step whose rounding is not pinned by numpy is the dot product (BLAS): it is
// evaluated as fma(p2, R2j, fma(p1, R1j, p0 * R0j)).  HBM-bound: 12 B read + 33 B written per point, two passes.
#include "common.cuh"

namespace {

// order-preserving map float -> uint32 so that atomicMin / atomicMax work on floats
__device__ __forceinline__ uint32_t f2o(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float o2f(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__device__ __forceinline__ void rotate_scale(const float* __restrict__ p, const float* __restrict__ R, float scale,
                                             float (&c)[3]) {
  const float p0 = __ldg(p), p1 = __ldg(p + 1), p2 = __ldg(p + 2);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float r = __fmaf_rn(p2, R[6 + j], __fmaf_rn(p1, R[3 + j], __fmul_rn(p0, R[j])));
    c[j] = __fmul_rn(r, scale);
  }
}

__device__ __forceinline__ int sample_of(const int64_t* __restrict__ offs, int B, int64_t i) {
  int lo = 0, hi = B - 1;  // last sample whose first point is <= i
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(offs + mid) <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void k_init_minmax(uint32_t* __restrict__ mm, int B) {
  mm3d_griddep_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 6 * B; i += gridDim.x * blockDim.x)
    mm[i] = (i % 6) < 3 ? 0xFFFFFFFFu : 0u;  // [min x y z | max x y z] in ordered-uint form
}

__global__ void k_minmax(const float* __restrict__ pts, const int64_t* __restrict__ offs, int B, int64_t n,
                         const float* __restrict__ rot, float scale, uint32_t* __restrict__ mm) {
  mm3d_griddep_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = sample_of(offs, B, i);
    float R[9], c[3];
#pragma unroll
    for (int q = 0; q < 9; ++q) R[q] = __ldg(rot + 9 * b + q);
    rotate_scale(pts + 3 * i, R, scale, c);
    // one atomic per warp and component where the whole warp is in one sample (the common case)
    const int b0 = __shfl_sync(__activemask(), b, __ffs(__activemask()) - 1);
    const bool uniform = __all_sync(__activemask(), b == b0);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      uint32_t lo = f2o(c[j]), hi = lo;
      if (uniform) {
        lo = __reduce_min_sync(__activemask(), lo);
        hi = __reduce_max_sync(__activemask(), hi);
        if ((threadIdx.x & 31) == __ffs(__activemask()) - 1) {
          atomicMin(mm + 6 * b + j, lo);
          atomicMax(mm + 6 * b + 3 + j, hi);
        }
      } else {
        atomicMin(mm + 6 * b + j, lo);
        atomicMax(mm + 6 * b + 3 + j, hi);
      }
    }
  }
}

__global__ void k_coords(const float* __restrict__ pts, const int64_t* __restrict__ offs, int B, int64_t n,
                         const float* __restrict__ rot, float scale, int full_scale, const double* __restrict__ u,
                         const uint32_t* __restrict__ mm, int64_t* __restrict__ coords, uint8_t* __restrict__ keep,
                         float* __restrict__ min_value, double* __restrict__ offset) {
  mm3d_griddep_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = sample_of(offs, B, i);
    float R[9], c[3];
#pragma unroll
    for (int q = 0; q < 9; ++q) R[q] = __ldg(rot + 9 * b + q);
    rotate_scale(pts + 3 * i, R, scale, c);
    long long out[4];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float mn = o2f(mm[6 * b + j]), mx = o2f(mm[6 * b + 3 + j]);
      float v = __fsub_rn(c[j], mn);
      double off = 0.0;
      if (u) {
        // numpy: full_scale - coords.max(0) - 0.001 stays float32 (python scalars are weak), the product with
        // the float64 draws and the in-place += are float64, stored back as float32
        float room = __fsub_rn(__fsub_rn((float)full_scale, __fsub_rn(mx, mn)), 0.001f);
        room = room > 0.f ? room : 0.f;
        off = (double)room * __ldg(u + 3 * b + j);
        v = (float)((double)v + off);
      }
      if (i == __ldg(offs + b)) {  // the sample's first point records the sample's min / offset
        min_value[3 * b + j] = mn;
        offset[3 * b + j] = off;
      }
      const long long q = (long long)v;  // astype(int64): truncation
      out[j] = q;
      ok = ok && q >= 0 && q < full_scale;
    }
    out[3] = b;
    reinterpret_cast<longlong2*>(coords)[2 * i] = make_longlong2(out[0], out[1]);
    reinterpret_cast<longlong2*>(coords)[2 * i + 1] = make_longlong2(out[2], out[3]);
    keep[i] = ok ? 1 : 0;
  }
}

}  // namespace

extern "C" size_t mm3d_scale_points_workspace_bytes(int B) { return mm3d_align(sizeof(uint32_t) * 6 * (size_t)(B > 0 ? B : 1)); }

extern "C" int mm3d_scale_points(const float* points, const int64_t* sample_offsets, int B, int64_t n, const float* rot,
                                 float scale, int full_scale, const double* transl_u, int64_t* coords, uint8_t* keep,
                                 float* min_value, double* offset, void* ws, size_t ws_bytes, mm3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MM3D_REQUIRE(B > 0 && n >= 0 && full_scale > 0, MM3D_ERR_INVALID, "scale_points: bad sizes");
  MM3D_REQUIRE(ws && ws_bytes >= mm3d_scale_points_workspace_bytes(B), MM3D_ERR_WORKSPACE, "scale_points: workspace too small");
  MM3D_REQUIRE((((uintptr_t)coords) & 15) == 0, MM3D_ERR_INVALID, "scale_points: coords must be 16-byte aligned");
  if (n == 0) return MM3D_OK;
  MM3D_REQUIRE(points && sample_offsets && rot && coords && keep && min_value && offset, MM3D_ERR_INVALID,
               "scale_points: null pointer");
  uint32_t* mm = (uint32_t*)ws;
  MM3D_CUDA(mm3d_launch_pdl(k_init_minmax, dim3(1), dim3(256), 0, stream, mm, B));
  MM3D_CUDA(mm3d_launch_pdl(k_minmax, dim3(mm3d_grid(n, 256)), dim3(256), 0, stream, points, sample_offsets, B, n, rot,
                            scale, mm));
  MM3D_CUDA(mm3d_launch_pdl(k_coords, dim3(mm3d_grid(n, 256)), dim3(256), 0, stream, points, sample_offsets, B, n, rot,
                            scale, full_scale, transl_u, (const uint32_t*)mm, coords, keep, min_value, offset));
  mm3d_count_launches(3);
  MM3D_CHECK_LAUNCH("mm3d_scale_points");
  return MM3D_OK;
}
