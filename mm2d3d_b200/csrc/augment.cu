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
#include "augment.cuh"

namespace {

__global__ void k_init_minmax(uint32_t* __restrict__ mm, int B) {
  mm3d_griddep_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 6 * B; i += gridDim.x * blockDim.x)
    mm[i] = (i % 6) < 3 ? 0xFFFFFFFFu : 0u;  // [min x y z | max x y z] in ordered-uint form
}

__global__ void k_minmax(const float* __restrict__ pts, const int64_t* __restrict__ offs, int B, int64_t n,
                         const float* __restrict__ rot, float scale, uint32_t* __restrict__ mm) {
  mm3d_griddep_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = mm3d_sample_of(offs, B, i);
    float R[9], c[3];
#pragma unroll
    for (int q = 0; q < 9; ++q) R[q] = __ldg(rot + 9 * b + q);
    mm3d_rotate_scale(pts + 3 * i, R, scale, c);
    // one atomic per warp and component where the whole warp is in one sample (the common case)
    const int b0 = __shfl_sync(__activemask(), b, __ffs(__activemask()) - 1);
    const bool uniform = __all_sync(__activemask(), b == b0);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      uint32_t lo = mm3d_f2o(c[j]), hi = lo;
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
    const int b = mm3d_sample_of(offs, B, i);
    long long q[3];
    const bool ok = mm3d_point_voxel(pts, offs, b, i, rot, scale, full_scale, u, mm, min_value, offset, q);
    reinterpret_cast<longlong2*>(coords)[2 * i] = make_longlong2(q[0], q[1]);
    reinterpret_cast<longlong2*>(coords)[2 * i + 1] = make_longlong2(q[2], (long long)b);
    keep[i] = ok ? 1 : 0;
  }
}

}  // namespace

int mm3d_launch_minmax(const float* points, const int64_t* sample_offsets, int B, int64_t n, const float* rot, float scale,
                       uint32_t* mm, cudaStream_t stream) {
  MM3D_CUDA(mm3d_launch_pdl(k_init_minmax, dim3(1), dim3(256), 0, stream, mm, B));
  MM3D_CUDA(mm3d_launch_pdl(k_minmax, dim3(mm3d_grid(n, 256)), dim3(256), 0, stream, points, sample_offsets, B, n, rot,
                            scale, mm));
  mm3d_count_launches(2);
  return MM3D_OK;
}

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
  int rc = mm3d_launch_minmax(points, sample_offsets, B, n, rot, scale, mm, stream);
  if (rc) return rc;
  MM3D_CUDA(mm3d_launch_pdl(k_coords, dim3(mm3d_grid(n, 256)), dim3(256), 0, stream, points, sample_offsets, B, n, rot,
                            scale, full_scale, transl_u, (const uint32_t*)mm, coords, keep, min_value, offset));
  mm3d_count_launches(1);
  MM3D_CHECK_LAUNCH("mm3d_scale_points");
  return MM3D_OK;
}
