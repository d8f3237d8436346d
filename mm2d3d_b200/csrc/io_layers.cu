// InputLayer / OutputLayer feature kernels and the 2D->3D lift: HBM-bound gathers/scatters.
//
// InputLayer (mode 4): voxel row = mean of its points' features  (SURVEY A.2);
// OutputLayer: point row = copy of its voxel's row, backward = sum over the voxel's points (A.7);
// lift: out[n,:] = fmap[b(n), :, row(n), col(n)] on an NCHW map (2d_net/model.py:131-137).
// One thread per (row, channel-vector); consecutive threads walk a row, so the row side of
// every gather/scatter is fully coalesced and the indexed side moves whole 16/32-byte pieces.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

int* mm3d_device_err_flag();  // conv_tc.cu

namespace {

// ------------------------------------------------------------------ InputLayer
__global__ void k_input_fwd(const float* __restrict__ feats, const int32_t* __restrict__ p2v,
                            const int32_t* __restrict__ npts, int64_t n_points, int c, int mode,
                            float* __restrict__ out) {
  mm3d_griddep_wait();
  const int64_t total = n_points * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / c;
    const int ch = (int)(i - p * c);
    const int32_t v = __ldg(p2v + p);
    const int32_t cnt = __ldg(npts + v);
    const float f = __ldg(feats + i);
    if (cnt == 1) {
      out[(int64_t)v * c + ch] = f;  // sole contributor: plain store, no RMW
    } else {
      atomicAdd(out + (int64_t)v * c + ch, mode == 4 ? f / (float)cnt : f);
    }
  }
}

__global__ void k_input_bwd(const float* __restrict__ d_vox, const int32_t* __restrict__ p2v,
                            const int32_t* __restrict__ npts, int64_t n_points, int c, int mode,
                            float* __restrict__ d_feats) {
  mm3d_griddep_wait();
  const int64_t total = n_points * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / c;
    const int ch = (int)(i - p * c);
    const int32_t v = __ldg(p2v + p);
    float g = __ldg(d_vox + (int64_t)v * c + ch);
    if (mode == 4) g /= (float)__ldg(npts + v);
    d_feats[i] = g;
  }
}

// ------------------------------------------------------------------ OutputLayer
template <int VEC>
__global__ void k_output_fwd(const float* __restrict__ vox, const int32_t* __restrict__ p2v,
                             int64_t n_points, int cv, float* __restrict__ out) {
  mm3d_griddep_wait();
  const int64_t total = n_points * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / cv;
    const int ch = (int)(i - p * cv);
    const int32_t v = __ldg(p2v + p);
    if (VEC == 4) {
      reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(vox) + (int64_t)v * cv + ch);
    } else {
      out[i] = __ldg(vox + (int64_t)v * cv + ch);
    }
  }
}

__global__ void k_output_bwd(const float* __restrict__ d_out, const int32_t* __restrict__ p2v,
                             int64_t n_points, int c, float* __restrict__ d_vox) {
  mm3d_griddep_wait();
  const int64_t total = n_points * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / c;
    const int ch = (int)(i - p * c);
    atomicAdd(d_vox + (int64_t)__ldg(p2v + p) * c + ch, __ldg(d_out + i));
  }
}

// ------------------------------------------------------------------ lift
__device__ __forceinline__ int find_sample(const int64_t* __restrict__ offs, int B, int64_t n) {
  int lo = 0, hi = B;  // offs[lo] <= n < offs[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(offs + mid) <= n) lo = mid; else hi = mid;
  }
  return lo;
}

// one thread per (point, channel), channel fastest: the [N, C] side is coalesced; the map side
// reads C planes at one pixel (stride H*W) -- inherent to gathering from NCHW.
template <class T>
__global__ void k_lift_fwd(const T* __restrict__ fmap, int B, int C, int H, int W,
                           const int64_t* __restrict__ idx, const int64_t* __restrict__ offs, int64_t n,
                           T* __restrict__ out, int* err) {
  const int64_t total = n * C;
  const int64_t hw = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / C;
    const int ch = (int)(i - p * C);
    const int b = find_sample(offs, B, p);
    int64_t r = __ldg(idx + 2 * p), col = __ldg(idx + 2 * p + 1);
    if (r < 0) r += H;      // torch advanced indexing wraps negative indices
    if (col < 0) col += W;
    if ((uint64_t)r >= (uint64_t)H || (uint64_t)col >= (uint64_t)W) {
      // torch raises IndexError here (2d_net/model.py:131-137); a kernel cannot: zero row + sticky error word 1
      out[i] = T(0.f);
      if (ch == 0) mm3d_raise(err, 1);
      continue;
    }
    out[i] = fmap[((int64_t)b * C + ch) * hw + r * W + col];
  }
}

__device__ __forceinline__ void lift_atomic_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void lift_atomic_add(__half* p, __half v) { atomicAdd(p, v); }
__device__ __forceinline__ void lift_atomic_add(__nv_bfloat16* p, __nv_bfloat16 v) { atomicAdd(p, v); }

// backward: scatter-add into the (zero-filled) map gradient; several points may share a pixel
template <class T>
__global__ void k_lift_bwd(const T* __restrict__ d_out, int B, int C, int H, int W,
                           const int64_t* __restrict__ idx, const int64_t* __restrict__ offs, int64_t n,
                           T* __restrict__ d_fmap, int* err) {
  const int64_t total = n * C;
  const int64_t hw = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / C;
    const int ch = (int)(i - p * C);
    const int b = find_sample(offs, B, p);
    int64_t r = __ldg(idx + 2 * p), col = __ldg(idx + 2 * p + 1);
    if (r < 0) r += H;
    if (col < 0) col += W;
    if ((uint64_t)r >= (uint64_t)H || (uint64_t)col >= (uint64_t)W) {  // skipped + sticky error word 1
      if (ch == 0) mm3d_raise(err, 1);
      continue;
    }
    lift_atomic_add(d_fmap + ((int64_t)b * C + ch) * hw + r * W + col, d_out[i]);
  }
}

}  // namespace

extern "C" int mm3d_input_fwd(const float* feats, const int32_t* p2v, const int32_t* npts, int64_t n_points,
                              int64_t n_vox, int c, int mode, float* out_vox, mm3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MM3D_REQUIRE(mode == 3 || mode == 4, MM3D_ERR_UNSUPPORTED, "InputLayer mode %d not implemented (3=sum, 4=mean)", mode);
  MM3D_REQUIRE(c > 0 && n_points >= 0 && n_vox >= 0, MM3D_ERR_INVALID, "bad sizes");
  if (n_vox > 0) MM3D_CUDA(cudaMemsetAsync(out_vox, 0, sizeof(float) * (size_t)n_vox * c, stream));
  if (n_points > 0)
    MM3D_CUDA(mm3d_launch_pdl(k_input_fwd, dim3(mm3d_grid(n_points * c, 256)), dim3(256), 0, stream, feats, p2v, npts, n_points, c, mode, out_vox));
  mm3d_count_launches(n_points > 0 ? 1 : 0);
  MM3D_CHECK_LAUNCH("mm3d_input_fwd");
  return MM3D_OK;
}

extern "C" int mm3d_input_bwd(const float* d_vox, const int32_t* p2v, const int32_t* npts, int64_t n_points,
                              int c, int mode, float* d_feats, mm3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MM3D_REQUIRE(mode == 3 || mode == 4, MM3D_ERR_UNSUPPORTED, "InputLayer mode %d not implemented", mode);
  if (n_points > 0)
    MM3D_CUDA(mm3d_launch_pdl(k_input_bwd, dim3(mm3d_grid(n_points * c, 256)), dim3(256), 0, stream, d_vox, p2v, npts, n_points, c, mode, d_feats));
  mm3d_count_launches(n_points > 0 ? 1 : 0);
  MM3D_CHECK_LAUNCH("mm3d_input_bwd");
  return MM3D_OK;
}

extern "C" int mm3d_output_fwd(const float* vox, const int32_t* p2v, int64_t n_points, int c, float* out,
                               mm3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_points > 0) {
    const bool vec = (c % 4 == 0) && (((uintptr_t)vox | (uintptr_t)out) & 15) == 0;
    if (vec)
      MM3D_CUDA(mm3d_launch_pdl(k_output_fwd<4>, dim3(mm3d_grid(n_points * (c / 4), 256)), dim3(256), 0, stream, vox, p2v, n_points, c / 4, out));
    else
      MM3D_CUDA(mm3d_launch_pdl(k_output_fwd<1>, dim3(mm3d_grid(n_points * c, 256)), dim3(256), 0, stream, vox, p2v, n_points, c, out));
  }
  mm3d_count_launches(n_points > 0 ? 1 : 0);
  MM3D_CHECK_LAUNCH("mm3d_output_fwd");
  return MM3D_OK;
}

extern "C" int mm3d_output_bwd(const float* d_out, const int32_t* p2v, int64_t n_points, int64_t n_vox, int c,
                               float* d_vox, mm3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_vox > 0) MM3D_CUDA(cudaMemsetAsync(d_vox, 0, sizeof(float) * (size_t)n_vox * c, stream));
  if (n_points > 0)
    MM3D_CUDA(mm3d_launch_pdl(k_output_bwd, dim3(mm3d_grid(n_points * c, 256)), dim3(256), 0, stream, d_out, p2v, n_points, c, d_vox));
  mm3d_count_launches(n_points > 0 ? 1 : 0);
  MM3D_CHECK_LAUNCH("mm3d_output_bwd");
  return MM3D_OK;
}

extern "C" int mm3d_lift2d_fwd(const void* fmap, int dtype, int B, int C, int H, int W, const int64_t* idx,
                               const int64_t* sample_offsets, int64_t n, void* out, mm3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MM3D_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && n >= 0, MM3D_ERR_INVALID, "bad lift sizes");
  if (n == 0) return MM3D_OK;
  int* err = mm3d_device_err_flag();
  const int grid = mm3d_grid(n * C, 256);
  switch (dtype) {
    case 0: k_lift_fwd<float><<<grid, 256, 0, stream>>>((const float*)fmap, B, C, H, W, idx, sample_offsets, n, (float*)out, err); break;
    case 1: k_lift_fwd<__half><<<grid, 256, 0, stream>>>((const __half*)fmap, B, C, H, W, idx, sample_offsets, n, (__half*)out, err); break;
    case 2: k_lift_fwd<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)fmap, B, C, H, W, idx, sample_offsets, n, (__nv_bfloat16*)out, err); break;
    default: MM3D_REQUIRE(false, MM3D_ERR_UNSUPPORTED, "lift dtype %d (0=f32,1=f16,2=bf16)", dtype);
  }
  mm3d_count_launches(1);
  MM3D_CHECK_LAUNCH("mm3d_lift2d_fwd");
  return MM3D_OK;
}

extern "C" int mm3d_lift2d_bwd(const void* d_out, int dtype, int B, int C, int H, int W, const int64_t* idx,
                               const int64_t* sample_offsets, int64_t n, void* d_fmap, mm3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MM3D_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && n >= 0, MM3D_ERR_INVALID, "bad lift sizes");
  if (n == 0) return MM3D_OK;
  int* err = mm3d_device_err_flag();
  const int grid = mm3d_grid(n * C, 256);
  switch (dtype) {
    case 0: k_lift_bwd<float><<<grid, 256, 0, stream>>>((const float*)d_out, B, C, H, W, idx, sample_offsets, n, (float*)d_fmap, err); break;
    case 1: k_lift_bwd<__half><<<grid, 256, 0, stream>>>((const __half*)d_out, B, C, H, W, idx, sample_offsets, n, (__half*)d_fmap, err); break;
    case 2: k_lift_bwd<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)d_out, B, C, H, W, idx, sample_offsets, n, (__nv_bfloat16*)d_fmap, err); break;
    default: MM3D_REQUIRE(false, MM3D_ERR_UNSUPPORTED, "lift dtype %d (0=f32,1=f16,2=bf16)", dtype);
  }
  mm3d_count_launches(1);
  MM3D_CHECK_LAUNCH("mm3d_lift2d_bwd");
  return MM3D_OK;
}

// ------------------------------------------------------------------------------------------------
// Point values -> image: the loaders' sparse depth / 2D label maps
//     depth = zeros(H, W); depth[idx[:,0], idx[:,1]] = z            (lib/dataset/nuscenes_dataloader.py:275-278)
// numpy assigns repeated indices in order, so the LAST point of a pixel wins: pass 1 records the largest
// point index per pixel (atomicMax), pass 2 writes that point's value -- deterministic, same result.
// ------------------------------------------------------------------------------------------------
namespace {

__global__ void k_raster_fill(float* __restrict__ out, int32_t* __restrict__ winner, int64_t total, float fill) {
  mm3d_griddep_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    out[i] = fill;
    winner[i] = -1;
  }
}

__device__ __forceinline__ int64_t raster_pixel(const int64_t* __restrict__ idx, const int64_t* __restrict__ offs, int B,
                                                int H, int W, int64_t i) {
  int lo = 0, hi = B - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(offs + mid) <= i) lo = mid; else hi = mid - 1;
  }
  const int64_t r = __ldg(idx + 2 * i), c = __ldg(idx + 2 * i + 1);
  if ((uint64_t)r >= (uint64_t)H || (uint64_t)c >= (uint64_t)W) return -1;
  return ((int64_t)lo * H + r) * W + c;
}

__global__ void k_raster_winner(const int64_t* __restrict__ idx, const int64_t* __restrict__ offs, int B, int H, int W,
                                int64_t n, int32_t* __restrict__ winner) {
  mm3d_griddep_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t px = raster_pixel(idx, offs, B, H, W, i);
    if (px >= 0) atomicMax(winner + px, (int32_t)i);
  }
}

__global__ void k_raster_write(const int64_t* __restrict__ idx, const int64_t* __restrict__ offs, int B, int H, int W,
                               int64_t n, const float* __restrict__ vals, const int32_t* __restrict__ winner,
                               float* __restrict__ out) {
  mm3d_griddep_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t px = raster_pixel(idx, offs, B, H, W, i);
    if (px >= 0 && winner[px] == (int32_t)i) out[px] = __ldg(vals + i);
  }
}

}  // namespace

extern "C" size_t mm3d_raster2d_workspace_bytes(int B, int H, int W) { return mm3d_align(sizeof(int32_t) * (size_t)B * H * W); }

extern "C" int mm3d_raster2d(const int64_t* idx, const int64_t* sample_offsets, int B, int H, int W, int64_t n,
                             const float* vals, float fill, float* out, void* ws, size_t ws_bytes, mm3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MM3D_REQUIRE(B > 0 && H > 0 && W > 0 && n >= 0 && n < (1ll << 31), MM3D_ERR_INVALID, "raster2d: bad sizes");
  MM3D_REQUIRE(out && ws && ws_bytes >= mm3d_raster2d_workspace_bytes(B, H, W), MM3D_ERR_WORKSPACE, "raster2d: workspace too small");
  const int64_t total = (int64_t)B * H * W;
  int32_t* winner = (int32_t*)ws;
  MM3D_CUDA(mm3d_launch_pdl(k_raster_fill, dim3(mm3d_grid(total, 256)), dim3(256), 0, stream, out, winner, total, fill));
  if (n > 0) {
    MM3D_REQUIRE(idx && sample_offsets && vals, MM3D_ERR_INVALID, "raster2d: null pointer");
    MM3D_CUDA(mm3d_launch_pdl(k_raster_winner, dim3(mm3d_grid(n, 256)), dim3(256), 0, stream, idx, sample_offsets, B, H, W, n, winner));
    MM3D_CUDA(mm3d_launch_pdl(k_raster_write, dim3(mm3d_grid(n, 256)), dim3(256), 0, stream, idx, sample_offsets, B, H, W, n, vals,
                              (const int32_t*)winner, out));
  }
  mm3d_count_launches(n > 0 ? 3 : 1);
  MM3D_CHECK_LAUNCH("mm3d_raster2d");
  return MM3D_OK;
}
