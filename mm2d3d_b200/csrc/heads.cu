// Point-wise prologue and loss around the 3D network:
//
//  * RGB mask (3d_net/model.py:46-48):   s = sigmoid(x . w + b);  x *= s        x [n, c] point features, c small
//  * cross-modal loss (train.py:157-184): mean over points of  sum_c q_c (log q_c - log p_c),
//        p = softmax(pred), q = softmax(target.detach())      pred, target [n, C] logits
//
// Both are a handful of small element-wise / row-reduction launches in the reference; here each direction is one
// launch, one thread per point (rows are 12-80 bytes).  HBM-bound: every tensor is read or written once.
#include "common.cuh"

namespace {

constexpr int kMaxC = 64;

__global__ void k_rgb_mask_fwd(const float* x, int64_t n, int c, const float* __restrict__ w,
                               const float* __restrict__ b, float* y, float* __restrict__ s_out) {
  mm3d_griddep_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float xr[8];
    float acc = __ldg(b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xr[j] = j < c ? x[i * c + j] : 0.f;
      if (j < c) acc = fmaf(xr[j], __ldg(w + j), acc);
    }
    const float s = 1.f / (1.f + expf(-acc));
    s_out[i] = s;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < c) y[i * c + j] = xr[j] * s;
  }
}

// dx = dy s + t w,  t = (dy . x) s (1 - s);   dw = sum_n t x,  db = sum_n t   (FP64 totals, one atomic per CTA)
__global__ void k_rgb_mask_bwd(const float* __restrict__ x, const float* __restrict__ s_in, const float* __restrict__ dy,
                               int64_t n, int c, const float* __restrict__ w, float* __restrict__ dx,
                               double* __restrict__ dwb) {
  mm3d_griddep_wait();
  __shared__ double red[kMaxC + 1];
  for (int j = threadIdx.x; j <= c; j += blockDim.x) red[j] = 0.0;
  __syncthreads();
  double loc[9];  // c <= 8 on this path (checked by the host)
#pragma unroll
  for (int j = 0; j < 9; ++j) loc[j] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = __ldg(s_in + i);
    float g = 0.f;
    for (int j = 0; j < c; ++j) g = fmaf(__ldg(dy + i * c + j), __ldg(x + i * c + j), g);
    const float t = g * s * (1.f - s);
    if (dx)
      for (int j = 0; j < c; ++j) dx[i * c + j] = fmaf(t, __ldg(w + j), __ldg(dy + i * c + j) * s);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < c) loc[j] += (double)(t * __ldg(x + i * c + j));
    loc[8] += (double)t;
  }
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    double v = loc[j];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && (j < c || j == 8)) atomicAdd(&red[j == 8 ? c : j], v);
  }
  __syncthreads();
  for (int j = threadIdx.x; j <= c; j += blockDim.x) atomicAdd(dwb + j, red[j]);
}

__global__ void k_f64_to_f32(const double* __restrict__ src, float* __restrict__ dst, int n, double scale) {
  mm3d_griddep_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = (float)(src[i] * scale);
}

// one row: log-softmax of pred and target, KL term, optionally the gradient (p - q) * gscale
__device__ __forceinline__ double kl_row(const float* __restrict__ pr, const float* __restrict__ tg, int C, float* dpred,
                                         float gscale) {
  float mp = -INFINITY, mt = -INFINITY;
  for (int j = 0; j < C; ++j) { mp = fmaxf(mp, __ldg(pr + j)); mt = fmaxf(mt, __ldg(tg + j)); }
  float sp = 0.f, st = 0.f;
  for (int j = 0; j < C; ++j) { sp += expf(__ldg(pr + j) - mp); st += expf(__ldg(tg + j) - mt); }
  const float lsp = logf(sp), lst = logf(st);
  double kl = 0.0;
  for (int j = 0; j < C; ++j) {
    const float logp = __ldg(pr + j) - mp - lsp, logq = __ldg(tg + j) - mt - lst;
    const float q = expf(logq);
    if (q > 0.f) kl += (double)(q * (logq - logp));
    if (dpred) dpred[j] = (expf(logp) - q) * gscale;
  }
  return kl;
}

__global__ void k_kl_fwd(const float* __restrict__ pred, const float* __restrict__ target, int64_t n, int C,
                         double* __restrict__ total) {
  mm3d_griddep_wait();
  double loc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    loc += kl_row(pred + i * C, target + i * C, C, nullptr, 0.f);
  for (int o = 16; o; o >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, o);
  __shared__ double red;
  if (threadIdx.x == 0) red = 0.0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) atomicAdd(&red, loc);
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(total, red);
}

__global__ void k_kl_bwd(const float* __restrict__ pred, const float* __restrict__ target, int64_t n, int C,
                         const float* __restrict__ dloss, float inv_n, float* __restrict__ dpred) {
  mm3d_griddep_wait();
  const float gs = __ldg(dloss) * inv_n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    kl_row(pred + i * C, target + i * C, C, dpred + i * C, gs);
}

}  // namespace

extern "C" {

// y = x * sigmoid(x . w + b) row-wise; s_out [n] keeps the gate for the backward pass.  y may be x (in place,
// like the reference's `feats *= mask`): a row is read completely before it is written.
int mm3d_rgb_mask_fwd(const float* x, int64_t n, int c, const float* w, const float* b, float* y, float* s_out,
                      mm3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MM3D_REQUIRE(n >= 0 && c > 0 && c <= 8, MM3D_ERR_INVALID, "rgb_mask: 1..8 feature channels supported, got %d", c);
  if (n == 0) return MM3D_OK;
  MM3D_REQUIRE(x && w && b && y && s_out, MM3D_ERR_INVALID, "rgb_mask: null pointer");
  MM3D_CUDA(mm3d_launch_pdl(k_rgb_mask_fwd, dim3(mm3d_grid(n, 256)), dim3(256), 0, stream, x, n, c, w, b, y, s_out));
  mm3d_count_launches(1);
  MM3D_CHECK_LAUNCH("mm3d_rgb_mask_fwd");
  return MM3D_OK;
}

// x: the UNMASKED features; dx may be NULL (the features are data); dw [c], db [1]; ws: (c + 1) doubles
int mm3d_rgb_mask_bwd(const float* x, const float* s, const float* dy, int64_t n, int c, const float* w, float* dx,
                      float* dw, float* db, void* ws, size_t ws_bytes, mm3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MM3D_REQUIRE(n >= 0 && c > 0 && c <= 8, MM3D_ERR_INVALID, "rgb_mask: 1..8 feature channels supported, got %d", c);
  MM3D_REQUIRE(dw && db && ws && ws_bytes >= sizeof(double) * (size_t)(c + 1), MM3D_ERR_WORKSPACE, "rgb_mask: workspace too small");
  double* tot = (double*)ws;
  MM3D_CUDA(cudaMemsetAsync(tot, 0, sizeof(double) * (size_t)(c + 1), stream));
  if (n > 0) {
    MM3D_REQUIRE(x && s && dy && w, MM3D_ERR_INVALID, "rgb_mask: null pointer");
    MM3D_CUDA(mm3d_launch_pdl(k_rgb_mask_bwd, dim3(mm3d_grid(n, 256, 2)), dim3(256), 0, stream, x, s, dy, n, c, w, dx, tot));
  }
  MM3D_CUDA(mm3d_launch_pdl(k_f64_to_f32, dim3(1), dim3(32), 0, stream, (const double*)tot, dw, c, 1.0));
  MM3D_CUDA(mm3d_launch_pdl(k_f64_to_f32, dim3(1), dim3(32), 0, stream, (const double*)(tot + c), db, 1, 1.0));
  mm3d_count_launches(n > 0 ? 3 : 2);
  MM3D_CHECK_LAUNCH("mm3d_rgb_mask_bwd");
  return MM3D_OK;
}

// loss[0] = mean_n sum_c q (log q - log p),  p = softmax(pred[n, :]), q = softmax(target[n, :]);  ws: one double
int mm3d_kl_logits_fwd(const float* pred, const float* target, int64_t n, int C, float* loss, void* ws, size_t ws_bytes,
                       mm3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MM3D_REQUIRE(n > 0 && C > 0 && C <= 4096, MM3D_ERR_INVALID, "kl_logits: bad sizes");
  MM3D_REQUIRE(pred && target && loss && ws && ws_bytes >= sizeof(double), MM3D_ERR_WORKSPACE, "kl_logits: workspace too small");
  MM3D_CUDA(cudaMemsetAsync(ws, 0, sizeof(double), stream));
  MM3D_CUDA(mm3d_launch_pdl(k_kl_fwd, dim3(mm3d_grid(n, 256, 2)), dim3(256), 0, stream, pred, target, n, C, (double*)ws));
  MM3D_CUDA(mm3d_launch_pdl(k_f64_to_f32, dim3(1), dim3(32), 0, stream, (const double*)ws, loss, 1, 1.0 / (double)n));
  mm3d_count_launches(2);
  MM3D_CHECK_LAUNCH("mm3d_kl_logits_fwd");
  return MM3D_OK;
}

// dpred = dloss[0] * (softmax(pred) - softmax(target)) / n   (the target side is detached in the reference)
int mm3d_kl_logits_bwd(const float* pred, const float* target, int64_t n, int C, const float* dloss, float* dpred,
                       mm3d_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MM3D_REQUIRE(n > 0 && C > 0 && C <= 4096, MM3D_ERR_INVALID, "kl_logits: bad sizes");
  MM3D_REQUIRE(pred && target && dloss && dpred, MM3D_ERR_INVALID, "kl_logits: null pointer");
  MM3D_CUDA(mm3d_launch_pdl(k_kl_bwd, dim3(mm3d_grid(n, 256)), dim3(256), 0, stream, pred, target, n, C, dloss,
                            1.f / (float)n, dpred));
  mm3d_count_launches(1);
  MM3D_CHECK_LAUNCH("mm3d_kl_logits_bwd");
  return MM3D_OK;
}

}  // extern "C"
