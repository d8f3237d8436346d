// BatchNorm + (leaky) ReLU over the active rows of a sparse tensor (replaces
// SparseConvNet's BatchNormalization_f_train / _f_test / _b, which launch <= 16 thread blocks).
//
// HBM-bound: forward = read x (stats) + read x again (L2) + write y; backward = read x, dy + again (L2) +
// write dx.  Training forward and backward are ONE launch each: per-CTA partial sums -> FP64 atomics ->
// grid-wide barrier (all CTAs co-resident) -> every CTA normalises the rows it summed.
// Layout: a CTA's 256 threads tile (rows_pass x CV) where CV = C / VEC channel vectors, so a
// warp reads consecutive float4s of consecutive rows (fully coalesced) and every thread keeps
// one fixed channel vector; per-thread FP32 partials are combined in FP64 (shared, then global
// atomics into 2*C doubles), which makes var = E[x^2] - mean^2 safe.
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
// Same-address atomics from ~300 CTAs serialise in L2 (~27 cycles each): the per-channel totals therefore go to one
// of kSlots copies (CTA b adds into copy b % kSlots, every CTA sums the copies after the barrier), the barrier has
// one counter per slot plus a top counter for the slots' last arrivers, and nobody cleans up behind a launch:
// the workspace holds TWO such regions and a launch zeroes the one it does not use -- the next launch on the stream
// uses that one (`parity` alternates per launch).
constexpr int kSlots = 8;

template <int VEC> struct Vec;
template <> struct Vec<4> { using T = float4; };
template <> struct Vec<1> { using T = float; };

template <int VEC> __device__ __forceinline__ void vload(const float* p, float (&v)[VEC]);
template <> __device__ __forceinline__ void vload<4>(const float* p, float (&v)[4]) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void vload<1>(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
template <int VEC> __device__ __forceinline__ void vstore(float* p, const float (&v)[VEC]);
template <> __device__ __forceinline__ void vstore<4>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void vstore<1>(float* p, const float (&v)[1]) { *p = v[0]; }

// Store for a tensor that a tensor-core convolution gathers: mode 0 = as is, 1 = rounded to TF32 (nearest), 2 = two
// planes `plane` floats apart: hi = tf32(v), lo = tf32(v - hi) (the error-compensated TF32x3 mode), 3 = the
// TF32-rounded FP32 plane (the weight-gradient operand) and, `plane` floats behind the tensor's base, a BF16 plane
// (the gathered operand of the BF16 mode; element i of the tensor at BF16 index i)
template <int VEC>
__device__ __forceinline__ void store_planes(float* base, float* p, float (&v)[VEC], int mode, int64_t plane) {
  if (mode == 0) { vstore<VEC>(p, v); return; }
  if (mode == 3) {
    __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(base + plane) + (p - base);
    if (VEC == 4) {
      const __nv_bfloat162 b01 = __floats2bfloat162_rn(v[0], v[1]), b23 = __floats2bfloat162_rn(v[2 % VEC], v[3 % VEC]);
      uint2 u;
      u.x = *reinterpret_cast<const unsigned int*>(&b01);
      u.y = *reinterpret_cast<const unsigned int*>(&b23);
      *reinterpret_cast<uint2*>(b) = u;
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) b[j] = __float2bfloat16_rn(v[j]);
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = mm3d_rna_tf32(v[j]);
    vstore<VEC>(p, v);
    return;
  }
  float lo[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const float hi = mm3d_rna_tf32(v[j]);
    lo[j] = mm3d_rna_tf32(v[j] - hi);
    v[j] = hi;
  }
  vstore<VEC>(p, v);
  if (mode == 2) vstore<VEC>(p + plane, lo);
}

// dynamic shared: 2*C doubles (accumulators) -- also reused as 2*C floats of scale/shift
extern __shared__ double s_acc[];

// Block reduction of per-thread FP32 partials (VEC channels each, thread layout rows_pass x cv):
// partials go to shared memory, then one thread per channel sums the rows_pass values in FP64 and
// issues ONE global FP64 atomic per channel and CTA.
template <int VEC>
__device__ __forceinline__ void block_flush(const float (&s)[VEC], const float (&q)[VEC], int c, int cv, int rows_pass,
                                            int r, int v, double* __restrict__ sums) {
  float* part = reinterpret_cast<float*>(s_acc);  // [2][rows_pass][c]
  if (r < rows_pass) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      part[r * c + v * VEC + j] = s[j];
      part[(rows_pass + r) * c + v * VEC + j] = q[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * c; i += kThreads) {
    const int which = i >= c, ch = i - which * c;
    double acc = 0.0;
    for (int rr = 0; rr < rows_pass; ++rr) acc += (double)part[(which * rows_pass + rr) * c + ch];
    atomicAdd(sums + i, acc);
  }
}

// ---- inference forward: normalise with the running statistics (+ optional copy of them to save_*)
template <int VEC>
__global__ void __launch_bounds__(kThreads)
k_bn_apply(const float* __restrict__ x, float* __restrict__ y, int64_t n, int c,
           const float* __restrict__ gamma, const float* __restrict__ beta,
           float* running_mean, float* running_var, float* save_mean, float* save_invstd,
           const double* __restrict__ sums, float eps, float momentum, float leak, int training, int round_tf32) {
  float* s_mean = reinterpret_cast<float*>(s_acc);
  float* s_scale = s_mean + c;
  for (int i = threadIdx.x; i < c; i += kThreads) {
    float mean, invstd;
    if (training) {
      const double m = sums[i] / (double)n;
      double var = sums[c + i] / (double)n - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      invstd = (float)(1.0 / sqrt(var + (double)eps));
      if (blockIdx.x == 0) {
        save_mean[i] = mean;
        save_invstd[i] = invstd;
        const double unbiased = var * ((double)n / (double)(n > 1 ? n - 1 : 1));
        running_mean[i] = momentum * running_mean[i] + (1.f - momentum) * mean;
        running_var[i] = momentum * running_var[i] + (1.f - momentum) * (float)unbiased;
      }
    } else {
      mean = running_mean[i];
      invstd = rsqrtf(running_var[i] + eps);
      if (blockIdx.x == 0 && save_mean) { save_mean[i] = mean; save_invstd[i] = invstd; }
    }
    s_mean[i] = mean;
    s_scale[i] = invstd * gamma[i];
  }
  __syncthreads();
  const int cv = c / VEC;
  const int rows_pass = kThreads / cv;
  const int r = threadIdx.x / cv, v = threadIdx.x - r * cv;
  if (r >= rows_pass) return;
  float mean[VEC], scale[VEC], bet[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    mean[j] = s_mean[v * VEC + j];
    scale[j] = s_scale[v * VEC + j];
    bet[j] = __ldg(beta + v * VEC + j);
  }
  for (int64_t row = (int64_t)blockIdx.x * rows_pass + r; row < n; row += (int64_t)gridDim.x * rows_pass) {
    float t[VEC];
    vload<VEC>(x + row * c + v * VEC, t);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float o = fmaf(t[j] - mean[j], scale[j], bet[j]);
      t[j] = o > 0.f ? o : o * leak;
    }
    store_planes<VEC>(y, y + row * c + v * VEC, t, round_tf32, n * c);
  }
}

// ---- grid-wide barrier for the single-launch kernels below.  All CTAs of these grids are co-resident (bn_grid caps
// the grid by the occupancy API x the device's SM count), so waiting for the others is safe; the wait is bounded anyway
// and raises the device error word instead of hanging.  sync[0..kSlots) count the arrivals of the CTAs of each slot,
// sync[kSlots] the slots that are complete: an arrival contends with ~grid / kSlots others, not with all of them.
__device__ __forceinline__ bool grid_arrive_and_wait(unsigned int* sync, int* err) {
  __shared__ int s_ok;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int slot = blockIdx.x % kSlots;
    const unsigned int expect = (gridDim.x + kSlots - 1 - slot) / kSlots;  // CTAs b < gridDim.x with b % kSlots == slot
    const unsigned int slots_used = gridDim.x < (unsigned)kSlots ? gridDim.x : (unsigned)kSlots;
    if (atomicAdd(sync + slot, 1u) == expect - 1) {
      __threadfence();
      atomicAdd(sync + kSlots, 1u);
    }
    int ok = 0;
    for (unsigned int i = 0; i < (1u << 22); ++i) {
      if (*reinterpret_cast<volatile unsigned int*>(sync + kSlots) >= slots_used) { ok = 1; break; }
      __nanosleep(32);
    }
    if (!ok) mm3d_raise(err);
    s_ok = ok;
    __threadfence();
  }
  __syncthreads();
  return s_ok != 0;
}
// zero the workspace region the NEXT launch will use (this launch does not touch it otherwise)
__device__ __forceinline__ void zero_other(double* other, int words) {
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < words; i += gridDim.x * kThreads) other[i] = 0.0;
}
// total of one accumulator over the slot copies
__device__ __forceinline__ double slot_total(const double* sums, int c2, int i) {
  double t = 0.0;
#pragma unroll
  for (int sl = 0; sl < kSlots; ++sl) t += __ldcg(sums + sl * c2 + i);
  return t;
}

// ---- training forward in ONE launch: statistics, grid barrier, normalise + ReLU.  A CTA re-reads exactly the
// rows it summed, so the second read comes from L2 / L1 instead of HBM.
template <int VEC>
__global__ void __launch_bounds__(kThreads)
k_bn_fwd_fused(const float* __restrict__ x, const float* __restrict__ x_hi, int c_lo, float* __restrict__ y, int64_t n, int c,
               const float* __restrict__ gamma, const float* __restrict__ beta,
               float* running_mean, float* running_var, float* save_mean, float* save_invstd,
               double* sums, unsigned int* sync, double* other, int other_words, float eps, float momentum, float leak,
               int round_tf32, int* err) {
  mm3d_griddep_launch();
  mm3d_griddep_wait();
  zero_other(other, other_words);
  const int cv = c / VEC;
  const int rows_pass = kThreads / cv;
  const int r = threadIdx.x / cv, v = threadIdx.x - r * cv;
  const int64_t stride = (int64_t)gridDim.x * rows_pass;
  // x may be given as two column blocks [x | x_hi] (JoinTable without materialising the concatenation): this
  // thread's channel vector of row i is xb[i * xld]
  const bool hi = c_lo > 0 && v * VEC >= c_lo;
  const float* xb = (x_hi && hi) ? x_hi + (v * VEC - c_lo) : x + v * VEC;
  const int64_t xld = x_hi ? (hi ? c - c_lo : c_lo) : c;
  {
    float s[VEC], q[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) s[j] = q[j] = 0.f;
    if (r < rows_pass) {
      int64_t row = (int64_t)blockIdx.x * rows_pass + r;
      for (; row + 3 * stride < n; row += 4 * stride) {  // four independent loads in flight
        float t[4][VEC];
#pragma unroll
        for (int u = 0; u < 4; ++u) vload<VEC>(xb + (row + u * stride) * xld, t[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int j = 0; j < VEC; ++j) { s[j] += t[u][j]; q[j] = fmaf(t[u][j], t[u][j], q[j]); }
      }
      for (; row < n; row += stride) {
        float t[VEC];
        vload<VEC>(xb + row * xld, t);
#pragma unroll
        for (int j = 0; j < VEC; ++j) { s[j] += t[j]; q[j] = fmaf(t[j], t[j], q[j]); }
      }
    }
    block_flush<VEC>(s, q, c, cv, rows_pass, r, v, sums + (blockIdx.x % kSlots) * 2 * c);
  }
  grid_arrive_and_wait(sync, err);
  float* s_mean = reinterpret_cast<float*>(s_acc);
  float* s_scale = s_mean + c;
  __syncthreads();  // block_flush's use of the shared buffer is over
  for (int i = threadIdx.x; i < c; i += kThreads) {
    const double m = slot_total(sums, 2 * c, i) / (double)n;
    double var = slot_total(sums, 2 * c, c + i) / (double)n - m * m;
    if (var < 0.0) var = 0.0;
    const float mean = (float)m, invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (blockIdx.x == 0) {
      save_mean[i] = mean;
      save_invstd[i] = invstd;
      const double unbiased = var * ((double)n / (double)(n > 1 ? n - 1 : 1));
      running_mean[i] = momentum * running_mean[i] + (1.f - momentum) * mean;
      running_var[i] = momentum * running_var[i] + (1.f - momentum) * (float)unbiased;
    }
    s_mean[i] = mean;
    s_scale[i] = invstd * gamma[i];
  }
  __syncthreads();
  if (r >= rows_pass) return;
  float mean[VEC], scale[VEC], bet[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    mean[j] = s_mean[v * VEC + j];
    scale[j] = s_scale[v * VEC + j];
    bet[j] = __ldg(beta + v * VEC + j);
  }
  int64_t row = (int64_t)blockIdx.x * rows_pass + r;
  for (; row + 3 * stride < n; row += 4 * stride) {  // four independent loads in flight
    float t[4][VEC];
#pragma unroll
    for (int u = 0; u < 4; ++u) vload<VEC>(xb + (row + u * stride) * xld, t[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float o = fmaf(t[u][j] - mean[j], scale[j], bet[j]);
        t[u][j] = o > 0.f ? o : o * leak;
      }
      store_planes<VEC>(y, y + (row + u * stride) * c + v * VEC, t[u], round_tf32, n * c);
    }
  }
  for (; row < n; row += stride) {
    float t[VEC];
    vload<VEC>(xb + row * xld, t);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float o = fmaf(t[j] - mean[j], scale[j], bet[j]);
      t[j] = o > 0.f ? o : o * leak;
    }
    store_planes<VEC>(y, y + row * c + v * VEC, t, round_tf32, n * c);
  }
}

// ---- backward in ONE launch: reductions, grid barrier, dx (+ d_gamma / d_beta by CTA 0)
template <int VEC>
__global__ void __launch_bounds__(kThreads)
k_bn_bwd_fused(const float* __restrict__ x, const float* __restrict__ x_hi, int c_lo, const float* __restrict__ dy,
               float* __restrict__ dx, float* __restrict__ dx_hi, int64_t n, int c,
               const float* __restrict__ gamma, const float* __restrict__ beta,
               const float* __restrict__ save_mean, const float* __restrict__ save_invstd, float leak,
               double* sums, unsigned int* sync, double* other, int other_words, float* d_gamma, float* d_beta, int training,
               int round_flags, const float* __restrict__ add, int64_t add_ld, int* err) {
  mm3d_griddep_launch();
  mm3d_griddep_wait();
  zero_other(other, other_words);
  const int cv = c / VEC;
  const int rows_pass = kThreads / cv;
  const int r = threadIdx.x / cv, v = threadIdx.x - r * cv;
  const int64_t stride = (int64_t)gridDim.x * rows_pass;
  // x may be given as two column blocks [x | x_hi] (JoinTable without materialising the concatenation): this
  // thread's channel vector of row i is xb[i * xld]
  const bool hi = c_lo > 0 && v * VEC >= c_lo;
  const float* xb = (x_hi && hi) ? x_hi + (v * VEC - c_lo) : x + v * VEC;
  const int64_t xld = x_hi ? (hi ? c - c_lo : c_lo) : c;
  float* dxb = (dx_hi && hi) ? dx_hi + (v * VEC - c_lo) : dx + v * VEC;
  const int64_t dxld = dx_hi ? (hi ? c - c_lo : c_lo) : c;
  // round_flags bit 0: dx (or its low column block) feeds a TF32 convolution as d_out -> store RNA-rounded values;
  // bit 1: same for the high column block dx_hi
  // bit 2: those outputs carry a hi and a lo plane (TF32x3 mode), the lo plane n * (row length) floats behind
  // bit 3: ... a TF32-rounded FP32 plane and a BF16 plane behind it (BF16 mode)
  const int rnd = (round_flags & ((dx_hi && hi) ? 2 : 1)) ? ((round_flags & 8) ? 3 : (round_flags & 4) ? 2 : 1) : 0;
  float* dx0 = (dx_hi && hi) ? dx_hi : dx;  // base of the tensor this thread writes
  const int64_t dplane = n * dxld;
  float mean[VEC], invstd[VEC], scale[VEC], bet[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) mean[j] = invstd[j] = scale[j] = bet[j] = 0.f;
  if (r < rows_pass) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int ch = v * VEC + j;
      mean[j] = __ldg(save_mean + ch);
      invstd[j] = __ldg(save_invstd + ch);
      scale[j] = invstd[j] * __ldg(gamma + ch);
      bet[j] = __ldg(beta + ch);
    }
  }
  {
    float s[VEC], q[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) s[j] = q[j] = 0.f;
    if (r < rows_pass) {
      int64_t row = (int64_t)blockIdx.x * rows_pass + r;
      for (; row < n; row += 2 * stride) {  // two rows (four loads) in flight
        float t[2][VEC], g[2][VEC];
        const bool second = row + stride < n;
        vload<VEC>(xb + row * xld, t[0]);
        vload<VEC>(dy + row * c + v * VEC, g[0]);
        if (second) {
          vload<VEC>(xb + (row + stride) * xld, t[1]);
          vload<VEC>(dy + (row + stride) * c + v * VEC, g[1]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u == 1 && !second) break;
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            const float xc = t[u][j] - mean[j];
            const float o = fmaf(xc, scale[j], bet[j]);
            const float d = o > 0.f ? g[u][j] : g[u][j] * leak;
            s[j] += d;
            q[j] = fmaf(d, xc * invstd[j], q[j]);
          }
        }
      }
    }
    block_flush<VEC>(s, q, c, cv, rows_pass, r, v, sums + (blockIdx.x % kSlots) * 2 * c);
  }
  grid_arrive_and_wait(sync, err);
  // totals over the slot copies -> shared memory (block_flush's use of the buffer is over)
  double* s_tot = s_acc;
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * c; i += kThreads) s_tot[i] = slot_total(sums, 2 * c, i);
  __syncthreads();
  if (blockIdx.x == 0)  // (frozen affine parameters: NULL gradient pointers)
    for (int i = threadIdx.x; i < c; i += kThreads) {
      if (d_beta) d_beta[i] = (float)s_tot[i];
      if (d_gamma) d_gamma[i] = (float)s_tot[c + i];
    }
  float md[VEC], mdx[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) md[j] = mdx[j] = 0.f;
  if (r < rows_pass && training) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      md[j] = (float)(s_tot[v * VEC + j] / (double)n);
      mdx[j] = (float)(s_tot[c + v * VEC + j] / (double)n);
    }
  }
  if (r >= rows_pass) return;
  int64_t row = (int64_t)blockIdx.x * rows_pass + r;
  for (; row + stride < n; row += 2 * stride) {  // two rows (four loads) in flight
    float t[2][VEC], g[2][VEC];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      vload<VEC>(xb + (row + u * stride) * xld, t[u]);
      vload<VEC>(dy + (row + u * stride) * c + v * VEC, g[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float xc = t[u][j] - mean[j];
        const float o = fmaf(xc, scale[j], bet[j]);
        const float d = o > 0.f ? g[u][j] : g[u][j] * leak;
        g[u][j] = scale[j] * (d - md[j] - xc * invstd[j] * mdx[j]);
      }
      if (add) {  // dx += another gradient of the same tensor (the U-Net's skip connection), before any rounding
        float a_[VEC];
        vload<VEC>(add + (row + u * stride) * add_ld + v * VEC, a_);
#pragma unroll
        for (int j = 0; j < VEC; ++j) g[u][j] += a_[j];
      }
      store_planes<VEC>(dx0, dxb + (row + u * stride) * dxld, g[u], rnd, dplane);
    }
  }
  for (; row < n; row += stride) {
    float t[VEC], g[VEC];
    vload<VEC>(xb + row * xld, t);
    vload<VEC>(dy + row * c + v * VEC, g);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float xc = t[j] - mean[j];
      const float o = fmaf(xc, scale[j], bet[j]);
      const float d = o > 0.f ? g[j] : g[j] * leak;
      g[j] = scale[j] * (d - md[j] - xc * invstd[j] * mdx[j]);
    }
    if (add) {
      float a_[VEC];
      vload<VEC>(add + row * add_ld + v * VEC, a_);
#pragma unroll
      for (int j = 0; j < VEC; ++j) g[j] += a_[j];
    }
    store_planes<VEC>(dx0, dxb + row * dxld, g, rnd, dplane);
  }
}

// Single-launch kernels wait on each other, so every CTA of the grid must be able to be resident at once: the grid
// is capped at (CTAs per SM the occupancy API reports for this kernel and shared-memory size, at most 2) x (the
// device's real SM count).  Kernels of other streams (the weight-gradient CTAs of the side stream) may delay a
// CTA's start but never depend on this grid, so the barrier always completes; the wait is bounded anyway.
// CTAs per SM the grid may count on.  MM3D_BN_PER_SM overrides it for A/B measurements (bench.py --ab).
static int bn_per_sm_cap() {
  const char* e = getenv("MM3D_BN_PER_SM");
  const int v = e ? atoi(e) : 0;
  return v >= 1 && v <= 8 ? v : 2;
}

template <typename Kernel>
int bn_grid(Kernel kernel, int64_t n, int cv, size_t smem) {
  const int rows_pass = kThreads / cv;
  int64_t g = mm3d_cdiv(n, (int64_t)rows_pass * 4);  // >= 4 rows per thread when there is enough work
  if (g < 1) g = 1;
  // the query is cached per device and kernel instantiation for the 16 KB upper bound of these kernels' dynamic
  // shared memory (2 * rows_pass * c floats <= 8 KB for every supported c); larger requests are queried directly
  constexpr size_t kSmemBound = 16 * 1024;
  static int cached[64] = {0};
  int& slot = cached[mm3d_device_slot()];
  int per_sm = smem <= kSmemBound ? slot : 0;
  if (per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem <= kSmemBound ? kSmemBound : smem) != cudaSuccess || per_sm < 1) {
      (void)cudaGetLastError();
      per_sm = 1;
    }
    if (smem <= kSmemBound) slot = per_sm;
  }
  if (per_sm > bn_per_sm_cap()) per_sm = bn_per_sm_cap();
  const int64_t cap = (int64_t)mm3d_sm_count() * per_sm;
  return (int)(g < cap ? g : cap);
}

}  // namespace

int* mm3d_device_err_flag();  // conv_tc.cu

// workspace: two regions of [kSlots copies of 2*c doubles of totals | kSlots + 1 barrier counters]
static size_t bn_region_bytes(int c) { return mm3d_align(sizeof(double) * (kSlots * 2 * (size_t)c + 8)); }
extern "C" size_t mm3d_bnrelu_workspace_bytes(int c) { return 2 * bn_region_bytes(c); }
static int g_bn_parity = 0;  // (host threads launching BatchNorm concurrently on one device must use separate workspaces)

#define BN_DISPATCH(KERNEL, ...)                                                        \
  do {                                                                                  \
    const int grid = mm3d_grid(n, kThreads / cv);                                       \
    if (vec4) KERNEL<4><<<grid, kThreads, smem, stream>>>(__VA_ARGS__);                 \
    else      KERNEL<1><<<grid, kThreads, smem, stream>>>(__VA_ARGS__);                 \
  } while (0)
#define BN_DISPATCH_PDL(KERNEL, ...)                                                                        \
  do {                                                                                                      \
    if (vec4) { const int grid = bn_grid(KERNEL<4>, n, cv, smem);                                           \
                MM3D_CUDA(mm3d_launch_pdl(KERNEL<4>, dim3(grid), dim3(kThreads), smem, stream, __VA_ARGS__)); } \
    else      { const int grid = bn_grid(KERNEL<1>, n, cv, smem);                                           \
                MM3D_CUDA(mm3d_launch_pdl(KERNEL<1>, dim3(grid), dim3(kThreads), smem, stream, __VA_ARGS__)); } \
  } while (0)

// ws_clean: the workspace is known to be all zero (the kernels leave it that way), skip the memset
int mm3d_bnrelu_fwd_impl(const float* x, const float* x_hi, int c_lo, float* y, int64_t n, int c, const float* gamma,
                         const float* beta,
                         float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                         float eps, float momentum, float leakiness, int training,
                         void* ws, size_t ws_bytes, bool ws_clean, int round_tf32, cudaStream_t stream) {
  MM3D_REQUIRE(c > 0 && n >= 0, MM3D_ERR_INVALID, "bad sizes");
  const bool vec4 = (c % 4 == 0) && (c_lo % 4 == 0) && ((((uintptr_t)x | (uintptr_t)x_hi | (uintptr_t)y) & 15) == 0);
  const int cv = vec4 ? c / 4 : c;
  MM3D_REQUIRE(cv <= kThreads, MM3D_ERR_UNSUPPORTED, "BatchNorm with %d channels not supported", c);
  MM3D_REQUIRE(ws_bytes >= mm3d_bnrelu_workspace_bytes(c) && ws, MM3D_ERR_WORKSPACE, "bnrelu workspace too small");
  MM3D_REQUIRE(!x_hi || (training && c_lo > 0 && c_lo < c), MM3D_ERR_INVALID, "split input needs training mode and 0 < c_lo < c");
  if (n == 0) return MM3D_OK;
  // region `parity` is this launch's (all zero: memset below, or zeroed by the previous launch), the other one is
  // zeroed by this launch for the next.  Region size follows the workspace size (the executor shares one workspace
  // among layers of different widths), so consecutive launches agree on where the regions are.
  const size_t region = ws_bytes / 2 / 256 * 256;
  const int parity = (ws_clean && training) ? (g_bn_parity ^= 1) : 0;  // (only launches that use the workspace alternate)
  double* sums = (double*)((char*)ws + (size_t)parity * region);
  double* other = (double*)((char*)ws + (size_t)(parity ^ 1) * region);
  unsigned int* sync = (unsigned int*)(sums + kSlots * 2 * c);
  const int other_words = (int)(region / 8);
  const size_t smem = sizeof(double) * 2 * c > sizeof(float) * 2 * (kThreads / cv) * (size_t)c
                          ? sizeof(double) * 2 * c : sizeof(float) * 2 * (kThreads / cv) * (size_t)c;
  if (training) {
    MM3D_REQUIRE(save_mean && save_invstd && running_mean && running_var, MM3D_ERR_INVALID, "training needs stat buffers");
    if (!ws_clean) MM3D_CUDA(cudaMemsetAsync(ws, 0, 2 * region, stream));
    BN_DISPATCH_PDL(k_bn_fwd_fused, x, x_hi, c_lo, y, n, c, gamma, beta, running_mean, running_var, save_mean, save_invstd, sums, sync,
                other, other_words, eps, momentum, leakiness, round_tf32, mm3d_device_err_flag());
  } else {
    BN_DISPATCH(k_bn_apply, x, y, n, c, gamma, beta, running_mean, running_var, save_mean, save_invstd, sums, eps,
                momentum, leakiness, training, round_tf32);
  }
  mm3d_count_launches(1);
  MM3D_CHECK_LAUNCH("mm3d_bnrelu_fwd");
  return MM3D_OK;
}

int mm3d_bnrelu_bwd_impl(const float* x, const float* x_hi, int c_lo, const float* dy, float* dx, float* dx_hi,
                         int64_t n, int c, const float* gamma, const float* beta, const float* save_mean, const float* save_invstd,
                         float* d_gamma, float* d_beta, float leakiness, int training,
                         void* ws, size_t ws_bytes, bool ws_clean, int round_flags, const float* add, int64_t add_ld,
                         cudaStream_t stream) {
  MM3D_REQUIRE(c > 0 && n >= 0, MM3D_ERR_INVALID, "bad sizes");
  MM3D_REQUIRE(!add || !dx_hi, MM3D_ERR_INVALID, "bnrelu_bwd: an addend is only supported for a single output block");
  const bool vec4 = (c % 4 == 0) && (c_lo % 4 == 0) && (!add || add_ld % 4 == 0) &&
                    ((((uintptr_t)x | (uintptr_t)x_hi | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)dx_hi | (uintptr_t)add) & 15) == 0);
  const int cv = vec4 ? c / 4 : c;
  MM3D_REQUIRE(cv <= kThreads, MM3D_ERR_UNSUPPORTED, "BatchNorm with %d channels not supported", c);
  MM3D_REQUIRE(ws_bytes >= mm3d_bnrelu_workspace_bytes(c) && ws, MM3D_ERR_WORKSPACE, "bnrelu workspace too small");
  if (n == 0) {
    if (d_gamma) MM3D_CUDA(cudaMemsetAsync(d_gamma, 0, sizeof(float) * c, stream));
    if (d_beta) MM3D_CUDA(cudaMemsetAsync(d_beta, 0, sizeof(float) * c, stream));
    return MM3D_OK;
  }
  const size_t region = ws_bytes / 2 / 256 * 256;
  const int parity = ws_clean ? (g_bn_parity ^= 1) : 0;
  double* sums = (double*)((char*)ws + (size_t)parity * region);
  double* other = (double*)((char*)ws + (size_t)(parity ^ 1) * region);
  unsigned int* sync = (unsigned int*)(sums + kSlots * 2 * c);
  const int other_words = (int)(region / 8);
  if (!ws_clean) MM3D_CUDA(cudaMemsetAsync(ws, 0, 2 * region, stream));
  const size_t smem = sizeof(double) * 2 * c > sizeof(float) * 2 * (kThreads / cv) * (size_t)c
                          ? sizeof(double) * 2 * c : sizeof(float) * 2 * (kThreads / cv) * (size_t)c;
  BN_DISPATCH_PDL(k_bn_bwd_fused, x, x_hi, c_lo, dy, dx, dx_hi, n, c, gamma, beta, save_mean, save_invstd, leakiness, sums, sync, other, other_words, d_gamma,
              d_beta, training, round_flags, add, add_ld, mm3d_device_err_flag());
  mm3d_count_launches(1);
  MM3D_CHECK_LAUNCH("mm3d_bnrelu_bwd");
  return MM3D_OK;
}

extern "C" int mm3d_bnrelu_fwd(const float* x, float* y, int64_t n, int c, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                               float eps, float momentum, float leakiness, int training,
                               void* ws, size_t ws_bytes, mm3d_stream_t stream) {
  return mm3d_bnrelu_fwd_impl(x, nullptr, 0, y, n, c, gamma, beta, running_mean, running_var, save_mean, save_invstd, eps, momentum,
                              leakiness, training, ws, ws_bytes, false, 0, (cudaStream_t)stream);
}

extern "C" int mm3d_bnrelu_bwd(const float* x, const float* dy, float* dx, int64_t n, int c, const float* gamma,
                               const float* beta, const float* save_mean, const float* save_invstd,
                               float* d_gamma, float* d_beta, float leakiness, int training,
                               void* ws, size_t ws_bytes, mm3d_stream_t stream) {
  return mm3d_bnrelu_bwd_impl(x, nullptr, 0, dy, dx, nullptr, n, c, gamma, beta, save_mean, save_invstd, d_gamma, d_beta, leakiness, training,
                              ws, ws_bytes, false, 0, nullptr, 0, (cudaStream_t)stream);
}
