// Whole-network executor: UNetSCN forward and backward as ONE C-ABI call each.
//
// The per-layer entry points are enough for a drop-in, but driving ~110 autograd nodes from
// Python costs ~8 ms of host time per step -- as much as the GPU work itself.  The executor walks
// the fixed U-Net topology of 3d_net/scn_unet.py:55-84 (VGG blocks, block_reps == 1) natively:
// it carves every activation out of one caller-provided workspace with a bump allocator that
// forward and backward replay identically, and issues the same kernels the modules use.
//
//   level l (p = m(l+1), q = m(l+2)):
//     X_l --BN--> A_l --SMC--> Y_l                                   ("pre" block)
//     if l < L-1:  Y_l --BN--> B_l --Conv 2/2--> X_{l+1} ... R_{l+1} --BN--> E_l --Deconv 2/2--> F_l
//                  J_l = [Y_l | F_l] --BN--> G_l --SMC(2p->p)--> R_l   ("post" block)
//     else         R_l = Y_l
//   stem: V (InputLayer mean) --SMC(in->m)--> X_0 ;  head: R_0 --BN--> Z --OutputLayer--> out
//
// Parameter pointers come in the module tree's order (Appendix B of SURVEY.md):
//   stem.w | per level: pre_bn{gamma,beta,rmean,rvar} pre.w [dn_bn{4} dn.w <deeper level> up_bn{4}
//   up.w post_bn{4} post.w] | head_bn{4}
#include <cuda_bf16.h>
#include <stdlib.h>

#include <cstring>
#include <vector>

#include "common.cuh"

// bnrelu.cu
int mm3d_bnrelu_fwd_impl(const float* x, const float* x_hi, int c_lo, float* y, int64_t n, int c, const float* gamma,
                         const float* beta, float* running_mean, float* running_var, float* save_mean,
                         float* save_invstd, float eps, float momentum, float leakiness, int training, void* ws,
                         size_t ws_bytes, bool ws_clean, int round_tf32, cudaStream_t stream);
int mm3d_bnrelu_bwd_impl(const float* x, const float* x_hi, int c_lo, const float* dy, float* dx, float* dx_hi,
                         int64_t n, int c, const float* gamma, const float* beta, const float* save_mean,
                         const float* save_invstd, float* d_gamma, float* d_beta, float leakiness, int training,
                         void* ws, size_t ws_bytes, bool ws_clean, int round_flags, const float* add, int64_t add_ld,
                         cudaStream_t stream);

// conv_tc.cu
size_t mm3d_conv_tc_workspace_bytes(int c_in, int c_out, int K);
int mm3d_conv_tc_supported(int c_in, int c_out, int K);
int mm3d_conv_tc_build_images(const float* const* weights, float* const* images, const int* K, const int* c_in,
                              const int* c_out, const int* flags, int n, cudaStream_t stream);
int mm3d_conv_fwd_tc_img(const float* in, int64_t n_in, int c_in, float* out, int64_t n_out, int c_out,
                         const float* wimg, int K, const void* plan, int64_t plan_cap, int accumulate, cudaStream_t stream,
                         const float* wimg_lo = nullptr, int bf16 = 0);

namespace {

struct LevelMeta {
  int64_t n;
  const int32_t* nbr;
  int64_t tstride;
  const int32_t* parent;
  const uint8_t* off;
  const int32_t* child;
  const void *plan_smc, *plan_down, *plan_up;  // row plans of the three tables (tensor-core modes)
  int64_t plan_cap;
};

struct Bump {
  char* base;
  size_t off, cap;
  bool ok;
  float* f(int64_t n, int64_t c) {
    size_t bytes = mm3d_align(sizeof(float) * (size_t)n * (size_t)c);
    size_t o = off;
    off += bytes;
    if (off > cap) ok = false;
    return base ? reinterpret_cast<float*>(base + o) : nullptr;
  }
};

struct LevelBufs {
  float *X, *A, *Y, *B, *E, *F, *J, *G, *R;
  float *s_pre, *s_dn, *s_up, *s_post;  // BN save_mean | save_invstd ([2][C])
};

struct Net {
  int L, m, cin, cin_k;  // cin_k = stem input channels as seen by the kernels (padded to 16 in TC modes)
  int mode;
  std::vector<LevelMeta> lv;
  std::vector<LevelBufs> b;
  float *V, *Vp, *Z, *s_head;
  int64_t n_points;
  int planes(int l) const { return m * (l + 1); }
  // TF32x3 mode: every tensor a convolution gathers carries a hi and a lo plane ([rows, c] floats each)
  // BF16 mode: an FP32 plane (TF32-rounded: the weight-gradient operand) and a BF16 plane behind it (what forward and
  // dgrad gather; it only fills the first half of its plane)
  int pl() const { return (mode == MM3D_MODE_TF32X3 || mode == MM3D_MODE_BF16) ? 2 : 1; }
  // how a producer stores a tensor that a convolution gathers (bnrelu.cu store_planes / k_pad_cols)
  int rmode() const { return mode == MM3D_MODE_TF32X3 ? 2 : mode == MM3D_MODE_BF16 ? 3 : mode == MM3D_MODE_TF32 ? 1 : 0; }
  bool x3() const { return mode == MM3D_MODE_TF32X3; }
  bool bf16() const { return mode == MM3D_MODE_BF16; }
};

// the same carving in forward, backward and the size query
void carve(Net& net, Bump& bp) {
  const int64_t n0 = net.lv[0].n;
  net.V = bp.f(n0, net.cin);
  const int pl = net.pl();
  net.Vp = (net.cin_k != net.cin || pl == 2) ? bp.f(n0, (int64_t)net.cin_k * pl) : net.V;
  net.b.assign(net.L, LevelBufs());
  for (int l = 0; l < net.L; ++l) {
    const int64_t n = net.lv[l].n;
    const int p = net.planes(l);
    LevelBufs& B = net.b[l];
    B.X = bp.f(n, p);
    B.A = bp.f(n, (int64_t)p * pl);
    B.Y = bp.f(n, p);
    B.s_pre = bp.f(2, p);
    if (l + 1 < net.L) {
      const int q = net.planes(l + 1);
      B.B = bp.f(n, (int64_t)p * pl);
      B.s_dn = bp.f(2, p);
      B.E = bp.f(net.lv[l + 1].n, (int64_t)q * pl);
      B.s_up = bp.f(2, q);
      B.F = bp.f(n, p);
      B.J = bp.f(n, 2 * p);
      B.G = bp.f(n, (int64_t)2 * p * pl);
      B.s_post = bp.f(2, 2 * p);
      B.R = bp.f(n, p);
    } else {
      B.R = B.Y;
    }
  }
  net.Z = bp.f(n0, net.m);
  net.s_head = bp.f(2, net.m);
}

__global__ void k_copy_cols(const float* __restrict__ src, int64_t n, int c_src, float* __restrict__ dst, int c_dst,
                            int dst_col0, int ncols, int src_col0) {
  mm3d_griddep_launch();
  mm3d_griddep_wait();
  // dst[r, dst_col0 + j] = src[r, src_col0 + j], j < ncols  (other dst columns untouched)
  if (((ncols | c_src | c_dst | dst_col0 | src_col0) & 3) == 0) {  // whole float4s (every layer of the network)
    const int nv = ncols >> 2;
    const int64_t total = n * nv;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = i / nv;
      const int j = (int)(i - r * nv) << 2;
      *reinterpret_cast<float4*>(dst + r * c_dst + dst_col0 + j) =
          __ldg(reinterpret_cast<const float4*>(src + r * c_src + src_col0 + j));
    }
    return;
  }
  const int64_t total = n * ncols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ncols;
    const int j = (int)(i - r * ncols);
    dst[r * c_dst + dst_col0 + j] = __ldg(src + r * c_src + src_col0 + j);
  }
}

// JoinTable: J = [A | B] (both [n, p], p a multiple of 4) in one pass
__global__ void k_concat2(const float* __restrict__ a, const float* __restrict__ b, int64_t n, int p, float* __restrict__ j_out) {
  mm3d_griddep_launch();
  mm3d_griddep_wait();
  const int pv = p >> 2;
  const int64_t total = n * 2 * pv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (2 * pv);
    const int v = (int)(i - r * 2 * pv);  // float4 column of J
    const float* src = v < pv ? a + r * p + (v << 2) : b + r * p + ((v - pv) << 2);
    *reinterpret_cast<float4*>(j_out + r * 2 * p + (v << 2)) = __ldg(reinterpret_cast<const float4*>(src));
  }
}

__global__ void k_pad_cols(const float* __restrict__ src, int64_t n, int c_src, float* __restrict__ dst, int c_dst,
                           int round_tf32) {
  mm3d_griddep_launch();
  mm3d_griddep_wait();
  // dst[r, j] = j < c_src ? src[r, j] : 0   (c_dst >= c_src: pad;  c_dst < c_src: slice); optionally RNA-rounded to
  // TF32 (the padded stem input is the A operand of a kind::tf32 MMA)
  const int64_t total = n * c_dst;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c_dst;
    const int j = (int)(i - r * c_dst);
    float v = j < c_src ? __ldg(src + r * c_src + j) : 0.f;
    if (round_tf32) {
      const float hi = mm3d_rna_tf32(v);
      if (round_tf32 == 2) dst[total + i] = mm3d_rna_tf32(v - hi);  // lo plane (TF32x3 mode)
      if (round_tf32 == 3) reinterpret_cast<__nv_bfloat16*>(dst + total)[i] = __float2bfloat16_rn(v);  // BF16 plane
      v = hi;
    }
    dst[i] = v;
  }
}

// out = RNA-rounded-to-TF32(in), whole float4s + tail (module-by-module path: operands of a TF32 convolution that
// were not produced by one of this library's rounding producers)
__global__ void k_round_tf32(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
  mm3d_griddep_launch();
  mm3d_griddep_wait();
  const int64_t nv = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    v.x = mm3d_rna_tf32(v.x); v.y = mm3d_rna_tf32(v.y); v.z = mm3d_rna_tf32(v.z); v.w = mm3d_rna_tf32(v.w);
    reinterpret_cast<float4*>(out)[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) out[(nv << 2) + threadIdx.x] = mm3d_rna_tf32(__ldg(in + (nv << 2) + threadIdx.x));
}

// Backward runs the weight gradients on a second stream: a layer's wgrad only feeds the optimiser, while its
// dgrad is on the critical chain, and neither kernel fills the GPU alone at the deeper levels.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[64];
  int n_ev = 0, next = 0;
  bool ok = false;
};
SideStream* side_stream() {
  static SideStream pool[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& s = pool[dev];
  if (!s.ok) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (s.n_ev = 0; s.n_ev < 64; ++s.n_ev)
      if (cudaEventCreateWithFlags(&s.ev[s.n_ev], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    s.ok = true;
  }
  return &s;
}

struct Ctx {
  Net* net;
  void* const* params;
  void* const* grads;  // backward only
  void* scratch;
  size_t scratch_bytes;
  cudaStream_t stream;
  float eps, momentum;
  int training;
  int pi;  // running parameter index
  int rc;
  SideStream* side = nullptr;  // backward only; NULL = everything on `stream`
  void* bn_ws = nullptr;       // BatchNorm totals + barrier counters: zeroed once per call, the kernels keep it clean
  size_t bn_ws_bytes = 0;
  // tensor-core modes: the weight images of all layers of this direction, built by one launch up front
  void* img_ws = nullptr;
  size_t img_ws_bytes = 0;
  const char *grad_lo = nullptr, *grad_hi = nullptr;  // span of the caller's gradient buffers when contiguous
  int wgrad_acc = 0;  // 1: the weight-gradient buffers were zeroed by one memset up front, wgrad kernels accumulate
  std::vector<const float*> img_key;  // weight pointer ...
  std::vector<const float*> img_val;  // ... -> its image
  std::vector<const float*> img_lo;   // ... -> the image of its TF32 remainder (TF32x3 mode)
};

// stream for a layer's weight gradient: the side stream once everything enqueued on the main stream so far
// (the layer's d_out) is done
cudaStream_t wgrad_stream(Ctx& c) {
  if (!c.side || c.rc) return c.stream;
  cudaEvent_t e = c.side->ev[c.side->next++ % c.side->n_ev];
  if (cudaEventRecord(e, c.stream) != cudaSuccess || cudaStreamWaitEvent(c.side->stream, e, 0) != cudaSuccess) {
    mm3d_set_error("unet backward: could not fork the weight-gradient stream");
    c.rc = MM3D_ERR_CUDA;
    return c.stream;
  }
  return c.side->stream;
}

// Timing experiments only: a library built with -DMM3D_ABLATION reads MM3D_ABL_SKIP, a list of kernel families the
// executor then does not launch (any of "bn", "wgrad", "conv" = forward + dgrad); results are garbage.  The
// product build has no such switch.
#ifdef MM3D_ABLATION
bool abl_skip(const char* what) {
  static const char* e = getenv("MM3D_ABL_SKIP");
  return e && strstr(e, what) != nullptr;
}
#else
constexpr bool abl_skip(const char*) { return false; }
#endif

#define EX(call)                   \
  do {                             \
    if (c.rc == 0) c.rc = (call);  \
  } while (0)

#ifdef MM3D_TRACE
// development builds: an event after every operation of the main stream; mm3d_debug_dump_marks() prints the time
// between consecutive events of the last forward + backward
struct Mark { const char* name; int level; cudaEvent_t ev; };
static std::vector<Mark> g_marks;
static bool g_marks_on = false;
static void trace_mark(cudaStream_t s, const char* name, int level) {
  if (!g_marks_on) return;
  Mark m{name, level, nullptr};
  cudaEventCreate(&m.ev);
  cudaEventRecord(m.ev, s);
  g_marks.push_back(m);
}
#define MARK(name, level) trace_mark(c.stream, name, level)
#else
#define MARK(name, level) do {} while (0)
#endif

const float* P(Ctx& c, int i) { return (const float*)c.params[i]; }
float* Gp(Ctx& c, int i) { return (float*)c.grads[i]; }

// x_hi != NULL: the input is the column blocks [x | x_hi] ([n, c_lo] and [n, ch - c_lo]) -- JoinTable without
// materialising the concatenation (training mode only)
// Tensor-core modes: every tensor that a kind::tf32 MMA reads as its gathered operand (activations in forward and
// wgrad, d_out in dgrad and wgrad) is stored RNA-rounded to TF32 by the kernel that PRODUCES it, so that the
// tensor core's truncation of the low mantissa bits is exact.  `to_conv`: the output feeds a convolution.
bool tc_mode(const Ctx& c) { return c.net->mode != MM3D_MODE_FP32; }

void bn_fwd(Ctx& c, int pidx, const float* x, float* y, int64_t n, int ch, float* save, const float* x_hi = nullptr,
            int c_lo = 0, bool to_conv = true) {
  if (abl_skip("bn")) return;
  EX(mm3d_bnrelu_fwd_impl(x, x_hi, c_lo, y, n, ch, P(c, pidx), P(c, pidx + 1), (float*)c.params[pidx + 2], (float*)c.params[pidx + 3],
                          save, save + ch, c.eps, c.momentum, 0.f, c.training, c.bn_ws, c.bn_ws_bytes, true,
                          tc_mode(c) && to_conv ? c.net->rmode() : 0, c.stream));
}
// round: bit 0 = dx (its low column block when split) feeds a convolution as d_out, bit 1 = dx_hi does
// add (rows add_ld floats apart): another gradient of the same tensor, summed into dx before it is stored
void bn_bwd(Ctx& c, int pidx, const float* x, const float* dy, float* dx, int64_t n, int ch, const float* save,
            int round, const float* x_hi = nullptr, int c_lo = 0, float* dx_hi = nullptr, const float* add = nullptr,
            int64_t add_ld = 0) {
  if (abl_skip("bn")) return;
  EX(mm3d_bnrelu_bwd_impl(x, x_hi, c_lo, dy, dx, dx_hi, n, ch, P(c, pidx), P(c, pidx + 1), save, save + ch, Gp(c, pidx), Gp(c, pidx + 1), 0.f,
                          c.training, c.bn_ws, c.bn_ws_bytes, true, tc_mode(c) ? (round | (c.net->x3() ? 4 : c.net->bf16() ? 8 : 0)) : 0, add, add_ld,
                          c.stream));
}
enum Kind { SMC, DOWN, UP };

const float* find_img(const Ctx& c, const float* w) {
  for (size_t i = 0; i < c.img_key.size(); ++i)
    if (c.img_key[i] == w) return c.img_val[i];
  return nullptr;
}
const float* find_img_lo(const Ctx& c, const float* w) {
  for (size_t i = 0; i < c.img_key.size() && i < c.img_lo.size(); ++i)
    if (c.img_key[i] == w) return c.img_lo[i];
  return nullptr;
}
// One rule-table convolution from prebuilt weight images.  TF32x3 mode: hi.Whi + lo.Whi + hi.Wlo inside one launch (`in`
// carries the lo plane n_in * c_in floats behind the hi one, both planes are gathered into every ring stage).
int conv_img(const Ctx& c, const float* in, int64_t n_in, int c_in, float* out, int64_t n_out, int c_out, const float* im,
             const float* im_lo, int K, const void* plan, int64_t plan_cap) {
  if (c.net->x3() && !im_lo) { mm3d_set_error("tf32x3: missing lo weight image"); return MM3D_ERR_INVALID; }
  if (c.net->bf16())  // the BF16 plane sits n_in * c_in floats behind the FP32 one
    return mm3d_conv_fwd_tc_img(in + n_in * (int64_t)c_in, n_in, c_in, out, n_out, c_out, im, K, plan, plan_cap, 0, c.stream, nullptr, 1);
  return mm3d_conv_fwd_tc_img(in, n_in, c_in, out, n_out, c_out, im, K, plan, plan_cap, 0, c.stream, c.net->x3() ? im_lo : nullptr);
}
// forward of layer type `kind` whose FINE level is l
void conv_fwd(Ctx& c, Kind kind, int l, const float* in, int c_in, float* out, int c_out, const float* w) {
  if (abl_skip("conv")) return;
  const LevelMeta& f = c.net->lv[l];
  if (const float* im = find_img(c, w)) {  // prebuilt weight image: the tcgen05 kernel directly
    const int64_t nc = l + 1 < c.net->L ? c.net->lv[l + 1].n : 0;
    const float* il = find_img_lo(c, w);
    if (kind == SMC) EX(conv_img(c, in, f.n, c_in, out, f.n, c_out, im, il, 27, f.plan_smc, f.plan_cap));
    else if (kind == DOWN) EX(conv_img(c, in, f.n, c_in, out, nc, c_out, im, il, 8, f.plan_down, f.plan_cap));
    else EX(conv_img(c, in, nc, c_in, out, f.n, c_out, im, il, 8, f.plan_up, f.plan_cap));
    return;
  }
  if (kind == SMC)
    EX(mm3d_conv_fwd(in, f.n, c_in, out, f.n, c_out, w, 27, f.nbr, f.tstride, nullptr, f.plan_smc, f.plan_cap, 0,
                     c.net->mode, c.scratch, c.scratch_bytes, c.stream));
  else if (kind == DOWN)
    EX(mm3d_conv_fwd(in, f.n, c_in, out, c.net->lv[l + 1].n, c_out, w, 8, f.child, f.tstride, nullptr, f.plan_down,
                     f.plan_cap, 0, c.net->mode, c.scratch, c.scratch_bytes, c.stream));
  else
    EX(mm3d_conv_fwd(in, c.net->lv[l + 1].n, c_in, out, f.n, c_out, w, 8, f.parent, 0, f.off, f.plan_up, f.plan_cap, 0,
                     c.net->mode, c.scratch, c.scratch_bytes, c.stream));
}
// dgrad (d_in from d_out) and wgrad of the same layer; c_in / c_out are the FORWARD layer's
void conv_bwd(Ctx& c, Kind kind, int l, const float* in, int c_in, const float* d_out, int c_out, const float* w,
              float* d_in, float* d_w) {
  const LevelMeta& f = c.net->lv[l];
  const int64_t nc = l + 1 < c.net->L ? c.net->lv[l + 1].n : 0;
  const int md = c.net->mode;
  cudaStream_t ws = wgrad_stream(c);  // d_out is complete on the main stream at this point
  const float* im = find_img(c, w);   // prebuilt dgrad image (tensor-core modes)
  const float* il = find_img_lo(c, w);
  // parameter gradients inside the caller's (pre-zeroed) flat buffer accumulate; temporaries are overwritten
  const int acc = c.wgrad_acc && c.grad_lo <= (const char*)d_w && (const char*)d_w < c.grad_hi;
  const bool do_wg = !abl_skip("wgrad") && d_w != nullptr, do_dg = !abl_skip("conv");  // (frozen weight: no d_w)
  // The weight gradient (side stream) is enqueued before the data gradient.  (Measured with the in-process A/B of
  // bench.py on the final kernels: the order makes no difference, 2.959 vs 2.958 ms per step -- the step is bound by
  // the total work of both streams, not by which kernel gets the SMs first.)
  if (kind == SMC) {
    if (do_wg)
      EX(mm3d_conv_wgrad(in, f.n, c_in, d_out, f.n, c_out, d_w, 27, f.nbr, f.tstride, nullptr, f.plan_smc, f.plan_cap, acc,
                         md, nullptr, 0, ws));
    if (d_in && im && do_dg)
      EX(conv_img(c, d_out, f.n, c_out, d_in, f.n, c_in, im, il, 27, f.plan_smc, f.plan_cap));
    else if (d_in && do_dg)
      EX(mm3d_conv_fwd(d_out, f.n, c_out, d_in, f.n, c_in, w, 27, f.nbr, f.tstride, nullptr, f.plan_smc, f.plan_cap,
                       MM3D_CONV_TRANSPOSE_W | MM3D_CONV_MIRROR_K, md, c.scratch, c.scratch_bytes, c.stream));
  } else if (kind == DOWN) {  // in: fine rows, d_out: coarse rows
    if (do_wg)
      EX(mm3d_conv_wgrad(in, f.n, c_in, d_out, nc, c_out, d_w, 8, f.child, f.tstride, nullptr, f.plan_down, f.plan_cap, acc,
                         md, nullptr, 0, ws));
    if (im && do_dg) EX(conv_img(c, d_out, nc, c_out, d_in, f.n, c_in, im, il, 8, f.plan_up, f.plan_cap));
    else if (do_dg)
      EX(mm3d_conv_fwd(d_out, nc, c_out, d_in, f.n, c_in, w, 8, f.parent, 0, f.off, f.plan_up, f.plan_cap,
                       MM3D_CONV_TRANSPOSE_W, md, c.scratch, c.scratch_bytes, c.stream));
  } else {  // UP: in: coarse rows, d_out: fine rows
    if (do_wg)
      EX(mm3d_conv_wgrad(in, nc, c_in, d_out, f.n, c_out, d_w, 8, f.parent, 0, f.off, f.plan_up, f.plan_cap, acc, md,
                         nullptr, 0, ws));
    if (im && do_dg) EX(conv_img(c, d_out, f.n, c_out, d_in, nc, c_in, im, il, 8, f.plan_down, f.plan_cap));
    else if (do_dg)
      EX(mm3d_conv_fwd(d_out, f.n, c_out, d_in, nc, c_in, w, 8, f.child, f.tstride, nullptr, f.plan_down, f.plan_cap,
                       MM3D_CONV_TRANSPOSE_W, md, c.scratch, c.scratch_bytes, c.stream));
  }
}

// main stream waits for everything enqueued on the side stream so far
void join_side(Ctx& c) {
  if (!c.side || c.rc) return;
  cudaEvent_t e = c.side->ev[c.side->next++ % c.side->n_ev];
  if (cudaEventRecord(e, c.side->stream) != cudaSuccess || cudaStreamWaitEvent(c.stream, e, 0) != cudaSuccess) {
    mm3d_set_error("unet backward: could not join the weight-gradient stream");
    c.rc = MM3D_ERR_CUDA;
  }
}

void launch_copy_cols(Ctx& c, const float* src, int64_t n, int c_src, float* dst, int c_dst, int col0, int ncols) {
  if (n == 0 || c.rc) return;
  if (mm3d_launch_pdl(k_copy_cols, dim3(mm3d_grid(n * ncols / 4 + 1, 256)), dim3(256), 0, c.stream, src, n, c_src, dst, c_dst, col0, ncols, 0) != cudaSuccess) c.rc = MM3D_ERR_CUDA;
  mm3d_count_launches(1);
}
void launch_pad_cols(Ctx& c, const float* src, int64_t n, int c_src, float* dst, int c_dst, int round_tf32 = 0) {
  if (n == 0 || c.rc) return;
  if (mm3d_launch_pdl(k_pad_cols, dim3(mm3d_grid(n * c_dst, 256)), dim3(256), 0, c.stream, src, n, c_src, dst, c_dst, round_tf32) != cudaSuccess) c.rc = MM3D_ERR_CUDA;
  mm3d_count_launches(1);
}

// parameter index bookkeeping: number of pointer slots a level (and everything below it) uses
int level_slots(int l, int L) { return l + 1 < L ? 5 + 5 + level_slots(l + 1, L) + 5 + 5 : 5; }

struct ConvInfo {
  int pidx, K, c_in, c_out, smc;
};
// every convolution of the network below level l (slot layout as in level_fwd)
void list_convs(const Net& net, int l, int pbase, std::vector<ConvInfo>& v) {
  const int p = net.planes(l);
  v.push_back({pbase + 4, 27, p, p, 1});
  if (l + 1 < net.L) {
    const int q = net.planes(l + 1);
    const int dn = pbase + 5, deeper = pbase + 10, up = deeper + level_slots(l + 1, net.L), post = up + 5;
    v.push_back({dn + 4, 8, p, q, 0});
    list_convs(net, l + 1, deeper, v);
    v.push_back({up + 4, 8, q, p, 0});
    v.push_back({post + 4, 27, 2 * p, p, 1});
  }
}

size_t img_region_bytes(const Net& net) {
  if (net.mode == MM3D_MODE_FP32) return 0;
  std::vector<ConvInfo> v;
  v.push_back({0, 27, net.cin_k, net.m, 1});
  list_convs(net, 0, 1, v);
  size_t fwd = 0, bwd = 0;
  for (const ConvInfo& ci : v) {
    fwd += mm3d_align(mm3d_conv_tc_workspace_bytes(ci.c_in, ci.c_out, ci.K));
    bwd += mm3d_align(mm3d_conv_tc_workspace_bytes(ci.c_out, ci.c_in, ci.K));
  }
  return (fwd > bwd ? fwd : bwd) * (size_t)net.pl() + 256;
}

// One launch builds the weight images of every layer for this direction (dgrad: W^T, mirrored for 3^3 layers).
int build_images(Ctx& c, bool backward, const float* w_stem) {
  const Net& net = *c.net;
  if (net.mode == MM3D_MODE_FP32 || !c.img_ws) return MM3D_OK;
  std::vector<ConvInfo> v;
  v.push_back({0, 27, net.cin_k, net.m, 1});
  list_convs(net, 0, 1, v);
  std::vector<const float*> w;
  std::vector<float*> img;
  std::vector<int> K, ci, co, fl;
  char* at = (char*)c.img_ws;
  for (const ConvInfo& x : v) {
    const int a = backward ? x.c_out : x.c_in, b = backward ? x.c_in : x.c_out;
    if (!mm3d_conv_tc_supported(a, b, x.K)) continue;
    const size_t bytes = mm3d_align(mm3d_conv_tc_workspace_bytes(a, b, x.K));
    MM3D_REQUIRE((size_t)(at - (char*)c.img_ws) + bytes <= c.img_ws_bytes, MM3D_ERR_WORKSPACE, "weight image region too small");
    w.push_back(x.pidx == 0 ? w_stem : (const float*)c.params[x.pidx]);
    img.push_back((float*)at);
    K.push_back(x.K); ci.push_back(a); co.push_back(b);
    fl.push_back((backward ? (MM3D_CONV_TRANSPOSE_W | (x.smc ? MM3D_CONV_MIRROR_K : 0)) : 0) | (net.bf16() ? MM3D_CONV_BF16 : 0));
    at += bytes;
  }
  int rc = mm3d_conv_tc_build_images(w.data(), img.data(), K.data(), ci.data(), co.data(), fl.data(), (int)w.size(), c.stream);
  if (rc) return rc;
  c.img_key.assign(w.begin(), w.end());
  c.img_val.assign(img.begin(), img.end());
  c.img_lo.clear();
  if (net.x3()) {  // the images of the weights' TF32 remainders, behind the hi images
    std::vector<float*> lo;
    for (size_t i = 0; i < w.size(); ++i) {
      const size_t bytes = mm3d_align(mm3d_conv_tc_workspace_bytes(ci[i], co[i], K[i]));
      MM3D_REQUIRE((size_t)(at - (char*)c.img_ws) + bytes <= c.img_ws_bytes, MM3D_ERR_WORKSPACE, "weight image region too small");
      lo.push_back((float*)at);
      at += bytes;
      fl[i] |= MM3D_CONV_WEIGHT_LO;
    }
    rc = mm3d_conv_tc_build_images(w.data(), lo.data(), K.data(), ci.data(), co.data(), fl.data(), (int)w.size(), c.stream);
    if (rc) return rc;
    c.img_lo.assign(lo.begin(), lo.end());
  }
  return MM3D_OK;
}

void level_fwd(Ctx& c, int l, int pbase) {
  Net& net = *c.net;
  LevelBufs& B = net.b[l];
  const int64_t n = net.lv[l].n;
  const int p = net.planes(l);
  bn_fwd(c, pbase, B.X, B.A, n, p, B.s_pre);
  MARK("bn_fwd pre", l);
  conv_fwd(c, SMC, l, B.A, p, B.Y, p, P(c, pbase + 4));
  MARK("smc_fwd pre", l);
  if (l + 1 < net.L) {
    const int q = net.planes(l + 1);
    const int dn = pbase + 5, deeper = pbase + 10, up = deeper + level_slots(l + 1, net.L), post = up + 5;
    bn_fwd(c, dn, B.Y, B.B, n, p, B.s_dn);
    MARK("bn_fwd dn", l);
    conv_fwd(c, DOWN, l, B.B, p, net.b[l + 1].X, q, P(c, dn + 4));
    MARK("down_fwd", l);
    level_fwd(c, l + 1, deeper);
    bn_fwd(c, up, net.b[l + 1].R, B.E, net.lv[l + 1].n, q, B.s_up);
    MARK("bn_fwd up", l);
    conv_fwd(c, UP, l, B.E, q, B.F, p, P(c, up + 4));
    MARK("up_fwd", l);
    if ((p & 3) == 0 && c.training) {
      // JoinTable is never materialised: BatchNorm reads the two column blocks [Y | F] directly
      bn_fwd(c, post, B.Y, B.G, n, 2 * p, B.s_post, B.F, p);
      MARK("bn_fwd post", l);
    } else {
      if ((p & 3) == 0) {
        if (n && !c.rc) {
          if (mm3d_launch_pdl(k_concat2, dim3(mm3d_grid(n * p / 2 + 1, 256)), dim3(256), 0, c.stream, (const float*)B.Y,
                              (const float*)B.F, n, p, B.J) != cudaSuccess)
            c.rc = MM3D_ERR_CUDA;
          mm3d_count_launches(1);
        }
      } else {
        launch_copy_cols(c, B.Y, n, p, B.J, 2 * p, 0, p);
        launch_copy_cols(c, B.F, n, p, B.J, 2 * p, p, p);
      }
      bn_fwd(c, post, B.J, B.G, n, 2 * p, B.s_post);
    }
    conv_fwd(c, SMC, l, B.G, 2 * p, B.R, p, P(c, post + 4));
    MARK("smc_fwd post", l);
  }
}

// d_R: gradient w.r.t. the level's output R_l; returns (in d_X) the gradient w.r.t. X_l.
// Gradient temporaries come from the backward bump allocator `g`.
void level_bwd(Ctx& c, Bump& g, int l, int pbase, const float* d_R, float* d_X) {
  Net& net = *c.net;
  LevelBufs& B = net.b[l];
  const int64_t n = net.lv[l].n;
  const int p = net.planes(l);
  const float* d_Y = d_R;
  const size_t mark = g.off;
  if (l + 1 < net.L) {
    const int q = net.planes(l + 1);
    const int64_t nc = net.lv[l + 1].n;
    const int dn = pbase + 5, deeper = pbase + 10, up = deeper + level_slots(l + 1, net.L), post = up + 5;
    float* d_G = g.f(n, 2 * p);
    conv_bwd(c, SMC, l, B.G, 2 * p, d_R, p, P(c, post + 4), d_G, Gp(c, post + 4));
    MARK("smc_dgrad post", l);
    const bool split = (p & 3) == 0 && c.training;  // as in the forward: [Y | F] was never concatenated
    float* d_J = g.f(n, split ? p : 2 * p);         // split: only the skip half d_J[:, :p]
    const int pl = net.pl();  // conv d_out tensors carry a second plane in the TF32x3 and BF16 modes
    float* d_F = g.f(n, (int64_t)p * pl);
    float* d_Yskip = g.f(n, (int64_t)p * pl);
    if (split) {
      bn_bwd(c, post, B.Y, d_G, d_J, n, 2 * p, B.s_post, 2, B.F, p, d_F);  // d_F is the deconvolution's d_out
      MARK("bn_bwd post", l);
    } else {
      bn_bwd(c, post, B.J, d_G, d_J, n, 2 * p, B.s_post, 0);  // (eval-mode backward: d_F stays unrounded)
      // d_F = d_J[:, p:]; the skip half is combined with the branch gradient further down
      if (n && !c.rc) {
        if (mm3d_launch_pdl(k_copy_cols, dim3(mm3d_grid(n * p / 4 + 1, 256)), dim3(256), 0, c.stream, (const float*)d_J, n, 2 * p, d_F, p, 0, p, p) != cudaSuccess) c.rc = MM3D_ERR_CUDA;
        mm3d_count_launches(1);
        if (tc_mode(c)) launch_pad_cols(c, d_F, n, p, d_F, p, net.rmode());  // round in place (+ second plane): d_F is the deconvolution's d_out
      }
    }
    float* d_E = g.f(nc, q);
    conv_bwd(c, UP, l, B.E, q, d_F, p, P(c, up + 4), d_E, Gp(c, up + 4));
    MARK("up_dgrad", l);
    float* d_Rn = g.f(nc, (int64_t)q * pl);
    bn_bwd(c, up, net.b[l + 1].R, d_E, d_Rn, nc, q, B.s_up, 1);
    MARK("bn_bwd up", l);
    float* d_Xn = g.f(nc, (int64_t)q * pl);
    level_bwd(c, g, l + 1, deeper, d_Rn, d_Xn);
    float* d_B = g.f(n, p);
    conv_bwd(c, DOWN, l, B.B, p, d_Xn, q, P(c, dn + 4), d_B, Gp(c, dn + 4));
    MARK("down_dgrad", l);
    // d_Y = d_J[:, :p] (the skip half of the join's gradient) + the branch gradient: summed inside the BatchNorm
    // backward of the down-branch, stored rounded (it is the pre-block convolution's d_out)
    bn_bwd(c, dn, B.Y, d_B, d_Yskip, n, p, B.s_dn, 1, nullptr, 0, nullptr, d_J, split ? p : 2 * p);
    MARK("bn_bwd dn", l);
    d_Y = d_Yskip;
  }
  float* d_A = g.f(n, p);
  conv_bwd(c, SMC, l, B.A, p, d_Y, p, P(c, pbase + 4), d_A, Gp(c, pbase + 4));
  MARK("smc_dgrad pre", l);
  bn_bwd(c, pbase, B.X, d_A, d_X, n, p, B.s_pre, 1);
  MARK("bn_bwd pre", l);
  if (!c.side) g.off = mark;  // temporaries of this level are dead once d_X is written -- unless the side stream
                              // may still be reading a d_out (the workspace bound assumes no reuse anyway)
}

int fill_net(Net& net, int in_channels, int m, int num_planes, int mode, const int64_t* level_desc, int64_t n_points) {
  MM3D_REQUIRE(num_planes >= 1 && num_planes <= 16 && m > 0 && in_channels > 0, MM3D_ERR_INVALID, "bad network shape");
  MM3D_REQUIRE(mode == MM3D_MODE_FP32 || mode == MM3D_MODE_TF32 || mode == MM3D_MODE_TF32X3 || mode == MM3D_MODE_BF16, MM3D_ERR_UNSUPPORTED,
               "conv mode %d not implemented in this build", mode);
  net.L = num_planes; net.m = m; net.cin = in_channels; net.mode = mode; net.n_points = n_points;
  // tensor-core kernels gather rows in 64-byte pieces: pad the stem input to a multiple of 16 channels
  net.cin_k = in_channels;
  if (mode != MM3D_MODE_FP32) net.cin_k = (in_channels + 15) / 16 * 16;
  net.lv.resize(num_planes);
  for (int l = 0; l < num_planes; ++l) {
    const int64_t* d = level_desc + MM3D_LEVEL_DESC_WORDS * l;
    net.lv[l] = LevelMeta{d[0], (const int32_t*)d[1], d[2], (const int32_t*)d[3], (const uint8_t*)d[4], (const int32_t*)d[5],
                          (const void*)d[6], (const void*)d[7], (const void*)d[8], d[9]};
  }
  return MM3D_OK;
}

size_t bwd_temp_bytes(const Net& net) {
  // upper bound of the live gradient temporaries along the recursion
  size_t total = 0;
  for (int l = 0; l < net.L; ++l) {
    const size_t n = (size_t)net.lv[l].n, p = (size_t)net.planes(l);
    size_t rows = 0;
    if (l + 1 < net.L) {
      const size_t nc = (size_t)net.lv[l + 1].n, q = (size_t)net.planes(l + 1);
      rows = n * (2 * p + 2 * p + p + p + p + p) + nc * (3 * q);
    }
    rows += n * p;
    total += mm3d_align(sizeof(float) * rows) + 16 * 256;
  }
  const size_t n0 = (size_t)net.lv[0].n;
  total += mm3d_align(4 * n0 * (size_t)net.m) * 3 + mm3d_align(4 * n0 * (size_t)(net.cin + net.cin_k)) * 2 + 8 * 256;
  total += mm3d_align(4 * (size_t)27 * net.cin_k * net.m) * 2;
  return total * (size_t)net.pl();
}

}  // namespace

extern "C" {

// level_desc: MM3D_LEVEL_DESC_WORDS int64 per level = {n rows, nbr table ptr, table stride, parent ptr, off ptr,
// child ptr, plan of the 3^3 table, plan of the child table, plan of the (parent, off) table, plan capacity}
MM3D_API int64_t mm3d_unet_num_params(int num_planes) { return 1 + level_slots(0, num_planes) + 4; }

MM3D_API size_t mm3d_unet_act_bytes(int in_channels, int m, int num_planes, int mode, const int64_t* level_desc,
                                    int64_t n_points) {
  Net net;
  if (fill_net(net, in_channels, m, num_planes, mode, level_desc, n_points)) return 0;
  Bump bp{nullptr, 0, ~(size_t)0, true};
  carve(net, bp);
  return bp.off + 256;
}

MM3D_API size_t mm3d_unet_bwd_bytes(int in_channels, int m, int num_planes, int mode, const int64_t* level_desc,
                                    int64_t n_points) {
  Net net;
  if (fill_net(net, in_channels, m, num_planes, mode, level_desc, n_points)) return 0;
  return bwd_temp_bytes(net);
}

MM3D_API size_t mm3d_unet_scratch_bytes(int in_channels, int m, int num_planes, int mode) {
  size_t best = mm3d_bnrelu_workspace_bytes(2 * m * num_planes);
  const int cin_k = (in_channels + 15) / 16 * 16;
  size_t s = mm3d_conv_workspace_bytes(0, 0, cin_k, m, 27, mode);
  if (s > best) best = s;
  for (int l = 0; l < num_planes; ++l) {
    const int p = m * (l + 1);
    s = mm3d_conv_workspace_bytes(0, 0, 2 * p, p, 27, mode);
    if (s > best) best = s;
    s = mm3d_conv_workspace_bytes(0, 0, p + m, p, 8, mode);
    if (s > best) best = s;
  }
  Net shape;
  shape.L = num_planes; shape.m = m; shape.cin = in_channels; shape.cin_k = cin_k; shape.mode = mode;
  return best + img_region_bytes(shape) + mm3d_bnrelu_workspace_bytes(2 * m * num_planes) +
         mm3d_align(4 * (size_t)27 * cin_k * m) + 256;
}

// carve the executor's private tail of the scratch buffer: [... | BatchNorm workspace | padded stem weight]
int carve_tail(Ctx& c, const Net& net) {
  const size_t wp_bytes = mm3d_align(4 * (size_t)27 * net.cin_k * net.m);
  c.bn_ws_bytes = mm3d_bnrelu_workspace_bytes(2 * net.m * net.L);
  c.img_ws_bytes = img_region_bytes(net) / 256 * 256;
  MM3D_REQUIRE(c.scratch_bytes >= wp_bytes + c.bn_ws_bytes + c.img_ws_bytes + 256, MM3D_ERR_WORKSPACE,
               "unet scratch too small");
  const size_t end = c.scratch_bytes / 256 * 256;
  c.bn_ws = (char*)c.scratch + end - wp_bytes - c.bn_ws_bytes;
  c.img_ws = c.img_ws_bytes ? (char*)c.bn_ws - c.img_ws_bytes : nullptr;
  c.scratch_bytes = end - wp_bytes - c.bn_ws_bytes - c.img_ws_bytes;  // what the convolutions may use
  MM3D_CUDA(cudaMemsetAsync(c.bn_ws, 0, c.bn_ws_bytes, c.stream));
  return MM3D_OK;
}

MM3D_API int mm3d_unet_forward(int in_channels, int m, int num_planes, int mode, int training, float eps, float momentum,
                               const int64_t* level_desc, int64_t n_points, const int32_t* p2v, const int32_t* npts,
                               const float* feats, float* out, void* const* params, void* act, size_t act_bytes,
                               void* scratch, size_t scratch_bytes, const mm3d_unet_mask* mask, mm3d_stream_t stream_) {
  Net net;
  int rc = fill_net(net, in_channels, m, num_planes, mode, level_desc, n_points);
  if (rc) return rc;
  Bump bp{(char*)act, 0, act_bytes, true};
  carve(net, bp);
  MM3D_REQUIRE(bp.ok, MM3D_ERR_WORKSPACE, "activation workspace too small: need %zu have %zu", bp.off, act_bytes);
  Ctx c{&net, params, nullptr, scratch, scratch_bytes, (cudaStream_t)stream_, eps, momentum, training, 0, 0};
  float* const wp_buf = (float*)((char*)scratch + scratch_bytes / 256 * 256 - mm3d_align(4 * (size_t)27 * net.cin_k * m));
  rc = carve_tail(c, net);
  if (rc) return rc;
  const int64_t n0 = net.lv[0].n;
  MARK("fwd start", -1);
  if (mask) {  // the RGB-mask prologue of Net3DSeg.forward folded into the point -> voxel scatter
    MM3D_REQUIRE(mask->wb && (n_points == 0 || mask->s), MM3D_ERR_INVALID, "unet forward: incomplete mask descriptor");
    EX(mm3d_input_masked_fwd(feats, p2v, npts, n_points, n0, in_channels, 4, mask->wb, net.V, mask->s, c.stream));
  } else {
    EX(mm3d_input_fwd(feats, p2v, npts, n_points, n0, in_channels, 4, net.V, c.stream));
  }
  MARK("input_fwd", -1);
  const float* w_stem = P(c, 0);
  if (net.Vp != net.V)  // pad the stem input to whole 64-byte pieces; rounded to TF32 (+ lo plane in TF32x3 mode)
    launch_pad_cols(c, net.V, n0, net.cin, net.Vp, net.cin_k, net.rmode());
  else if (tc_mode(c) && n0 > 0 && !c.rc) {
    // (stem input already a whole number of 64-byte pieces: round the InputLayer output in place)
    if (mm3d_launch_pdl(k_round_tf32, dim3(mm3d_grid(n0 * net.cin / 4 + 1, 256)), dim3(256), 0, c.stream, (const float*)net.V, net.V, n0 * (int64_t)net.cin) != cudaSuccess) c.rc = MM3D_ERR_CUDA;
    mm3d_count_launches(1);
  }
  if (net.cin_k != net.cin) {  // zero input channels for the stem weight too
    float* wp = wp_buf;
    launch_pad_cols(c, w_stem, 27, net.cin * m, wp, net.cin_k * m);
    w_stem = wp;
  }
  EX(build_images(c, false, w_stem));
  MARK("pad + weight images", -1);
  conv_fwd(c, SMC, 0, net.Vp, net.cin_k, net.b[0].X, m, w_stem);
  MARK("stem fwd", -1);
  level_fwd(c, 0, 1);
  const int head = 1 + level_slots(0, net.L);
  bn_fwd(c, head, net.b[0].R, net.Z, n0, m, net.s_head, nullptr, 0, /*to_conv=*/false);
  MARK("bn_fwd head", -1);
  EX(mm3d_output_fwd(net.Z, p2v, n_points, m, out, c.stream));
  MARK("output_fwd", -1);
  return c.rc;
}

// grads: same slot order as params (running-stat slots ignored); d_feats may be NULL.
MM3D_API int mm3d_unet_backward(int in_channels, int m, int num_planes, int mode, int training,
                                const int64_t* level_desc, int64_t n_points, const int32_t* p2v, const int32_t* npts,
                                const float* d_out, float* d_feats, void* const* params, void* const* grads,
                                void* act, size_t act_bytes, void* tmp, size_t tmp_bytes, void* scratch,
                                size_t scratch_bytes, const mm3d_unet_mask* mask, mm3d_stream_t stream_) {
  Net net;
  int rc = fill_net(net, in_channels, m, num_planes, mode, level_desc, n_points);
  if (rc) return rc;
  Bump bp{(char*)act, 0, act_bytes, true};
  carve(net, bp);
  MM3D_REQUIRE(bp.ok, MM3D_ERR_WORKSPACE, "activation workspace too small");
  MM3D_REQUIRE(tmp_bytes >= bwd_temp_bytes(net), MM3D_ERR_WORKSPACE, "backward workspace too small");
  Ctx c{&net, params, grads, scratch, scratch_bytes, (cudaStream_t)stream_, 0.f, 0.f, training, 0, 0};
  c.side = getenv("MM3D_NO_SIDE_STREAM") ? nullptr : side_stream();
  float* const wp_buf = (float*)((char*)scratch + scratch_bytes / 256 * 256 - mm3d_align(4 * (size_t)27 * net.cin_k * m));
  rc = carve_tail(c, net);
  if (rc) return rc;
  Bump g{(char*)tmp, 0, tmp_bytes, true};
  const int64_t n0 = net.lv[0].n;
  const int head = 1 + level_slots(0, net.L);
  {  // One memset instead of one per layer when the gradient buffers are one contiguous block (the shipped host
     // code hands out views of one flat tensor): slot sizes follow from the network shape.
    std::vector<ConvInfo> v;
    v.push_back({0, 27, net.cin, net.m, 1});
    list_convs(net, 0, 1, v);
    const char *lo = nullptr, *hi = nullptr;
    size_t bytes = 0;
    bool all = true;
    for (const ConvInfo& x : v) {
      const char* g0 = (const char*)grads[x.pidx];
      if (!g0) { all = false; break; }
      const size_t nb_ = sizeof(float) * (size_t)x.K * x.c_in * x.c_out;
      if (!lo || g0 < lo) lo = g0;
      if (!hi || g0 + nb_ > hi) hi = g0 + nb_;
      bytes += nb_;
    }
    // BatchNorm gradients (2 * channels each) may sit in between: allow the span to be at most their size larger
    size_t bn_bytes = 0;
    for (int l = 0; l < net.L; ++l) {
      const size_t pl = (size_t)net.planes(l);
      bn_bytes += sizeof(float) * 2 * (l + 1 < net.L ? 4 * pl + (size_t)net.planes(l + 1) : pl);  // pre, dn, up, post
    }
    bn_bytes += sizeof(float) * 2 * (size_t)net.m;  // head
    if (all && lo && (size_t)(hi - lo) <= bytes + bn_bytes) {
      MM3D_CUDA(cudaMemsetAsync((void*)lo, 0, (size_t)(hi - lo), c.stream));
      c.grad_lo = lo; c.grad_hi = hi; c.wgrad_acc = 1;
    }
  }
  {  // stem weight as the kernels see it (zero-padded input channels), then every layer's dgrad weight image
    const float* w_stem_k = P(c, 0);
    if (net.cin_k != net.cin) {
      launch_pad_cols(c, w_stem_k, 27, net.cin * m, wp_buf, net.cin_k * m);
      w_stem_k = wp_buf;
    }
    EX(build_images(c, true, w_stem_k));
  }
  MARK("bwd start (memset, images)", -1);
  float* d_Z = g.f(n0, m);
  EX(mm3d_output_bwd(d_out, p2v, n_points, n0, m, d_Z, c.stream));
  MARK("output_bwd", -1);
  float* d_R0 = g.f(n0, (int64_t)m * net.pl());
  bn_bwd(c, head, net.b[0].R, d_Z, d_R0, n0, m, net.s_head, 1);
  MARK("bn_bwd head", -1);
  float* d_X0 = g.f(n0, (int64_t)m * net.pl());
  level_bwd(c, g, 0, 1, d_R0, d_X0);
  // stem
  const float* w_stem = P(c, 0);
  float* d_w = Gp(c, 0);
  if (mask)
    MM3D_REQUIRE(mask->wb && mask->d_wb && mask->ws && (n_points == 0 || (mask->s && mask->feats)), MM3D_ERR_INVALID,
                 "unet backward: incomplete mask descriptor");
  const bool need_dv = d_feats || mask;  // (the mask parameters get their gradient through the InputLayer)
  float* d_Vp = need_dv ? g.f(n0, net.cin_k) : nullptr;
  auto input_bwd = [&](const float* d_v) {
    if (mask)
      EX(mm3d_input_masked_bwd(d_v, mask->feats, mask->s, p2v, npts, n_points, net.cin, 4, mask->wb, d_feats, mask->d_wb,
                               mask->ws, mask->ws_bytes, c.stream));
    else
      EX(mm3d_input_bwd(d_v, p2v, npts, n_points, net.cin, 4, d_feats, c.stream));
  };
  if (net.cin_k != net.cin) {
    float* wp = wp_buf;  // padded at the start of this call
    float* d_wp = d_w ? g.f(27, (int64_t)net.cin_k * m) : nullptr;  // (frozen stem weight: no gradient)
    conv_bwd(c, SMC, 0, net.Vp, net.cin_k, d_X0, m, wp, d_Vp, d_wp);
    join_side(c);  // d_wp comes from the side stream
    if (d_w) launch_pad_cols(c, d_wp, 27, net.cin_k * m, d_w, net.cin * m);  // slice the real channels back out
    if (need_dv) {
      float* d_V = g.f(n0, net.cin);
      launch_pad_cols(c, d_Vp, n0, net.cin_k, d_V, net.cin);
      input_bwd(d_V);
    }
  } else {
    conv_bwd(c, SMC, 0, net.Vp, net.cin, d_X0, m, w_stem, d_Vp, d_w);
    if (need_dv) input_bwd(d_Vp);
  }
  MARK("stem bwd + input_bwd", -1);
  join_side(c);  // everything after this call on `stream` sees the weight gradients
  MARK("join wgrad stream", -1);
  MM3D_REQUIRE(g.ok, MM3D_ERR_WORKSPACE, "backward workspace overflow");
  return c.rc;
}

// out[i] = in[i] rounded to TF32 (round-to-nearest, ties away); in == out allowed.  The module-by-module path uses
// it on the operands of a TF32 convolution; the executor's producers round in place of it.
MM3D_API int mm3d_round_tf32(const float* in, float* out, int64_t n, mm3d_stream_t stream_) {
  MM3D_REQUIRE(n >= 0 && (n == 0 || (in && out)), MM3D_ERR_INVALID, "round_tf32: bad arguments");
  MM3D_REQUIRE((((uintptr_t)in | (uintptr_t)out) & 15) == 0, MM3D_ERR_INVALID, "round_tf32: pointers must be 16-byte aligned");
  if (n == 0) return MM3D_OK;
  MM3D_CUDA(mm3d_launch_pdl(k_round_tf32, dim3(mm3d_grid(n / 4 + 1, 256)), dim3(256), 0, (cudaStream_t)stream_, in, out, n));
  mm3d_count_launches(1);
  MM3D_CHECK_LAUNCH("mm3d_round_tf32");
  return MM3D_OK;
}

#ifdef MM3D_TRACE
MM3D_API void mm3d_debug_marks(int on) {
  for (Mark& m : g_marks) cudaEventDestroy(m.ev);
  g_marks.clear();
  g_marks_on = on != 0;
}
MM3D_API void mm3d_debug_dump_marks(void) {
  cudaDeviceSynchronize();
  float total = 0.f;
  for (size_t i = 1; i < g_marks.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_marks[i - 1].ev, g_marks[i].ev);
    total += ms;
    fprintf(stderr, "MARK %-28s L%-2d %8.1f us   (t = %8.1f)\n", g_marks[i].name, g_marks[i].level, ms * 1e3f, total * 1e3f);
  }
}
#endif

// out[0..n) = tf32(in) as float32; behind it, at float index n, bf16(in) as n BF16 elements: the operand planes of
// MM3D_MODE_BF16 (`out` holds 2 n floats like the TF32x3 planes; the BF16 plane fills the first half of the second one)
MM3D_API int mm3d_split_bf16(const float* in, float* out, int64_t n, mm3d_stream_t stream_) {
  MM3D_REQUIRE(n >= 0 && (n == 0 || (in && out)), MM3D_ERR_INVALID, "split_bf16: bad arguments");
  if (n == 0) return MM3D_OK;
  MM3D_CUDA(mm3d_launch_pdl(k_pad_cols, dim3(mm3d_grid(n, 256)), dim3(256), 0, (cudaStream_t)stream_, in, n, 1, out, 1, 3));
  mm3d_count_launches(1);
  MM3D_CHECK_LAUNCH("mm3d_split_bf16");
  return MM3D_OK;
}

// out[0..n) = tf32(in), out[n..2n) = tf32(in - tf32(in)): the operand planes of MM3D_MODE_TF32X3
MM3D_API int mm3d_split_tf32(const float* in, float* out, int64_t n, mm3d_stream_t stream_) {
  MM3D_REQUIRE(n >= 0 && (n == 0 || (in && out)), MM3D_ERR_INVALID, "split_tf32: bad arguments");
  if (n == 0) return MM3D_OK;
  // (k_pad_cols with equal widths: a copy that writes the rounded value and the lo plane `n` floats behind it)
  MM3D_CUDA(mm3d_launch_pdl(k_pad_cols, dim3(mm3d_grid(n, 256)), dim3(256), 0, (cudaStream_t)stream_, in, n, 1, out, 1, 2));
  mm3d_count_launches(1);
  MM3D_CHECK_LAUNCH("mm3d_split_tf32");
  return MM3D_OK;
}

}  // extern "C"
