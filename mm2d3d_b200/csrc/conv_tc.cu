// tcgen05 / TMEM rule-table convolution (forward and dgrad) for sm_100a, driven by a row plan.
//
//   out[j,:] = sum_k in[tbl(j,k),:] . W'[k]          j in a tile of 128 output rows (plan.cuh)
//
// Output-stationary implicit GEMM with block-sparse K.  A CTA owns tiles of 128 output rows in
// plan order; the reduction runs over the tile's NON-EMPTY offsets only (tile_mask), each offset
// cut into K-blocks of 32 input channels (= one 128-byte shared-memory row; the last block of an
// offset may be 16 channels wide).  For every K-block the 128 gathered rows (zero where the
// neighbour is absent) form a K-major SWIZZLE_128B tile that tcgen05.mma (M=128, N=C_out,
// kind::tf32) multiplies with the matching block of the pre-swizzled weight image, accumulating all
// offsets of the tile in one TMEM accumulator.  No atomics, no read-modify-write of `out`, each
// output row stored once; the result does not depend on which rows share a tile.
//
// Warp roles ((S+5) warps):   warps 0..S-1    gather producers, warp w OWNS ring stage w: 16-byte
//                                             cp.async per (row, chunk), zero fill for absent rows,
//                                             plus the bulk copy of the K-block's weight block
//                             warps S..S+3    epilogue (TMEM lanes -> registers -> out[perm[row]])
//                             warp  S+4       MMA issuer (one lane) + TMEM allocator
// Pipelines (mbarriers):      ring     a_full[S] (32 cp.async arrivals + expect_tx of the weight block)
//                                      / a_empty[S] (tcgen05.commit)
//                             accum    acc_full[2] (tcgen05.commit)       / acc_empty[2] (128 arrivals)
// The accumulator is double buffered in TMEM, so the epilogue of tile i overlaps the MMAs of i+1.
#include <stdlib.h>

#include "plan.cuh"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kTileM = 128;
constexpr int kKBlock = 32;                 // tf32 elements per 128-byte row of a K-block
constexpr int kStageBytes = kTileM * 128;   // one A stage = one K-block of 128 rows = 16 KB
constexpr int kEntBytes = kTileM * 4;       // the K-block's 128 table entries
constexpr int kEpilogue = 128;
constexpr int kMaxStages = 8;
constexpr int kMaxThreads = (kMaxStages + 5) * 32;

struct TcParams {
  const float* in;
  float* out;
  const float* wimg;  // [K][nb][n_pad][32] tf32, rows 128-byte swizzled
  const int32_t* perm;
  const uint32_t* tile_mask;
  const int32_t* order;
  const int32_t* tbl;
  int64_t tstride;
  int c_in, c_out, K;
  int nb;       // K-blocks per offset = ceil(c_in / 32)
  int last_w;   // 16-byte chunks of an offset's last block (8, or 4 when c_in % 32 == 16)
  int S;        // ring stages = producer warps
  int n_pad, num_tiles, tmem_cols;
  int n_local;  // tiles per CTA (upper bound) = ceil(num_tiles / grid)
  int* err;
};

// Weight image: one [n_pad][32] block per (offset, channel block); element (n, e) of block (k, j) is
// Wsel(k)[32 j + e][n] (0 beyond c_in / c_out), tf32-rounded, the 16-byte chunks of every 128-byte
// row XOR-swizzled with (n & 7): exactly the bytes a SWIZZLE_128B K-major B tile has in shared
// memory, so the kernel bulk-copies it verbatim.
__global__ void k_weight_image(const float* __restrict__ w, float* __restrict__ img, int K, int c_in, int c_out, int nb,
                               int n_pad, int transposed, int mirror) {
  const int64_t total = (int64_t)K * nb * n_pad * kKBlock;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int e_sw = (int)(i % kKBlock);
    const int n = (int)((i / kKBlock) % n_pad);
    const int blk = (int)(i / ((int64_t)kKBlock * n_pad));
    const int k = blk / nb, j = blk - k * nb;
    const int chunk = (e_sw >> 2) ^ (n & 7);  // un-swizzle: which logical chunk lives here
    const int ci = j * kKBlock + chunk * 4 + (e_sw & 3);
    float v = 0.f;
    if (ci < c_in && n < c_out) {
      const int ks = mirror ? K - 1 - k : k;
      v = transposed ? __ldg(w + ((int64_t)ks * c_out + n) * c_in + ci)   // forward weight is [K][c_out][c_in]
                     : __ldg(w + ((int64_t)ks * c_in + ci) * c_out + n);  // [K][c_in][c_out]
    }
    img[i] = to_tf32(v);
  }
}

// Several weight images in one launch (the whole network's, built once per direction by the executor).
struct WImgItem {
  const float* w;
  float* img;
  int K, c_in, c_out, nb, n_pad, transposed, mirror, block0;
};
constexpr int kMaxImgBatch = 32;
struct WImgBatch {
  int n;
  WImgItem item[kMaxImgBatch];
};
__global__ void k_weight_images(const __grid_constant__ WImgBatch b) {
  mm3d_griddep_wait();
  int i = 0;
  while (i + 1 < b.n && (int)blockIdx.x >= b.item[i + 1].block0) ++i;
  const WImgItem& it = b.item[i];
  const int64_t total = (int64_t)it.K * it.nb * it.n_pad * kKBlock;
  const int nblk = (i + 1 < b.n ? b.item[i + 1].block0 : (int)gridDim.x) - it.block0;
  for (int64_t e = (int64_t)((int)blockIdx.x - it.block0) * blockDim.x + threadIdx.x; e < total; e += (int64_t)nblk * blockDim.x) {
    const int e_sw = (int)(e % kKBlock);
    const int n = (int)((e / kKBlock) % it.n_pad);
    const int blk = (int)(e / ((int64_t)kKBlock * it.n_pad));
    const int k = blk / it.nb, j = blk - k * it.nb;
    const int chunk = (e_sw >> 2) ^ (n & 7);
    const int ci = j * kKBlock + chunk * 4 + (e_sw & 3);
    float v = 0.f;
    if (ci < it.c_in && n < it.c_out) {
      const int ks = it.mirror ? it.K - 1 - k : k;
      v = it.transposed ? __ldg(it.w + ((int64_t)ks * it.c_out + n) * it.c_in + ci)
                        : __ldg(it.w + ((int64_t)ks * it.c_in + ci) * it.c_out + n);
    }
    it.img[e] = to_tf32(v);
  }
}

// Walks the CTA's items in order: its tiles (dealt from the plan's cost-ordered list, kept in shared
// memory with their masks); per tile the set bits of its mask (ascending offset); per offset the nb
// channel blocks.  Every role runs its own copy.
struct ItemWalk {
  const uint32_t* lmask;  // shared memory: [n_local] masks, then [n_local] tile indices
  int n_local, nb;
  int lt, tile, k, j;
  uint32_t rem;
  __device__ __forceinline__ void init(const TcParams& p, const uint32_t* local_masks, int n_loc) {
    lmask = local_masks; n_local = n_loc; nb = p.nb;
    lt = 0; tile = 0; j = 0; rem = 0; k = 0;
    if (lt < n_local) { rem = lmask[0]; tile = (int)lmask[n_local]; k = __ffs(rem) - 1; }
  }
  __device__ __forceinline__ bool valid() const { return lt < n_local; }
  __device__ __forceinline__ void next() {
    if (++j < nb) return;
    j = 0;
    rem &= rem - 1;
    if (rem) { k = __ffs(rem) - 1; return; }
    ++lt;
    if (lt < n_local) { rem = lmask[lt]; tile = (int)lmask[n_local + lt]; k = __ffs(rem) - 1; }
  }
  // true when the current item is the last of its tile
  __device__ __forceinline__ bool last_of_tile() const { return j == nb - 1 && (rem & (rem - 1)) == 0; }
};

__global__ void __launch_bounds__(kMaxThreads, 1)
k_conv_tc(const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int S = p.S;
  // carve: [A stages][B stages][entry rows][barriers][tmem ptr][abort]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t b_bytes = (uint32_t)p.n_pad * 128u;
  const uint32_t b_stride = (b_bytes + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + (uint32_t)S * kStageBytes;
  const uint32_t e_base = b_base + (uint32_t)S * b_stride;  // one weight block and one entry row per A stage
  const uint32_t m_base = e_base + (uint32_t)S * kEntBytes;  // masks of this CTA's tiles
  const uint32_t bar_base = m_base + (((uint32_t)p.n_local * 8u + 15u) & ~15u);
  uint32_t* lmask = reinterpret_cast<uint32_t*>(smem + (m_base - smem_base));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (bar_base - smem_base));
  auto a_full = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (uint32_t)(kMaxStages + s); };
  auto acc_full = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kMaxStages + s); };
  auto acc_empty = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kMaxStages + 2 + s); };
  constexpr int kNumBars = 2 * kMaxStages + 4;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + kNumBars);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(bars + kNumBars + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup.  Nothing before mm3d_griddep_wait() touches global memory (the previous kernel of the
  // stream may still be running: programmatic dependent launch).
  mm3d_griddep_launch();
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(a_full(s), 33);  // 32 cp.async arrivals (gathered rows) + 1 expect_tx arrival (weight block)
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), kEpilogue);
    }
    *abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == S + 4) tmem_alloc(smem_u32((const void*)tmem_slot), (uint32_t)p.tmem_cols);
  mm3d_griddep_wait();
  // this CTA's tiles (mm3d_plan_local_tile) and their masks live in shared memory: [masks][tile indices]
  int n_local = p.n_local;
  if (mm3d_plan_local_tile(p.order, p.num_tiles, (int)gridDim.x, (int)blockIdx.x, n_local - 1) < 0) --n_local;
  for (int i = threadIdx.x; i < n_local; i += blockDim.x) {
    const int t = mm3d_plan_local_tile(p.order, p.num_tiles, (int)gridDim.x, (int)blockIdx.x, i);
    lmask[i] = __ldg(p.tile_mask + t);
    lmask[n_local + i] = (uint32_t)t;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < S) {
    // =================================================================== gather producer: owns stage `warp`
    const uint32_t stage = a_base + (uint32_t)warp * kStageBytes;
    const uint32_t ent = e_base + (uint32_t)warp * kEntBytes;
    ItemWalk it;
    it.init(p, lmask, n_local);
    for (int i = 0; i < warp && it.valid(); ++i) it.next();  // this warp takes every S-th item
    uint32_t round = 0;
    int4 e = make_int4(-1, -1, -1, -1);  // entries of rows 4*lane .. 4*lane+3 of the current item
    if (it.valid()) e = __ldg(reinterpret_cast<const int4*>(p.tbl + (int64_t)it.k * p.tstride + (int64_t)it.tile * kTileM) + lane);
    while (it.valid()) {
      const int k = it.k, j = it.j;
      for (int i = 0; i < S && it.valid(); ++i) it.next();
      int4 en = make_int4(-1, -1, -1, -1);  // the next item's entries are in flight while this one is copied
      if (it.valid()) en = __ldg(reinterpret_cast<const int4*>(p.tbl + (int64_t)it.k * p.tstride + (int64_t)it.tile * kTileM) + lane);
      if (!mbar_wait(a_empty(warp), (round & 1u) ^ 1u, abort_flag)) goto done;
      if (lane == 0) {  // this K-block's weight block rides on the same barrier as the gathered rows
        mbar_arrive_expect_tx(a_full(warp), b_bytes);
        bulk_g2s(b_base + (uint32_t)warp * b_stride, p.wimg + ((size_t)k * p.nb + j) * p.n_pad * kKBlock, b_bytes, a_full(warp));
      }
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ent + (uint32_t)lane * 16u), "r"(e.x), "r"(e.y), "r"(e.z), "r"(e.w) : "memory");
      __syncwarp();
      const float* src0 = p.in + j * kKBlock;
      const bool half = (j == p.nb - 1) && p.last_w == 4;
      if (!half) gather_block<8, false>(stage, ent, p.in, src0, (uint32_t)p.c_in, lane);
      else       gather_block<4, false>(stage, ent, p.in, src0, (uint32_t)p.c_in, lane);
      cp_async_arrive(a_full(warp));
      __syncwarp();  // the entry row is rewritten by the next item
      e = en;
      ++round;
    }
  } else if (warp < S + 4) {
    // =================================================================== epilogue
    const int ew = warp & 3;  // a warp may only touch TMEM lanes 32*(warp%4) .. +31
    for (uint32_t tile_iter = 0; tile_iter < (uint32_t)n_local; ++tile_iter) {
      const int ab = (int)(tile_iter & 1u);
      const int tile = (int)lmask[n_local + tile_iter];
      const int row = __ldg(p.perm + (int64_t)tile * kTileM + ew * 32 + lane);
      if (!mbar_wait_sleep(acc_full(ab), (tile_iter >> 1) & 1u, abort_flag, 100)) goto done;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(ab * p.n_pad);
      for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
        float acc[16];
        tmem_ld16(taddr + (uint32_t)c0, acc);
        if (row >= 0) {
          float* dst = p.out + (int64_t)row * p.c_out + c0;
          if ((p.c_out & 3) == 0) {
#pragma unroll
            for (int jj = 0; jj < 16; jj += 4)
              if (c0 + jj < p.c_out) *reinterpret_cast<float4*>(dst + jj) = make_float4(acc[jj], acc[jj + 1], acc[jj + 2], acc[jj + 3]);
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj)
              if (c0 + jj < p.c_out) dst[jj] = acc[jj];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty(ab));
    }
  } else if (warp == S + 4) {
    // =================================================================== MMA issuer: the whole warp walks the
    // items (warp-uniform control flow), one elected lane issues
    {
      const uint32_t idesc = make_idesc_tf32(kTileM, p.n_pad);
      const uint64_t desc0 = make_desc_sw128(0);
      ItemWalk it;
      it.init(p, lmask, n_local);
      uint32_t tile_iter = 0;
      int s = 0;
      uint32_t ph = 0;
      while (it.valid()) {
        const int ab = (int)(tile_iter & 1u);
        if (!mbar_wait(acc_empty(ab), ((tile_iter >> 1) & 1u) ^ 1u, abort_flag)) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * p.n_pad);
        bool first = true, ok = true;
        while (true) {
          const bool last = it.last_of_tile();
          const int ksteps = (it.j == p.nb - 1 ? p.last_w : 8) >> 1;
          if (!mbar_wait(a_full(s), ph, abort_flag)) { ok = false; break; }
          tc_fence_after();
          if (elect_one()) {
            const uint64_t a_desc = desc0 + desc_addr(a_base + (uint32_t)s * kStageBytes);
            const uint64_t b_desc = desc0 + desc_addr(b_base + (uint32_t)s * b_stride);
            umma_tf32(d_tmem, a_desc, b_desc, idesc, first ? 0u : 1u);
            umma_tf32(d_tmem, a_desc + 2, b_desc + 2, idesc, 1u);  // +32 bytes per K-step of 8
            if (ksteps == 4) {
              umma_tf32(d_tmem, a_desc + 4, b_desc + 4, idesc, 1u);
              umma_tf32(d_tmem, a_desc + 6, b_desc + 6, idesc, 1u);
            }
            umma_commit(a_empty(s));  // frees the stage and its weight block
            if (last) umma_commit(acc_full(ab));
          }
          __syncwarp();
          first = false;
          if (++s == S) { s = 0; ph ^= 1u; }
          it.next();
          if (last) break;
        }
        if (!ok) break;
        ++tile_iter;
      }
    }
  }
done:
  asm volatile("cp.async.wait_all;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && *abort_flag) mm3d_raise(p.err);
  if (warp == S + 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

}  // namespace

// ---- host entry points used by capi.cu -------------------------------------------------------

static int tc_geometry(int c_in, int c_out, int K, int* nb, int* n_pad) {
  *nb = (c_in + 31) / 32;
  *n_pad = (c_out + 15) / 16 * 16;
  // input rows are cut into 128-byte blocks with an optional 64-byte tail
  return (*n_pad <= 256 && (c_in % 16) == 0 && K <= 32) ? 0 : 1;
}

// 1 when the tcgen05 kernel handles this shape
int mm3d_conv_tc_supported(int c_in, int c_out, int K) {
  int nb, n_pad;
  return tc_geometry(c_in, c_out, K, &nb, &n_pad) == 0;
}

size_t mm3d_conv_tc_workspace_bytes(int c_in, int c_out, int K) {
  int nb, n_pad;
  if (tc_geometry(c_in, c_out, K, &nb, &n_pad)) return 0;
  return mm3d_align((size_t)K * nb * n_pad * 128);  // weight image
}

// Sticky per-device error words (mm3d_raise, common.cuh) in MAPPED PINNED host memory: a kernel whose bounded wait
// timed out (never in a correct build; it turns a would-be GPU hang into a reportable error) or that met an
// out-of-image lift index stores 1 there, and the host reads the words without any CUDA call or synchronisation --
// the executor polls them at every forward / backward call.
static int* g_err_host[64] = {nullptr};
static int* g_err_dev[64] = {nullptr};
int* mm3d_device_err_flag() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!g_err_dev[dev]) {
    int* h = nullptr;
    if (cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    for (int i = 0; i < 16; ++i) h[i] = 0;
    int* d = nullptr;
    if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) { (void)cudaGetLastError(); cudaFreeHost(h); return nullptr; }
    g_err_host[dev] = h;
    g_err_dev[dev] = d;
  }
  return g_err_dev[dev];
}

static int g_sm_count[64] = {0};
int mm3d_sm_count() {
  const int dev = mm3d_device_slot();
  if (!g_sm_count[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = MM3D_NUM_SMS; }
    g_sm_count[dev] = n;
  }
  return g_sm_count[dev];
}

// Bits: 1 = a kernel pipeline / grid barrier timed out, 2 = a lift index outside the image.  Reads and clears the
// current device's words; no synchronisation (callers that want everything enqueued so far synchronise first).
extern "C" int mm3d_take_device_error(void) {
  if (!mm3d_device_err_flag()) return -1;
  volatile int* h = g_err_host[mm3d_device_slot()];
  int v = 0;
  if (h[0]) { v |= 1; h[0] = 0; }
  if (h[1]) { v |= 2; h[1] = 0; }
  return v;
}

// images of several layers in one launch; descs[i] = {weight, image buffer (mm3d_conv_tc_workspace_bytes), K, c_in,
// c_out, flags}
int mm3d_conv_tc_build_images(const float* const* weights, float* const* images, const int* K, const int* c_in,
                              const int* c_out, const int* flags, int n, cudaStream_t stream) {
  for (int first = 0; first < n; first += kMaxImgBatch) {
    WImgBatch b;
    b.n = 0;
    int blocks = 0;
    for (int i = first; i < n && i < first + kMaxImgBatch; ++i) {
      WImgItem& it = b.item[b.n++];
      MM3D_REQUIRE(tc_geometry(c_in[i], c_out[i], K[i], &it.nb, &it.n_pad) == 0, MM3D_ERR_UNSUPPORTED,
                   "tcgen05 conv: unsupported shape c_in %d c_out %d K %d", c_in[i], c_out[i], K[i]);
      it.w = weights[i]; it.img = images[i]; it.K = K[i]; it.c_in = c_in[i]; it.c_out = c_out[i];
      it.transposed = (flags[i] & MM3D_CONV_TRANSPOSE_W) ? 1 : 0;
      it.mirror = (flags[i] & MM3D_CONV_MIRROR_K) ? 1 : 0;
      it.block0 = blocks;
      int nb_blocks = (int)mm3d_cdiv((int64_t)it.K * it.nb * it.n_pad * kKBlock, 256 * 8);
      blocks += nb_blocks < 1 ? 1 : nb_blocks;
    }
    if (blocks == 0) continue;
    MM3D_CUDA(mm3d_launch_pdl(k_weight_images, dim3(blocks), dim3(256), 0, stream, b));
    mm3d_count_launches(1);
    MM3D_CHECK_LAUNCH("mm3d_conv_tc_build_images");
  }
  return MM3D_OK;
}

int mm3d_conv_fwd_tc_img(const float* in, int64_t n_in, int c_in, float* out, int64_t n_out, int c_out,
                         const float* wimg, int K, const void* plan, int64_t plan_cap, cudaStream_t stream);

int mm3d_conv_fwd_tc(const float* in, int64_t n_in, int c_in, float* out, int64_t n_out, int c_out,
                     const float* weight, int K, const void* plan, int64_t plan_cap, int flags, void* ws,
                     size_t ws_bytes, cudaStream_t stream) {
  int nb, n_pad;
  MM3D_REQUIRE(tc_geometry(c_in, c_out, K, &nb, &n_pad) == 0, MM3D_ERR_UNSUPPORTED,
               "tcgen05 conv: unsupported shape c_in %d c_out %d K %d", c_in, c_out, K);
  MM3D_REQUIRE(ws && ws_bytes >= mm3d_conv_tc_workspace_bytes(c_in, c_out, K), MM3D_ERR_WORKSPACE,
               "tcgen05 conv: workspace too small");
  const bool tr = (flags & MM3D_CONV_TRANSPOSE_W) != 0, mir = (flags & MM3D_CONV_MIRROR_K) != 0;
  MM3D_REQUIRE(tr || !mir, MM3D_ERR_UNSUPPORTED, "MIRROR_K without TRANSPOSE_W not implemented");
  if (n_out == 0) return MM3D_OK;
  float* wimg = (float*)ws;
  k_weight_image<<<mm3d_grid((int64_t)K * nb * n_pad * kKBlock, 256), 256, 0, stream>>>(weight, wimg, K, c_in, c_out, nb,
                                                                                    n_pad, tr ? 1 : 0, mir ? 1 : 0);
  mm3d_count_launches(1);
  return mm3d_conv_fwd_tc_img(in, n_in, c_in, out, n_out, c_out, wimg, K, plan, plan_cap, stream);
}

// the convolution proper, from a prebuilt weight image
int mm3d_conv_fwd_tc_img(const float* in, int64_t n_in, int c_in, float* out, int64_t n_out, int c_out,
                         const float* wimg, int K, const void* plan, int64_t plan_cap, cudaStream_t stream) {
  int nb, n_pad;
  MM3D_REQUIRE(tc_geometry(c_in, c_out, K, &nb, &n_pad) == 0, MM3D_ERR_UNSUPPORTED,
               "tcgen05 conv: unsupported shape c_in %d c_out %d K %d", c_in, c_out, K);
  MM3D_REQUIRE(n_out < (1ll << 31) && n_in * (int64_t)c_in < (1ll << 32), MM3D_ERR_UNSUPPORTED,
               "tcgen05 conv: tensor too large for 32-bit element offsets");
  MM3D_REQUIRE(plan && plan_cap >= n_out, MM3D_ERR_INVALID, "tcgen05 conv: needs a row plan covering n_out rows");
  MM3D_REQUIRE((((uintptr_t)in | (uintptr_t)out | (uintptr_t)wimg) & 15) == 0, MM3D_ERR_INVALID,
               "tcgen05 conv: pointers must be 16-byte aligned");
  if (n_out == 0) return MM3D_OK;
  const Mm3dPlanView pv = mm3d_plan_view(plan, plan_cap);
  TcParams p;
  p.in = in; p.out = out; p.wimg = wimg;
  p.perm = pv.perm; p.tile_mask = pv.tile_mask; p.order = pv.order; p.tbl = pv.tbl; p.tstride = pv.stride;
  p.c_in = c_in; p.c_out = c_out; p.K = K; p.nb = nb;
  p.last_w = (c_in % 32) == 16 ? 4 : 8;
  p.n_pad = n_pad;
  p.num_tiles = (int)mm3d_cdiv(n_out, kTileM);
  p.tmem_cols = pow2_cols(2 * n_pad);
  p.err = mm3d_device_err_flag();
  const uint32_t b_stride = ((uint32_t)n_pad * 128u + 1023u) & ~1023u;
  const size_t per_stage = (size_t)kStageBytes + b_stride + kEntBytes;
  // as many stages as give two CTAs per SM; one CTA per SM (deeper ring) when the weight blocks are
  // large or there are no more tiles than SMs anyway
  const int two = (int)((112 * 1024) / per_stage), one = (int)((222 * 1024) / per_stage);
  int S = (p.num_tiles > MM3D_NUM_SMS && two >= 3) ? two : one;
  if (S > kMaxStages) S = kMaxStages;
  p.S = S;
  size_t smem = 1024 + (size_t)S * per_stage + 8 * (2 * kMaxStages + 4) + 64;
  static int regs_dev[64] = {0};
  int& regs = regs_dev[mm3d_device_slot()];
  if (!regs) {
    MM3D_CUDA(cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    MM3D_CUDA(cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    cudaFuncAttributes fa;
    MM3D_CUDA(cudaFuncGetAttributes(&fa, k_conv_tc));
    regs = fa.numRegs > 0 ? fa.numRegs : 64;
  }
  // persistent CTAs: as many as fit (registers, shared memory, threads, TMEM columns); tiles round-robin
  const int threads = (S + 5) * 32;
  const int regs_alloc = (regs + 7) / 8 * 8;
  int per_sm = 65536 / (regs_alloc * threads);
  const int by_smem = (int)((227 * 1024) / (smem + 1024));
  const int by_tmem = 512 / p.tmem_cols;
  if (per_sm > by_smem) per_sm = by_smem;
  if (per_sm > 2048 / threads) per_sm = 2048 / threads;
  if (per_sm > by_tmem) per_sm = by_tmem;
  if (per_sm < 1) per_sm = 1;
  int grid = MM3D_NUM_SMS * per_sm;
  if (grid > p.num_tiles) grid = p.num_tiles;
  p.n_local = (p.num_tiles + grid - 1) / grid;
  smem += ((size_t)p.n_local * 8 + 15) / 16 * 16;
  MM3D_REQUIRE(smem <= 226 * 1024, MM3D_ERR_UNSUPPORTED, "tcgen05 conv: too many rows per CTA for the tile-mask cache");
  MM3D_CUDA(mm3d_launch_pdl(k_conv_tc, dim3(grid), dim3(threads), smem, stream, p));
  mm3d_count_launches(1);
  MM3D_CHECK_LAUNCH("mm3d_conv_fwd_tc");
  return MM3D_OK;
}
