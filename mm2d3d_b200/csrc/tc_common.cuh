// PTX wrappers shared by the tcgen05 kernels (sm_100a): mbarriers, cp.async, bulk copies, TMEM
// allocation / loads, tcgen05.mma / commit / fences, UMMA descriptors.
#pragma once

#include "common.cuh"

namespace tc {

constexpr uint32_t kSpinLimit = 1u << 22;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a broken pipeline sets *abort (shared) and the kernel's error flag instead of
// hanging the GPU; every other wait sees the abort flag and leaves too.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  for (uint32_t i = 0; i < kSpinLimit; ++i) {
    if (mbar_try_wait(bar, parity)) return true;
    if ((i & 1023) == 1023 && *abort_flag) return false;
  }
  *abort_flag = 1;
  return false;
}
// Same, for waits that are expected to be long (an epilogue waiting for a whole tile or kernel): sleeps between
// polls so that the warp does not take issue slots from the producers of its scheduler.
__device__ __forceinline__ bool mbar_wait_sleep(uint32_t bar, uint32_t parity, volatile int* abort_flag, uint32_t ns) {
  for (uint32_t i = 0; i < kSpinLimit; ++i) {
    if (mbar_try_wait(bar, parity)) return true;
    __nanosleep(ns);
    if ((i & 255) == 255 && *abort_flag) return false;
  }
  *abort_flag = 1;
  return false;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One lane of a converged warp.  tcgen05.mma / commit take their operands from uniform registers: issued under
// elect.sync inside warp-uniform control flow the descriptor arithmetic stays in the uniform datapath, while a
// lone `if (lane == 0)` thread pays a per-instruction register -> uniform-register hand-over (measured: 102 vs
// 45 cycles per tcgen05.mma, tools/ubench/umma_rate.cu).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same with BF16 operands (kind::f16, K = 16 elements = 32 bytes per instruction, FP32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// TMA row gather: 4 rows (row0..row3, any order) x the tensor map's box width starting at column `col` land as 4
// consecutive 128-byte rows at dst (swizzled as the tensor map says); out-of-bounds rows / columns arrive as zeros;
// the mbarrier receives complete_tx for the full 4 x box bytes.
__device__ __forceinline__ void tma_gather4(uint32_t dst_smem, const void* tmap, int col, int row0, int row1, int row2, int row3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
          dst_smem),
      "l"(tmap), "r"(col), "r"(row0), "r"(row1), "r"(row2), "r"(row3), "r"(bar)
      : "memory");
}
// 16-byte asynchronous copy global -> shared; src_bytes < 16 zero-fills the rest (0 = pure zero fill)
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// Same with the copy's `ignore-src` predicate: `ignore` writes 16 zero bytes and reads nothing (src must still be a valid
// address).  ptxas turns the variable src-size form above into ~6 extra instructions per copy (pointer / size
// arithmetic for the partial-copy case); the predicate form is one ISETP.
// .cg (bypass L1, allocate in L2 only): measured on B200, the gathers of these kernels are limited by the rate at
// which the SM's L1 takes row lines (~0.35 lines per cycle with .ca, whatever their useful width); with .cg the
// level-0 forward drops 33 -> 27.6 us and the whole step 3.09 -> 2.97 ms.  A gathered row is used once per CTA, so
// nothing is lost by not caching it in L1.
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, bool ignore) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t"
      "cp.async.cg.shared.global [%0], [%1], 16, p;\n\t}" ::"r"(dst_smem),
      "l"(src), "r"((uint32_t)ignore)
      : "memory");
}
// the mbarrier receives one arrival when all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// Gather one K-block: the 128 table entries sit in shared memory at `ent`; row r of the stage receives chunks
// [0, W) of input row ent[r] starting at `src0` (or zeros when ent[r] < 0).  W = 8: 8 lanes per row, 4 rows per
// pass; W = 4: 4 lanes per row, 8 rows per pass (chunks 4..7 of the stage are then never read).  Entries are
// read eight passes at a time so that their shared-memory latency is paid once per batch, not once per row.
// wvalid < W: only the first wvalid chunks of a row exist (the rest of the stage row is zero-filled).
template <int W, bool BASE32>
__device__ __forceinline__ void gather_block(uint32_t stage, uint32_t ent, const float* __restrict__ base,
                                             const float* __restrict__ src0, uint32_t row_floats, int lane, int wvalid = W) {
  constexpr int kLanesPerRow = W, kRowsPerPass = 32 / W, kBatch = 8;
  const int c = lane & (kLanesPerRow - 1), rsub = lane / kLanesPerRow;
  const float* srcc = src0 + c * 4;
#pragma unroll 1
  for (int r0 = 0; r0 < 128; r0 += kRowsPerPass * kBatch) {
    int rows[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u)
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rows[u]) : "r"(ent + (uint32_t)(r0 + u * kRowsPerPass + rsub) * 4u) : "memory");
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int r = r0 + u * kRowsPerPass + rsub;
      const uint32_t off = BASE32 ? (uint32_t)((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4)) : (uint32_t)((c ^ (r & 7)) << 4);
      // (absent rows read nothing: the address of row 0 only has to be valid)
      cp_async16_zfill(stage + (uint32_t)r * 128u + off, srcc + (size_t)(uint32_t)max(rows[u], 0) * row_floats, rows[u] < 0 || c >= wvalid);
    }
  }
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100), SWIZZLE_128B:
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=SWIZZLE_128B(2) [61,64).
// K-major operand: rows of 128 bytes, SBO = 1024 B between 8-row groups, LBO unused.
// MN-major operand: 128 contiguous bytes of M/N per K index, 8 K indices per 1024-byte atom,
//                   LBO = bytes between 32-element M/N blocks, SBO = bytes between 8-deep K groups.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes = 16, uint32_t sbo_bytes = 1024) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major operands of 32-bit types (tf32) must use SWIZZLE_128B_BASE32B (layout type 1): rows of 128
// bytes whose four 32-byte granules are XOR-ed with (row & 3) (cute Swizzle<2,5,2>), 4 K-rows per atom.
// LBO = bytes between 32-element M/N blocks, SBO = bytes between 4-row K groups.
__device__ __forceinline__ uint64_t make_desc_sw128_base32(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// Descriptors of one operand differ only in the start-address field: build the constant part once and add
// (byte address >> 4); stepping the operand by `bytes` adds (bytes >> 4) (no carry: addresses < 256 KB).
__device__ __forceinline__ uint64_t desc_addr(uint32_t saddr) { return (uint64_t)((saddr & 0x3FFFF) >> 4); }
// byte offset of 16-byte chunk c (0..7) of row r inside a 128-byte row under that swizzle
__device__ __forceinline__ uint32_t swz_base32(int c, int r) {
  return (uint32_t)((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4));
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=B=TF32 [7,10)=[10,13)=2,
// a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc_tf32(int m, int n, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// A = B = BF16 (kind::f16 formats: 0 = F16, 1 = BF16), D = F32
__device__ __forceinline__ uint32_t make_idesc_bf16(int m, int n, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

}  // namespace tc
