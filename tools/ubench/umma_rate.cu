// Microbenchmark: issue rate of tcgen05.mma flavours on sm_100a (operands = whatever is in shared memory).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../include -I ../../mm2d3d_b200/csrc umma_rate.cu -o umma_rate
#include <cstdio>
#include <vector>
#include "tc_common.cuh"
using namespace tc;

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
               "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}

// mode 0: tf32 K-major SW128 (fwd kernel); 1: tf32 MN-major SW128_BASE32 (wgrad kernel); 2: bf16 MN-major SW128; 3: bf16 K-major
template <int VARIANT>
__global__ void k(int mode, int M, int N, int reps, int per_commit, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) reinterpret_cast<float*>(smem_raw)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&slot), 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (VARIANT == 0 ? threadIdx.x == 0 : threadIdx.x < 32) {
    uint32_t idesc;
    uint64_t a, b;
    if (mode == 0) { idesc = make_idesc_tf32(M, N, 0, 0); a = make_desc_sw128(base); b = make_desc_sw128(base + 65536); }
    else if (mode == 1) { idesc = make_idesc_tf32(M, N, 1, 1); a = make_desc_sw128_base32(base, 16384, 512); b = make_desc_sw128_base32(base + 65536, 16384, 512); }
    else {
      // kind::f16 with bf16 operands: D=F32 (1<<4), A=B=BF16 (1<<7 | 1<<10)
      const uint32_t mn = mode == 2 ? 1u : 0u;
      idesc = (1u << 4) | (1u << 7) | (1u << 10) | (mn << 15) | (mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      a = mode == 2 ? make_desc_sw128(base, 8192, 1024) : make_desc_sw128(base);
      b = mode == 2 ? make_desc_sw128(base + 65536, 8192, 1024) : make_desc_sw128(base + 65536);
    }
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (VARIANT == 0) {
        for (int i = 0; i < per_commit; ++i) {
          if (mode >= 2) umma_f16(tm, a + (i & 7) * 2, b + (i & 7) * 2, idesc, 1u);
          else umma_tf32(tm, a + (i & 3) * (mode == 1 ? 64 : 2), b + (i & 3) * (mode == 1 ? 64 : 2), idesc, 1u);
        }
        umma_commit(smem_u32(&bar));
      } else {
        if (elect_one()) {
          for (int i = 0; i < per_commit; ++i) {
            if (mode >= 2) umma_f16(tm, a + (i & 7) * 2, b + (i & 7) * 2, idesc, 1u);
            else umma_tf32(tm, a + (i & 3) * (mode == 1 ? 64 : 2), b + (i & 3) * (mode == 1 ? 64 : 2), idesc, 1u);
          }
          umma_commit(smem_u32(&bar));
        }
        __syncwarp();
      }
      while (!mbar_try_wait(smem_u32(&bar), ph)) {}
      ph ^= 1u;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 256); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8 * 256);
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[] = {"tf32 K-major", "tf32 MN-major(base32)", "bf16 MN-major", "bf16 K-major"};
  for (int variant = 0; variant < 2; ++variant)
  for (int mode = 0; mode < 2; ++mode)
    for (int M : {64, 128})
      for (int N : {16, 128}) {
        for (int per : {4, 16, 64}) {
          const int reps = 200;
          if (variant == 0) k<0><<<1, 128, 197 * 1024>>>(mode, M, N, reps, per, d);
          else k<1><<<1, 128, 197 * 1024>>>(mode, M, N, reps, per, d);
          long long c = 0;
          cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) { printf("%s M%d N%d: %s\n", names[mode], M, N, cudaGetErrorString(e)); return 1; }
          printf("variant %d %-22s M=%3d N=%3d mma/commit=%2d : %7.1f cycles per MMA (incl. commit round trip %.0f per batch)\n", variant, names[mode], M, N, per,
                 (double)c / (reps * per), (double)c / reps);
        }
      }
  return 0;
}
