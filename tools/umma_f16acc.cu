// Microbenchmark / layout probe: tcgen05.mma kind::f16 with an fp16 ACCUMULATOR (instruction-descriptor c_format = 0).
// One MMA, M = 128, N = 64, K = 16 on known operands (D[m][n] = 0.25 * (m % 7) * n), then TMEM is read back (a) raw,
// 32 bits per column, and (b) with tcgen05.ld ... .pack::16b, to learn how 16-bit accumulators sit in tensor memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I speech-enhancement-via-hybrid-vision-transformer-project_b200/csrc -o tools/umma_f16acc tools/umma_f16acc.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "common.cuh"
#include "kernels.h"
using namespace hvit;

__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t sbo_bytes, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
__device__ __forceinline__ void ld16_pack(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(int f16acc, uint32_t* raw /*[128][64]*/, uint32_t* packed /*[128][2][16]*/) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __half* A = reinterpret_cast<__half*>(smem);
  __half* Bm = reinterpret_cast<__half*>(smem + 16384);
  for (int i = threadIdx.x; i < 8192; i += 128) { A[i] = __float2half(0.f); Bm[i] = __float2half(0.f); }
  __syncthreads();
  // un-swizzled K-major: addr(row, k) = (k / 8) * 4096 + (row / 8) * 128 + (row % 8) * 16 + (k % 8) * 2
  for (int m = threadIdx.x; m < 128; m += 128) A[((m / 8) * 128 + (m % 8) * 16) / 2] = __float2half(0.25f * (m % 7));
  for (int n = threadIdx.x; n < 64; n += 128) Bm[((n / 8) * 128 + (n % 8) * 16) / 2] = __float2half(static_cast<float>(n));
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  // clear the columns we will look at (tcgen05.st of zeros is not needed: the MMA overwrites with accumulate = 0)
  if (warp == 1 && elect_one()) {
    uint32_t idesc = make_idesc_16(128, 64, 0, 0, 1);
    if (f16acc) idesc &= ~(3u << 4);  // c_format: 0 = F16, 1 = F32
    umma_bf16(tmem, desc_nosw(smem_u32(A), 128, 4096), desc_nosw(smem_u32(Bm), 128, 4096), idesc, 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t ta = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  uint32_t r[32];
  for (int c = 0; c < 64; c += 32) {
    tmem_ld32(ta + c, r);
    tmem_ld_wait(r);
    for (int i = 0; i < 32; ++i) raw[(warp * 32 + lane) * 64 + c + i] = r[i];
  }
  uint32_t q[16];
  for (int h = 0; h < 2; ++h) {
    ld16_pack(ta + h * 32, q);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) packed[((warp * 32 + lane) * 2 + h) * 16 + i] = q[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  uint32_t *raw, *packed;
  cudaMallocManaged(&raw, 128 * 64 * 4);
  cudaMallocManaged(&packed, 128 * 32 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  for (int f16acc = 0; f16acc < 2; ++f16acc) {
    probe<<<1, 128, 32768>>>(f16acc, raw, packed);
    cudaError_t e = cudaDeviceSynchronize();
    printf("== accumulator %s: %s\n", f16acc ? "fp16" : "fp32", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    for (int m : {1, 2, 9}) {
      printf("row %d (expect D[m][n] = %.2f * n) raw columns 0..11:", m, 0.25 * (m % 7));
      for (int c = 0; c < 12; ++c) printf(" %08x", raw[m * 64 + c]);
      printf("\n   as fp32:");
      for (int c = 0; c < 8; ++c) printf(" %g", *reinterpret_cast<float*>(&raw[m * 64 + c]));
      printf("\n   low half as fp16:");
      for (int c = 0; c < 8; ++c) printf(" %g", __half2float(*reinterpret_cast<__half*>(&raw[m * 64 + c])));
      printf("\n   pack::16b x16 at column 0, registers 0..7:");
      for (int i = 0; i < 8; ++i) {
        const uint32_t v = packed[(m * 2 + 0) * 16 + i];
        const __half lo = *reinterpret_cast<const __half*>(&v);
        const uint16_t hs = v >> 16;
        printf(" (%g,%g)", __half2float(lo), __half2float(*reinterpret_cast<const __half*>(&hs)));
      }
      printf("\n   pack::16b x16 at column 32, registers 0..3:");
      for (int i = 0; i < 4; ++i) {
        const uint32_t v = packed[(m * 2 + 1) * 16 + i];
        const __half lo = *reinterpret_cast<const __half*>(&v);
        const uint16_t hs = v >> 16;
        printf(" (%g,%g)", __half2float(lo), __half2float(*reinterpret_cast<const __half*>(&hs)));
      }
      printf("\n");
    }
  }
  return 0;
}
