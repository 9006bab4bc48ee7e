// Microbenchmark: tcgen05.mma issue-to-completion rate as a function of the A-operand shared-memory layout.
// One CTA per SM, cta_group::1, M = 128, K = 16 per instruction, 16-bit operands; 2048 MMAs back to back on fixed
// (uninitialised) shared memory, one commit, cycles per MMA.  A layouts:
//   sw128      128B-swizzled K-major tile (what TMA-loaded GEMM tiles use)
//   nosw/128   un-swizzled K-major, core matrices (8 rows x 16 B) 128-byte aligned, SBO = 128
//   nosw/160+16  un-swizzled, SBO = 160 B and start + 16 B: the halo-tile taps of igemm_halo_kernel (unaligned core
//              matrices)
// B is always a 128B-swizzled K-major tile.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I speech-enhancement-via-hybrid-vision-transformer-project_b200/csrc -o tools/umma_bench tools/umma_bench.cu
#include <cuda.h>
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

__global__ void __launch_bounds__(128, 1) bench(int mode, int N, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = make_idesc_16(128, N, 0, 0, 1);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 65536);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int k = i & 3;
      uint64_t ad;
      if (mode == 0) ad = make_smem_desc_sw128(a + k * 32, 1024, 16);
      else if (mode == 1) ad = desc_nosw(a + k * 2 * 4096, 128, 4096);
      else if (mode == 2) ad = desc_nosw(a + 16 + k * 2 * 2880, 160, 2880);
      else ad = desc_nosw(a + k * 2 * 2880, 160, 2880);
      const uint64_t bd = make_smem_desc_sw128(b + k * 32, 1024, 16);
      umma_bf16(tmem, ad, bd, idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 148 * 8);
  const int smem = 65536 + 32768 + 1024, iters = 2048;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[4] = {"sw128       ", "nosw/128    ", "nosw/160+16 ", "nosw/160+0  "};
  const int Ns[3] = {64, 128, 256};
  for (int n = 0; n < 3; ++n)
    for (int m = 0; m < 4; ++m) {
      bench<<<148, 128, smem>>>(m, Ns[n], iters, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s N=%d: %s\n", names[m], Ns[n], cudaGetErrorString(e)); return 1; }
      double s = 0;
      for (int i = 0; i < 148; ++i) s += out[i];
      printf("A %s N=%3d: %.1f cycles per MMA (floor %d)\n", names[m], Ns[n], s / 148 / iters, Ns[n] / 2);
    }
  return 0;
}
