// Developer microbenchmark: TMEM -> register read throughput (tcgen05.ld 32x32b) per SM, for W warps per CTA.
#include <cstdio>
#include <cuda_runtime.h>
#include "../speech-enhancement-via-hybrid-vision-transformer-project_b200/csrc/common.cuh"
using namespace hvit;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <int MODE>
__global__ void k(long long* out, int reps, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (MODE == 0) {  // one x32 load, wait, consume
#pragma unroll 1
      for (int c = 0; c < 16; ++c) {
        uint32_t v[32];
        tmem_ld32(base + c * 32, v);
        tmem_ld_wait(v);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
      }
    } else if (MODE == 1) {  // two x32 loads in flight
#pragma unroll 1
      for (int c = 0; c < 16; c += 2) {
        uint32_t v[32], w[32];
        tmem_ld32(base + c * 32, v);
        tmem_ld32(base + c * 32 + 32, w);
        tmem_ld_wait(v);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
        tmem_ld_wait(w);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= w[i];
      }
    } else {  // x16 loads
#pragma unroll 1
      for (int c = 0; c < 32; ++c) {
        uint32_t v[16];
        tmem_ld16(base + c * 16, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= v[i];
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
  long long* d; uint32_t* s;
  cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
  const int reps = 200;
  for (int mode = 0; mode < 3; ++mode)
    for (int W : {1, 4, 8, 16}) {
      for (int it = 0; it < 2; ++it) {
        if (mode == 0) k<0><<<148, 32 * W>>>(d, reps, s);
        else if (mode == 1) k<1><<<148, 32 * W>>>(d, reps, s);
        else k<2><<<148, 32 * W>>>(d, reps, s);
        cudaDeviceSynchronize();
      }
      long long h[148];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double m = 0; for (int i = 0; i < 148; ++i) m += h[i]; m /= 148;
      const double bytes = double(W) * reps * 16 * 4096;
      printf("mode %d warps %2d: %.0f cycles, %.1f B/cycle/SM (%s)\n", mode, W, m, bytes / m, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
