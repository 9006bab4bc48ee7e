// Developer microbenchmark: cost of one GEMM-epilogue chunk (32 TMEM columns -> scale/shift -> 16-bit -> swizzled
// staging row) per warp, with W epilogue warps per CTA, isolating TMEM load / math / shared-memory stores.
#include <cstdio>
#include <cuda_runtime.h>
#include "../speech-enhancement-via-hybrid-vision-transformer-project_b200/csrc/common.cuh"
using namespace hvit;

template <int MODE>
__global__ void k(long long* out, int reps, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  float* cst = reinterpret_cast<float*>(smem + 65536);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) cst[i] = 1.0f + i * 1e-3f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const int m = (warp & 3) * 32 + lane, sw = m & 7;
  uint8_t* row = smem + (warp >> 2) * 16384 + m * 128;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      uint32_t v[32];
      if (MODE & 1) {
        tmem_ld32(base + c * 32, v);
        tmem_ld_wait(v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = acc + i;
      }
      float f[32];
      const float* sc = cst + c * 32;
      const float* sh = cst + 512 + c * 32;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (MODE & 2) {
          const float4 a = *reinterpret_cast<const float4*>(sc + 4 * q), b = *reinterpret_cast<const float4*>(sh + 4 * q);
          f[4 * q + 0] = fmaf(__uint_as_float(v[4 * q + 0]), a.x, b.x); f[4 * q + 1] = fmaf(__uint_as_float(v[4 * q + 1]), a.y, b.y);
          f[4 * q + 2] = fmaf(__uint_as_float(v[4 * q + 2]), a.z, b.z); f[4 * q + 3] = fmaf(__uint_as_float(v[4 * q + 3]), a.w, b.w);
        } else {
          f[4 * q + 0] = __uint_as_float(v[4 * q + 0]); f[4 * q + 1] = __uint_as_float(v[4 * q + 1]);
          f[4 * q + 2] = __uint_as_float(v[4 * q + 2]); f[4 * q + 3] = __uint_as_float(v[4 * q + 3]);
        }
      }
      if (MODE & 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 pk;
          pk.x = pack_f16x2(f[8 * q + 0], f[8 * q + 1]); pk.y = pack_f16x2(f[8 * q + 2], f[8 * q + 3]);
          pk.z = pack_f16x2(f[8 * q + 4], f[8 * q + 5]); pk.w = pack_f16x2(f[8 * q + 6], f[8 * q + 7]);
          *reinterpret_cast<uint4*>(row + ((((c & 1) * 4 + q) ^ sw) << 4)) = pk;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= __float_as_uint(f[i]);
      }
      if (MODE & 8) { fence_proxy_async_smem(); named_bar_sync(1 + (warp >> 2), 128); }
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
  const char* names[16] = {"none", "tmem", "math", "tmem+math", "pack+sts", "tmem+pack+sts", "math+pack+sts", "tmem+math+pack+sts",
                           "", "", "", "", "", "", "", "all+fence+bar"};
  auto run = [&](int mode, int W) {
    const int smem = 65536 + 4096;
    for (int it = 0; it < 2; ++it) {
      switch (mode) {
        case 1: cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<1><<<148, 32 * W, smem>>>(d, reps, s); break;
        case 2: cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<2><<<148, 32 * W, smem>>>(d, reps, s); break;
        case 3: cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<3><<<148, 32 * W, smem>>>(d, reps, s); break;
        case 4: cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<4><<<148, 32 * W, smem>>>(d, reps, s); break;
        case 5: cudaFuncSetAttribute(k<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<5><<<148, 32 * W, smem>>>(d, reps, s); break;
        case 6: cudaFuncSetAttribute(k<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<6><<<148, 32 * W, smem>>>(d, reps, s); break;
        case 7: cudaFuncSetAttribute(k<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<7><<<148, 32 * W, smem>>>(d, reps, s); break;
        case 15: cudaFuncSetAttribute(k<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k<15><<<148, 32 * W, smem>>>(d, reps, s); break;
      }
      cudaDeviceSynchronize();
    }
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < 148; ++i) m += h[i]; m /= 148;
    printf("%-22s warps %2d: %7.1f cycles per chunk per warp (%s)\n", names[mode], W, m / (reps * 8.0), cudaGetErrorString(cudaGetLastError()));
  };
  for (int mode : {1, 2, 3, 4, 5, 6, 7, 15}) for (int W : {4, 8}) run(mode, W);
  return 0;
}
