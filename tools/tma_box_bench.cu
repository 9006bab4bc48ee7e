// Microbenchmark: cost of one 23 KB halo-tile TMA load as a function of the box's inner row length.
//   A: 5-D box (8 ch, 10, 18, 8 chunks, 1)  - 16-byte inner rows (what igemm_halo_kernel uses: chunk-plane layout)
//   B: 4-D box (64 ch, 10, 18, 1), no swizzle - 128-byte inner rows
//   C: same as B with the 128-byte swizzle
// Two loads in flight per CTA, 148 CTAs; prints cycles per load per CTA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_box_bench tools/tma_box_bench.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n));
}
__device__ __forceinline__ void expect_tx(uint64_t* b, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* b, uint32_t ph) {
  asm volatile(
      "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(s32(b)),
      "r"(ph)
      : "memory");
}

template <int RANK>
__global__ void bench(const __grid_constant__ CUtensorMap map, int iters, int B, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
    const long long t0 = clock64();
    for (int i = 0; i < iters + 2; ++i) {
      const int s = i & 1;
      if (i >= 2) wait(&bar[s], ((i - 2) >> 1) & 1);
      if (i < iters) {
        const int t = blockIdx.x * iters + i;
        const int w0 = (t % 31) * 8 - 1, h0 = ((t / 31) % 8) * 16 - 1, b = (t / 248) % B;
        expect_tx(&bar[s], 23040);
        if (RANK == 5)
          asm volatile(
              "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
                  s32(smem + s * 23552)),
              "l"(&map), "r"(s32(&bar[s])), "r"(0), "r"(w0), "r"(h0), "r"(0), "r"(b)
              : "memory");
        else
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                  s32(smem + s * 23552)),
              "l"(&map), "r"(s32(&bar[s])), "r"(0), "r"(w0), "r"(h0), "r"(b)
              : "memory");
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int B = 64, H = 128, W = 250, C = 64, iters = 400;
  __half* img;
  cudaMalloc(&img, size_t(B) * H * W * C * 2);
  cudaMemset(img, 0, size_t(B) * H * W * C * 2);
  long long* out;
  cudaMallocManaged(&out, 148 * 8);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fp, 12000, cudaEnableDefault, &q);
  EncodeFn enc = reinterpret_cast<EncodeFn>(fp);
  CUtensorMap mA, mB, mC;
  {
    cuuint64_t gd[5] = {8, (cuuint64_t)W, (cuuint64_t)H, 8, (cuuint64_t)B};
    cuuint64_t gs[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, 16, (cuuint64_t)H * W * C * 2};
    cuuint32_t bx[5] = {8, 10, 18, 8, 1}, es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, img, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) printf("encode A failed %d\n", (int)r);
  }
  for (int sw = 0; sw < 2; ++sw) {
    cuuint64_t gd[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gs[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t bx[4] = {64, 10, 18, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(sw ? &mC : &mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, img, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) printf("encode B/C failed %d\n", (int)r);
  }
  const int smem = 2 * 23552 + 1024;
  cudaFuncSetAttribute(bench<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[3] = {"A 5-D 16-byte rows ", "B 4-D 128-byte rows", "C 4-D 128B swizzle "};
  for (int rep = 0; rep < 2; ++rep)
    for (int v = 0; v < 3; ++v) {
      if (v == 0) bench<5><<<148, 32, smem>>>(mA, iters, B, out);
      else bench<4><<<148, 32, smem>>>(v == 1 ? mB : mC, iters, B, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", names[v], cudaGetErrorString(e)); return 1; }
      double s = 0;
      for (int i = 0; i < 148; ++i) s += out[i];
      printf("%s: %.0f cycles per 23 KB load per CTA (%.1f B/clk/SM)\n", names[v], s / 148 / iters, 23040.0 * 148 * iters / s);
    }
  return 0;
}
