// Developer microbenchmark: MUFU.EX2 (fp32 and packed fp16: two MUFU.EX2.F16 + PRMT) / F2FP / SHF+IADD / SHFL issue rates per SM sub-partition (warps per CTA = W).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(long long* out, float* sink, int reps) {
  float v[16];
  for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 0.001f + i;
  unsigned u[16];
  for (int i = 0; i < 16; ++i) u[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (MODE == 1) { unsigned p; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(v[i]), "f"(v[(i + 1) & 15])); v[i] = __uint_as_float(p); }
      if (MODE == 2) { u[i] = (u[i] << 3) + 0x8000u; asm volatile("" : "+r"(u[i])); }
      if (MODE == 3) { asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i])); }
      if (MODE == 5) { asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i])); }
      if (MODE == 6) { asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(u[i])); }
      if (MODE == 4) { asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(v[(i + 1) & 15]), "f"(v[(i + 2) & 15])); }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  float s = 0; for (int i = 0; i < 16; ++i) s += v[i] + u[i];
  if (s == 1.2345f) sink[0] = s;
}
int main() {
  long long* d; float* s; cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
  const int reps = 1000;
  const char* names[7] = {"MUFU.EX2", "F2FP.f16x2", "SHL+IADD", "FFMA", "FMNMX3", "ex2.f16x2", "SHFL.BFLY"};
  for (int mode = 0; mode < 7; ++mode)
    for (int W : {4, 8, 16}) {
      for (int it = 0; it < 2; ++it) {
        if (mode == 0) k<0><<<148, 32 * W>>>(d, s, reps);
        if (mode == 1) k<1><<<148, 32 * W>>>(d, s, reps);
        if (mode == 2) k<2><<<148, 32 * W>>>(d, s, reps);
        if (mode == 3) k<3><<<148, 32 * W>>>(d, s, reps);
        if (mode == 4) k<4><<<148, 32 * W>>>(d, s, reps);
        if (mode == 5) k<5><<<148, 32 * W>>>(d, s, reps);
        if (mode == 6) k<6><<<148, 32 * W>>>(d, s, reps);
        cudaDeviceSynchronize();
      }
      long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double m = 0; for (int i = 0; i < 148; ++i) m += h[i]; m /= 148;
      printf("%-10s warps/CTA %2d: %.2f cycles per warp-instruction per SMSP\n", names[mode], W, m / (reps * 16.0 * (W / 4.0)));
    }
  return 0;
}
