// fp32 CUDA-core kernels: the `precision="fp32"` accuracy mode (north star: max-rel error <= 1e-4 against the
// reference's fp32 forward) and the on-device cross-check for the tcgen05 kernels.  Same implicit-GEMM problem
// description (IgemmParams) as gemm_tc.cu, FFMA arithmetic, fp32 activations.
#include "common.cuh"
#include "kernels.h"

namespace hvit {
namespace {

constexpr int SBM = 64, SBN = 64, SBK = 16;

struct RowCtx {
  int b, oy, ox;
  bool valid;
  long long out_row;
};

__device__ __forceinline__ RowCtx decode_row(const IgemmParams& p, long long m, long long mtot) {
  RowCtx r;
  r.valid = m < mtot;
  r.b = 0; r.oy = 0; r.ox = 0; r.out_row = m;
  if (!r.valid) return r;
  if (p.mode == IG_PLAIN) return r;
  if (p.mode == IG_PATCH) {
    const int np = p.Hp * p.Wp;
    r.b = static_cast<int>(m / np);
    const int t = static_cast<int>(m - static_cast<long long>(r.b) * np);
    r.oy = t / p.Wp;
    r.ox = t - r.oy * p.Wp;
    return r;
  }
  const int Ho = p.mode == IG_UP2 ? 2 * p.H : p.H, Wo = p.mode == IG_UP2 ? 2 * p.W : p.W;
  r.b = static_cast<int>(m / (static_cast<long long>(Ho) * Wo));
  const int t = static_cast<int>(m - static_cast<long long>(r.b) * Ho * Wo);
  r.oy = t / Wo;
  r.ox = t - r.oy * Wo;
  r.out_row = (static_cast<long long>(r.b) * p.HoPitch + r.oy) * p.Wo + r.ox;
  return r;
}

// pointer to A(m, k..k+3) or nullptr when the tap falls in the zero padding
__device__ __forceinline__ const float* a_ptr(const IgemmParams& p, const float* A, int lda, const RowCtx& r,
                                              long long m, int k) {
  if (!r.valid) return nullptr;
  if (p.mode == IG_PLAIN) return A + m * lda + k;
  const int tap = k / p.Cin, c = k - tap * p.Cin;
  int iy, ix, pitch = p.H;
  if (p.mode == IG_CONV3) {
    iy = r.oy + tap / 3 - 1;
    ix = r.ox + tap % 3 - 1;
    if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) return nullptr;
  } else if (p.mode == IG_UP2) {
    const int uy = r.oy + tap / 3 - 1, ux = r.ox + tap % 3 - 1;
    if (uy < 0 || uy >= 2 * p.H || ux < 0 || ux >= 2 * p.W) return nullptr;
    iy = uy >> 1;
    ix = ux >> 1;
  } else {
    iy = r.oy * p.patch + tap / p.patch;
    ix = r.ox * p.patch + tap % p.patch;
    pitch = p.Hq * p.patch;
  }
  return A + ((static_cast<long long>(r.b) * pitch + iy) * p.W + ix) * lda + c;
}

__global__ void __launch_bounds__(256) igemm_f32_kernel(const IgemmParams p, const float* __restrict__ A, int lda,
                                                        const float* __restrict__ Wt, long long mtot) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const long long m0 = static_cast<long long>(blockIdx.x) * SBM;
  const int n0 = blockIdx.y * SBN;
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const long long am = m0 + lr;
  const RowCtx arow = decode_row(p, am, mtot);
  const int bn = n0 + lr;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += SBK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* ap = a_ptr(p, A, lda, arow, am, k0 + lk);
    if (ap != nullptr) av = *reinterpret_cast<const float4*>(ap);
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bn < p.N) bv = *reinterpret_cast<const float4*>(Wt + static_cast<long long>(bn) * p.K + k0 + lk);
    As[lk][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
    Bs[lk][lr] = bv.x; Bs[lk + 1][lr] = bv.y; Bs[lk + 2][lr] = bv.z; Bs[lk + 3][lr] = bv.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* out = reinterpret_cast<float*>(p.out);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    const RowCtx r = decode_row(p, m, mtot);
    if (!r.valid) continue;
    const int col = n0 + tx * 4;
    if (col >= p.N) continue;
    float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (p.scale != nullptr) v[j] *= p.scale[col + j];
      if (p.shift != nullptr) v[j] += p.shift[col + j];
      if (p.act == ACT_RELU) v[j] = fmaxf(v[j], 0.f);
      else if (p.act == ACT_GELU) v[j] = gelu_erf(v[j]);
    }
    if (p.residual != nullptr) {
      const long long rr = p.res_mod > 0 ? (r.out_row % p.res_mod) : r.out_row;
      const float4 r4 = *reinterpret_cast<const float4*>(p.residual + rr * p.ldr + col);
      v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
    }
    *reinterpret_cast<float4*>(out + r.out_row * p.ldc + col) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// Attention on CUDA cores: one warp per (clip, head, query), head_dim = 64 (two dims per lane), online softmax.
// Optionally materialises the softmax probabilities (HybridViT.forward(return_attentions=True),
// reference attention.py:113-114).
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float2 ld2(const T* p);
template <>
__device__ __forceinline__ float2 ld2<float>(const float* p) {
  return *reinterpret_cast<const float2*>(p);
}
template <>
__device__ __forceinline__ float2 ld2<bf16>(const bf16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
template <>
__device__ __forceinline__ float2 ld2<__half>(const __half* p) {
  return __half22float2(*reinterpret_cast<const __half2*>(p));
}
__device__ __forceinline__ void st2(__half* p, float a, float b) {
  *reinterpret_cast<uint32_t*>(p) = pack_f16x2(a, b);
}
__device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void st2(bf16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

template <typename T, bool WRITE_O>
__global__ void __launch_bounds__(256) attn_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                                                        float* __restrict__ probs, int B, int N, int heads, int D,
                                                        float scale, const int* __restrict__ geo) {
  const int lane = threadIdx.x & 31;
  const long long wg = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (wg >= static_cast<long long>(B) * heads * N) return;
  const int q = static_cast<int>(wg % N);
  const int h = static_cast<int>((wg / N) % heads);
  const int b = static_cast<int>(wg / (static_cast<long long>(N) * heads));
  const int ld = 3 * D;
  const T* base = qkv + static_cast<long long>(b) * N * ld + h * 64 + 2 * lane;
  const float2 qv = ld2<T>(base + static_cast<long long>(q) * ld);
  float mx = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  const int Nb = geo != nullptr ? geo[b * GEO_STRIDE + GEO_NTOK] : N;  // variable-length batch: valid key prefix of clip b
  for (int j = 0; j < Nb; ++j) {
    const float2 kv = ld2<T>(base + static_cast<long long>(j) * ld + D);
    const float s = warp_sum(qv.x * kv.x + qv.y * kv.y) * scale;
    const float mn = fmaxf(mx, s);
    const float alpha = expf(mx - mn);
    const float pj = expf(s - mn);
    l = l * alpha + pj;
    if (WRITE_O) {
      const float2 vv = ld2<T>(base + static_cast<long long>(j) * ld + 2 * D);
      o0 = o0 * alpha + pj * vv.x;
      o1 = o1 * alpha + pj * vv.y;
    }
    mx = mn;
  }
  const float inv = 1.f / l;
  if (WRITE_O) st2(out + (static_cast<long long>(b) * N + q) * D + h * 64 + 2 * lane, o0 * inv, o1 * inv);
  if (probs != nullptr) {
    float* prow = probs + ((static_cast<long long>(b) * heads + h) * N + q) * N;
    for (int j0 = 0; j0 < N; j0 += 32) {
      float mine = -INFINITY;
      for (int jj = 0; jj < 32 && j0 + jj < N; ++jj) {
        const float2 kv = ld2<T>(base + static_cast<long long>(j0 + jj) * ld + D);
        const float s = warp_sum(qv.x * kv.x + qv.y * kv.y) * scale;
        if (lane == jj) mine = s;
      }
      if (j0 + lane < N) prow[j0 + lane] = expf(mine - mx) * inv;
    }
  }
}

}  // namespace

int launch_igemm_f32(const IgemmParams& p, const float* A, int lda, const float* Wt, cudaStream_t stream) {
  if (p.K % SBK != 0 || p.N % 4 != 0 || (p.mode != IG_PLAIN && p.Cin % 4 != 0) || !p.out_f32 || p.pool) {
    set_error("igemm_f32: unsupported problem N=%d K=%d Cin=%d", p.N, p.K, p.Cin);
    return -1;
  }
  long long mtot;
  if (p.mode == IG_PLAIN) mtot = p.M;
  else if (p.mode == IG_PATCH) mtot = static_cast<long long>(p.B) * p.Hp * p.Wp;
  else if (p.mode == IG_UP2) mtot = static_cast<long long>(p.B) * 4 * p.H * p.W;
  else mtot = static_cast<long long>(p.B) * p.H * p.W;
  dim3 grid(static_cast<unsigned>((mtot + SBM - 1) / SBM), static_cast<unsigned>((p.N + SBN - 1) / SBN));
  igemm_f32_kernel<<<grid, 256, 0, stream>>>(p, A, lda, Wt, mtot);
  return check_launch("igemm_f32");
}

int launch_attn_f32(const float* qkv, float* out, float* probs, int B, int N, int heads, int D, float scale,
                    cudaStream_t stream, const int* geo) {
  if (D != heads * 64) {
    set_error("attention: head_dim must be 64 (D=%d heads=%d)", D, heads);
    return -1;
  }
  const long long warps = static_cast<long long>(B) * heads * N;
  attn_simt_kernel<float, true><<<static_cast<unsigned>((warps + 7) / 8), 256, 0, stream>>>(qkv, out, probs, B, N,
                                                                                            heads, D, scale, geo);
  return check_launch("attn_f32");
}

int launch_attn_probs_16(const void* qkv, int f16, float* probs, int B, int N, int heads, int D, float scale,
                         cudaStream_t stream) {
  if (D != heads * 64) {
    set_error("attention: head_dim must be 64 (D=%d heads=%d)", D, heads);
    return -1;
  }
  const long long warps = static_cast<long long>(B) * heads * N;
  const unsigned grid = static_cast<unsigned>((warps + 7) / 8);
  if (f16)
    attn_simt_kernel<__half, false><<<grid, 256, 0, stream>>>(reinterpret_cast<const __half*>(qkv), nullptr, probs, B,
                                                              N, heads, D, scale, nullptr);
  else
    attn_simt_kernel<bf16, false><<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(qkv), nullptr, probs, B, N,
                                                            heads, D, scale, nullptr);
  return check_launch("attn_probs_16");
}

}  // namespace hvit
