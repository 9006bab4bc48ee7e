// Bandwidth-bound kernels of the enhance path: peak normalisation, STFT (+magnitude, per-clip max), stem
// conv+BN+ReLU+pool, LayerNorm, bilinear skip sampling, 64->1 head conv + tanh, final bilinear resize and the
// iSTFT (inverse FFT, window, overlap-add, window-sum-square normalisation).
// Reference arithmetic: inference/enhancer.py:55-135, models/components.py:15-99,160-167, models/hybrid_vit.py:367-389,458-465.
#include "common.cuh"
#include "kernels.h"

namespace hvit {
namespace {

constexpr int NFFT = 512;
constexpr int HOP = 128;
constexpr int NBIN = NFFT / 2 + 1;  // 257
constexpr int FR = 16;              // frames per block (one warp each)
constexpr int XS = NFFT + 1;        // padded frame stride in float2 (bank-conflict-free transposes)

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}
__device__ __forceinline__ float guard_scalar(unsigned bits) {  // "if > 1e-8 use it else 1.0" (enhancer.py:74-79,97-101)
  const float v = __uint_as_float(bits);
  return v > 1e-8f ? v : 1.0f;
}
__device__ __forceinline__ float hann512(int i) { return 0.5f - 0.5f * cospif(static_cast<float>(i) * (1.0f / 256.0f)); }
__device__ __forceinline__ int brev9(int i) { return static_cast<int>(__brev(static_cast<unsigned>(i)) >> 23); }

// In-place radix-2 DIT FFT of one 512-point frame held (bit-reversed) in shared memory, executed by one warp.
// tw[k] = exp(-2*pi*i*k/512), k < 256.
__device__ __forceinline__ void fft512_warp(float2* x, const float2* tw, int lane, bool inverse) {
#pragma unroll 1
  for (int s = 1; s <= 9; ++s) {
    const int half = 1 << (s - 1);
    const int tstep = NFFT >> s;
#pragma unroll
    for (int j = lane; j < NFFT / 2; j += 32) {
      const int grp = j >> (s - 1), pos = j & (half - 1);
      const int i0 = (grp << s) + pos, i1 = i0 + half;
      float2 w = tw[pos * tstep];
      if (inverse) w.y = -w.y;
      const float2 a = x[i0], b = x[i1];
      const float2 t = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
      x[i0] = make_float2(a.x + t.x, a.y + t.y);
      x[i1] = make_float2(a.x - t.x, a.y - t.y);
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void fill_twiddles(float2* tw) {
  for (int k = threadIdx.x; k < NFFT / 2; k += blockDim.x) {
    float s, c;
    sincospif(-static_cast<float>(k) * (1.0f / 256.0f), &s, &c);
    tw[k] = make_float2(c, s);
  }
}

// ------------------------------------------------------------------ peak |x| per clip (enhancer.py:72-79)
__global__ void peak_kernel(const float* __restrict__ wave, int n, unsigned* __restrict__ max_bits) {
  const int b = blockIdx.y;
  const float* w = wave + static_cast<long long>(b) * n;
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(max_bits + b, __float_as_uint(m));
}
__global__ void fill_u32_kernel(unsigned* p, int n, unsigned v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ------------------------------------------------------------------ STFT + |.| + per-clip max (enhancer.py:82-101)
__global__ void __launch_bounds__(FR * 32) stft_kernel(const float* __restrict__ wave, int n, int T,
                                                       const unsigned* __restrict__ max_bits,
                                                       float2* __restrict__ spec, float* __restrict__ mag,
                                                       unsigned* __restrict__ mag_max_bits) {
  extern __shared__ float2 sm[];
  float2* xs = sm;                // [FR][XS]
  float2* tw = sm + FR * XS;      // [256]
  const int b = blockIdx.y, t0 = blockIdx.x * FR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  fill_twiddles(tw);
  const float mv = guard_scalar(max_bits[b]);
  const int t = t0 + warp;
  float2* x = xs + warp * XS;
  if (t < T) {
    const float* w = wave + static_cast<long long>(b) * n;
    for (int i = lane; i < NFFT; i += 32) {
      const int src = t * HOP - NFFT / 2 + i;  // centred frame, zero padding
      float v = 0.f;
      if (src >= 0 && src < n) v = (w[src] / mv) * hann512(i);
      x[brev9(i)] = make_float2(v, 0.f);
    }
  }
  __syncthreads();  // twiddles + frames visible
  if (t < T) fft512_warp(x, tw, lane, false);
  __syncthreads();
  float lmax = 0.f;
  for (int idx = threadIdx.x; idx < NBIN * FR; idx += blockDim.x) {
    const int tl = idx & (FR - 1), f = idx / FR;
    if (t0 + tl < T) {
      const float2 z = xs[tl * XS + f];
      const long long o = (static_cast<long long>(b) * NBIN + f) * T + t0 + tl;
      spec[o] = z;
      const float m = hypotf(z.x, z.y);
      mag[o] = m;
      lmax = fmaxf(lmax, m);
    }
  }
  lmax = warp_max(lmax);
  if (lane == 0) atomicMax(mag_max_bits + b, __float_as_uint(lmax));
}

// ------------------------------------------------------------------ iSTFT part 1: spectrum -> windowed frames
// E = (model_out * mag_max) * S/|S|  (== mag * exp(1j*angle(S)), enhancer.py:115-119), irfft-512, * Hann.
__global__ void __launch_bounds__(FR * 32) istft_frames_kernel(const float* __restrict__ model_out,
                                                               const float2* __restrict__ spec,
                                                               const unsigned* __restrict__ mag_max_bits, int T,
                                                               float* __restrict__ frames) {
  extern __shared__ float2 sm[];
  float2* xs = sm;
  float2* tw = sm + FR * XS;
  const int b = blockIdx.y, t0 = blockIdx.x * FR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  fill_twiddles(tw);
  const float mm = guard_scalar(mag_max_bits[b]);
  for (int idx = threadIdx.x; idx < NBIN * FR; idx += blockDim.x) {
    const int tl = idx & (FR - 1), f = idx / FR;
    if (t0 + tl < T) {
      const long long o = (static_cast<long long>(b) * NBIN + f) * T + t0 + tl;
      const float2 z = spec[o];
      const float a = hypotf(z.x, z.y);
      const float e = model_out[o] * mm;
      float2 E = a > 0.f ? make_float2(e * (z.x / a), e * (z.y / a)) : make_float2(e, 0.f);
      if (f == 0 || f == NFFT / 2) E.y = 0.f;  // c2r transforms ignore the imaginary part of DC / Nyquist
      float2* x = xs + tl * XS;
      x[brev9(f)] = E;
      if (f > 0 && f < NFFT / 2) x[brev9(NFFT - f)] = make_float2(E.x, -E.y);
    }
  }
  __syncthreads();
  const int t = t0 + warp;
  if (t < T) {
    float2* x = xs + warp * XS;
    fft512_warp(x, tw, lane, true);
    float* fr = frames + (static_cast<long long>(b) * T + t) * NFFT;
    for (int i = lane; i < NFFT; i += 32) fr[i] = x[i].x * (1.0f / NFFT) * hann512(i);
  }
}

// iSTFT part 2: overlap-add, trim n_fft/2, divide by the window sum-square envelope, de-normalise.
__global__ void istft_ola_kernel(const float* __restrict__ frames, const unsigned* __restrict__ max_bits, int n, int T,
                                 float* __restrict__ wave_out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = i + NFFT / 2;
  int tlo = (j - (NFFT - 1) + HOP - 1) / HOP;
  if (j - (NFFT - 1) < 0) tlo = 0;
  int thi = j / HOP;
  if (thi > T - 1) thi = T - 1;
  float acc = 0.f, wss = 0.f;
  for (int t = tlo; t <= thi; ++t) {
    const int k = j - t * HOP;
    acc += frames[(static_cast<long long>(b) * T + t) * NFFT + k];
    const float w = hann512(k);
    wss += w * w;
  }
  if (wss > 1.17549435e-38f) acc /= wss;
  wave_out[static_cast<long long>(b) * n + i] = acc * guard_scalar(max_bits[b]);
}

// ------------------------------------------------------------------ stem: Conv3x3(1->C, no bias)+BN+ReLU[+MaxPool2]
template <typename T>
__device__ __forceinline__ void store16(T* dst, const float* v);
template <>
__device__ __forceinline__ void store16<float>(float* dst, const float* v) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
}
template <>
__device__ __forceinline__ void store16<bf16>(bf16* dst, const float* v) {
#pragma unroll
  for (int j = 0; j < 16; j += 8) {
    uint4 q;
    q.x = pack_bf16x2(v[j], v[j + 1]);
    q.y = pack_bf16x2(v[j + 2], v[j + 3]);
    q.z = pack_bf16x2(v[j + 4], v[j + 5]);
    q.w = pack_bf16x2(v[j + 6], v[j + 7]);
    *reinterpret_cast<uint4*>(dst + j) = q;
  }
}

template <>
__device__ __forceinline__ void store16<__half>(__half* dst, const float* v) {
#pragma unroll
  for (int j = 0; j < 16; j += 8) {
    uint4 q;
    q.x = pack_f16x2(v[j], v[j + 1]);
    q.y = pack_f16x2(v[j + 2], v[j + 3]);
    q.z = pack_f16x2(v[j + 4], v[j + 5]);
    q.w = pack_f16x2(v[j + 6], v[j + 7]);
    *reinterpret_cast<uint4*>(dst + j) = q;
  }
}

template <typename T, int POOL>
__global__ void stem_kernel(const float* __restrict__ x, const unsigned* __restrict__ mag_max_bits,
                            const float* __restrict__ w9c, const float* __restrict__ scale,
                            const float* __restrict__ shift, T* __restrict__ out, int H, int W, int C, int Ho, int Wo) {
  constexpr int WIN = POOL + 2;
  constexpr int TW = 32 * POOL + 2;
  extern __shared__ float sf[];
  float* in_s = sf;                    // [WIN][TW]
  float* w_s = in_s + WIN * TW;        // [9][C]
  float* sc_s = w_s + 9 * C;           // [C]
  float* sh_s = sc_s + C;              // [C]
  const int b = blockIdx.z, oy = blockIdx.y, ox0 = blockIdx.x * 32;
  const int tid = threadIdx.y * 32 + threadIdx.x, nthr = blockDim.x * blockDim.y;
  const float mm = mag_max_bits != nullptr ? guard_scalar(mag_max_bits[b]) : 1.0f;
  for (int i = tid; i < WIN * TW; i += nthr) {
    const int r = i / TW, cc = i - r * TW;
    const int iy = oy * POOL - 1 + r, ix = ox0 * POOL - 1 + cc;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      v = x[(static_cast<long long>(b) * H + iy) * W + ix];
      if (mag_max_bits != nullptr) v = v / mm;
    }
    in_s[i] = v;
  }
  for (int i = tid; i < 9 * C; i += nthr) w_s[i] = w9c[i];
  for (int i = tid; i < C; i += nthr) {
    sc_s[i] = scale[i];
    sh_s[i] = shift[i];
  }
  __syncthreads();
  const int ox = ox0 + threadIdx.x;
  if (ox >= Wo) return;
  float win[WIN][WIN];
#pragma unroll
  for (int r = 0; r < WIN; ++r)
#pragma unroll
    for (int c = 0; c < WIN; ++c) win[r][c] = in_s[r * TW + threadIdx.x * POOL + c];
  const int cbase = threadIdx.y * 16;
  float res[16];
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    float acc[POOL * POOL][4];
#pragma unroll
    for (int q = 0; q < POOL * POOL; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float4 wv = *reinterpret_cast<const float4*>(&w_s[tap * C + cbase + c4 * 4]);
      const int ky = tap / 3, kx = tap % 3;
#pragma unroll
      for (int py = 0; py < POOL; ++py)
#pragma unroll
        for (int px = 0; px < POOL; ++px) {
          const float a = win[py + ky][px + kx];
          float* q = acc[py * POOL + px];
          q[0] = fmaf(a, wv.x, q[0]); q[1] = fmaf(a, wv.y, q[1]); q[2] = fmaf(a, wv.z, q[2]); q[3] = fmaf(a, wv.w, q[3]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = cbase + c4 * 4 + j;
      float best = 0.f;  // ReLU output >= 0
#pragma unroll
      for (int q = 0; q < POOL * POOL; ++q) best = fmaxf(best, fmaf(acc[q][j], sc_s[c], sh_s[c]));
      res[c4 * 4 + j] = best;
    }
  }
  store16<T>(out + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * C + cbase, res);
}

template <typename T>
__device__ __forceinline__ void ld4(const T* p, float* v);
template <typename T>
__device__ __forceinline__ void st4(T* p, const float* v);
template <>
__device__ __forceinline__ void ld4<float>(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <>
__device__ __forceinline__ void ld4<bf16>(const bf16* p, float* v) {
  const uint2 a = *reinterpret_cast<const uint2*>(p);
  const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a.x));
  const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a.y));
  v[0] = x.x; v[1] = x.y; v[2] = y.x; v[3] = y.y;
}
template <>
__device__ __forceinline__ void ld4<__half>(const __half* p, float* v) {
  const uint2 a = *reinterpret_cast<const uint2*>(p);
  const float2 x = __half22float2(*reinterpret_cast<const __half2*>(&a.x));
  const float2 y = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
  v[0] = x.x; v[1] = x.y; v[2] = y.x; v[3] = y.y;
}
template <>
__device__ __forceinline__ void st4<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void st4<bf16>(bf16* p, const float* v) {
  uint2 pk;
  pk.x = pack_bf16x2(v[0], v[1]);
  pk.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = pk;
}
template <>
__device__ __forceinline__ void st4<__half>(__half* p, const float* v) {
  uint2 pk;
  pk.x = pack_f16x2(v[0], v[1]);
  pk.y = pack_f16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = pk;
}

// ------------------------------------------------------------------ LayerNorm (warp per row)
template <typename T>
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bb,
                                 T* __restrict__ out, int rows, int D, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<long long>(row) * D;
  float s = 0.f;
  for (int i = lane * 4; i < D; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + i);
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
  for (int i = lane * 4; i < D; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + i);
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) / D + eps);
  T* orow = out + static_cast<long long>(row) * D;
  for (int i = lane * 4; i < D; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + i);
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + i));
    const float4 be = __ldg(reinterpret_cast<const float4*>(bb + i));
    const float y0 = (v.x - mean) * rstd * gg.x + be.x, y1 = (v.y - mean) * rstd * gg.y + be.y;
    const float y2 = (v.z - mean) * rstd * gg.z + be.z, y3 = (v.w - mean) * rstd * gg.w + be.w;
    const float yv[4] = {y0, y1, y2, y3};
    st4<T>(orow + i, yv);
  }
}

// ------------------------------------------------------------------ bilinear helpers (torch align_corners=False)
struct Lerp {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Lerp make_lerp(int dst, int in_size, int out_size) {
  const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  Lerp r;
  r.i0 = static_cast<int>(src);
  if (r.i0 > in_size - 1) r.i0 = in_size - 1;
  r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
  r.l1 = src - static_cast<float>(r.i0);
  r.l0 = 1.0f - r.l1;
  return r;
}

// skip feature [B, Hs(pitch), Ws, C] -> bilinear sample at decoder resolution [B*Hd*Wd, C]
// (the 1x1 projection is applied AFTER sampling; exact because bilinear weights sum to 1 - hybrid_vit.py:377-386)
template <typename T>
__global__ void skip_sample_kernel(const T* __restrict__ src, int Hs, int HsPitch, int Ws, int C, int Hd, int Wd,
                                   T* __restrict__ dst, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = C / 4;
  const int c = static_cast<int>(idx % cv) * 4;
  long long r = idx / cv;
  const int wd = static_cast<int>(r % Wd);
  r /= Wd;
  const int hd = static_cast<int>(r % Hd);
  const int b = static_cast<int>(r / Hd);
  const Lerp ly = make_lerp(hd, Hs, Hd), lx = make_lerp(wd, Ws, Wd);
  const T* base = src + static_cast<long long>(b) * HsPitch * Ws * C + c;
  float v00[4], v01[4], v10[4], v11[4], o[4];
  ld4<T>(base + (static_cast<long long>(ly.i0) * Ws + lx.i0) * C, v00);
  ld4<T>(base + (static_cast<long long>(ly.i0) * Ws + lx.i1) * C, v01);
  ld4<T>(base + (static_cast<long long>(ly.i1) * Ws + lx.i0) * C, v10);
  ld4<T>(base + (static_cast<long long>(ly.i1) * Ws + lx.i1) * C, v11);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    o[j] = ly.l0 * (lx.l0 * v00[j] + lx.l1 * v01[j]) + ly.l1 * (lx.l0 * v10[j] + lx.l1 * v11[j]);
  st4<T>(dst + ((static_cast<long long>(b) * Hd + hd) * Wd + wd) * C + c, o);
}

// ------------------------------------------------------------------ head: Conv3x3(C->1, no bias) + tanh, fp32 accumulate
template <typename T>
__global__ void head_kernel(const T* __restrict__ x, const float* __restrict__ w9c, int B, int H, int W, int C,
                            float* __restrict__ logits, float* __restrict__ out_tanh) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long pix = gid >> 3;
  const int sub = static_cast<int>(gid & 7);
  const long long npix = static_cast<long long>(B) * H * W;
  const bool live = pix < npix;
  float acc = 0.f;
  if (live) {
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    const int b = static_cast<int>(pix / (static_cast<long long>(W) * H));
    for (int tap = 0; tap < 9; ++tap) {
      const int iy = h + tap / 3 - 1, ix = w + tap % 3 - 1;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const T* px = x + ((static_cast<long long>(b) * H + iy) * W + ix) * C;
      for (int c = sub * 4; c < C; c += 32) {
        float v[4];
        ld4<T>(px + c, v);
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w9c + tap * C + c));
        acc = fmaf(v[0], wv.x, acc); acc = fmaf(v[1], wv.y, acc); acc = fmaf(v[2], wv.z, acc); acc = fmaf(v[3], wv.w, acc);
      }
    }
  }
  acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 1);
  acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 2);
  acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 4);
  if (live && sub == 0) {
    if (logits != nullptr) logits[pix] = acc;
    out_tanh[pix] = tanhf(acc);
  }
}

// ------------------------------------------------------------------ final bilinear resize [B,Hs,Ws] -> [B,Hd,Wd]
__global__ void resize_kernel(const float* __restrict__ src, int Hs, int Ws, float* __restrict__ dst, int Hd, int Wd,
                              long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wd = static_cast<int>(idx % Wd);
  const int hd = static_cast<int>((idx / Wd) % Hd);
  const long long b = idx / (static_cast<long long>(Wd) * Hd);
  const Lerp ly = make_lerp(hd, Hs, Hd), lx = make_lerp(wd, Ws, Wd);
  const float* s = src + b * Hs * Ws;
  const float v00 = s[ly.i0 * Ws + lx.i0], v01 = s[ly.i0 * Ws + lx.i1];
  const float v10 = s[ly.i1 * Ws + lx.i0], v11 = s[ly.i1 * Ws + lx.i1];
  dst[idx] = ly.l0 * (lx.l0 * v00 + lx.l1 * v01) + ly.l1 * (lx.l0 * v10 + lx.l1 * v11);
}

// ------------------------------------------------------------------ 2x2 max-pool, NHWC fp32 (fp32 mode only)
__global__ void maxpool2_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, int C, int Ho,
                                int Wo, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = C / 4;
  const int c = static_cast<int>(idx % cv) * 4;
  long long r = idx / cv;
  const int ow = static_cast<int>(r % Wo);
  r /= Wo;
  const int oh = static_cast<int>(r % Ho);
  const long long b = r / Ho;
  const float* s = src + ((b * H + 2 * oh) * W + 2 * ow) * C + c;
  const float4 a = *reinterpret_cast<const float4*>(s), bq = *reinterpret_cast<const float4*>(s + C);
  const float4 cq = *reinterpret_cast<const float4*>(s + static_cast<long long>(W) * C);
  const float4 d = *reinterpret_cast<const float4*>(s + static_cast<long long>(W) * C + C);
  float4 o;
  o.x = fmaxf(fmaxf(a.x, bq.x), fmaxf(cq.x, d.x));
  o.y = fmaxf(fmaxf(a.y, bq.y), fmaxf(cq.y, d.y));
  o.z = fmaxf(fmaxf(a.z, bq.z), fmaxf(cq.z, d.z));
  o.w = fmaxf(fmaxf(a.w, bq.w), fmaxf(cq.w, d.w));
  *reinterpret_cast<float4*>(dst + ((b * Ho + oh) * Wo + ow) * C + c) = o;
}

constexpr int FFT_SMEM = (FR * XS + NFFT / 2) * sizeof(float2);

}  // namespace

// ================================================================== launchers
int launch_peak(const float* wave, int B, int n, float* max_val, int normalize, cudaStream_t s) {
  unsigned* bits = reinterpret_cast<unsigned*>(max_val);
  if (!normalize) {
    fill_u32_kernel<<<(B + 255) / 256, 256, 0, s>>>(bits, B, 0x3F800000u);  // 1.0f
    return check_launch("fill(max_val)");
  }
  fill_u32_kernel<<<(B + 255) / 256, 256, 0, s>>>(bits, B, 0u);
  if (n > 0) {
    dim3 grid(8, B);
    peak_kernel<<<grid, 256, 0, s>>>(wave, n, bits);
  }
  return check_launch("peak");
}

int launch_stft(const float* wave, int B, int n, int T, const float* max_val, float2* spec, float* mag,
                unsigned* mag_max_bits, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FFT_SMEM);
    cudaFuncSetAttribute(istft_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FFT_SMEM);
    configured = true;
  }
  fill_u32_kernel<<<(B + 255) / 256, 256, 0, s>>>(mag_max_bits, B, 0u);
  dim3 grid((T + FR - 1) / FR, B);
  stft_kernel<<<grid, FR * 32, FFT_SMEM, s>>>(wave, n, T, reinterpret_cast<const unsigned*>(max_val), spec, mag,
                                               mag_max_bits);
  return check_launch("stft");
}

static void fft_smem_config() {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FFT_SMEM);
    cudaFuncSetAttribute(istft_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FFT_SMEM);
    configured = true;
  }
}

int launch_istft_frames(const float* model_out, const float2* spec, const unsigned* mag_max_bits, float* frames, int B,
                        int T, cudaStream_t s) {
  fft_smem_config();
  dim3 grid((T + FR - 1) / FR, B);
  istft_frames_kernel<<<grid, FR * 32, FFT_SMEM, s>>>(model_out, spec, mag_max_bits, T, frames);
  return check_launch("istft_frames");
}

int launch_istft_ola(const float* frames, const float* max_val, float* wave_out, int B, int n, int T, cudaStream_t s) {
  if (n <= 0) return 0;
  dim3 g2((n + 255) / 256, B);
  istft_ola_kernel<<<g2, 256, 0, s>>>(frames, reinterpret_cast<const unsigned*>(max_val), n, T, wave_out);
  return check_launch("istft_ola");
}

template <typename T>
static void stem_dispatch(const float* x, const unsigned* mm, const float* w, const float* scale, const float* shift,
                          void* out, int H, int W, int C, int pool, dim3 grid, dim3 block, int smem, cudaStream_t s) {
  const int Ho = H / pool, Wo = W / pool;
  if (pool == 2)
    stem_kernel<T, 2><<<grid, block, smem, s>>>(x, mm, w, scale, shift, reinterpret_cast<T*>(out), H, W, C, Ho, Wo);
  else
    stem_kernel<T, 1><<<grid, block, smem, s>>>(x, mm, w, scale, shift, reinterpret_cast<T*>(out), H, W, C, Ho, Wo);
}

int launch_stem(const float* x, const unsigned* mag_max_bits, const float* w, const float* scale, const float* shift,
                void* out, int dt, int B, int H, int W, int C, int pool, cudaStream_t s) {
  if (C % 16 != 0 || C > 512 || (pool != 1 && pool != 2)) {
    set_error("stem: unsupported C=%d pool=%d", C, pool);
    return -1;
  }
  const int Ho = H / pool, Wo = W / pool;
  dim3 block(32, C / 16), grid((Wo + 31) / 32, Ho, B);
  const int smem = ((pool + 2) * (32 * pool + 2) + 11 * C) * sizeof(float);
  if (dt == DT_BF16) stem_dispatch<bf16>(x, mag_max_bits, w, scale, shift, out, H, W, C, pool, grid, block, smem, s);
  else if (dt == DT_F16) stem_dispatch<__half>(x, mag_max_bits, w, scale, shift, out, H, W, C, pool, grid, block, smem, s);
  else stem_dispatch<float>(x, mag_max_bits, w, scale, shift, out, H, W, C, pool, grid, block, smem, s);
  return check_launch("stem");
}

int launch_layernorm(const float* x, const float* g, const float* b, void* out, int dt, int rows, int D,
                     float eps, cudaStream_t s) {
  if (D % 4 != 0) {
    set_error("layernorm: D %% 4 != 0");
    return -1;
  }
  const int grid = (rows + 7) / 8;
  if (dt == DT_BF16)
    layernorm_kernel<bf16><<<grid, 256, 0, s>>>(x, g, b, reinterpret_cast<bf16*>(out), rows, D, eps);
  else if (dt == DT_F16)
    layernorm_kernel<__half><<<grid, 256, 0, s>>>(x, g, b, reinterpret_cast<__half*>(out), rows, D, eps);
  else
    layernorm_kernel<float><<<grid, 256, 0, s>>>(x, g, b, reinterpret_cast<float*>(out), rows, D, eps);
  return check_launch("layernorm");
}

int launch_skip_sample(const void* src, int dt, int B, int Hs, int HsPitch, int Ws, int C, int Hd, int Wd,
                       void* dst, cudaStream_t s) {
  const long long total = static_cast<long long>(B) * Hd * Wd * (C / 4);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (dt == DT_BF16)
    skip_sample_kernel<bf16><<<grid, 256, 0, s>>>(reinterpret_cast<const bf16*>(src), Hs, HsPitch, Ws, C, Hd, Wd,
                                                  reinterpret_cast<bf16*>(dst), total);
  else if (dt == DT_F16)
    skip_sample_kernel<__half><<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(src), Hs, HsPitch, Ws, C, Hd, Wd,
                                                    reinterpret_cast<__half*>(dst), total);
  else
    skip_sample_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(src), Hs, HsPitch, Ws, C, Hd, Wd,
                                                   reinterpret_cast<float*>(dst), total);
  return check_launch("skip_sample");
}

int launch_head(const void* x, int dt, const float* w, int B, int H, int W, int C, float* logits,
                float* out_tanh, cudaStream_t s) {
  if (C % 4 != 0) {
    set_error("head: C %% 4 != 0");
    return -1;
  }
  const long long threads = static_cast<long long>(B) * H * W * 8;
  const unsigned grid = static_cast<unsigned>((threads + 255) / 256);
  if (dt == DT_BF16)
    head_kernel<bf16><<<grid, 256, 0, s>>>(reinterpret_cast<const bf16*>(x), w, B, H, W, C, logits, out_tanh);
  else if (dt == DT_F16)
    head_kernel<__half><<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(x), w, B, H, W, C, logits, out_tanh);
  else
    head_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(x), w, B, H, W, C, logits, out_tanh);
  return check_launch("head");
}

int launch_resize(const float* src, int B, int Hs, int Ws, float* dst, int Hd, int Wd, cudaStream_t s) {
  const long long total = static_cast<long long>(B) * Hd * Wd;
  resize_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(src, Hs, Ws, dst, Hd, Wd, total);
  return check_launch("resize");
}

int launch_maxpool2(const float* src, float* dst, int B, int H, int W, int C, cudaStream_t s) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo * (C / 4);
  maxpool2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(src, dst, H, W, C, Ho, Wo, total);
  return check_launch("maxpool2");
}

}  // namespace hvit
