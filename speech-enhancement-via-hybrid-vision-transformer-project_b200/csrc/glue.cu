// Bandwidth-bound kernels of the enhance path: peak normalisation, STFT (+magnitude, per-clip max), stem
// conv+BN+ReLU+pool, LayerNorm, bilinear skip sampling, 64->1 head conv + tanh, final bilinear resize and the
// iSTFT (inverse FFT, window, overlap-add, window-sum-square normalisation).
// Reference arithmetic: inference/enhancer.py:55-135, models/components.py:15-99,160-167, models/hybrid_vit.py:367-389,458-465.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace hvit {
namespace {

constexpr int NFFT = 512;
constexpr int HOP = 128;
constexpr int NBIN = NFFT / 2 + 1;  // 257
constexpr int FR = 8;               // warps per block; every warp transforms a PAIR of frames
constexpr int FRAMES = 2 * FR;      // frames per block
constexpr int XS = 545;             // frame stride in float2: 512 + 32 in-frame padding + 1 (odd mod 16 -> the
                                    // 16-frame transposes are bank-conflict free)
__device__ __forceinline__ int fpad(int i) { return i + (i >> 4); }  // in-frame padding against Stockham store conflicts

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
// Tables shared by the STFT / iSTFT kernels, filled once per process by fft_tables_kernel (read through L1):
//   g_tw512[n] = exp(-2*pi*i*n/512), g_hann512[n] = periodic Hann window (scipy.signal.get_window('hann', 512)).
__device__ float2 g_tw512[NFFT];
__device__ float g_hann512[NFFT];
__global__ void fft_tables_kernel() {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < NFFT) {
    float sn, cs;
    sincospif(-static_cast<float>(k) * (1.0f / 256.0f), &sn, &cs);
    g_tw512[k] = make_float2(cs, sn);
    g_hann512[k] = 0.5f - 0.5f * cospif(static_cast<float>(k) * (1.0f / 256.0f));
  }
}
__device__ __forceinline__ float hann512(int i) { return __ldg(&g_hann512[i]); }
__device__ __forceinline__ int brev9(int i) { return static_cast<int>(__brev(static_cast<unsigned>(i)) >> 23); }

// Complex arithmetic on the packed fp32 pipes (FADD2 / FMUL2 / FFMA2: one issue slot per complex add) - the FFT
// kernels are bound by instruction issue.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  // (a.x b.x - a.y b.y, a.x b.y + a.y b.x)
  return __ffma2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x), __fmul2_rn(make_float2(a.x, a.x), b));
}

// 8-point DFT in registers (three radix-2 levels).  INV selects the conjugate twiddles.
template <bool INV>
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  constexpr float S = INV ? 1.0f : -1.0f;   // exp(S * i * theta)
  constexpr float H = 0.70710678118654752440f;
  // a + (S*i) z = (a.x - S z.y, a.y + S z.x);  a - (S*i) z = (a.x + S z.y, a.y - S z.x)
  auto add_jz = [](float2 a, float2 z) { return __ffma2_rn(make_float2(z.y, z.x), make_float2(-S, S), a); };
  auto sub_jz = [](float2 a, float2 z) { return __ffma2_rn(make_float2(z.y, z.x), make_float2(S, -S), a); };
  const float2 a0 = cadd(v[0], v[4]), a1 = csub(v[0], v[4]);
  const float2 a2 = cadd(v[2], v[6]), a3 = csub(v[2], v[6]);
  const float2 a4 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
  const float2 a6 = cadd(v[3], v[7]), a7 = csub(v[3], v[7]);
  const float2 b0 = cadd(a0, a2), b2 = csub(a0, a2);
  const float2 b1 = add_jz(a1, a3), b3 = sub_jz(a1, a3);
  const float2 c0 = cadd(a4, a6), c2 = csub(a4, a6);
  float2 c1 = add_jz(a5, a7), c3 = sub_jz(a5, a7);
  c1 = __fmul2_rn(add_jz(c1, c1), make_float2(H, H));      // * exp(S*i*pi/4)  = H (1 + S i)
  c3 = __fmul2_rn(sub_jz(c3, c3), make_float2(-H, -H));    // * exp(S*i*3pi/4) = H (-1 + S i)
  v[0] = cadd(b0, c0); v[4] = csub(b0, c0);
  v[1] = cadd(b1, c1); v[5] = csub(b1, c1);
  v[2] = add_jz(b2, c2); v[6] = sub_jz(b2, c2);            // c2 * exp(S*i*pi/2) = (S i) c2
  v[3] = cadd(b3, c3); v[7] = csub(b3, c3);
}

// 512-point complex FFT of one frame in shared memory by ONE warp: radix-8 Stockham, 3 passes, natural order in and
// out (indices through fpad()).  Each lane does two 8-point butterflies per pass (64 per pass); all 16 inputs are
// read into registers before any output is written, so the transform is in place with only __syncwarp().
// tw[n] = exp(-2*pi*i*n/512), n < 512.  (Indexing verified against numpy.fft in a host prototype, DESIGN.md.)
template <bool INV>
__device__ __forceinline__ void fft512_warp(float2* x, int lane) {
  const float2* tw = g_tw512;
#pragma unroll
  for (int stage = 0; stage < 3; ++stage) {
    const int Ns = stage == 0 ? 1 : (stage == 1 ? 8 : 64);
    const int tmul = 64 / Ns;  // 512 / (Ns * 8)
    float2 v[2][8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h;
#pragma unroll
      for (int r = 0; r < 8; ++r) v[h][r] = x[fpad(j + r * 64)];
      if (stage > 0) {
        const int k = j & (Ns - 1);
#pragma unroll
        for (int r = 1; r < 8; ++r) {
          float2 w = __ldg(&tw[r * k * tmul]);
          if (INV) w.y = -w.y;
          v[h][r] = cmul(v[h][r], w);
        }
      }
      dft8<INV>(v[h]);
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h;
      const int k = j & (Ns - 1);
      const int base = (j - k) * 8 + k;
#pragma unroll
      for (int r = 0; r < 8; ++r) x[fpad(base + r * Ns)] = v[h][r];
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ peak |x| per clip (enhancer.py:72-79)
__global__ void peak_kernel(const float* __restrict__ wave, int n, unsigned* __restrict__ max_bits) {
  griddep_launch_dependents();
  griddep_wait();
  const int b = blockIdx.y;
  const float* w = wave + static_cast<long long>(b) * n;
  float m = 0.f;
  // 16-byte loads where the clip base allows it (4 independent maxima per thread), scalar tail / fallback
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  int i0 = 0;
  if ((reinterpret_cast<uintptr_t>(w) & 15) == 0) {
    const int n4 = n >> 2;
    float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < n4; i += nthr) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(w) + i);
      m4.x = fmaxf(m4.x, fabsf(v.x)); m4.y = fmaxf(m4.y, fabsf(v.y));
      m4.z = fmaxf(m4.z, fabsf(v.z)); m4.w = fmaxf(m4.w, fabsf(v.w));
    }
    m = fmaxf(fmaxf(m4.x, m4.y), fmaxf(m4.z, m4.w));
    i0 = n4 << 2;
  }
  for (int i = i0 + tid; i < n; i += nthr) m = fmaxf(m, fabsf(w[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(max_bits + b, __float_as_uint(m));
}
__global__ void fill_u32_kernel(unsigned* p, int n, unsigned v) {
  griddep_launch_dependents();
  griddep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ------------------------------------------------------------------ STFT + |.| + per-clip max (enhancer.py:82-101)
// Every warp transforms TWO real frames with one complex FFT (frame a in the real parts, frame b in the imaginary
// parts): A[k] = (Z[k] + conj Z[N-k]) / 2, B[k] = (Z[k] - conj Z[N-k]) / 2i.  Half the butterflies per frame.
__global__ void __launch_bounds__(FR * 32, 4) stft_kernel(const float* __restrict__ wave, int n_pitch, int T_pitch,
                                                       const unsigned* __restrict__ max_bits,
                                                       float2* __restrict__ spec, float* __restrict__ mag,
                                                       unsigned* __restrict__ mag_max_bits, const int* __restrict__ geo) {
  griddep_launch_dependents();
  griddep_wait();
  extern __shared__ float2 sm[];
  float2* xs = sm;                // [FR][XS]: one complex buffer per frame pair
  const int b = blockIdx.y, t0 = blockIdx.x * FRAMES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_mv = 1.0f / guard_scalar(max_bits[b]);
  // variable-length batch: this clip's own sample / frame counts (its frames >= T are written as zero magnitude, so
  // the padded clip looks to the encoder exactly like the reference's zero padding at the right border)
  const int n = geo != nullptr ? geo[b * GEO_STRIDE + GEO_N] : n_pitch;
  const int T = geo != nullptr ? geo[b * GEO_STRIDE + GEO_T] : T_pitch;
  const int ta = t0 + 2 * warp;
  float2* x = xs + warp * XS;
  if (ta < T) {
    const float* w = wave + static_cast<long long>(b) * n_pitch;
    const bool has_b = ta + 1 < T;
    const int s0 = ta * HOP - NFFT / 2;  // first sample of frame a (frame b starts HOP later); centred, zero padded
    if (s0 >= 0 && s0 + HOP + NFFT <= n && has_b) {
      // interior pair (all but the first / last frames of a clip): no range checks
      const float* wa = w + s0;
#pragma unroll 4
      for (int i = lane; i < NFFT; i += 32) {
        const float h = hann512(i);
        x[fpad(i)] = make_float2((wa[i] * inv_mv) * h, (wa[i + HOP] * inv_mv) * h);
      }
    } else {
      for (int i = lane; i < NFFT; i += 32) {
        const int sa = s0 + i, sb = sa + HOP;
        const float h = hann512(i);
        float va = 0.f, vb = 0.f;
        if (sa >= 0 && sa < n) va = (w[sa] * inv_mv) * h;
        if (has_b && sb >= 0 && sb < n) vb = (w[sb] * inv_mv) * h;
        x[fpad(i)] = make_float2(va, vb);
      }
    }
  }
  __syncwarp();
  if (ta < T) fft512_warp<false>(x, lane);
  __syncthreads();
  // output: thread -> (frame tl, bins fq, fq + 16, ...); 16 consecutive frames of a bin are one 128-byte store
  float lmax = 0.f;
  {
    const int tl = threadIdx.x & (FRAMES - 1), fq = threadIdx.x / FRAMES;
    if (t0 + tl >= T && t0 + tl < T_pitch) {   // padding frames of a shorter clip in a variable-length batch
      const long long o0 = (static_cast<long long>(b) * NBIN + fq) * T_pitch + t0 + tl;
      const long long step = static_cast<long long>(FR * 32 / FRAMES) * T_pitch;
      long long o = o0;
      for (int f = fq; f < NBIN; f += FR * 32 / FRAMES, o += step) {
        if (spec != nullptr) spec[o] = make_float2(0.f, 0.f);
        mag[o] = 0.f;
      }
    }
    if (t0 + tl < T) {
      const float2* xw = xs + (tl >> 1) * XS;
      const bool second = (tl & 1) != 0;
      const long long o0 = (static_cast<long long>(b) * NBIN + fq) * T_pitch + t0 + tl;
      float2* sp = spec + o0;
      float* mp = mag + o0;
      const long long step = static_cast<long long>(FR * 32 / FRAMES) * T_pitch;
      for (int f = fq; f < NBIN; f += FR * 32 / FRAMES, sp += step, mp += step) {
        const float2 z1 = xw[fpad(f)], z2 = xw[fpad((NFFT - f) & (NFFT - 1))];
        const float2 z = second ? make_float2(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x))
                                : make_float2(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
        if (spec != nullptr) *sp = z;   // (the enhance path never stores the complex spectrogram)
        const float m = sqrtf(z.x * z.x + z.y * z.y);
        *mp = m;
        lmax = fmaxf(lmax, m);
      }
    }
  }
  lmax = warp_max(lmax);
  if (lane == 0) atomicMax(mag_max_bits + b, __float_as_uint(lmax));
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

// ------------------------------------------------------------------ iSTFT part 1: spectrum -> windowed frames
// E = (model_out * mag_max) * S/|S|  (== mag * exp(1j*angle(S)), enhancer.py:115-119), irfft-512, * Hann.
// When `lowres` is given the model output is bilinearly sampled from the decoder's [B,Hs,Ws] tanh map on the fly
// (HybridViT's final F.interpolate, hybrid_vit.py:458-465, fused here) and also written to model_out.
__global__ void __launch_bounds__(FR * 32, 4) istft_frames_kernel(float* __restrict__ model_out,
                                                               const float* __restrict__ lowres, int Hs, int Ws,
                                                               const float2* __restrict__ spec,
                                                               const unsigned* __restrict__ mag_max_bits, int T,
                                                               float* __restrict__ frames) {
  griddep_launch_dependents();
  griddep_wait();
  extern __shared__ float2 sm[];
  float2* xs = sm;
  const int b = blockIdx.y, t0 = blockIdx.x * FRAMES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float mm = guard_scalar(mag_max_bits[b]);
  // Two frames per inverse FFT: Z = Ea + i Eb over the full (Hermitian-filled) spectrum; the real part of the
  // transform is frame a, the imaginary part frame b.  Thread -> (frame pair q, bins fq, fq + 32, ...): the time
  // interpolation weights are per thread, the frequency ones per iteration (shared by both frames).
  {
    const int q = threadIdx.x & (FR - 1), fq = threadIdx.x / FR;
    const int qa = t0 + 2 * q;
    if (qa < T) {
      const bool has_b = qa + 1 < T;
      const Lerp lxa = make_lerp(qa, Ws > 0 ? Ws : 1, T), lxb = make_lerp(has_b ? qa + 1 : qa, Ws > 0 ? Ws : 1, T);
      const float* src = lowres != nullptr ? lowres + static_cast<long long>(b) * Hs * Ws : nullptr;
      float2* x = xs + q * XS;
      const long long o0 = (static_cast<long long>(b) * NBIN + fq) * T + qa;
      const float2* sp = spec + o0;
      float* mp = model_out + o0;
      const long long step = static_cast<long long>(FR * 32 / FR) * T;
      // E = (model_out * mag_max) * S / |S| of one bin (z = S, mo = model output)
      auto bin = [&](float2 z, float mo, int f) -> float2 {
        const float zz = z.x * z.x + z.y * z.y;
        const float e = mo * mm;
        float2 E;
        if (zz >= 1.17549435e-38f) {  // e * S / |S| with one MUFU.RSQ (2^-22 relative) instead of sqrt + divide
          const float ea = e * rsqrtf(zz);
          E = make_float2(ea * z.x, ea * z.y);
        } else {                       // |S|^2 underflows or S == 0 (angle 0): exact path
          const float a = sqrtf(zz);
          const float ea = a > 0.f ? e / a : 0.f;
          E = a > 0.f ? make_float2(ea * z.x, ea * z.y) : make_float2(e, 0.f);
        }
        if (f == 0 || f == NFFT / 2) E.y = 0.f;  // c2r transforms ignore the imaginary part of DC / Nyquist
        return E;
      };
      for (int f = fq; f < NBIN; f += 32, sp += step, mp += step) {
        const float2 za = sp[0];
        const float2 zb = has_b ? sp[1] : make_float2(0.f, 0.f);
        float moa, mob = 0.f;
        if (src != nullptr) {
          const Lerp ly = make_lerp(f, Hs, NBIN);
          const float* r0 = src + ly.i0 * Ws;
          const float* r1 = src + ly.i1 * Ws;
          moa = ly.l0 * (lxa.l0 * r0[lxa.i0] + lxa.l1 * r0[lxa.i1]) + ly.l1 * (lxa.l0 * r1[lxa.i0] + lxa.l1 * r1[lxa.i1]);
          mp[0] = moa;
          if (has_b) {
            mob = ly.l0 * (lxb.l0 * r0[lxb.i0] + lxb.l1 * r0[lxb.i1]) + ly.l1 * (lxb.l0 * r1[lxb.i0] + lxb.l1 * r1[lxb.i1]);
            mp[1] = mob;
          }
        } else {
          moa = mp[0];
          if (has_b) mob = mp[1];
        }
        const float2 Ea = bin(za, moa, f);
        const float2 Eb = has_b ? bin(zb, mob, f) : make_float2(0.f, 0.f);
        x[fpad(f)] = make_float2(Ea.x - Eb.y, Ea.y + Eb.x);
        if (f > 0 && f < NFFT / 2) x[fpad(NFFT - f)] = make_float2(Ea.x + Eb.y, Eb.x - Ea.y);
      }
    }
  }
  __syncthreads();
  const int ta = t0 + 2 * warp;
  if (ta < T) {
    float2* x = xs + warp * XS;
    fft512_warp<true>(x, lane);
    float* fa = frames + (static_cast<long long>(b) * T + ta) * NFFT;
    const bool has_b = ta + 1 < T;
    for (int i = lane; i < NFFT; i += 32) {
      const float2 v = x[fpad(i)];
      const float h = hann512(i) * (1.0f / NFFT);
      fa[i] = v.x * h;
      if (has_b) fa[NFFT + i] = v.y * h;
    }
  }
}

// iSTFT part 2: overlap-add, trim n_fft/2, divide by the window sum-square envelope, de-normalise.
__global__ void istft_ola_kernel(const float* __restrict__ frames, const unsigned* __restrict__ max_bits, int n, int T,
                                 float* __restrict__ wave_out) {
  griddep_launch_dependents();
  griddep_wait();
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

// ------------------------------------------------------------------ fused back end of AudioEnhancer.enhance
// enhancer.py:115-133: E = (model_out * mag_max) * exp(1j * angle(S)), librosa.istft (irfft-512, periodic Hann,
// overlap-add, trim n_fft/2, divide by the window sum-square envelope), * max_val - in ONE kernel that touches HBM only
// for its algorithmic bytes: the noisy waveform in, the decoder's low-resolution tanh map in, the waveform out.
//   * the noisy phase is RECOMPUTED here (forward FFT of the same windowed frames as stft_kernel - identical code,
//     identical bits), so the complex spectrogram [B,257,T] is never written or read (66 MB each way at 64 x 4 s);
//   * HybridViT's final bilinear resize (hybrid_vit.py:458-465) is sampled per bin from the [B,Hs,Ws] map;
//   * overlap-add runs out of shared memory: a block transforms FRAMES = 16 consecutive frames and owns the
//     FRAMES - 3 = 13 hops of output that only those frames touch (every sample is covered by <= 4 frames), so
//     neighbouring blocks recompute 3 frames instead of exchanging partial sums - no [B,T,512] frame buffer
//     (66 MB written + read before), no atomics.
// Per warp: load + window two frames (real / imaginary part of one complex buffer) -> forward FFT -> per bin: untangle
// the two spectra, E = mo * mag_max * S / |S| for both, re-tangle as Z = Ea + i Eb (in place: bin f and N - f belong to
// the same lane) -> inverse FFT; then the block overlap-adds its hop range.
// FRI = warps per block (each transforms a pair of frames): 2 * FRI frames per block, 2 * FRI - 3 hops of output.
// FRI = 16: 29 of 32 transformed frames are "new" (1.10x redundant FFT work; 1.23x at FRI = 8).
template <int FRI>
__global__ void __launch_bounds__(FRI * 32, FRI == 16 ? 2 : 4) enhance_istft_kernel(const float* __restrict__ wave, int n_pitch, int T_pitch,
                                                                const unsigned* __restrict__ max_bits,
                                                                const unsigned* __restrict__ mag_max_bits,
                                                                const float* __restrict__ lowres, int Hs, int Ws_pitch,
                                                                float* __restrict__ model_out,
                                                                float* __restrict__ wave_out,
                                                                const int* __restrict__ geo, int geo_ws_idx) {
  griddep_launch_dependents();
  griddep_wait();
  constexpr int NFR = 2 * FRI, HOPS = NFR - 3;
  extern __shared__ float2 sm[];
  float2* xs = sm;                                            // [FRI][XS]
  int2* lyi = reinterpret_cast<int2*>(xs + FRI * XS);         // [NBIN] source rows (i0, i1) of the frequency resize
  float* lyw = reinterpret_cast<float*>(lyi + NBIN);          // [NBIN] weight of row i1
  float* cols = lyw + NBIN + 3;                               // [NFR][Hs] decoder map resized along time, per frame
  const int b = blockIdx.y, t0 = blockIdx.x * HOPS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float mv = guard_scalar(max_bits[b]);
  const float inv_mv = 1.0f / mv;
  const float mm = guard_scalar(mag_max_bits[b]);
  // variable-length batch: this clip's own sample count, frame count and decoder-output width (the resize of the
  // reference maps ITS [Hs, Ws_b] map onto ITS [257, T_b] spectrogram); samples beyond n_b are written as zero
  const int n = geo != nullptr ? geo[b * GEO_STRIDE + GEO_N] : n_pitch;
  const int T = geo != nullptr ? geo[b * GEO_STRIDE + GEO_T] : T_pitch;
  const int Ws = geo != nullptr ? geo[b * GEO_STRIDE + geo_ws_idx] : Ws_pitch;
  for (int f = threadIdx.x; f < NBIN; f += FRI * 32) {
    const Lerp ly = make_lerp(f, Hs, NBIN);
    lyi[f] = make_int2(ly.i0, ly.i1);
    lyw[f] = ly.l1;
  }
  {
    // HybridViT's final bilinear resize (hybrid_vit.py:458-465), time axis first: cols[t][h] = the decoder map's row h
    // interpolated at frame t0 + t - every (bin, frame) below then needs two shared-memory reads instead of four global
    // ones (same association as the stand-alone resize: ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11))
    const float* src = lowres + static_cast<long long>(b) * Hs * Ws_pitch;
    for (int i = threadIdx.x; i < NFR * Hs; i += FRI * 32) {
      const int t = i / Hs, h = i - t * Hs;
      float v = 0.f;
      if (t0 + t < T) {
        const Lerp lx = make_lerp(t0 + t, Ws, T);
        const float* r = src + h * Ws_pitch;
        v = lx.l0 * __ldg(r + lx.i0) + lx.l1 * __ldg(r + lx.i1);
      }
      cols[i] = v;
    }
  }
  const int ta = t0 + 2 * warp;
  const bool live = ta < T, has_b = ta + 1 < T;
  float2* x = xs + warp * XS;
  if (live) {
    // ---- frames ta, ta + 1 of the peak-normalised noisy clip (same arithmetic as stft_kernel)
    const float* w = wave + static_cast<long long>(b) * n_pitch;
    const int s0 = ta * HOP - NFFT / 2;
    if (s0 >= 0 && s0 + HOP + NFFT <= n && has_b) {
      const float* wa = w + s0;
#pragma unroll 4
      for (int i = lane; i < NFFT; i += 32) {
        const float h = hann512(i);
        x[fpad(i)] = make_float2((wa[i] * inv_mv) * h, (wa[i + HOP] * inv_mv) * h);
      }
    } else {
      for (int i = lane; i < NFFT; i += 32) {
        const int sa = s0 + i, sb = sa + HOP;
        const float h = hann512(i);
        float va = 0.f, vb = 0.f;
        if (sa >= 0 && sa < n) va = (w[sa] * inv_mv) * h;
        if (has_b && sb >= 0 && sb < n) vb = (w[sb] * inv_mv) * h;
        x[fpad(i)] = make_float2(va, vb);
      }
    }
    __syncwarp();
    fft512_warp<false>(x, lane);
  }
  __syncthreads();  // resize tables complete
  if (live) {
    const float* ca = cols + (2 * warp) * Hs;
    const float* cb = ca + Hs;
    // E = (model_out * mag_max) * S / |S| of one bin (z = S, mo = model output)
    auto bin = [&](float2 z, float mo, int f) -> float2 {
      const float zz = z.x * z.x + z.y * z.y;
      const float e = mo * mm;
      float2 E;
      if (zz >= 1.17549435e-38f) {  // e * S / |S| with one MUFU.RSQ (2^-22 relative) instead of sqrt + divide
        const float ea = e * rsqrtf(zz);
        E = make_float2(ea * z.x, ea * z.y);
      } else {                       // |S|^2 underflows or S == 0 (angle 0): exact path
        const float a = sqrtf(zz);
        const float ea = a > 0.f ? e / a : 0.f;
        E = a > 0.f ? make_float2(ea * z.x, ea * z.y) : make_float2(e, 0.f);
      }
      if (f == 0 || f == NFFT / 2) E.y = 0.f;  // c2r transforms ignore the imaginary part of DC / Nyquist
      return E;
    };
    for (int f = lane; f < NBIN; f += 32) {
      const int fm = (NFFT - f) & (NFFT - 1);
      const float2 z1 = x[fpad(f)], z2 = x[fpad(fm)];
      const float2 za = make_float2(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
      const float2 zb = make_float2(0.5f * (z1.y + z2.y), 0.5f * (z2.x - z1.x));
      const int2 ro = lyi[f];
      const float wy1 = lyw[f], wy0 = 1.0f - wy1;
      const float moa = wy0 * ca[ro.x] + wy1 * ca[ro.y];
      const float mob = has_b ? wy0 * cb[ro.x] + wy1 * cb[ro.y] : 0.f;
      if (model_out != nullptr) {  // debug / test builds of the plan only: the resized model output [B,257,T]
        float* mp = model_out + (static_cast<long long>(b) * NBIN + f) * T_pitch + ta;
        mp[0] = moa;
        if (has_b) mp[1] = mob;
      }
      const float2 Ea = bin(za, moa, f);
      const float2 Eb = has_b ? bin(zb, mob, f) : make_float2(0.f, 0.f);
      x[fpad(f)] = make_float2(Ea.x - Eb.y, Ea.y + Eb.x);
      if (f > 0 && f < NFFT / 2) x[fpad(fm)] = make_float2(Ea.x + Eb.y, Eb.x - Ea.y);
    }
    __syncwarp();
    fft512_warp<true>(x, lane);
  }
  __syncthreads();
  // ---- overlap-add of this block's hop range straight out of shared memory, envelope division, de-normalisation
  const int j_lo = blockIdx.x == 0 ? NFFT / 2 : HOP * (t0 + 3);
  const int j_hi = min(HOP * (t0 + NFR), NFFT / 2 + n_pitch);
  float* out = wave_out + static_cast<long long>(b) * n_pitch - NFFT / 2;
  for (int j = j_lo + threadIdx.x; j < j_hi; j += FRI * 32) {
    if (j >= NFFT / 2 + n) {  // padding of a shorter clip (variable-length batch)
      out[j] = 0.f;
      continue;
    }
    const int h = j >> 7;
    const int thi = min(h, T - 1), tlo = max(h - 3, 0);
    float acc = 0.f, wss = 0.f;
    for (int t = tlo; t <= thi; ++t) {
      const int k = j - t * HOP, tl = t - t0;
      const float2 v = xs[(tl >> 1) * XS + fpad(k)];
      const float w = hann512(k);
      acc += ((tl & 1) ? v.y : v.x) * (w * (1.0f / NFFT));
      wss += w * w;
    }
    if (wss > 1.17549435e-38f) acc /= wss;
    out[j] = acc * mv;
  }
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
  griddep_launch_dependents();
  griddep_wait();
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
// V float4 per lane (D = 128 * V): the row is read once into registers; mean / centred variance in fp32 via warp
// shuffles (same two-pass arithmetic as torch), gamma/beta through the read-only path.  V == 0: generic D.
template <typename T, int V>
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bb,
                                 T* __restrict__ out, int rows, int D, float eps) {
  griddep_launch_dependents();
  griddep_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<long long>(row) * D;
  T* orow = out + static_cast<long long>(row) * D;
  if (V > 0) {
    float4 v[V > 0 ? V : 1];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      v[k] = *reinterpret_cast<const float4*>(xr + k * 128 + lane * 4);
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      v[k].x -= mean; v[k].y -= mean; v[k].z -= mean; v[k].w -= mean;
      q += (v[k].x * v[k].x + v[k].y * v[k].y) + (v[k].z * v[k].z + v[k].w * v[k].w);
    }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int i = k * 128 + lane * 4;
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g + i));
      const float4 be = __ldg(reinterpret_cast<const float4*>(bb + i));
      const float yv[4] = {fmaf(v[k].x * rstd, gg.x, be.x), fmaf(v[k].y * rstd, gg.y, be.y),
                           fmaf(v[k].z * rstd, gg.z, be.z), fmaf(v[k].w * rstd, gg.w, be.w)};
      st4<T>(orow + i, yv);
    }
    return;
  }
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
  for (int i = lane * 4; i < D; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + i);
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + i));
    const float4 be = __ldg(reinterpret_cast<const float4*>(bb + i));
    const float yv[4] = {(v.x - mean) * rstd * gg.x + be.x, (v.y - mean) * rstd * gg.y + be.y,
                         (v.z - mean) * rstd * gg.z + be.z, (v.w - mean) * rstd * gg.w + be.w};
    st4<T>(orow + i, yv);
  }
}

// LayerNorm folded into the GEMMs (IgemmParams::ln_*): what is left of a stand-alone nn.LayerNorm is the 16-bit copy of
// the row and its statistics in the producer epilogue's format - `slots` (mean, M2) pairs per row, each standing for
// D / slots columns; here every slot carries the row's exact two-pass mean and an equal share of its M2.  Used for the
// first block (the residual stream comes from the patch embedding / the variable-length token compaction).
template <typename T, int V>
__global__ void rowstats_kernel(const float* __restrict__ x, T* __restrict__ x16, float2* __restrict__ stats, int rows,
                                int D, int slots) {
  griddep_launch_dependents();
  griddep_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<long long>(row) * D;
  T* orow = x16 + static_cast<long long>(row) * D;
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    v[k] = *reinterpret_cast<const float4*>(xr + k * 128 + lane * 4);
    s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    const float yv[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
    st4<T>(orow + k * 128 + lane * 4, yv);
  }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  q = warp_sum(q);
  if (lane < slots) stats[static_cast<long long>(row) * slots + lane] = make_float2(mean, q / slots);
}

// Weights of a linear layer that follows a LayerNorm, for the folded form (one warp per output feature n):
//   W''[n,k] = 16-bit(gamma[k] W[n,k] - mean_k(gamma W[n,:]))   (centred over k: sum_k x_k W''[n,k] = sum_k (x_k - mean x) gamma_k W[n,k])
//   c[n]     = bias[n] + sum_k beta[k] W[n,k]
template <typename T>
__global__ void ln_fold_kernel(const T* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ bias, T* __restrict__ wp, float* __restrict__ c, int N, int K) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float gs = 0.f, cs = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = static_cast<float>(w[static_cast<long long>(n) * K + k]);
    gs = fmaf(gamma[k], wv, gs);
    cs = fmaf(beta[k], wv, cs);
  }
  gs = warp_sum(gs);
  cs = warp_sum(cs);
  const float mean = gs / K;
  for (int k = lane; k < K; k += 32) {
    const float wv = static_cast<float>(w[static_cast<long long>(n) * K + k]);
    wp[static_cast<long long>(n) * K + k] = static_cast<T>(fmaf(gamma[k], wv, -mean));
  }
  if (lane == 0) c[n] = (bias != nullptr ? bias[n] : 0.f) + cs;
}

// skip feature [B, Hs(pitch), Ws, C] -> bilinear sample at decoder resolution [B*Hd*Wd, C]
// (the 1x1 projection is applied AFTER sampling; exact because bilinear weights sum to 1 - hybrid_vit.py:377-386)
template <typename T>
__global__ void skip_sample_kernel(const T* __restrict__ src, int Hs, int HsPitch, int Ws, int C, int Hd, int Wd,
                                   T* __restrict__ dst, long long total, const int* __restrict__ geo, int gsrc, int gdst) {
  griddep_launch_dependents();
  griddep_wait();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = C / 4;
  const int c = static_cast<int>(idx % cv) * 4;
  long long r = idx / cv;
  const int wd = static_cast<int>(r % Wd);
  r /= Wd;
  const int hd = static_cast<int>(r % Hd);
  const int b = static_cast<int>(r / Hd);
  T* out = dst + ((static_cast<long long>(b) * Hd + hd) * Wd + wd) * C + c;
  int Ws_b = Ws, Wd_b = Wd;
  if (geo != nullptr) {  // variable-length batch: this clip's own source / destination widths; zero beyond
    Ws_b = geo[b * GEO_STRIDE + gsrc];
    Wd_b = geo[b * GEO_STRIDE + gdst];
    if (wd >= Wd_b) {
      const float z[4] = {0.f, 0.f, 0.f, 0.f};
      st4<T>(out, z);
      return;
    }
  }
  const Lerp ly = make_lerp(hd, Hs, Hd), lx = make_lerp(wd, Ws_b, Wd_b);
  const T* base = src + static_cast<long long>(b) * HsPitch * Ws * C + c;
  float v00[4], v01[4], v10[4], v11[4], o[4];
  ld4<T>(base + (static_cast<long long>(ly.i0) * Ws + lx.i0) * C, v00);
  ld4<T>(base + (static_cast<long long>(ly.i0) * Ws + lx.i1) * C, v01);
  ld4<T>(base + (static_cast<long long>(ly.i1) * Ws + lx.i0) * C, v10);
  ld4<T>(base + (static_cast<long long>(ly.i1) * Ws + lx.i1) * C, v11);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    o[j] = ly.l0 * (lx.l0 * v00[j] + lx.l1 * v01[j]) + ly.l1 * (lx.l0 * v10[j] + lx.l1 * v11[j]);
  st4<T>(out, o);
}

// ------------------------------------------------------------------ head: Conv3x3(C->1, no bias) + tanh, fp32 accumulate
// One thread per output pixel, 32x8 pixel tiles per block (neighbouring threads re-read the same input pixels
// through L1), 16-byte loads, weights broadcast from shared memory.
template <typename T>
__device__ __forceinline__ void ld8(const T* p, float* v);
template <>
__device__ __forceinline__ void ld8<float>(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void ld8<bf16>(const bf16* p, float* v) {
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
template <>
__device__ __forceinline__ void ld8<__half>(const __half* p, float* v) {
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

// 16-bit variant of skip_sample_kernel with 8 channels (one 16-byte load per tap, one 16-byte store) per thread
template <typename T>
__global__ void skip_sample8_kernel(const T* __restrict__ src, int Hs, int HsPitch, int Ws, int C, int Hd, int Wd,
                                    T* __restrict__ dst, long long total, const int* __restrict__ geo, int gsrc, int gdst) {
  griddep_launch_dependents();
  griddep_wait();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = C / 8;
  const int c = static_cast<int>(idx % cv) * 8;
  long long r = idx / cv;
  const int wd = static_cast<int>(r % Wd);
  r /= Wd;
  const int hd = static_cast<int>(r % Hd);
  const int b = static_cast<int>(r / Hd);
  T* out = dst + ((static_cast<long long>(b) * Hd + hd) * Wd + wd) * C + c;
  int Ws_b = Ws, Wd_b = Wd;
  if (geo != nullptr) {  // variable-length batch: this clip's own source / destination widths; zero beyond
    Ws_b = geo[b * GEO_STRIDE + gsrc];
    Wd_b = geo[b * GEO_STRIDE + gdst];
    if (wd >= Wd_b) {
      *reinterpret_cast<uint4*>(out) = make_uint4(0, 0, 0, 0);
      return;
    }
  }
  const Lerp ly = make_lerp(hd, Hs, Hd), lx = make_lerp(wd, Ws_b, Wd_b);
  const T* base = src + static_cast<long long>(b) * HsPitch * Ws * C + c;
  float v00[8], v01[8], v10[8], v11[8], o[8];
  ld8<T>(base + (static_cast<long long>(ly.i0) * Ws + lx.i0) * C, v00);
  ld8<T>(base + (static_cast<long long>(ly.i0) * Ws + lx.i1) * C, v01);
  ld8<T>(base + (static_cast<long long>(ly.i1) * Ws + lx.i0) * C, v10);
  ld8<T>(base + (static_cast<long long>(ly.i1) * Ws + lx.i1) * C, v11);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    o[j] = ly.l0 * (lx.l0 * v00[j] + lx.l1 * v01[j]) + ly.l1 * (lx.l0 * v10[j] + lx.l1 * v11[j]);
  constexpr int f16 = sizeof(T) == 2 && !std::is_same<T, bf16>::value ? 1 : 0;
  uint4 pk;
  pk.x = pack_16x2(o[0], o[1], f16); pk.y = pack_16x2(o[2], o[3], f16);
  pk.z = pack_16x2(o[4], o[5], f16); pk.w = pack_16x2(o[6], o[7], f16);
  *reinterpret_cast<uint4*>(out) = pk;
}

// 8 lanes per output pixel, each lane owns 8 channels (one 16-byte load per tap): a warp instruction reads 4 whole
// 128-byte pixel vectors (4 L1 wavefronts instead of 32 with a pixel-per-lane mapping - ncu showed that version
// L1-bound at 84 %).  The lane's 72 weights live in registers; the partial sums meet with three shuffles.
template <typename T>
__global__ void __launch_bounds__(256) head_kernel(const T* __restrict__ x, const float* __restrict__ w9c, int B, int H,
                                                   int W, int C, float* __restrict__ logits,
                                                   float* __restrict__ out_tanh) {
  griddep_launch_dependents();
  griddep_wait();
  const int lane8 = threadIdx.x & 7;
  const long long npix = static_cast<long long>(B) * H * W;
  const long long stride = static_cast<long long>(gridDim.x) * 32;  // pixels per grid sweep (32 per block)
  for (long long pix0 = static_cast<long long>(blockIdx.x) * 32 + (threadIdx.x >> 3); pix0 - (threadIdx.x >> 3) < npix;
       pix0 += stride) {
    const bool live = pix0 < npix;
    const long long pix = live ? pix0 : 0;
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    const int b = static_cast<int>(pix / (static_cast<long long>(W) * H));
    float acc = 0.f;
    for (int c = lane8 * 8; c < C; c += 64) {  // 64 channels per pass of the 8 lanes
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int iy = h + tap / 3 - 1, ix = w + tap % 3 - 1;
        if (!live || iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
        float v[8];
        ld8<T>(x + ((static_cast<long long>(b) * H + iy) * W + ix) * C + c, v);
        const float4 wa = __ldg(reinterpret_cast<const float4*>(w9c + tap * C + c));
        const float4 wb = __ldg(reinterpret_cast<const float4*>(w9c + tap * C + c + 4));
        acc = fmaf(v[0], wa.x, acc); acc = fmaf(v[1], wa.y, acc); acc = fmaf(v[2], wa.z, acc); acc = fmaf(v[3], wa.w, acc);
        acc = fmaf(v[4], wb.x, acc); acc = fmaf(v[5], wb.y, acc); acc = fmaf(v[6], wb.z, acc); acc = fmaf(v[7], wb.w, acc);
      }
    }
    acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 1);
    acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 2);
    acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 4);
    if (live && lane8 == 0) {
      if (logits != nullptr) logits[pix] = acc;
      out_tanh[pix] = tanhf(acc);
    }
  }
}

// Row-streaming head for the 16-bit modes with C = 64: one block per (clip, strip of HEAD_STRIP output rows) slides a
// window of input rows (contiguous W*C*2 bytes each in NHWC) through a 4-slot shared-memory ring filled by 16-byte
// async copies - every input row is read once per strip, fully coalesced, and row h+2 streams in while row h is
// computed.  8 lanes per output pixel (8 channels = one conflict-free 16-byte shared-memory load per tap) accumulate
// in fp32 against register-resident weights.  (The per-pixel global-load kernel above was latency / L1 bound:
// 116 us for 65 MB of input.)
constexpr int HEAD_STRIP = 8;  // default strip height; the launcher shortens it when that balances the grid better
template <typename T>
__global__ void __launch_bounds__(256) head_rows_kernel(const T* __restrict__ x, const float* __restrict__ w9c, int H, int W,
                                                        float* __restrict__ logits, float* __restrict__ out_tanh, int strip) {
  griddep_launch_dependents();
  griddep_wait();
  constexpr int C = 64;
  extern __shared__ __align__(16) uint8_t hsm[];
  const int h0 = blockIdx.x * strip, b = blockIdx.y;
  const int h1 = min(h0 + strip, H);
  const int row_bytes = W * C * 2;
  // input row ih -> ring slot (ih + 1) & 3; rows outside the image are zero (conv padding)
  auto load_row = [&](int ih) {
    uint8_t* dst = hsm + ((ih + 1) & 3) * row_bytes;
    if (ih >= 0 && ih < H) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(x + (static_cast<long long>(b) * H + ih) * W * C);
      for (int o = threadIdx.x * 16; o < row_bytes; o += 256 * 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + o)), "l"(src + o) : "memory");
    } else {
      for (int o = threadIdx.x * 16; o < row_bytes; o += 256 * 16) *reinterpret_cast<uint4*>(dst + o) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_row(h0 - 1);
  load_row(h0);
  load_row(h0 + 1);
  const int lane8 = threadIdx.x & 7;
  float wr[9][8];  // this lane's 8 channels of the 9 taps
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w9c + tap * C + lane8 * 8));
    const float4 c = __ldg(reinterpret_cast<const float4*>(w9c + tap * C + lane8 * 8 + 4));
    wr[tap][0] = a.x; wr[tap][1] = a.y; wr[tap][2] = a.z; wr[tap][3] = a.w;
    wr[tap][4] = c.x; wr[tap][5] = c.y; wr[tap][6] = c.z; wr[tap][7] = c.w;
  }
  for (int h = h0; h < h1; ++h) {
    load_row(h + 2);  // slot of row h-2, whose last readers passed the barrier at the end of the previous iteration
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // rows <= h+1 have landed (this thread's copies)
    __syncthreads();                                       // ... and everybody else's
    // Each 8-lane group walks a run of 4 consecutive pixels with a 3-column window of fp32 values in registers:
    // every input vector is loaded and converted once per output row and group (1.5 instead of 9 loads + converts
    // per pixel and lane).  The lanes' partial sums of the 4 pixels meet in a butterfly (4 shuffles instead of 12).
    for (int wb0 = (threadIdx.x >> 5) * 16; wb0 < W; wb0 += 128) {  // (warp-uniform trip count: full-mask shuffles below)
      const int wb = wb0 + ((threadIdx.x >> 3) & 3) * 4;
      const T* r0 = reinterpret_cast<const T*>(hsm + ((h + 0) & 3) * row_bytes) + lane8 * 8;
      const T* r1 = reinterpret_cast<const T*>(hsm + ((h + 1) & 3) * row_bytes) + lane8 * 8;
      const T* r2 = reinterpret_cast<const T*>(hsm + ((h + 2) & 3) * row_bytes) + lane8 * 8;
      float win[3][3][8];  // [column slot][row][channel]
      auto load_col = [&](int slot, int ix) {
        if (ix >= 0 && ix < W) {
          ld8<T>(r0 + ix * C, win[slot][0]);
          ld8<T>(r1 + ix * C, win[slot][1]);
          ld8<T>(r2 + ix * C, win[slot][2]);
        } else {
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int e = 0; e < 8; ++e) win[slot][r][e] = 0.f;
        }
      };
      load_col(0, wb - 1);
      load_col(1, wb);
      float part[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        load_col((i + 2) % 3, wb + i + 1);
        float2 acc2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const float(&v)[8] = win[(i + tap % 3) % 3][tap / 3];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            acc2[e] = __ffma2_rn(make_float2(v[2 * e], v[2 * e + 1]), make_float2(wr[tap][2 * e], wr[tap][2 * e + 1]), acc2[e]);
        }
        part[i] = ((acc2[0].x + acc2[0].y) + (acc2[1].x + acc2[1].y)) + ((acc2[2].x + acc2[2].y) + (acc2[3].x + acc2[3].y));
      }
      // butterfly over the 8 lanes: afterwards every lane holds the finished pixel wb + (lane8 >> 1)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool up = (lane8 & 4) != 0;
        const float send = up ? part[i] : part[i + 2];
        const float keep = up ? part[i + 2] : part[i];
        part[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 4);
      }
      {
        const bool up = (lane8 & 2) != 0;
        const float send = up ? part[0] : part[1];
        const float keep = up ? part[1] : part[0];
        part[0] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 2);
      }
      part[0] += __shfl_xor_sync(0xFFFFFFFFu, part[0], 1);
      const int w0 = wb + (lane8 >> 1);
      if (w0 < W && (lane8 & 1) == 0) {
        const long long pix = (static_cast<long long>(b) * H + h) * W + w0;
        if (logits != nullptr) logits[pix] = part[0];
        out_tanh[pix] = tanhf(part[0]);
      }
    }
    __syncthreads();  // row h-1's slot may be overwritten by the next iteration's copy
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------ final bilinear resize [B,Hs,Ws] -> [B,Hd,Wd]
__global__ void resize_kernel(const float* __restrict__ src, int Hs, int Ws, float* __restrict__ dst, int Hd, int Wd,
                              long long total) {
  griddep_launch_dependents();
  griddep_wait();
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

// ------------------------------------------------------------------ variable-length batches (see kernels.h)
__global__ void varlen_geometry_kernel(const int* __restrict__ n_valid, int B, VarlenCfg c, int* __restrict__ geo) {
  griddep_launch_dependents();
  griddep_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int n = n_valid[b];
  n = n < c.n_min ? c.n_min : (n > c.n_max ? c.n_max : n);
  int* g = geo + b * GEO_STRIDE;
  const int T = 1 + n / HOP;
  g[GEO_N] = n;
  g[GEO_T] = T;
  int w = T;
  for (int i = 0; i < c.n_enc; ++i) {
    w /= c.enc_pool[i];
    g[GEO_ENC + i] = w;
  }
  const int wp = w / c.patch;
  g[GEO_WP] = wp;
  g[GEO_NTOK] = wp * c.Hp;
  int wx = wp;
  for (int i = 0; i < c.n_dec; ++i) {
    g[GEO_CAT + i] = wx;
    wx *= c.dec_up[i];
  }
}

// one block per (pixel row, clip): zero the bytes of pixels [W_b, Wmax)
__global__ void zero_cols_kernel(uint8_t* __restrict__ buf, int Hpitch, int Wmax, int pix_bytes, const int* __restrict__ geo,
                                 int geo_idx) {
  griddep_launch_dependents();
  griddep_wait();
  const int h = blockIdx.x, b = blockIdx.y;
  const int wb = geo[b * GEO_STRIDE + geo_idx];
  if (wb >= Wmax) return;
  uint8_t* row = buf + ((static_cast<long long>(b) * Hpitch + h) * Wmax + wb) * pix_bytes;
  const long long bytes = static_cast<long long>(Wmax - wb) * pix_bytes;   // multiple of 16 (pix_bytes is)
  for (long long o = static_cast<long long>(threadIdx.x) * 16; o < bytes; o += static_cast<long long>(blockDim.x) * 16)
    *reinterpret_cast<uint4*>(row + o) = make_uint4(0, 0, 0, 0);
}

// grid (Np, B), D / 4 threads: x[b][n] = grid[b][n / Wp_b][n % Wp_b] + pos[n]  (n < N_b), 0 otherwise
__global__ void tokens_compact_kernel(const float* __restrict__ grid, const float* __restrict__ pos, float* __restrict__ x,
                                      int Hp, int Wp, int D, const int* __restrict__ geo) {
  griddep_launch_dependents();
  griddep_wait();
  const int n = blockIdx.x, b = blockIdx.y;
  const int wpb = geo[b * GEO_STRIDE + GEO_WP], ntok = geo[b * GEO_STRIDE + GEO_NTOK];
  float4* dst = reinterpret_cast<float4*>(x + (static_cast<long long>(b) * Hp * Wp + n) * D);
  if (n >= ntok) {
    for (int i = threadIdx.x; i < D / 4; i += blockDim.x) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const int hq = n / wpb, wq = n - hq * wpb;
  const float4* src = reinterpret_cast<const float4*>(grid + ((static_cast<long long>(b) * Hp + hq) * Wp + wq) * D);
  const float4* pp = reinterpret_cast<const float4*>(pos + static_cast<long long>(n) * D);
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
    const float4 a = src[i], q = __ldg(pp + i);
    dst[i] = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
  }
}

// grid (Hp * Wp, B): cat[b][h][w][0:Cx] = rows[b][h * Wp_b + w][:]  (w < Wp_b), 0 otherwise; 16-byte vectors
__global__ void tofm_expand_kernel(const uint8_t* __restrict__ rows, uint8_t* __restrict__ cat, int Hp, int Wp, int row_bytes,
                                   int cat_pix_bytes, const int* __restrict__ geo) {
  griddep_launch_dependents();
  griddep_wait();
  const int pix = blockIdx.x, b = blockIdx.y;
  const int hq = pix / Wp, wq = pix - hq * Wp;
  const int wpb = geo[b * GEO_STRIDE + GEO_WP];
  uint4* dst = reinterpret_cast<uint4*>(cat + (static_cast<long long>(b) * Hp * Wp + pix) * cat_pix_bytes);
  if (wq >= wpb) {
    for (int i = threadIdx.x; i < row_bytes / 16; i += blockDim.x) dst[i] = make_uint4(0, 0, 0, 0);
    return;
  }
  const uint4* src = reinterpret_cast<const uint4*>(rows + (static_cast<long long>(b) * Hp * Wp + hq * wpb + wq) * row_bytes);
  for (int i = threadIdx.x; i < row_bytes / 16; i += blockDim.x) dst[i] = src[i];
}

constexpr int FFT_SMEM = FR * XS * sizeof(float2);

}  // namespace

// ================================================================== launchers
int launch_peak(const float* wave, int B, int n, float* max_val, int normalize, cudaStream_t s) {
  unsigned* bits = reinterpret_cast<unsigned*>(max_val);
  if (!normalize) {
    launch_pdl(fill_u32_kernel, dim3((B + 255) / 256), dim3(256), 0, s, bits, B, 0x3F800000u);  // 1.0f
    return check_launch("fill(max_val)");
  }
  launch_pdl(fill_u32_kernel, dim3((B + 255) / 256), dim3(256), 0, s, bits, B, 0u);
  if (n > 0) {
    dim3 grid(16, B);
    launch_pdl(peak_kernel, dim3(grid), dim3(256), 0, s, wave, n, bits);
  }
  return check_launch("peak");
}

// one-time table fill; called from plan creation / the stand-alone entry points (never inside a graph capture)
int ensure_fft_tables(cudaStream_t s) {
  static PerDeviceOnce once;  // the __device__ tables exist once per device
  if (once.first()) {
    fft_tables_kernel<<<2, 256, 0, s>>>();
    const int r = check_launch("fft_tables");
    if (r) {
      once.retry();
      return r;
    }
    // later launches on OTHER streams read the tables: order them after the fill (unless `s` is being captured, in which
    // case the fill is part of the captured work and the caller's stream order covers it)
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone) {
      if (cudaStreamSynchronize(s) != cudaSuccess) {
        once.retry();
        return check_launch("fft_tables(sync)");
      }
    }
    cudaGetLastError();
  }
  return 0;
}

int launch_stft(const float* wave, int B, int n, int T, const float* max_val, float2* spec, float* mag,
                unsigned* mag_max_bits, cudaStream_t s, const int* geo) {
  launch_pdl(fill_u32_kernel, dim3((B + 255) / 256), dim3(256), 0, s, mag_max_bits, B, 0u);
  dim3 grid((T + FRAMES - 1) / FRAMES, B);
  launch_pdl(stft_kernel, dim3(grid), dim3(FR * 32), FFT_SMEM, s, wave, n, T, reinterpret_cast<const unsigned*>(max_val), spec, mag,
                                               mag_max_bits, geo);
  return check_launch("stft");
}

static void fft_smem_config() {}

template <int FRI>
static int launch_enhance_istft_t(const float* wave_in, const float* max_val, const unsigned* mag_max_bits, const float* lowres,
                                  int Hs, int Ws, float* model_out, float* wave_out, int B, int n, int T, cudaStream_t s,
                                  const int* geo, int geo_ws_idx) {
  constexpr int HOPS = 2 * FRI - 3;
  const size_t smem = static_cast<size_t>(FRI) * XS * sizeof(float2) + NBIN * (sizeof(int2) + sizeof(float)) + 16 +
                      static_cast<size_t>(2 * FRI) * Hs * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("enhance_istft: decoder map too tall (%d rows)", Hs);
    return -1;
  }
  static PerDeviceOnce once;
  if (once.first() && cudaFuncSetAttribute(enhance_istft_kernel<FRI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
    once.retry();
    return check_launch("enhance_istft(attr)");
  }
  const int hops = (NFFT / 2 - 1 + n) >> 7;                        // index of the last hop that holds output
  const int blocks = hops >= 3 ? (hops - 3) / HOPS + 1 : 1;
  launch_pdl(enhance_istft_kernel<FRI>, dim3(blocks, B), dim3(FRI * 32), smem, s, wave_in, n, T,
             reinterpret_cast<const unsigned*>(max_val), mag_max_bits, lowres, Hs, Ws, model_out, wave_out, geo, geo_ws_idx);
  return check_launch("enhance_istft");
}

int launch_enhance_istft(const float* wave_in, const float* max_val, const unsigned* mag_max_bits, const float* lowres,
                         int Hs, int Ws, float* model_out, float* wave_out, int B, int n, int T, cudaStream_t s,
                         const int* geo, int geo_ws_idx) {
  if (n <= 0) return 0;
  // 16 frames per block by default (measured at 64 x 4 s: 106-109 us; 32-frame blocks - HVIT_ISTFT_FR=16, 10 % less
  // redundant FFT work but half the blocks per SM - 112-114 us)
  static const int fri = getenv("HVIT_ISTFT_FR") != nullptr ? atoi(getenv("HVIT_ISTFT_FR")) : 8;
  if (fri == 16)
    return launch_enhance_istft_t<16>(wave_in, max_val, mag_max_bits, lowres, Hs, Ws, model_out, wave_out, B, n, T, s, geo, geo_ws_idx);
  return launch_enhance_istft_t<8>(wave_in, max_val, mag_max_bits, lowres, Hs, Ws, model_out, wave_out, B, n, T, s, geo, geo_ws_idx);
}

int launch_istft_frames(float* model_out, const float* lowres, int Hs, int Ws, const float2* spec,
                        const unsigned* mag_max_bits, float* frames, int B, int T, cudaStream_t s) {
  dim3 grid((T + FRAMES - 1) / FRAMES, B);
  launch_pdl(istft_frames_kernel, dim3(grid), dim3(FR * 32), FFT_SMEM, s, model_out, lowres, Hs, Ws, spec, mag_max_bits, T, frames);
  return check_launch("istft_frames");
}

int launch_istft_ola(const float* frames, const float* max_val, float* wave_out, int B, int n, int T, cudaStream_t s) {
  if (n <= 0) return 0;
  dim3 g2((n + 255) / 256, B);
  launch_pdl(istft_ola_kernel, dim3(g2), dim3(256), 0, s, frames, reinterpret_cast<const unsigned*>(max_val), n, T, wave_out);
  return check_launch("istft_ola");
}

template <typename T>
static void stem_dispatch(const float* x, const unsigned* mm, const float* w, const float* scale, const float* shift,
                          void* out, int H, int W, int C, int pool, dim3 grid, dim3 block, int smem, cudaStream_t s) {
  const int Ho = H / pool, Wo = W / pool;
  if (pool == 2)
    launch_pdl(stem_kernel<T, 2>, dim3(grid), dim3(block), smem, s, x, mm, w, scale, shift, reinterpret_cast<T*>(out), H, W, C, Ho, Wo);
  else
    launch_pdl(stem_kernel<T, 1>, dim3(grid), dim3(block), smem, s, x, mm, w, scale, shift, reinterpret_cast<T*>(out), H, W, C, Ho, Wo);
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

template <typename T>
static void ln_dispatch(const float* x, const float* g, const float* b, void* out, int rows, int D, float eps,
                        cudaStream_t s) {
  const int grid = (rows + 7) / 8;
  T* o = reinterpret_cast<T*>(out);
  if (D == 512) launch_pdl(layernorm_kernel<T, 4>, dim3(grid), dim3(256), 0, s, x, g, b, o, rows, D, eps);
  else if (D == 768) launch_pdl(layernorm_kernel<T, 6>, dim3(grid), dim3(256), 0, s, x, g, b, o, rows, D, eps);
  else if (D == 1024) launch_pdl(layernorm_kernel<T, 8>, dim3(grid), dim3(256), 0, s, x, g, b, o, rows, D, eps);
  else if (D == 256) launch_pdl(layernorm_kernel<T, 2>, dim3(grid), dim3(256), 0, s, x, g, b, o, rows, D, eps);
  else if (D == 128) launch_pdl(layernorm_kernel<T, 1>, dim3(grid), dim3(256), 0, s, x, g, b, o, rows, D, eps);
  else launch_pdl(layernorm_kernel<T, 0>, dim3(grid), dim3(256), 0, s, x, g, b, o, rows, D, eps);
}

int launch_layernorm(const float* x, const float* g, const float* b, void* out, int dt, int rows, int D,
                     float eps, cudaStream_t s) {
  if (D % 4 != 0) {
    set_error("layernorm: D %% 4 != 0");
    return -1;
  }
  if (dt == DT_BF16) ln_dispatch<bf16>(x, g, b, out, rows, D, eps, s);
  else if (dt == DT_F16) ln_dispatch<__half>(x, g, b, out, rows, D, eps, s);
  else ln_dispatch<float>(x, g, b, out, rows, D, eps, s);
  return check_launch("layernorm");
}

int launch_rowstats(const float* x, void* x16, float* stats, int dt, int rows, int D, int slots, cudaStream_t s) {
  if (dt == DT_F32 || D % 128 != 0 || D > 1024 || slots < 1 || slots > 32) {
    set_error("rowstats: needs a 16-bit output, D a multiple of 128 (<= 1024) and 1..32 slots (D=%d slots=%d)", D, slots);
    return -1;
  }
  const int grid = (rows + 7) / 8;
  float2* st = reinterpret_cast<float2*>(stats);
#define HVIT_RS(TT, VV) launch_pdl(rowstats_kernel<TT, VV>, dim3(grid), dim3(256), 0, s, x, reinterpret_cast<TT*>(x16), st, rows, D, slots)
  const int V = D / 128;
  if (dt == DT_F16) {
    switch (V) {
      case 1: HVIT_RS(__half, 1); break; case 2: HVIT_RS(__half, 2); break; case 3: HVIT_RS(__half, 3); break;
      case 4: HVIT_RS(__half, 4); break; case 5: HVIT_RS(__half, 5); break; case 6: HVIT_RS(__half, 6); break;
      case 7: HVIT_RS(__half, 7); break; default: HVIT_RS(__half, 8); break;
    }
  } else {
    switch (V) {
      case 1: HVIT_RS(bf16, 1); break; case 2: HVIT_RS(bf16, 2); break; case 3: HVIT_RS(bf16, 3); break;
      case 4: HVIT_RS(bf16, 4); break; case 5: HVIT_RS(bf16, 5); break; case 6: HVIT_RS(bf16, 6); break;
      case 7: HVIT_RS(bf16, 7); break; default: HVIT_RS(bf16, 8); break;
    }
  }
#undef HVIT_RS
  return check_launch("rowstats");
}

int launch_ln_fold(const void* w, const float* gamma, const float* beta, const float* bias, void* wp, float* c, int dt,
                   int N, int K, cudaStream_t s) {
  const int grid = (N + 7) / 8;
  if (dt == DT_F16)
    ln_fold_kernel<__half><<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(w), gamma, beta, bias,
                                                reinterpret_cast<__half*>(wp), c, N, K);
  else if (dt == DT_BF16)
    ln_fold_kernel<bf16><<<grid, 256, 0, s>>>(reinterpret_cast<const bf16*>(w), gamma, beta, bias,
                                              reinterpret_cast<bf16*>(wp), c, N, K);
  else {
    set_error("ln_fold: 16-bit weights only");
    return -1;
  }
  return check_launch("ln_fold");
}

int launch_skip_sample(const void* src, int dt, int B, int Hs, int HsPitch, int Ws, int C, int Hd, int Wd,
                       void* dst, cudaStream_t s, const int* geo, int gsrc, int gdst) {
  if (dt != DT_F32 && C % 8 == 0) {
    const long long total8 = static_cast<long long>(B) * Hd * Wd * (C / 8);
    const unsigned grid8 = static_cast<unsigned>((total8 + 255) / 256);
    if (dt == DT_BF16)
      launch_pdl(skip_sample8_kernel<bf16>, dim3(grid8), dim3(256), 0, s, reinterpret_cast<const bf16*>(src), Hs, HsPitch, Ws, C, Hd,
                 Wd, reinterpret_cast<bf16*>(dst), total8, geo, gsrc, gdst);
    else
      launch_pdl(skip_sample8_kernel<__half>, dim3(grid8), dim3(256), 0, s, reinterpret_cast<const __half*>(src), Hs, HsPitch, Ws,
                 C, Hd, Wd, reinterpret_cast<__half*>(dst), total8, geo, gsrc, gdst);
    return check_launch("skip_sample");
  }
  const long long total = static_cast<long long>(B) * Hd * Wd * (C / 4);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (dt == DT_BF16)
    launch_pdl(skip_sample_kernel<bf16>, dim3(grid), dim3(256), 0, s, reinterpret_cast<const bf16*>(src), Hs, HsPitch, Ws, C, Hd, Wd,
                                                  reinterpret_cast<bf16*>(dst), total, geo, gsrc, gdst);
  else if (dt == DT_F16)
    launch_pdl(skip_sample_kernel<__half>, dim3(grid), dim3(256), 0, s, reinterpret_cast<const __half*>(src), Hs, HsPitch, Ws, C, Hd, Wd,
                                                    reinterpret_cast<__half*>(dst), total, geo, gsrc, gdst);
  else
    launch_pdl(skip_sample_kernel<float>, dim3(grid), dim3(256), 0, s, reinterpret_cast<const float*>(src), Hs, HsPitch, Ws, C, Hd, Wd,
                                                   reinterpret_cast<float*>(dst), total, geo, gsrc, gdst);
  return check_launch("skip_sample");
}

int launch_head(const void* x, int dt, const float* w, int B, int H, int W, int C, float* logits,
                float* out_tanh, cudaStream_t s) {
  if (C % 8 != 0) {
    set_error("head: C must be a multiple of 8");
    return -1;
  }
  const long long npix = static_cast<long long>(B) * H * W;
  const size_t rows_smem = static_cast<size_t>(4) * W * C * 2;
  if (C == 64 && dt != DT_F32 && rows_smem <= 72 * 1024 && B <= 65535 && getenv("HVIT_HEAD_SIMPLE") == nullptr) {
    static PerDeviceOnce once;
    if (once.first()) {
      cudaFuncSetAttribute(head_rows_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
      cudaFuncSetAttribute(head_rows_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    }
    // strip height: the row ring needs 63 KB per block (3 blocks per SM), so the grid is only a few blocks per slot;
    // pick the strip (8, 4 or 2 rows) whose last wave wastes the least time (shorter strips re-read more halo rows)
    const int sms = num_sms();
    const int per_sm = rows_smem > 0 ? static_cast<int>((200 * 1024) / rows_smem) : 1;
    const int slots = sms * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
    int strip = HEAD_STRIP;
    double best = 1e30;
    for (int st = HEAD_STRIP; st >= 2; st /= 2) {
      const long long blocks = static_cast<long long>((H + st - 1) / st) * B;
      const double waves = static_cast<double>((blocks + slots - 1) / slots);
      const double cost = waves * (st + 2);  // rows streamed per block
      if (cost < best) {
        best = cost;
        strip = st;
      }
    }
    if (const char* e = getenv("HVIT_HEAD_STRIP")) strip = atoi(e) > 0 ? atoi(e) : strip;
    if (dt == DT_BF16)
      launch_pdl(head_rows_kernel<bf16>, dim3((H + strip - 1) / strip, B), dim3(256), rows_smem, s, reinterpret_cast<const bf16*>(x), w, H, W,
                 logits, out_tanh, strip);
    else
      launch_pdl(head_rows_kernel<__half>, dim3((H + strip - 1) / strip, B), dim3(256), rows_smem, s, reinterpret_cast<const __half*>(x), w, H,
                 W, logits, out_tanh, strip);
    return check_launch("head_rows");
  }
  long long blocks = (npix + 31) / 32;
  if (blocks > 148 * 64) blocks = 148 * 64;
  const unsigned grid = static_cast<unsigned>(blocks);
  if (dt == DT_BF16)
    launch_pdl(head_kernel<bf16>, dim3(grid), dim3(256), 0, s, reinterpret_cast<const bf16*>(x), w, B, H, W, C, logits, out_tanh);
  else if (dt == DT_F16)
    launch_pdl(head_kernel<__half>, dim3(grid), dim3(256), 0, s, reinterpret_cast<const __half*>(x), w, B, H, W, C, logits, out_tanh);
  else
    launch_pdl(head_kernel<float>, dim3(grid), dim3(256), 0, s, reinterpret_cast<const float*>(x), w, B, H, W, C, logits, out_tanh);
  return check_launch("head");
}

int launch_resize(const float* src, int B, int Hs, int Ws, float* dst, int Hd, int Wd, cudaStream_t s) {
  const long long total = static_cast<long long>(B) * Hd * Wd;
  launch_pdl(resize_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, src, Hs, Ws, dst, Hd, Wd, total);
  return check_launch("resize");
}

int launch_maxpool2(const float* src, float* dst, int B, int H, int W, int C, cudaStream_t s) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo * (C / 4);
  maxpool2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(src, dst, H, W, C, Ho, Wo, total);
  return check_launch("maxpool2");
}

int launch_varlen_geometry(const int* n_valid, int B, const VarlenCfg& c, int* geo, cudaStream_t s) {
  launch_pdl(varlen_geometry_kernel, dim3((B + 127) / 128), dim3(128), 0, s, n_valid, B, c, geo);
  return check_launch("varlen_geometry");
}

int launch_zero_cols(void* buf, int B, int H, int Hpitch, int Wmax, int pix_bytes, const int* geo, int geo_idx,
                     cudaStream_t s) {
  if (pix_bytes % 16 != 0) {
    set_error("zero_cols: pixel size must be a multiple of 16 bytes");
    return -1;
  }
  launch_pdl(zero_cols_kernel, dim3(H, B), dim3(128), 0, s, reinterpret_cast<uint8_t*>(buf), Hpitch, Wmax, pix_bytes, geo, geo_idx);
  return check_launch("zero_cols");
}

int launch_tokens_compact(const float* grid, const float* pos, float* x, int B, int Hp, int Wp, int D, const int* geo,
                          cudaStream_t s) {
  launch_pdl(tokens_compact_kernel, dim3(Hp * Wp, B), dim3(128), 0, s, grid, pos, x, Hp, Wp, D, geo);
  return check_launch("tokens_compact");
}

int launch_tofm_expand(const void* rows, int dt, void* cat, int B, int Hp, int Wp, int Cx, int Ccat, const int* geo,
                       cudaStream_t s) {
  const int es = dt == DT_F32 ? 4 : 2;
  if ((Cx * es) % 16 != 0 || (Ccat * es) % 16 != 0) {
    set_error("tofm_expand: channel counts must be multiples of 16 bytes");
    return -1;
  }
  launch_pdl(tofm_expand_kernel, dim3(Hp * Wp, B), dim3(64), 0, s, reinterpret_cast<const uint8_t*>(rows),
             reinterpret_cast<uint8_t*>(cat), Hp, Wp, Cx * es, Ccat * es, geo);
  return check_launch("tofm_expand");
}

}  // namespace hvit
