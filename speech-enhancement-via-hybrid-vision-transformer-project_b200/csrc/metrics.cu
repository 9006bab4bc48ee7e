// On-device objective metrics of the evaluation caller (SURVEY.md section 8 f rank 3): SI-SDR, SNR, segmental SNR and
// log-spectral distance between a clean and an enhanced (or noisy) batch that is already on the device, so a scored
// clip makes no extra device -> host -> numpy trip.  Arithmetic follows the reference's evaluation/metrics.py:
//   compute_sisdr  :100-145  zero-mean both, alpha = <e,c> / (<c,c> + eps), 10 log10(|alpha c|^2 / (|e - alpha c|^2 + eps))
//   compute_snr    :148-184  10 log10(mean(c^2) / (mean((e - c)^2) + eps))
//   compute_segsnr :187-243  frames of 512 / hop 256 starting at range(0, n - 512, 256); per-frame 10 log10(ps / pn)
//                            clipped to [-10, 35], only frames with ps > eps and pn > eps; mean (0.0 when none)
//   compute_lsd    :246-296  |STFT| (512 / 128, centred, Hann) of both; mean over frames of sqrt(mean over bins of
//                            (ln(a + 1e-10) - ln(b + 1e-10))^2)
// Accumulation is fp64 (the reference's numpy float32 reductions are pairwise; fp64 sums are at least as accurate).
// PESQ / STOI stay host-side third-party packages, as in the reference.
#include "common.cuh"
#include "hvit.h"
#include "kernels.h"

namespace hvit {
namespace {

constexpr int MS = 16;  // doubles of accumulator state per clip
enum { A_SC = 0, A_SE = 1, A_CC = 2, A_EC = 3, A_EE = 4, A_C2 = 5, A_D2 = 6, A_SEG = 7, A_SEGN = 8, A_LSD = 9 };

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

__device__ __forceinline__ int clip_len(const int* n_valid, int b, int n_pitch) {
  if (n_valid == nullptr) return n_pitch;
  const int n = n_valid[b];
  return n < 0 ? 0 : (n > n_pitch ? n_pitch : n);
}

// pass 1: sum(c), sum(e), sum(c^2), sum((e - c)^2)
__global__ void metrics_pass1_kernel(const float* __restrict__ clean, const float* __restrict__ enh, int n_pitch,
                                     const int* __restrict__ n_valid, double* __restrict__ acc) {
  const int b = blockIdx.y;
  const int n = clip_len(n_valid, b, n_pitch);
  const float* c = clean + static_cast<long long>(b) * n_pitch;
  const float* e = enh + static_cast<long long>(b) * n_pitch;
  double sc = 0, se = 0, c2 = 0, d2 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double cv = c[i], ev = e[i], d = ev - cv;
    sc += cv; se += ev; c2 += cv * cv; d2 += d * d;
  }
  sc = warp_sum_d(sc); se = warp_sum_d(se); c2 = warp_sum_d(c2); d2 = warp_sum_d(d2);
  if ((threadIdx.x & 31) == 0) {
    double* a = acc + b * MS;
    atomicAdd(a + A_SC, sc); atomicAdd(a + A_SE, se); atomicAdd(a + A_C2, c2); atomicAdd(a + A_D2, d2);
  }
}

// pass 2: centred second moments <c',c'>, <e',c'>, <e',e'>
__global__ void metrics_pass2_kernel(const float* __restrict__ clean, const float* __restrict__ enh, int n_pitch,
                                     const int* __restrict__ n_valid, double* __restrict__ acc) {
  const int b = blockIdx.y;
  const int n = clip_len(n_valid, b, n_pitch);
  if (n == 0) return;
  const float* c = clean + static_cast<long long>(b) * n_pitch;
  const float* e = enh + static_cast<long long>(b) * n_pitch;
  // (the reference subtracts the float32 mean from float32 samples; fp64 here)
  const double mc = acc[b * MS + A_SC] / n, me = acc[b * MS + A_SE] / n;
  double cc = 0, ec = 0, ee = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double cv = c[i] - mc, ev = e[i] - me;
    cc += cv * cv; ec += ev * cv; ee += ev * ev;
  }
  cc = warp_sum_d(cc); ec = warp_sum_d(ec); ee = warp_sum_d(ee);
  if ((threadIdx.x & 31) == 0) {
    double* a = acc + b * MS;
    atomicAdd(a + A_CC, cc); atomicAdd(a + A_EC, ec); atomicAdd(a + A_EE, ee);
  }
}

// segmental SNR: one warp per frame of 512 samples, hop 256, frame starts range(0, n - 512, 256)
__global__ void metrics_segsnr_kernel(const float* __restrict__ clean, const float* __restrict__ enh, int n_pitch,
                                      const int* __restrict__ n_valid, double* __restrict__ acc) {
  const int b = blockIdx.y;
  const int n = clip_len(n_valid, b, n_pitch);
  const int frames = n > 512 ? (n - 512 + 255) / 256 : 0;
  const int fr = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (fr >= frames) return;
  const int lane = threadIdx.x & 31;
  const float* c = clean + static_cast<long long>(b) * n_pitch + fr * 256;
  const float* e = enh + static_cast<long long>(b) * n_pitch + fr * 256;
  double ps = 0, pn = 0;
  for (int i = lane; i < 512; i += 32) {
    const double cv = c[i], d = static_cast<double>(e[i]) - cv;
    ps += cv * cv; pn += d * d;
  }
  ps = warp_sum_d(ps) / 512.0;
  pn = warp_sum_d(pn) / 512.0;
  if (lane == 0 && ps > 1e-8 && pn > 1e-8) {
    double v = 10.0 * log10(ps / pn);
    v = v < -10.0 ? -10.0 : (v > 35.0 ? 35.0 : v);
    atomicAdd(acc + b * MS + A_SEG, v);
    atomicAdd(acc + b * MS + A_SEGN, 1.0);
  }
}

// log-spectral distance of one frame: block of 128 threads over 257 bins of |STFT| [B,257,T]
__global__ void metrics_lsd_kernel(const float* __restrict__ mag_a, const float* __restrict__ mag_b, int T_pitch,
                                   int n_pitch, const int* __restrict__ n_valid, double* __restrict__ acc) {
  const int b = blockIdx.y, t = blockIdx.x;
  const int n = clip_len(n_valid, b, n_pitch);
  if (n == 0 || t >= 1 + n / 128) return;
  __shared__ double part[4];
  double s = 0;
  for (int f = threadIdx.x; f < 257; f += blockDim.x) {
    const long long o = (static_cast<long long>(b) * 257 + f) * T_pitch + t;
    const float d = logf(mag_a[o] + 1e-10f) - logf(mag_b[o] + 1e-10f);   // float32 logs, like the reference
    s += static_cast<double>(d) * d;
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += part[w];
    atomicAdd(acc + b * MS + A_LSD, sqrt(tot / 257.0));
  }
}

__global__ void metrics_finalize_kernel(const double* __restrict__ acc, int B, int n_pitch, const int* __restrict__ n_valid,
                                        double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = clip_len(n_valid, b, n_pitch);
  const double* a = acc + b * MS;
  double* o = out + b * 4;
  if (n == 0) {
    o[0] = o[1] = o[2] = o[3] = 0.0;
    return;
  }
  const double eps = 1e-8;
  const double alpha = a[A_EC] / (a[A_CC] + eps);
  const double num = alpha * alpha * a[A_CC];
  const double den = a[A_EE] - 2.0 * alpha * a[A_EC] + alpha * alpha * a[A_CC];   // |e' - alpha c'|^2
  o[0] = 10.0 * log10(num / ((den > 0 ? den : 0.0) + eps));
  o[1] = 10.0 * log10((a[A_C2] / n) / (a[A_D2] / n + eps));
  o[2] = a[A_SEGN] > 0 ? a[A_SEG] / a[A_SEGN] : 0.0;
  o[3] = a[A_LSD] / (1 + n / 128);
}

}  // namespace
}  // namespace hvit

using namespace hvit;

namespace hvit {
namespace {
// Spectrogram losses of the validation forward (training/losses.py:15-85 SpectrogramLoss, :88-141 STOILoss as the
// reference defines it - one minus the cosine similarity of the flattened spectrograms -, :286-387 CombinedLoss):
// per sample b, fp64 sums {sum |p' - t'|, sum (p' - t')^2, sum p^2, sum t^2, sum p t}, p' / t' = ln(. + 1e-8) with
// log compression, the raw values otherwise.
__global__ void spec_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long n_per,
                                 int use_log, double* __restrict__ sums) {
  const int b = blockIdx.y;
  const float* p = pred + b * n_per;
  const float* t = target + b * n_per;
  double s1 = 0, s2 = 0, pp = 0, tt = 0, pt = 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_per;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float pv = p[i], tv = t[i];
    const float a = use_log ? logf(pv + 1e-8f) : pv, c = use_log ? logf(tv + 1e-8f) : tv;
    const double d = static_cast<double>(a) - static_cast<double>(c);
    s1 += fabs(d); s2 += d * d;
    pp += static_cast<double>(pv) * pv; tt += static_cast<double>(tv) * tv; pt += static_cast<double>(pv) * tv;
  }
  s1 = warp_sum_d(s1); s2 = warp_sum_d(s2); pp = warp_sum_d(pp); tt = warp_sum_d(tt); pt = warp_sum_d(pt);
  if ((threadIdx.x & 31) == 0) {
    double* a = sums + b * 5;
    atomicAdd(a + 0, s1); atomicAdd(a + 1, s2); atomicAdd(a + 2, pp); atomicAdd(a + 3, tt); atomicAdd(a + 4, pt);
  }
}
}  // namespace
}  // namespace hvit

extern "C" {

size_t hvit_metrics_scratch_bytes(int B, int n_samples) {
  if (B < 1 || n_samples < 1) return 0;
  const size_t T = 1 + static_cast<size_t>(n_samples) / 128;
  // two magnitude spectrograms, two [B] scalar arrays for the STFT kernel, accumulators
  return 2 * static_cast<size_t>(B) * 257 * T * 4 + 4 * static_cast<size_t>(B) * 4 + static_cast<size_t>(B) * MS * 8 + 1024;
}

int hvit_metrics(const float* clean_dev, const float* enhanced_dev, int B, int n_samples, const int* n_valid_dev,
                 void* scratch_dev, size_t scratch_bytes, double* out_dev, void* stream) {
  if (clean_dev == nullptr || enhanced_dev == nullptr || scratch_dev == nullptr || out_dev == nullptr || B < 1 ||
      n_samples < 1) {
    set_error("hvit_metrics: bad argument");
    return HVIT_E_ARG;
  }
  if (scratch_bytes < hvit_metrics_scratch_bytes(B, n_samples) || (reinterpret_cast<uintptr_t>(scratch_dev) & 255) != 0) {
    set_error("hvit_metrics: scratch too small or not 256-byte aligned (need %zu bytes)", hvit_metrics_scratch_bytes(B, n_samples));
    return HVIT_E_ALLOC;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int r = ensure_fft_tables(s);
  if (r) return r;
  const int T = 1 + n_samples / 128;
  const size_t spec = static_cast<size_t>(B) * 257 * T;
  uint8_t* p = reinterpret_cast<uint8_t*>(scratch_dev);
  double* acc = reinterpret_cast<double*>(p);                  // (first: 8-byte aligned)
  p += (static_cast<size_t>(B) * MS * 8 + 255) / 256 * 256;
  float* mag_a = reinterpret_cast<float*>(p);
  float* mag_b = mag_a + spec;
  float* ones = mag_b + spec;                                    // [B] peak scalars (1.0: no normalisation)
  unsigned* mmax = reinterpret_cast<unsigned*>(ones + B);        // [B] magnitude maxima (unused output of the STFT)
  if (cudaMemsetAsync(acc, 0, static_cast<size_t>(B) * MS * 8, s) != cudaSuccess) return check_launch("hvit_metrics(memset)");
  dim3 grid(32, B);
  metrics_pass1_kernel<<<grid, 256, 0, s>>>(clean_dev, enhanced_dev, n_samples, n_valid_dev, acc);
  metrics_pass2_kernel<<<grid, 256, 0, s>>>(clean_dev, enhanced_dev, n_samples, n_valid_dev, acc);
  const int max_frames = n_samples > 512 ? (n_samples - 512 + 255) / 256 : 0;
  if (max_frames > 0)
    metrics_segsnr_kernel<<<dim3((max_frames + 7) / 8, B), 256, 0, s>>>(clean_dev, enhanced_dev, n_samples, n_valid_dev, acc);
  r = check_launch("hvit_metrics(sums)");
  if (r) return r;
  // |STFT| of both signals with the enhance path's kernel (no peak normalisation)
  r = launch_peak(clean_dev, B, n_samples, ones, 0, s);
  if (r) return r;
  r = launch_stft(clean_dev, B, n_samples, T, ones, nullptr, mag_a, mmax, s);
  if (r) return r;
  r = launch_stft(enhanced_dev, B, n_samples, T, ones, nullptr, mag_b, mmax, s);
  if (r) return r;
  metrics_lsd_kernel<<<dim3(T, B), 128, 0, s>>>(mag_a, mag_b, T, n_samples, n_valid_dev, acc);
  metrics_finalize_kernel<<<(B + 127) / 128, 128, 0, s>>>(acc, B, n_samples, n_valid_dev, out_dev);
  return check_launch("hvit_metrics");
}


int hvit_spec_loss(const float* pred_dev, const float* target_dev, int B, long long n_per, int use_log,
                   double* sums_dev, void* stream) {
  using namespace hvit;
  if (pred_dev == nullptr || target_dev == nullptr || sums_dev == nullptr || B <= 0 || n_per <= 0) {
    set_error("hvit_spec_loss: bad arguments (B=%d n_per=%lld)", B, n_per);
    return HVIT_E_ARG;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(sums_dev, 0, static_cast<size_t>(B) * 5 * sizeof(double), s) != cudaSuccess)
    return check_launch("hvit_spec_loss(memset)");
  const long long per_block = 256 * 8;
  long long bx = (n_per + per_block - 1) / per_block;
  if (bx > 64) bx = 64;
  spec_loss_kernel<<<dim3(static_cast<unsigned>(bx), B), 256, 0, s>>>(pred_dev, target_dev, n_per, use_log, sums_dev);
  return check_launch("hvit_spec_loss");
}

}  // extern "C"
