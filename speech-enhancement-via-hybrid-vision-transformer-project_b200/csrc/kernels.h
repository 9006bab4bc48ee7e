// Internal launcher interface shared by the .cu translation units of libhvit_sm100.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hvit {

// ---------------------------------------------------------------------------------------------
// Implicit-GEMM problem description (used by both the tcgen05 bf16 kernel and the fp32 SIMT kernel)
//   C[m, n] = epilogue( sum_k A(m, k) * Wt[n, k] )
// A-operand addressing modes (activations are NHWC):
//   IG_PLAIN : A is a row-major [M, K] matrix
//   IG_CONV3 : 3x3 / pad 1 convolution, k = (ky*3+kx)*Cin + c, m = output pixel
//   IG_UP2   : nearest x2 upsample followed by 3x3 / pad 1 conv, evaluated as four 2x2 convolutions on the
//              low-resolution input (one per output parity class) with pre-summed weights;
//              k = (a*2+b)*Cin + c, weight rows [parity*N + n]
//   IG_PATCH : p x p / stride p convolution (patch embedding), k = (ky*p+kx)*Cin + c, m = token
// ---------------------------------------------------------------------------------------------
enum { IG_PLAIN = 0, IG_CONV3 = 1, IG_UP2 = 2, IG_PATCH = 3 };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };
// activation storage type of a plan: fp32 (CUDA-core mode) or one of the two 16-bit tensor-core operand types
enum { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };

// Division by a launch-time constant without the ~100-cycle integer-division sequence (the tile decode sits on the
// epilogue warps' critical path once per tile): q = umulhi(n, mul) >> shr for 0 <= n < 2^31, with
// mul = ceil(2^(31 + ceil_log2 d) / d), shr = ceil_log2 d - 1; d == 1 is the identity.
struct FastDiv {
  unsigned mul, shr;
  int d;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = d;
  f.mul = 0;
  f.shr = 0;
  if (d > 1) {
    int l = 0;
    while ((1ll << l) < d) ++l;
    const unsigned long long p2 = 1ull << (31 + l);
    f.mul = static_cast<unsigned>((p2 + static_cast<unsigned long long>(d) - 1) / static_cast<unsigned long long>(d));
    f.shr = static_cast<unsigned>(l - 1);
  }
  return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ void fast_divmod(const FastDiv& f, int n, int& q, int& r) {
  q = f.d == 1 ? n : static_cast<int>(__umulhi(static_cast<unsigned>(n), f.mul) >> f.shr);
  r = n - q * f.d;
}
#endif

struct IgemmParams {
  int mode;
  int M, N, K;          // IG_PLAIN: M rows. All modes: N output channels, K reduction length
  // conv geometry (input image, NHWC, channel stride == lda)
  int B, H, W, Cin;
  int Hq;               // IG_PATCH: rows of patch positions per image in the (padded) input buffer = Hpad / p
  int patch;            // IG_PATCH: patch size p
  int Hp, Wp;           // IG_PATCH: valid patch grid
  int Wt, Ht;           // spatial tile (Wt * Ht == 128)
  int tiles_w, tiles_h;
  // epilogue
  const float* scale;   // [N] or null (== 1)
  const float* shift;   // [N] or null (== 0): folded BN shift or linear bias
  int act;
  const float* residual;  // fp32 [*, ldr] added after activation, or null
  int ldr;
  int res_mod;          // > 0: residual row = out_row % res_mod (positional-embedding table)
  int res_inplace;      // set by the launcher: residual == out (same pitch) -> the epilogue adds into global memory
                        // with a TMA reduce-add instead of loading the residual tile
  int pool;             // 1: fused 2x2 max-pool (IG_CONV3 only)
  void* out;
  int ldc;              // elements between consecutive output rows
  int out_f32;          // 1: fp32 output, 0: 16-bit output
  int f16;              // tcgen05 path: 1 = fp16 operands / outputs, 0 = bf16
  int dbg;              // diagnostics only (HVIT_DBG): 1 skip global stores, 2 skip TMEM loads, 4 skip MMA issue
  int Ho, Wo;           // valid output image dims (conv modes)
  int HoPitch;          // image row pitch of the output buffer in pixel rows (>= Ho)
  int a_prefetch;       // set by the launcher: L2-prefetch distance (k-blocks) for an A operand that streams from HBM, 0 = off
  int halo;             // IG_CONV3 on the tensor-core path: input tile + halo staged once in shared memory (8x16 tile;
                        // maps.a is then the 5-D un-swizzled halo map), see igemm_halo_kernel
  long long* prof;      // diagnostics only (HVIT_PROF): per-CTA cycle counters [gridDim.x][16], or null
  // ---- LayerNorm folded into the GEMMs on either side of it (IG_PLAIN, tcgen05 path)
  // producer (fp32 output WITH a loaded residual: x_new = x + f(..)): besides x_new the epilogue writes the 16-bit copy
  // x16 and, per row and per 128-column slot (slot = 2 * n_tile + epilogue group), the slot's mean and centred sum of
  // squares M2 - the partial statistics nn.LayerNorm needs, merged exactly (Chan) by the consumer
  void* ln_x16_out;     // [M, ld16] 16-bit or null
  int ld16;
  float* ln_stats_out;  // [M, N / 128, 2] or null
  // consumer (16-bit output): A = x16, weights W''[n,k] = gamma[k] W[n,k] - mean_k(gamma W[n,:]) (scaled and centred over
  // k, so the contraction of x equals the contraction of x - mean(x)); the epilogue evaluates
  //   LN(x) W^T + b = rs_m * acc_mn + c_n,   c_n = b_n + sum_k beta_k W_nk   (c is passed as `shift`)
  // with rs_m = 1 / sqrt(var_m + eps) from ln_stats_in
  const float* ln_stats_in;  // [M, ln_slots, 2] or null
  int ln_slots;
  float ln_eps;
  FastDiv fd_ntn, fd_ppg, fd_tw, fd_th;  // set by launch_igemm_tc2: n tiles, pairs per group, tiles_w, tiles_h
};

// tcgen05 path. A / Wt are described by TMA tensor maps built on the host (api.cu).
// CTA-pair (cta_group::2, UMMA M = 256) kernel; tmap_b is built with a box of block_n / 2 rows.  The epilogue
// writes through shared memory with TMA stores: `out` holds one 4-D map per upsample parity (index 0 otherwise) with a
// 128-byte-wide box (64 16-bit or 32 fp32 columns); `res` is the fp32 residual map (same box) when p.residual is set.
struct IgemmMaps {
  CUtensorMap a, b, out[4], res;
};
int launch_igemm_tc2(const IgemmParams& p, const IgemmMaps& maps, int block_n, int num_sms, cudaStream_t stream);

// fp32 SIMT path (precision = fp32 mode, also the on-device cross-check of the tcgen05 kernels).
// A and Wt are fp32; for IG_UP2 the weights are the ORIGINAL 3x3 weights [N, 9*Cin] (direct y//2 gather).
int launch_igemm_f32(const IgemmParams& p, const float* A, int lda, const float* Wt, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// Attention. qkv: [B*N, 3*D] (q | k | v, head h at columns h*64 .. h*64+63 of each third), out: [B*N, D].
// ---------------------------------------------------------------------------------------------
// geo (nullable): per-clip valid token count geo[b][GEO_NTOK] - keys >= N_b are masked, whole key blocks beyond it skipped
int launch_attn_tc(const CUtensorMap& tmap_qkv, const CUtensorMap& tmap_out /*[B*N, D] 16-bit*/, int f16, int B, int N, int heads, int D, float scale,
                   cudaStream_t stream, long long* prof = nullptr, const int* geo = nullptr,
                   float* probs /*nullable [B,h,N,N]: attention maps from the same kernel, N <= 1280*/ = nullptr);
int launch_attn_f32(const float* qkv, float* out, float* probs /*nullable [B,h,N,N]*/, int B, int N, int heads, int D,
                    float scale, cudaStream_t stream, const int* geo = nullptr);
// probabilities only (return_attentions=True slow path) from bf16 qkv
int launch_attn_probs_16(const void* qkv16, int f16, float* probs, int B, int N, int heads, int D, float scale,
                         cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// Bandwidth-bound glue kernels.  The activation type is selected by `dt` (DT_F32 / DT_BF16 / DT_F16).
// ---------------------------------------------------------------------------------------------
int ensure_fft_tables(cudaStream_t s);
int launch_peak(const float* wave, int B, int n, float* max_val /*[B]*/, int normalize, cudaStream_t s);
int launch_stft(const float* wave, int B, int n, int T, const float* max_val, float2* spec /*[B,257,T] or null*/,
                float* mag /*[B,257,T]*/, unsigned* mag_max_bits /*[B]*/, cudaStream_t s, const int* geo = nullptr);
// model_out [B,257,T] is read when lowres == nullptr, otherwise it is WRITTEN with the bilinear resize of
// lowres [B,Hs,Ws] (fused final interpolate of HybridViT.forward).
int launch_istft_frames(float* model_out, const float* lowres, int Hs, int Ws, const float2* spec,
                        const unsigned* mag_max_bits, float* frames /*[B,T,512]*/, int B, int T, cudaStream_t s);
int launch_istft_ola(const float* frames, const float* max_val, float* wave_out, int B, int n, int T, cudaStream_t s);
// Fused back end of the enhance path: phase recomputed from the noisy waveform, final bilinear resize of the decoder's
// [B,Hs,Ws] tanh map, inverse FFT, overlap-add in shared memory, envelope, de-normalisation.  model_out (nullable)
// receives the resized model output [B,257,T] for tests.
int launch_enhance_istft(const float* wave_in, const float* max_val, const unsigned* mag_max_bits, const float* lowres,
                         int Hs, int Ws, float* model_out, float* wave_out, int B, int n, int T, cudaStream_t s,
                         const int* geo = nullptr, int geo_ws_idx = 0);
int launch_stem(const float* x /*[B,H,W]*/, const unsigned* mag_max_bits /*nullable*/, const float* w /*[9][C]*/,
                const float* scale, const float* shift, void* out, int dt, int B, int H, int W, int C, int pool,
                cudaStream_t s);
// tensor-core stem (C = 64, pool 2, 16-bit output): apack = [4][64][64] 16-bit position matrices built once per
// weight set by launch_stem_pack (BN scale folded in), see stem_tc.cu
int launch_stem_pack(const float* w9c, const float* scale, void* apack, int f16, cudaStream_t s);
// tmap_out: 4-D map (64 channels, Wo, Ho, B) of the NHWC output with a 128B-swizzled (64, 64, 1, 1) box
int launch_stem_tc(const float* x, const unsigned* mag_max_bits, const void* apack, const float* shift,
                   const CUtensorMap& tmap_out, int f16, int B, int H, int W, int num_sms, cudaStream_t s);
int launch_layernorm(const float* x, const float* g, const float* b, void* out, int dt, int rows, int D,
                     float eps, cudaStream_t s);
// LayerNorm folded into the neighbouring GEMMs (IgemmParams::ln_*): 16-bit copy + per-slot row statistics of an fp32
// matrix, and the gamma / beta folding of the consumer's weights
int launch_rowstats(const float* x, void* x16, float* stats /*[rows, slots, 2]*/, int dt, int rows, int D, int slots,
                    cudaStream_t s);
int launch_ln_fold(const void* w /*[N,K] 16-bit*/, const float* gamma, const float* beta, const float* bias /*nullable*/,
                   void* wp /*[N,K] 16-bit*/, float* c /*[N]*/, int dt, int N, int K, cudaStream_t s);
int launch_skip_sample(const void* src, int dt, int B, int Hs, int HsPitch, int Ws, int C, int Hd, int Wd,
                       void* dst, cudaStream_t s, const int* geo = nullptr, int geo_src_idx = 0, int geo_dst_idx = 0);
int launch_head(const void* x, int dt, const float* w /*[9][C]*/, int B, int H, int W, int C, float* logits,
                float* out_tanh, cudaStream_t s);
int launch_resize(const float* src, int B, int Hs, int Ws, float* dst, int Hd, int Wd, cudaStream_t s);
int launch_maxpool2(const float* src, float* dst, int B, int H, int W, int C, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// Variable-length batches (hvit_enhance_varlen): clips are zero-padded to the plan's n_samples; every clip is processed
// exactly as if it were alone (SURVEY.md section 8 f rank 2; reference plumbing: models/attention.py:94-98 key mask,
// data/dataset.py:297-347 zero-pad collate).  A one-block kernel derives the per-clip geometry table `geo`
// ([B][GEO_STRIDE] ints, device memory) from the valid sample counts; kernels that take a `geo` pointer use the
// per-clip sizes when it is non-null and the plan's sizes otherwise.
// ---------------------------------------------------------------------------------------------
constexpr int GEO_STRIDE = 32;
enum {
  GEO_N = 0,      // valid samples
  GEO_T = 1,      // STFT frames 1 + n / 128
  GEO_NTOK = 2,   // tokens Hp * Wp
  GEO_WP = 3,     // patch-grid width
  GEO_ENC = 4,    // + i: width of encoder block i's output
  GEO_CAT = 12    // + i: width of decoder block i's input (concat buffer i)
};
struct VarlenCfg {
  int n_enc, n_dec, patch, Hp;
  int enc_pool[8];
  int dec_up[8];
  int n_min, n_max;  // clamp range of the valid sample counts (n_min: smallest clip that still yields one patch column)
};
int launch_varlen_geometry(const int* n_valid, int B, const VarlenCfg& c, int* geo, cudaStream_t s);
// zero the pixel columns [W_b, Wmax) of an NHWC buffer [B, Hpitch, Wmax, pix_bytes] (rows h < H), W_b = geo[b][geo_idx]
int launch_zero_cols(void* buf, int B, int H, int Hpitch, int Wmax, int pix_bytes, const int* geo, int geo_idx,
                     cudaStream_t s);
// residual stream of a clip from the patch grid: x[b][n] = grid[b][n / Wp_b][n % Wp_b] + pos[n] for n < N_b (the
// reference's token order and positional rows for THAT clip's width), 0 for N_b <= n < Np
int launch_tokens_compact(const float* grid, const float* pos, float* x, int B, int Hp, int Wp, int D, const int* geo,
                          cudaStream_t s);
// to_feature_map output rows (compact token order) -> channels [0, Cx) of the NHWC concat buffer, zero beyond Wp_b
int launch_tofm_expand(const void* rows, int dt, void* cat, int B, int Hp, int Wp, int Cx, int Ccat, const int* geo,
                       cudaStream_t s);

// Kernel launch with programmatic stream serialization (PDL); HVIT_NO_PDL=1 falls back to plain stream order.
// While an L2 window is set (l2_window_set: the transformer's fp32 residual stream, re-read and updated in place by
// every LayerNorm / proj / fc2 of a step) each launch carries it as an access-policy-window attribute: lines of the
// window are kept as persisting L2 lines, everything else stays normal.
bool pdl_enabled();
struct L2Window {
  void* base;
  size_t bytes;
  float hit_ratio;
};
const L2Window& l2_window();
void l2_window_set(void* base, size_t bytes, float hit_ratio);  // bytes == 0 clears it
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (pdl_enabled()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  const L2Window& w = l2_window();
  if (w.bytes > 0) {
    at[na].id = cudaLaunchAttributeAccessPolicyWindow;
    at[na].val.accessPolicyWindow.base_ptr = w.base;
    at[na].val.accessPolicyWindow.num_bytes = w.bytes;
    at[na].val.accessPolicyWindow.hitRatio = w.hit_ratio;
    at[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    at[na].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
  if (e != cudaSuccess && w.bytes > 0) {
    // a driver that rejects the access-policy window (e.g. a smaller limit than the one queried) must not take the
    // whole step down: retry this launch once without it
    cudaGetLastError();
    cfg.numAttrs = na - 1;
    e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
  }
  return e;
}

// per-device state (api.cu): a process may drive several GPUs, so one-time initialisation (function attributes, FFT
// tables, SM count, L2 carve-out) is keyed by the current CUDA device
constexpr int kMaxDevices = 64;
int current_device();
int num_sms();
// true exactly once per (flag array, current device)
struct PerDeviceOnce {
  bool done[kMaxDevices] = {};
  bool first() {
    const int d = current_device();
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
  void retry() { done[current_device()] = false; }
};

// error plumbing (api.cu)
void set_error(const char* fmt, ...);
int check_launch(const char* what);

}  // namespace hvit
