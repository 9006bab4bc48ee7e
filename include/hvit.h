/* libhvit_sm100.so - C ABI of the B200 (sm_100a) HybridViT speech-enhancement inference path.
 *
 * This is the drop-in boundary for the reference's inference hot path.  The reference is pure Python and has no
 * FFI of its own; these entry points are what a binding for that path binds (see INTEGRATION.md for the ctypes
 * stub).  Each function cites the reference interface it replaces (paths relative to the reference repo).
 *
 * Conventions
 *   - every function returns an int status: 0 = ok, negative = HVIT_E_*; hvit_last_error() gives the message
 *     (thread-local).
 *   - all pointers named *_dev are CUDA device pointers owned by the caller; nothing is allocated, freed or
 *     retained beyond the lifetime of the plan that was given them; no hidden device allocation.
 *   - all work is enqueued asynchronously on the caller's stream (`stream` is a cudaStream_t passed as void*);
 *     no function synchronises the device, so calls can be captured into a CUDA graph.
 *   - activations are NHWC internally; the public tensors keep the reference layouts ([B,1,F,T], [B,n]).
 *   - there is no CPU fallback: on a device that is not sm_100 plan creation fails with HVIT_E_ARCH.
 */
#ifndef HVIT_H_
#define HVIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HVIT_OK 0
#define HVIT_E_SHAPE (-1)   /* unsupported or inconsistent shape / configuration */
#define HVIT_E_ARCH (-2)    /* device is not sm_100 */
#define HVIT_E_ALLOC (-3)   /* workspace too small / misaligned */
#define HVIT_E_LAUNCH (-4)  /* CUDA launch or driver error */
#define HVIT_E_ARG (-5)     /* null pointer or bad argument */

#define HVIT_PREC_FP32 0 /* fp32 CUDA-core arithmetic, fp32 activations (accuracy mode, <= 1e-4)           */
#define HVIT_PREC_BF16 1 /* bf16 tcgen05 tensor-core contractions, fp32 accumulation / residual / statistics */
#define HVIT_PREC_FP16 2 /* same kernels and tensor-core rate with fp16 operands / activations (3 more mantissa   */
                         /* bits; saturating conversions) - the 16-bit mode that meets the 1e-2 parity bar        */

#define HVIT_MAX_STAGES 8
#define HVIT_MAX_LAYERS 48

/* Mirrors HybridViT.__init__ (models/hybrid_vit.py:36-67) after create_hybrid_vit (hybrid_vit.py:492-525). */
typedef struct hvit_model_cfg {
  int n_enc;                          /* len(encoder_channels); block 0 is the 1->C stem               */
  int enc_channels[HVIT_MAX_STAGES];
  int enc_pool[HVIT_MAX_STAGES];      /* 1 (none) or 2                                                  */
  int embed_dim, num_heads, num_layers, mlp_hidden, patch_size;
  int n_dec;                          /* len(decoder_channels); last block is the C->1 tanh head        */
  int dec_channels[HVIT_MAX_STAGES];
  int dec_up[HVIT_MAX_STAGES];        /* 1 (none) or 2 (nearest)                                        */
  int use_skip;
  int precision;                      /* HVIT_PREC_*                                                    */
  float ln_eps;
} hvit_model_cfg;

/* Packed, device-resident weights (derived once from the reference state_dict; the
 * packing itself is host-side torch code in hvit_b200/models/packing.py).  "act" = bf16 / fp16 in HVIT_PREC_BF16 / HVIT_PREC_FP16, fp32 in
 * HVIT_PREC_FP32.  BatchNorm (eval) is folded into per-channel scale/shift.  Conv weights are [Cout][ky][kx][Cin].
 * In the 16-bit modes a decoder block with upsample 2 gets the four pre-summed 2x2 parity kernels [4][Cout][2][2][Cin];
 * in FP32 mode it keeps the original [Cout][3][3][Cin]. */
typedef struct hvit_weights {
  const float* stem_w;      /* [3][3][C0] fp32 */
  const float* stem_scale;  /* [C0] */
  const float* stem_shift;  /* [C0] */
  const void* enc_w[HVIT_MAX_STAGES];      /* act; index i = encoder block i (i >= 1) */
  const float* enc_scale[HVIT_MAX_STAGES];
  const float* enc_shift[HVIT_MAX_STAGES];
  const void* patch_w;      /* act [D][p][p][C] */
  const float* patch_b;     /* [D] */
  const float* pos_embed;   /* fp32 [pos_len][D] */
  int pos_len;
  const float* ln1_g[HVIT_MAX_LAYERS];
  const float* ln1_b[HVIT_MAX_LAYERS];
  const float* ln2_g[HVIT_MAX_LAYERS];
  const float* ln2_b[HVIT_MAX_LAYERS];
  const void* qkv_w[HVIT_MAX_LAYERS];   /* act [3D][D] */
  const float* qkv_b[HVIT_MAX_LAYERS];
  const void* proj_w[HVIT_MAX_LAYERS];  /* act [D][D] */
  const float* proj_b[HVIT_MAX_LAYERS];
  const void* fc1_w[HVIT_MAX_LAYERS];   /* act [hidden][D] */
  const float* fc1_b[HVIT_MAX_LAYERS];
  const void* fc2_w[HVIT_MAX_LAYERS];   /* act [D][hidden] */
  const float* fc2_b[HVIT_MAX_LAYERS];
  const float* lnf_g;
  const float* lnf_b;
  const void* tofm_w;       /* act [Cenc][D] */
  const float* tofm_b;
  const void* skip_w[HVIT_MAX_STAGES];  /* act [Cdec_i][Cenc_rev_i] (1x1 conv) */
  const float* skip_b[HVIT_MAX_STAGES];
  const void* dec_w[HVIT_MAX_STAGES];   /* act; blocks 0 .. n_dec-2 */
  const float* dec_scale[HVIT_MAX_STAGES];
  const float* dec_shift[HVIT_MAX_STAGES];
  const float* head_w;      /* [3][3][C] fp32, last decoder block */
} hvit_weights;

/* The reference state_dict (models/hybrid_vit.py:172-284) as fp32 DEVICE pointers in the
 * layouts torch stores them: conv weights [Cout][Cin][kh][kw], linear weights [out][in], BatchNorm weight / bias /
 * running_mean / running_var [C].  Input of hvit_pack_weights.  Index i of enc_* = encoder block i (0 = stem), of
 * dec_* = decoder block i (n_dec - 1 = the 1-channel head, no BatchNorm), of skip_* = skip_projections[i]. */
typedef struct hvit_ref_weights {
  const float* enc_conv_w[HVIT_MAX_STAGES];
  const float* enc_bn_w[HVIT_MAX_STAGES];
  const float* enc_bn_b[HVIT_MAX_STAGES];
  const float* enc_bn_mean[HVIT_MAX_STAGES];
  const float* enc_bn_var[HVIT_MAX_STAGES];
  const float* patch_w;   /* patch_embed.projection.weight [D][C][p][p] */
  const float* patch_b;
  const float* pos_embed; /* pos_encoding.pos_embed [1][pos_len][D] */
  int pos_len;
  const float* ln1_w[HVIT_MAX_LAYERS];
  const float* ln1_b[HVIT_MAX_LAYERS];
  const float* ln2_w[HVIT_MAX_LAYERS];
  const float* ln2_b[HVIT_MAX_LAYERS];
  const float* qkv_w[HVIT_MAX_LAYERS];
  const float* qkv_b[HVIT_MAX_LAYERS];
  const float* proj_w[HVIT_MAX_LAYERS];
  const float* proj_b[HVIT_MAX_LAYERS];
  const float* fc1_w[HVIT_MAX_LAYERS];
  const float* fc1_b[HVIT_MAX_LAYERS];
  const float* fc2_w[HVIT_MAX_LAYERS];
  const float* fc2_b[HVIT_MAX_LAYERS];
  const float* lnf_w;     /* transformer.norm */
  const float* lnf_b;
  const float* tofm_w;    /* to_feature_map */
  const float* tofm_b;
  const float* dec_conv_w[HVIT_MAX_STAGES];
  const float* dec_bn_w[HVIT_MAX_STAGES];
  const float* dec_bn_b[HVIT_MAX_STAGES];
  const float* dec_bn_mean[HVIT_MAX_STAGES];
  const float* dec_bn_var[HVIT_MAX_STAGES];
  const float* skip_w[HVIT_MAX_STAGES];  /* [Cdec_i][Cenc_rev_i][1][1] */
  const float* skip_b[HVIT_MAX_STAGES];
} hvit_ref_weights;

typedef struct hvit_plan hvit_plan;

const char* hvit_last_error(void);
int hvit_version(void);
/* 1 when the current CUDA device is compute capability 10.x, 0 otherwise, negative on CUDA error. */
int hvit_device_ok(void);

/* Weight packing on the device (replaces the parameter half of HybridViT.__init__ / load_state_dict for this path,
 * models/hybrid_vit.py:172-284, utils/checkpoint.py:127-161): BatchNorm(eval, eps 1e-5) folded to scale / shift (into
 * the conv weights in the 16-bit modes), conv kernels re-laid out K-major [Cout][ky][kx][Cin], "nearest x2 + 3x3"
 * decoder kernels pre-summed into four 2x2 parity kernels, conversion to the plan's operand type.  Everything the plan
 * reads is written into `packed_dev` (hvit_packed_weights_bytes(cfg) bytes, 256-byte aligned, caller-owned), so the
 * reference tensors may be freed once the stream has passed this call; `out` receives pointers into `packed_dev`. */
size_t hvit_packed_weights_bytes(const hvit_model_cfg* cfg, int pos_len);
int hvit_pack_weights(const hvit_model_cfg* cfg, const hvit_ref_weights* ref, void* packed_dev, size_t packed_bytes,
                      hvit_weights* out, void* stream);

/* Bytes of device workspace a plan needs.  n_samples > 0 adds the STFT/iSTFT buffers (enhance path) and
 * requires F == 257, T == 1 + n_samples / 128.  Returns 0 and sets the error string on bad input. */
size_t hvit_workspace_bytes(const hvit_model_cfg* cfg, int B, int F, int T, int n_samples);

/* Builds the launch plan (layer geometry, TMA tensor maps over `workspace_dev` and the weight buffers).
 * Replaces the module-construction half of HybridViT.__init__ / AudioEnhancer.__init__
 * (models/hybrid_vit.py:36-170, inference/enhancer.py:25-53).
 * `stream`: the one-time setup kernels (stem position matrices, FFT tables) are enqueued on it, after whatever
 * produced `weights` on that stream; the plan can be used on the same stream immediately, on another stream once
 * the caller has ordered it after this call.  No device-wide synchronisation. */
int hvit_plan_create(const hvit_model_cfg* cfg, const hvit_weights* weights, int B, int F, int T, int n_samples,
                     void* workspace_dev, size_t workspace_bytes, void* stream, hvit_plan** plan_out);
void hvit_plan_destroy(hvit_plan* plan);
/* Debug mode (tests): the plan additionally stores intermediates nothing downstream reads - the pre-tanh "logits" of the
 * head and, on the enhance path, the resized model output "model_out" [B,257,T].  Off by default (they cost HBM
 * traffic: 33 MB + 2 MB per 64 x 4 s batch). */
int hvit_plan_set_debug(hvit_plan* plan, int on);

/* HybridViT.forward (models/hybrid_vit.py:396-469), eval mode.
 *   x_dev: fp32 [B,1,F,T]   y_dev: fp32 [B,1,F,T]
 *   attn_probs_dev: null, or fp32 [num_layers][B][heads][N][N] to receive the softmax maps
 *                   (return_attentions=True, hybrid_vit.py:422-450; written by the tensor-core attention kernel itself
 *                   up to 1 280 tokens).
 * Stream semantics (this call and hvit_enhance / hvit_enhance_varlen): everything is enqueued on `stream`, except the
 * skip-path kernels, which run on a side stream owned by the plan, forked from `stream` after the encoder and joined
 * back into it (events) before the first decoder block - so all of the call's work is ordered before whatever the
 * caller enqueues on `stream` next, the call never synchronises, and it can be captured into a CUDA graph.  A plan
 * must not be run from two streams at the same time (one workspace). */
int hvit_forward(hvit_plan* plan, const float* x_dev, float* y_dev, float* attn_probs_dev, void* stream);

/* AudioEnhancer.enhance (inference/enhancer.py:55-135) for a batch of equal-length clips, everything on the
 * device: peak-normalise -> STFT(512/128/hann, centred) -> |.|, per-clip max-normalise -> HybridViT.forward ->
 * de-normalise, recombine with the noisy phase -> iSTFT(length=n) -> de-normalise.
 *   wave_in_dev / wave_out_dev: fp32 [B, n_samples]. */
int hvit_enhance(hvit_plan* plan, const float* wave_in_dev, float* wave_out_dev, int normalize, void* stream);

/* Variable-length batches (SURVEY.md section 8 f rank 2; the reference has the plumbing - key mask
 * models/attention.py:94-98, zero-pad collate data/dataset.py:297-347 - but never connects it).  wave_in_dev
 * [B, n_samples] holds clips zero-padded to the plan's n_samples, n_valid_dev [B] (device int32) their true lengths
 * (hvit_varlen_min_samples(plan) <= n <= n_samples; values outside are clamped).  Every clip is processed exactly as
 * AudioEnhancer.enhance (inference/enhancer.py:55-135) would process it alone: its own frame count, right-border zero
 * padding in every convolution, token order / positional rows / attention keys of its own patch grid, bilinear
 * resizes between its own widths, iSTFT to its own length.  wave_out_dev [B, n_samples]: samples >= n are zero. */
int hvit_enhance_varlen(hvit_plan* plan, const float* wave_in_dev, float* wave_out_dev, const int* n_valid_dev,
                        int normalize, void* stream);
int hvit_varlen_min_samples(const hvit_plan* plan);

/* Objective metrics of the evaluation caller on the device (evaluation/metrics.py:100-296, used by
 * evaluation/evaluator.py:119-231): SI-SDR, SNR, segmental SNR (512 / 256 frames, clipped to [-10, 35] dB) and
 * log-spectral distance (|STFT| 512 / 128) between clean_dev and enhanced_dev, both fp32 [B, n_samples] on the device
 * (n_valid_dev: nullable int32 [B] true lengths of zero-padded clips).  out_dev: [B][4] doubles
 * {sisdr, snr, segsnr, lsd}.  scratch_dev: hvit_metrics_scratch_bytes(B, n_samples) bytes, 256-byte aligned. */
size_t hvit_metrics_scratch_bytes(int B, int n_samples);
int hvit_metrics(const float* clean_dev, const float* enhanced_dev, int B, int n_samples, const int* n_valid_dev,
                 void* scratch_dev, size_t scratch_bytes, double* out_dev, void* stream);

/* Spectrogram losses of the validation forward (training/trainer.py:207-251 Trainer.validate with
 * training/losses.py:286-387 CombinedLoss): pred / target fp32 [B, n_per] on the device -> sums_dev [B,5] fp64 =
 * {sum |p' - t'|, sum (p' - t')^2, sum p^2, sum t^2, sum p t} per sample, p' = ln(p + 1e-8) when use_log (losses.py:46-57),
 * from which L1 / MSE (mean over all elements) and the reference's STOI proxy (1 - cosine similarity per sample,
 * losses.py:126-141) follow on the host. */
int hvit_spec_loss(const float* pred_dev, const float* target_dev, int B, long long n_per, int use_log, double* sums_dev,
                   void* stream);

/* Introspection for tests: byte offset (into the workspace), and dims of a named internal buffer.
 * Names: "enc<i>", "tokens", "ln", "qkv", "attn", "mlp", "cat<i>", "logits" (debug mode), "tanh", and for enhance
 * plans "model_out" (debug mode), "mag", "max_val", "mag_max".  dims receives up to 4 ints; returns the rank or
 * negative. */
int hvit_plan_buffer(const hvit_plan* plan, const char* name, size_t* offset, int* dims, int* elem_bytes);
/* Number of kernels one hvit_forward / hvit_enhance call launches. */
int hvit_plan_launch_count(const hvit_plan* plan, int enhance);
/* Measurement hooks (bench.py): the plan as a list of steps (enhance=1 includes peak/STFT/iSTFT), each with its
 * layer name, kernel family, algorithmic / executed FLOPs, compulsory HBM bytes and launch count; and one enhance
 * call with a CUDA event recorded on `stream` after every step (synchronises the stream, writes per-step
 * milliseconds to host memory). */
int hvit_plan_num_steps(const hvit_plan* plan, int enhance);
int hvit_plan_step_info(const hvit_plan* plan, int enhance, int i, char* name, int name_len, char* kernel,
                        int kernel_len, double* algo_flops, double* exec_flops, double* algo_bytes, int* launches);
int hvit_enhance_profiled(hvit_plan* plan, const float* wave_in_dev, float* wave_out_dev, int normalize, void* stream,
                          float* step_ms_host, int n_steps);
/* Token count N and patch grid of the plan. */
int hvit_plan_tokens(const hvit_plan* plan, int* hp, int* wp);

/* ---- per-kernel entry points (unit tests, microbenchmarks) ------------------------------------------------ */

/* C[M,N] = act(A[M,K] * W[N,K]^T * scale[n] + shift[n]) (+ residual), 16-bit operands on tcgen05
 * (f16 = 0: bf16, 1: fp16).  nn.Linear (attention.py:83,109; components.py:223-229; hybrid_vit.py:343).
 * out is 16-bit (same type as the operands) or fp32 (out_f32). */
int hvit_gemm_16(const void* a_dev, int lda, const void* w_dev, const float* scale_dev, const float* shift_dev,
                 int act, const float* residual_dev, int ldr, void* out_dev, int ldc, int out_f32, int M, int N,
                 int K, int f16, void* stream);
/* nn.LayerNorm folded into the linears on either side of it (attention.py:152-153,258-262 pre-norm blocks;
 * hybrid_vit.py:343 final norm + to_feature_map) - what the 16-bit plans run instead of a LayerNorm kernel:
 *   producer  x[M,N] += a[M,K] * w[N,K]^T + bias  (fp32, in place), plus x16_out = 16-bit(x) and, per row and per
 *             128-column slot, the slot's (mean, centred sum of squares) in stats_out [M, N/128, 2]; N % 256 == 0
 *   consumer  out = act(LayerNorm(x; gamma, beta, eps) * w[N,K]^T + bias) from x16 and the statistics; w_scratch
 *             ([N,K] 16-bit) and gc_scratch ([N] fp32) receive the gamma-scaled, k-centred weights and the c vector
 *   rowstats  x16 and statistics of an existing fp32 matrix (the first block's input) */
int hvit_linear_ln_producer_16(const void* a_dev, int lda, const void* w_dev, const float* bias_dev, float* x_dev,
                               void* x16_out_dev, float* stats_out_dev, int M, int N, int K, int f16, void* stream);
int hvit_linear_ln_consumer_16(const void* x16_dev, const float* stats_dev, int slots, const void* w_dev,
                               const float* gamma_dev, const float* beta_dev, const float* bias_dev, float eps, int act,
                               void* out_dev, int ldc, int M, int N, int K, int f16, void* w_scratch_dev,
                               float* gc_scratch_dev, void* stream);
int hvit_rowstats_16(const float* x_dev, void* x16_out_dev, float* stats_out_dev, int rows, int D, int slots, int f16,
                     void* stream);
/* Same contract on CUDA cores in fp32 (a, w, out fp32). */
int hvit_gemm_f32(const float* a_dev, int lda, const float* w_dev, const float* scale_dev, const float* shift_dev,
                  int act, const float* residual_dev, int ldr, float* out_dev, int ldc, int M, int N, int K,
                  void* stream);
/* 3x3 / pad 1 conv + per-channel scale/shift + ReLU (+ 2x2 max-pool) on NHWC bf16/fp16, tcgen05 implicit GEMM.
 * ConvBlock / TransposeConvBlock in eval mode (components.py:15-99,102-192).
 * up2 = 1: nearest x2 upsample first; w_dev then holds the four parity kernels [4][Cout][2][2][Cin]. */
int hvit_conv3x3_16(const void* x_dev, const void* w_dev, const float* scale_dev, const float* shift_dev, int relu,
                    int pool, int up2, void* out_dev, int B, int H, int W, int Cin, int Cout, int f16, void* stream);
int hvit_conv3x3_f32(const float* x_dev, const float* w_dev, const float* scale_dev, const float* shift_dev,
                     int relu, int up2, float* out_dev, int B, int H, int W, int Cin, int Cout, void* stream);
/* Encoder block 0 (components.py:15-99 as built at hybrid_vit.py:196-209): Conv3x3(1 -> C, pad 1, no bias) + folded
 * BatchNorm + ReLU + MaxPool(pool) on the fp32 spectrogram x [B,H,W] (divided by mag_max[b] when mag_max_dev != NULL,
 * enhancer.py:96-101), NHWC bf16/fp16 out [B,H/pool,W/pool,C].  w_dev fp32 [3][3][C].  C = 64, pool = 2 run on the
 * tensor cores (stem_tc.cu; use_tc = 0 forces the CUDA-core kernel).  scratch_dev: 64 KB, 16-byte aligned. */
int hvit_stem_16(const float* x_dev, const void* mag_max_dev, const float* w_dev, const float* scale_dev,
                 const float* shift_dev, void* out_dev, void* scratch_dev, int B, int H, int W, int C, int pool,
                 int f16, int use_tc, void* stream);
/* Last decoder block (components.py:160-167): Conv3x3(C -> 1, pad 1, no bias) on NHWC bf16/fp16 x [B,H,W,C], fp32
 * accumulation; logits_dev (nullable) receives the pre-activation, tanh_dev the tanh.  w_dev fp32 [3][3][C]. */
int hvit_head_16(const void* x_dev, const float* w_dev, float* logits_dev, float* tanh_dev, int B, int H, int W, int C,
                 int f16, void* stream);
/* PatchEmbedding + PositionalEncoding (models/components.py:282-307,310-386): p x p / stride p convolution with bias on
 * NHWC bf16/fp16 x [B,H,W,C] (H % p == 0; trailing W % p columns are dropped like the reference's conv does), plus
 * pos[:N] -> fp32 tokens [B*N, D], N = (H/p)*(W/p), token n = h'*(W/p) + w'.  w_dev [D][p][p][C] 16-bit, pos_dev fp32
 * [>= N][D]. */
int hvit_patch_embed_16(const void* x_dev, int B, int H, int W, int C, const void* w_dev, const float* bias_dev,
                        const float* pos_dev, int patch, int D, float* tokens_dev, int f16, void* stream);
/* Skip path of forward_decoder for one decoder block (models/hybrid_vit.py:367-389): 1x1 projection with bias of the
 * encoder feature src [B,Hs,Ws,Cs], bilinear resize (align_corners=False) to [Hd,Wd], written into channels
 * [c_off, c_off + Cdec) of the NHWC concat buffer cat [B,Hd,Wd,Ccat] (torch.cat never exists).  Evaluated as
 * sample-then-project, which is exact.  w_dev [Cdec][Cs] 16-bit; scratch_dev: B*Hd*Wd*Cs 16-bit elements. */
int hvit_skip_concat_16(const void* src_dev, int B, int Hs, int Ws, int Cs, const void* w_dev, const float* bias_dev,
                        int Cdec, void* cat_dev, int Hd, int Wd, int Ccat, int c_off, void* scratch_dev, int f16,
                        void* stream);
/* Multi-head self-attention core, head_dim 64 (attention.py:86-105). qkv: [B*N, 3D]; out: [B*N, D]. */
int hvit_attention_16(const void* qkv_dev, void* out_dev, int B, int N, int heads, int f16, void* stream);
int hvit_attention_f32(const float* qkv_dev, float* out_dev, float* probs_dev, int B, int N, int heads, void* stream);
/* nn.LayerNorm over the last dim (attention.py:152-153,258); x fp32, out_dtype 0 = fp32, 1 = bf16, 2 = fp16. */
int hvit_layernorm(const float* x_dev, const float* g_dev, const float* b_dev, void* out_dev, int out_dtype, int rows,
                   int D, float eps, void* stream);
/* compute_stft + compute_magnitude_phase + normalize_audio (utils/audio_processing.py:67-98,135-174).
 *   wave [B,n] -> max_val [B] (u32 bit patterns of the fp32 peaks; 1.0 when normalize=0), spec complex64 [B,257,T],
 *   mag fp32 [B,257,T] (of the peak-normalised audio), mag_max [B] (u32 bit patterns). */
int hvit_stft(const float* wave_dev, int B, int n, int normalize, void* max_val_dev, void* spec_dev, float* mag_dev,
              void* mag_max_dev, void* stream);
/* reconstruct_from_magnitude_phase + compute_istft (utils/audio_processing.py:101-132,177-193).
 *   mag_norm [B,257,T] (model output, multiplied by mag_max inside), phase taken from spec; frames_dev is a
 *   [B,T,512] fp32 scratch. */
int hvit_istft(const float* mag_norm_dev, const void* spec_dev, const void* mag_max_dev, const void* max_val_dev,
               float* frames_dev, float* wave_out_dev, int B, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HVIT_H_ */
