// C ABI of libhvit_sm100.so (declared in include/hvit.h): launch plan for HybridViT.forward /
// AudioEnhancer.enhance, TMA tensor-map construction, per-kernel test entry points, error plumbing.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "hvit.h"
#include "kernels.h"

namespace hvit {

// ------------------------------------------------------------------------------------------ errors
static thread_local char g_err[768] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HVIT_NO_PDL");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

static thread_local L2Window g_l2_window = {nullptr, 0, 0.f};
const L2Window& l2_window() { return g_l2_window; }
void l2_window_set(void* base, size_t bytes, float hit_ratio) { g_l2_window = L2Window{base, bytes, hit_ratio}; }

// ------------------------------------------------------------------------------------------ per-device state
// One process may drive several GPUs (a model on cuda:1 after one on cuda:0): everything cached below is keyed by the
// CUDA device that is current when it is asked for.
int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    dev = 0;
  }
  return dev < 0 ? 0 : (dev >= kMaxDevices ? kMaxDevices - 1 : dev);
}

struct DeviceState {
  int sms = 0;                 // multiprocessor count
  int max_persist = -1;        // cudaDevAttrMaxPersistingL2CacheSize (-1: not queried yet, 0: unavailable / disabled)
  int max_window = 0;          // cudaDevAttrMaxAccessPolicyWindowSize
  size_t carve = 0;            // current cudaLimitPersistingL2CacheSize set by this library
  size_t carve_before = 0;     // the limit found before the library first raised it (restored when the last plan goes)
  int pinned_plans = 0;        // live plans that use the carve-out
};
static DeviceState g_dev[kMaxDevices];

// Persisting-L2 window for the residual stream: returns the window size in bytes (<= the device's maximum access-policy
// window) and its hit ratio for a buffer of `bytes`; 0 bytes = feature unavailable or disabled with HVIT_NO_L2PIN=1.
static size_t l2_pin_window(size_t bytes, float* ratio) {
  *ratio = 0.f;
  const int dev = current_device();
  DeviceState& d = g_dev[dev];
  if (d.max_persist < 0) {
    d.max_persist = 0;
    const char* e = getenv("HVIT_NO_L2PIN");
    if (!(e != nullptr && e[0] == '1')) {
      int v = 0, wv = 0;
      if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, dev) == cudaSuccess && v > 0 &&
          cudaDeviceGetAttribute(&wv, cudaDevAttrMaxAccessPolicyWindowSize, dev) == cudaSuccess && wv > 0) {
        d.max_persist = v;
        d.max_window = wv;
      }
      cudaGetLastError();
    }
  }
  if (d.max_persist <= 0 || bytes == 0) return 0;
  // the window itself may not exceed the device limit (a larger num_bytes makes the launch attribute invalid): pin a
  // prefix of the buffer; the hit ratio then spreads the carve-out over that prefix
  const size_t window = bytes < static_cast<size_t>(d.max_window) ? bytes : static_cast<size_t>(d.max_window);
  size_t want = window < static_cast<size_t>(d.max_persist) ? window : static_cast<size_t>(d.max_persist);
  if (const char* e = getenv("HVIT_L2PIN_MB")) {  // tuning knob: carve-out size in MB (<= the device maximum)
    const size_t mb = static_cast<size_t>(atoi(e)) << 20;
    if (mb > 0 && mb < want) want = mb;
  }
  if (want > d.carve) {
    if (d.carve == 0) {
      size_t before = 0;
      if (cudaDeviceGetLimit(&before, cudaLimitPersistingL2CacheSize) == cudaSuccess) d.carve_before = before;
      cudaGetLastError();
    }
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) {
      cudaGetLastError();
      d.max_persist = 0;
      return 0;
    }
    d.carve = want;
  }
  const float r = static_cast<float>(static_cast<double>(d.carve) / static_cast<double>(window));
  *ratio = r > 1.f ? 1.f : r;
  return window;
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return HVIT_E_LAUNCH;
  }
  return HVIT_OK;
}

// ------------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    const cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
      cudaGetLastError();
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Element type of a tensor map: the two 16-bit operand types of the tensor-core path, or fp32 (epilogue outputs).
enum TmapType { TM_BF16 = 0, TM_F16 = 1, TM_F32 = 2 };
static TmapType tm16(int f16) { return f16 ? TM_F16 : TM_BF16; }

// 128-byte swizzle (unless no_swizzle), zero fill out of bounds. dims/box innermost first; strides (bytes) for dims 1..rank-1.
static int make_tmap(CUtensorMap* m, const void* base, TmapType type, int rank, const uint64_t* dims,
                     const uint64_t* strides, const uint32_t* box, int no_swizzle = 0) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return HVIT_E_LAUNCH;
  }
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides[i - 1];
  }
  const CUtensorMapDataType dt = type == TM_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : (type == TM_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  const CUresult r = fn(m, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                        gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        no_swizzle ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu,...] box=[%u,%u,...]",
              static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0);
    return HVIT_E_LAUNCH;
  }
  return HVIT_OK;
}

static int tmap_matrix(CUtensorMap* m, const void* base, int f16, long long rows, long long cols, long long ld, int box_rows) {
  const uint64_t dims[2] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(rows)};
  const uint64_t strides[1] = {static_cast<uint64_t>(ld) * 2};
  const uint32_t box[2] = {64, static_cast<uint32_t>(box_rows)};
  return make_tmap(m, base, tm16(f16), 2, dims, strides, box);
}

static int tmap_image(CUtensorMap* m, const void* base, int f16, int B, int H, int Hpitch, int W, int C, int Wt, int Ht) {
  const uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(W) * C * 2,
                               static_cast<uint64_t>(Hpitch) * W * C * 2};
  const uint32_t box[4] = {64, static_cast<uint32_t>(Wt), static_cast<uint32_t>(Ht), 1};
  return make_tmap(m, base, tm16(f16), 4, dims, strides, box);
}

static int tmap_patch(CUtensorMap* m, const void* base, int f16, int B, int Hq, int W, int C, int p, int Wp, int Wt, int Ht) {
  const uint64_t dims[5] = {static_cast<uint64_t>(C), static_cast<uint64_t>(p), static_cast<uint64_t>(Wp),
                            static_cast<uint64_t>(p), static_cast<uint64_t>(B) * Hq};
  const uint64_t strides[4] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(p) * C * 2,
                               static_cast<uint64_t>(W) * C * 2, static_cast<uint64_t>(p) * W * C * 2};
  const uint32_t box[5] = {64, 1, static_cast<uint32_t>(Wt), 1, static_cast<uint32_t>(Ht)};
  return make_tmap(m, base, tm16(f16), 5, dims, strides, box);
}

static int tmap_qkv(CUtensorMap* m, const void* base, int f16, int B, int N, int D) {
  const uint64_t dims[3] = {static_cast<uint64_t>(3) * D, static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
  const uint64_t strides[2] = {static_cast<uint64_t>(3) * D * 2, static_cast<uint64_t>(N) * 3 * D * 2};
  const uint32_t box[3] = {64, 128, 1};
  return make_tmap(m, base, tm16(f16), 3, dims, strides, box);
}

// attention output [B*N, D] 16-bit as a (D, N, B) map: 128-row x 64-column (one head) store boxes, clipped at N
static int tmap_attn_out(CUtensorMap* m, const void* base, int f16, int B, int N, int D) {
  const uint64_t dims[3] = {static_cast<uint64_t>(D), static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
  const uint64_t strides[2] = {static_cast<uint64_t>(D) * 2, static_cast<uint64_t>(N) * D * 2};
  const uint32_t box[3] = {64, 128, 1};
  return make_tmap(m, base, tm16(f16), 3, dims, strides, box);
}

// Number of row panels the MLP (fc1 -> GELU -> fc2) of a transformer block is split into (see build_steps).
// HVIT_MLP_PANELS overrides; default: as few panels as keep one panel's hidden activation under ~40 MB.
static int mlp_panels(int M, int hidden, int es, int tensor_core) {
  if (const char* e = getenv("HVIT_MLP_PANELS")) {
    const int v = atoi(e);
    if (v >= 1) return v;
  }
  (void)M; (void)hidden; (void)es; (void)tensor_core;
  return 1;
}

// LayerNorm folded into the GEMMs around it (16-bit modes): needs 256-column producer tiles (statistics slots of 128
// columns) and the register-resident row-statistics kernel.  OFF by default - measured a wash at 64 x 4 s (device leg
// 2.96 vs 3.01 ms per step, end-to-end leg 85.7k vs 86.4k audio-s/s; profiles/r2_experiments.json): the 13 LayerNorm
// launches it removes cost 37 us per layer, but the residual update then has to LOAD the old rows - 65 MB more L2 -> SM
// traffic per GEMM on kernels that already sit on that delivery rate (proj 32 -> 51 us, fc2 65 -> 78 us; the TMA
// reduce-add of the unfolded path never brings them into the SM) - and the consumers' epilogues grow (qkv +3, fc1 +10 us).
// HVIT_LN_FOLD=1 builds plans with it (read at plan creation; parity-tested both ways).
static bool ln_fold_enabled(int precision, int D, int layers) {
  const char* e = getenv("HVIT_LN_FOLD");
  const bool on = e != nullptr && e[0] == '1';
  return on && precision != HVIT_PREC_FP32 && layers > 0 && D % 256 == 0 && D <= 1024;
}

static int pick_block_n(int N) { return N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : 64); }

static void pick_tile(int H, int W, int* Wt, int* Ht) {
  const int cand[4][2] = {{128, 1}, {64, 2}, {32, 4}, {16, 8}};
  long long best = -1;
  for (int i = 0; i < 4; ++i) {
    const long long t = static_cast<long long>((W + cand[i][0] - 1) / cand[i][0]) * ((H + cand[i][1] - 1) / cand[i][1]);
    if (best < 0 || t < best) {
      best = t;
      *Wt = cand[i][0];
      *Ht = cand[i][1];
    }
  }
}

// 4-D output (and residual) maps of the TMA-store epilogue: dim 0 = channel / column (128-byte box), dims 1..3 =
// the tile's spatial / row coordinates as the kernel passes them (see igemm_tc2_kernel).
static int tmap_out4(CUtensorMap* m, const void* base, TmapType type, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                     uint64_t s1, uint64_t s2, uint64_t s3, uint32_t b1, uint32_t b2) {
  const uint64_t dims[4] = {d0, d1, d2, d3};
  const uint64_t strides[3] = {s1, s2, s3};
  const uint32_t box[4] = {type == TM_F32 ? 32u : 64u, b1, b2, 1};
  return make_tmap(m, base, type, 4, dims, strides, box);
}

static int make_out_maps(const IgemmParams& q, IgemmMaps* mp) {
  const TmapType ot = q.out_f32 ? TM_F32 : tm16(q.f16);
  const uint64_t es = q.out_f32 ? 4 : 2;
  const uint64_t ld = static_cast<uint64_t>(q.ldc) * es;
  uint8_t* out = reinterpret_cast<uint8_t*>(q.out);
  const uint64_t N = q.N;
  int r = HVIT_OK;
  memset(mp->out, 0, sizeof(mp->out));
  memset(&mp->res, 0, sizeof(mp->res));
  if (q.mode == IG_PLAIN) {
    r = tmap_out4(&mp->out[0], out, ot, N, q.M, 1, 1, ld, ld * q.M, ld * q.M, 128, 1);
    if (r == HVIT_OK && q.residual != nullptr) {
      const uint64_t lr = static_cast<uint64_t>(q.ldr) * 4;
      r = tmap_out4(&mp->res, q.residual, TM_F32, N, q.M, 1, 1, lr, lr * q.M, lr * q.M, 128, 1);
    }
  } else if (q.mode == IG_CONV3) {
    const uint64_t img = static_cast<uint64_t>(q.HoPitch) * q.Wo * ld;
    if (q.pool)
      r = tmap_out4(&mp->out[0], out, ot, N, q.Wo, q.Ho, q.B, ld, ld * q.Wo, img, q.Wt / 2, q.Ht / 2);
    else
      r = tmap_out4(&mp->out[0], out, ot, N, q.W, q.H, q.B, ld, ld * q.Wo, img, q.Wt, q.Ht);
  } else if (q.mode == IG_UP2) {
    const uint64_t img = static_cast<uint64_t>(q.HoPitch) * q.Wo * ld;
    for (int par = 0; par < 4 && r == HVIT_OK; ++par) {
      const uint64_t off = (static_cast<uint64_t>(par >> 1) * q.Wo + (par & 1)) * ld;
      r = tmap_out4(&mp->out[par], out + off, ot, N, q.W, q.H, q.B, 2 * ld, 2 * ld * q.Wo, img, q.Wt, q.Ht);
    }
  } else {  // IG_PATCH
    const uint64_t np = static_cast<uint64_t>(q.Hp) * q.Wp;
    r = tmap_out4(&mp->out[0], out, ot, N, q.Wp, q.Hp, q.B, ld, ld * q.Wp, ld * np, q.Wt, q.Ht);
    if (r == HVIT_OK && q.residual != nullptr) {
      const uint64_t lr = static_cast<uint64_t>(q.ldr) * 4;
      r = tmap_out4(&mp->res, q.residual, TM_F32, N, q.Wp, q.Hp, q.res_mod > 0 ? 1 : q.B, lr, lr * q.Wp, lr * np, q.Wt, q.Ht);
    }
  }
  return r;
}

int num_sms() {
  const int dev = current_device();
  DeviceState& d = g_dev[dev];
  if (d.sms == 0) {
    cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
    if (d.sms <= 0) d.sms = 148;
    if (const char* e = getenv("HVIT_NUM_SMS")) {  // diagnostics: run the persistent kernels on fewer SMs
      const int v = atoi(e);
      if (v >= 2 && v < d.sms) d.sms = v & ~1;
    }
  }
  return d.sms;
}

// ------------------------------------------------------------------------------------------ plan
struct Ctx {
  const float* x;            // model input [B,F,T]
  float* y;                  // model output [B,F,T]
  float* probs;              // optional attention maps
  const unsigned* mag_max;   // per-clip magnitude max (enhance) or null
  const float* wave_in;      // enhance only
  float* wave_out;           // enhance only
  int normalize;             // enhance only
  int fused_resize;          // enhance: the final bilinear resize is fused into the iSTFT frame kernel
  const int* geo;            // hvit_enhance_varlen: per-clip geometry table (kernels.h), null for equal-length batches
  cudaStream_t stream;
};
typedef std::function<int(const Ctx&)> Step;

// Book-keeping for measurement (bench.py roofline): what each step is and how much work it represents.
struct StepMeta {
  std::string name;      // layer name, e.g. "blocks.3.fc1"
  std::string kernel;    // kernel family, e.g. "igemm_tc"
  double algo_flops;     // multiply-add FLOPs of the reference graph for this op (torch FlopCounter convention)
  double exec_flops;     // FLOPs actually executed (differs when an exact algebraic shortcut is used)
  double algo_bytes;     // compulsory HBM traffic (inputs read once + outputs written once)
  int launches;
};

struct Buf {
  size_t off;
  int rank;
  int dims[4];
  int es;
};

struct EncGeo {
  int H, W, C, pitch;
};
struct CatGeo {
  int H, W, Cx, Ccat;
};

struct Geometry {
  int es, B, F, T, n_samples;
  EncGeo enc[HVIT_MAX_STAGES];
  int Hp, Wp, Np, M;
  CatGeo cat[HVIT_MAX_STAGES];  // cat[i] = input of decoder block i
  std::map<std::string, Buf> bufs;
  size_t total;
};

}  // namespace hvit

struct hvit_plan {
  hvit_model_cfg cfg;
  hvit_weights w;
  hvit::Geometry g;
  uint8_t* ws;
  std::vector<hvit::Step> steps;      // HybridViT.forward
  std::vector<hvit::StepMeta> meta;
  std::vector<hvit::Step> pre, post;  // enhance: peak/STFT before, iSTFT after
  std::vector<hvit::StepMeta> pre_meta, post_meta;
  int launches_forward;
  void* l2_pin_base = nullptr;        // residual stream kept in persisting L2 lines during the forward steps
  size_t l2_pin_bytes = 0;            // access-policy window (<= cudaDevAttrMaxAccessPolicyWindowSize), 0 = no pinning
  float l2_pin_ratio = 0.f;
  cudaStream_t setup_stream = nullptr; // stream of the one-time setup kernels (hvit_plan_create's argument)
  int device = -1;                    // device whose persisting-L2 carve-out this plan counts against (-1: none)
  int debug = 0;                      // keep test-only intermediates ("model_out", "logits"), see hvit_plan_set_debug
  // Skip path on a side stream: skip.i.sample / skip.i.proj depend only on the encoder outputs, so they are forked off
  // after the last encoder step and joined before the first decoder conv - three small, launch-bound sample + 1x1-GEMM
  // pairs (~110 us at 64 x 4 s) that run under the transformer's LayerNorm / attention launches instead of in line.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int side_fork = -1, side_join = -1;  // step indices: fork before step side_fork, join before step side_join
  std::vector<char> is_side;
  // tag the most recently pushed step(s)
  void tag(const std::string& name, const char* kernel, double aflops, double eflops, double bytes, int launches = 1) {
    while (meta.size() < steps.size()) meta.push_back(hvit::StepMeta{name, kernel, 0.0, 0.0, 0.0, 1});
    hvit::StepMeta& m = meta.back();
    m.name = name; m.kernel = kernel; m.algo_flops = aflops; m.exec_flops = eflops; m.algo_bytes = bytes;
    m.launches = launches;
  }
};

namespace hvit {

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int build_geometry(const hvit_model_cfg& c, int B, int F, int T, int n_samples, int pos_len, Geometry& g) {
  if (B < 1 || F < 1 || T < 1) {
    set_error("bad batch/spectrogram shape B=%d F=%d T=%d", B, F, T);
    return HVIT_E_SHAPE;
  }
  if (c.n_enc < 1 || c.n_enc > HVIT_MAX_STAGES || c.n_dec < 2 || c.n_dec > HVIT_MAX_STAGES || c.num_layers < 0 ||
      c.num_layers > HVIT_MAX_LAYERS) {
    set_error("unsupported depth: n_enc=%d n_dec=%d layers=%d", c.n_enc, c.n_dec, c.num_layers);
    return HVIT_E_SHAPE;
  }
  if (c.num_heads < 1 || c.embed_dim != c.num_heads * 64) {
    set_error("head_dim must be 64: embed_dim=%d num_heads=%d", c.embed_dim, c.num_heads);
    return HVIT_E_SHAPE;
  }
  if (c.patch_size < 1 || c.patch_size > 16) {
    set_error("unsupported patch_size %d", c.patch_size);
    return HVIT_E_SHAPE;
  }
  if (c.precision != HVIT_PREC_FP32 && c.precision != HVIT_PREC_BF16 && c.precision != HVIT_PREC_FP16) {
    set_error("unknown precision %d", c.precision);
    return HVIT_E_SHAPE;
  }
  const bool bf = c.precision != HVIT_PREC_FP32;  // 16-bit tensor-core path
  const int cmul = bf ? 64 : 16;
  g.es = bf ? 2 : 4;
  g.B = B; g.F = F; g.T = T; g.n_samples = n_samples;
  int H = F, W = T;
  for (int i = 0; i < c.n_enc; ++i) {
    const int pool = c.enc_pool[i];
    if (pool != 1 && pool != 2) {
      set_error("encoder pool size must be 1 or 2 (block %d: %d)", i, pool);
      return HVIT_E_SHAPE;
    }
    if (c.enc_channels[i] % cmul != 0 || (i == 0 && c.enc_channels[i] > 512)) {
      set_error("encoder channels must be multiples of %d (block %d: %d)", cmul, i, c.enc_channels[i]);
      return HVIT_E_SHAPE;
    }
    H /= pool;
    W /= pool;
    if (H < 1 || W < 1) {
      set_error("input %dx%d too small for the encoder", F, T);
      return HVIT_E_SHAPE;
    }
    g.enc[i] = {H, W, c.enc_channels[i], H};
  }
  EncGeo& last = g.enc[c.n_enc - 1];
  last.pitch = (last.H + c.patch_size - 1) / c.patch_size * c.patch_size;
  g.Hp = last.H / c.patch_size;
  g.Wp = last.W / c.patch_size;
  g.Np = g.Hp * g.Wp;
  if (g.Np < 1) {
    set_error("input %dx%d yields no patches", F, T);
    return HVIT_E_SHAPE;
  }
  if (pos_len > 0 && g.Np > pos_len) {
    set_error("token count %d exceeds the positional table (%d)", g.Np, pos_len);
    return HVIT_E_SHAPE;
  }
  g.M = B * g.Np;
  if (c.embed_dim % 64 != 0 || c.mlp_hidden % 64 != 0) {
    set_error("embed_dim / mlp hidden must be multiples of 64");
    return HVIT_E_SHAPE;
  }
  if (c.dec_channels[0] != last.C) {
    set_error("decoder_channels[0] (%d) must equal encoder_channels[-1] (%d)", c.dec_channels[0], last.C);
    return HVIT_E_SHAPE;
  }
  if (c.dec_channels[c.n_dec - 1] != 1 || c.dec_up[c.n_dec - 1] != 1) {
    set_error("the last decoder block must be a 1-channel head without upsampling");
    return HVIT_E_SHAPE;
  }
  int Hx = g.Hp, Wx = g.Wp, Cx = last.C;
  for (int i = 0; i < c.n_dec; ++i) {
    const bool final_blk = i == c.n_dec - 1;
    const bool has_skip = c.use_skip && !final_blk && i < c.n_enc;
    if (!final_blk && c.dec_channels[i] % cmul != 0) {
      set_error("decoder channels must be multiples of %d (block %d: %d)", cmul, i, c.dec_channels[i]);
      return HVIT_E_SHAPE;
    }
    g.cat[i] = {Hx, Wx, Cx, Cx + (has_skip ? c.dec_channels[i] : 0)};
    if (!final_blk) {
      const int up = c.dec_up[i];
      if (up != 1 && up != 2) {
        set_error("decoder upsample factor must be 1 or 2 (block %d: %d)", i, up);
        return HVIT_E_SHAPE;
      }
      Hx *= up;
      Wx *= up;
      Cx = c.dec_channels[i];
    }
  }

  size_t off = 0;
  auto add = [&](const std::string& name, int es, int rank, int d0, int d1, int d2, int d3, size_t elems) {
    Buf b;
    b.off = off;
    b.rank = rank;
    b.dims[0] = d0; b.dims[1] = d1; b.dims[2] = d2; b.dims[3] = d3;
    b.es = es;
    g.bufs[name] = b;
    off = align_up(off + elems * es, 1024);
  };
  char nm[32];
  size_t tmp_elems = 0;
  {
    int h = F, w = T;
    for (int i = 0; i < c.n_enc; ++i) {
      if (i > 0 && c.enc_pool[i] == 2) tmp_elems = std::max(tmp_elems, static_cast<size_t>(B) * h * w * c.enc_channels[i]);
      const EncGeo& e = g.enc[i];
      snprintf(nm, sizeof(nm), "enc%d", i);
      add(nm, g.es, 4, B, e.pitch, e.W, e.C, static_cast<size_t>(B) * e.pitch * e.W * e.C);
      h = e.H; w = e.W;
    }
  }
  if (!bf && tmp_elems > 0) add("conv_tmp", 4, 1, static_cast<int>(tmp_elems), 0, 0, 0, tmp_elems);
  if (bf) add("stem_a", 2, 1, 4 * 128 * 64, 0, 0, 0, 4 * 128 * 64);
  const size_t M = g.M;
  add("tokens", 4, 2, g.M, c.embed_dim, 0, 0, M * c.embed_dim);
  add("ln", g.es, 2, g.M, c.embed_dim, 0, 0, M * c.embed_dim);
  if (ln_fold_enabled(c.precision, c.embed_dim, c.num_layers)) {
    // folded LayerNorm: "ln" holds the 16-bit copy of the residual stream; per-row statistics; gamma-scaled weight
    // copies of qkv / fc1 (per layer) and to_feature_map, their g and c vectors
    const size_t D_ = c.embed_dim, wl = 3 * D_ * D_ + static_cast<size_t>(c.mlp_hidden) * D_;
    const size_t Cx0 = g.cat[0].Cx;
    add("lnstats", 4, 3, g.M, c.embed_dim / 128, 2, 0, M * (c.embed_dim / 128) * 2);
    add("lnfold_w", 2, 1, 0, 0, 0, 0, c.num_layers * wl + Cx0 * D_);
    add("lnfold_v", 4, 1, 0, 0, 0, 0, c.num_layers * (3 * D_ + c.mlp_hidden) + Cx0);
  }
  add("qkv", g.es, 2, g.M, 3 * c.embed_dim, 0, 0, M * 3 * c.embed_dim);
  add("attn", g.es, 2, g.M, c.embed_dim, 0, 0, M * c.embed_dim);
  add("mlp", g.es, 2, g.M, c.mlp_hidden, 0, 0, M * c.mlp_hidden);
  size_t samp_elems = 0;
  for (int i = 0; i < c.n_dec; ++i) {
    const CatGeo& k = g.cat[i];
    snprintf(nm, sizeof(nm), "cat%d", i);
    add(nm, g.es, 4, B, k.H, k.W, k.Ccat, static_cast<size_t>(B) * k.H * k.W * k.Ccat);
    if (k.Ccat > k.Cx)
      samp_elems = std::max(samp_elems, static_cast<size_t>(B) * k.H * k.W * c.enc_channels[c.n_enc - 1 - i]);
  }
  if (samp_elems > 0) add("samp", g.es, 1, static_cast<int>(samp_elems), 0, 0, 0, samp_elems);
  const CatGeo& hl = g.cat[c.n_dec - 1];
  add("logits", 4, 3, B, hl.H, hl.W, 0, static_cast<size_t>(B) * hl.H * hl.W);
  add("tanh", 4, 3, B, hl.H, hl.W, 0, static_cast<size_t>(B) * hl.H * hl.W);
  if (n_samples > 0) {
    if (F != 257 || T != 1 + n_samples / 128) {
      set_error("enhance plan needs F=257 and T=1+n/128 (got F=%d T=%d n=%d)", F, T, n_samples);
      return HVIT_E_SHAPE;
    }
    add("model_out", 4, 3, B, F, T, 0, static_cast<size_t>(B) * F * T);  // written only by plans in debug mode (tests)
    add("max_val", 4, 1, B, 0, 0, 0, B);
    add("mag_max", 4, 1, B, 0, 0, 0, B);
    add("mag", 4, 3, B, F, T, 0, static_cast<size_t>(B) * F * T);
    // variable-length batches (hvit_enhance_varlen): per-clip geometry table, the patch grid before it is compacted
    // into each clip's own token order, and to_feature_map's rows before they are scattered back onto the grid
    add("geo", 4, 2, B, GEO_STRIDE, 0, 0, static_cast<size_t>(B) * GEO_STRIDE);
    add("tokgrid", 4, 2, g.M, c.embed_dim, 0, 0, M * c.embed_dim);
    add("tofm_rows", g.es, 2, g.M, g.cat[0].Cx, 0, 0, M * g.cat[0].Cx);
  }
  g.total = off;
  return HVIT_OK;
}

template <typename T>
static T* at(hvit_plan* p, const char* name) {
  return reinterpret_cast<T*>(p->ws + p->g.bufs.at(name).off);
}

static IgemmParams ig_zero() {
  IgemmParams q;
  memset(&q, 0, sizeof(q));
  return q;
}

// plain GEMM step (both precisions)
struct LnFold {
  // consumer
  const float* stats_in = nullptr;
  const float* c = nullptr;   // folded bias vector (passed to the GEMM as its shift)
  int slots = 0;
  float eps = 0.f;
  // producer
  void* x16_out = nullptr;
  int ld16 = 0;
  float* stats_out = nullptr;
};
static int add_linear(hvit_plan* p, const std::string& name, const void* A, int lda, const void* W,
                      const float* shift, int act, const float* residual, int ldr, int res_mod, void* out, int ldc,
                      int out_f32, int M, int N, int K, double algo_flops = -1.0, const LnFold* lf = nullptr) {
  const double fl = 2.0 * M * N * K;
  const int es_ = p->cfg.precision == HVIT_PREC_FP32 ? 4 : 2;
  const double by = (static_cast<double>(M) * K + static_cast<double>(N) * K) * es_ +
                    static_cast<double>(M) * N * (out_f32 ? 4 : es_) * (residual != nullptr && res_mod == 0 ? 2 : 1);
  IgemmParams q = ig_zero();
  q.mode = IG_PLAIN;
  q.M = M; q.N = N; q.K = K;
  q.shift = shift; q.act = act; q.residual = residual; q.ldr = ldr; q.res_mod = res_mod;
  q.out = out; q.ldc = ldc; q.out_f32 = out_f32;
  q.f16 = p->cfg.precision == HVIT_PREC_FP16;
  if (lf != nullptr) {
    q.ln_stats_in = lf->stats_in; q.ln_slots = lf->slots; q.ln_eps = lf->eps;
    q.ln_x16_out = lf->x16_out; q.ld16 = lf->ld16; q.ln_stats_out = lf->stats_out;
  }
  if (p->cfg.precision != HVIT_PREC_FP32) {
    IgemmMaps mp;
    int bn = pick_block_n(N);
    if (K >= 1024 && out_f32 && lf == nullptr) {
      if (const char* e = getenv("HVIT_FC2_BN")) {  // experiment: narrower tiles for the long-K fp32 GEMM (fc2)
        const int v = atoi(e);
        if ((v == 64 || v == 128 || v == 256) && N % v == 0) bn = v;
      }
    }
    int r = tmap_matrix(&mp.a, A, q.f16, M, K, lda, 128);
    if (r) return r;
    r = tmap_matrix(&mp.b, W, q.f16, N, K, K, bn / 2);
    if (r) return r;
    r = make_out_maps(q, &mp);
    if (r) return r;
    const int sms = num_sms();
    p->steps.push_back([=](const Ctx& c) { return launch_igemm_tc2(q, mp, bn, sms, c.stream); });
    p->tag(name, "igemm_tc", algo_flops >= 0 ? algo_flops : fl, fl, by);
  } else {
    q.out_f32 = 1;
    const float* Af = reinterpret_cast<const float*>(A);
    const float* Wf = reinterpret_cast<const float*>(W);
    p->steps.push_back([=](const Ctx& c) { return launch_igemm_f32(q, Af, lda, Wf, c.stream); });
    p->tag(name, "igemm_f32", algo_flops >= 0 ? algo_flops : fl, fl, by);
  }
  return HVIT_OK;
}

// 3x3 conv (+ optional fused pool / x2 upsample) step
static int add_conv(hvit_plan* p, const std::string& name, const void* in, int B, int H, int W, int Cin,
                    const void* Wt, const float* scale,
                    const float* shift, int relu, int pool, int up2, void* out, int ldc, int HoPitch, int Cout,
                    float* conv_tmp) {
  IgemmParams q = ig_zero();
  q.mode = up2 ? IG_UP2 : IG_CONV3;
  q.B = B; q.H = H; q.W = W; q.Cin = Cin;
  q.N = Cout;
  q.scale = scale; q.shift = shift; q.act = relu ? ACT_RELU : ACT_NONE;
  q.out = out; q.ldc = ldc;
  const int Hfull = up2 ? 2 * H : H, Wfull = up2 ? 2 * W : W;
  const double algo_fl = 2.0 * B * Hfull * Wfull * Cout * 9.0 * Cin;
  const int es_ = p->cfg.precision == HVIT_PREC_FP32 ? 4 : 2;
  const double by = (static_cast<double>(B) * H * W * Cin + static_cast<double>(B) * Hfull * Wfull * Cout / (pool ? 4 : 1) +
                     9.0 * Cin * Cout) * es_;
  q.f16 = p->cfg.precision == HVIT_PREC_FP16;
  if (p->cfg.precision != HVIT_PREC_FP32) {
    q.K = (up2 ? 4 : 9) * Cin;
    q.pool = pool;
    q.out_f32 = 0;
    q.Ho = pool ? H / 2 : Hfull;
    q.Wo = pool ? W / 2 : Wfull;
    q.HoPitch = HoPitch;
    // 3x3 convs keep the input tile + halo in shared memory (igemm_halo_kernel); HVIT_NO_HALO=1 selects the
    // tap-shifted TMA boxes of igemm_tc2_kernel instead (A/B comparison)
    static const bool no_halo = getenv("HVIT_NO_HALO") != nullptr && getenv("HVIT_NO_HALO")[0] == '1';
    q.halo = (!no_halo && Cin % 64 == 0 && Cout <= 512 && (!up2 || pick_block_n(Cout) <= 128)) ? 1 : 0;
    if (q.halo) {
      q.Wt = 8; q.Ht = 16;
    } else if (pool) {
      q.Wt = 16; q.Ht = 8;
    } else {
      pick_tile(H, W, &q.Wt, &q.Ht);
    }
    q.tiles_w = (W + q.Wt - 1) / q.Wt;
    q.tiles_h = (H + q.Ht - 1) / q.Ht;
    IgemmMaps mp;
    const int bn = pick_block_n(Cout);
    int r;
    if (q.halo) {  // (8 channels, W, H, Cin / 8, B): one un-swizzled box = tile + halo of 64 channels as 8 chunk planes
      const uint64_t dims[5] = {8, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(Cin / 8),
                                static_cast<uint64_t>(B)};
      const uint64_t strides[4] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(W) * Cin * 2, 16,
                                   static_cast<uint64_t>(H) * W * Cin * 2};
      const uint32_t box[5] = {8, 10, 18, 8, 1};
      r = make_tmap(&mp.a, in, tm16(q.f16), 5, dims, strides, box, 1);
    } else {
      r = tmap_image(&mp.a, in, q.f16, B, H, H, W, Cin, q.Wt, q.Ht);
    }
    if (r) return r;
    r = tmap_matrix(&mp.b, Wt, q.f16, static_cast<long long>(up2 ? 4 : 1) * Cout, q.K, q.K, bn / 2);
    if (r) return r;
    r = make_out_maps(q, &mp);
    if (r) return r;
    const int sms = num_sms();
    p->steps.push_back([=](const Ctx& c) { return launch_igemm_tc2(q, mp, bn, sms, c.stream); });
    p->tag(name, "igemm_tc", algo_fl, up2 ? algo_fl * 4.0 / 9.0 : algo_fl, by);
  } else {
    q.K = 9 * Cin;
    q.out_f32 = 1;
    q.Ho = Hfull; q.Wo = Wfull;
    const float* inf = reinterpret_cast<const float*>(in);
    const float* wf = reinterpret_cast<const float*>(Wt);
    if (pool) {
      q.out = conv_tmp; q.ldc = Cout; q.HoPitch = Hfull;
      float* dst = reinterpret_cast<float*>(out);
      p->steps.push_back([=](const Ctx& c) { return launch_igemm_f32(q, inf, Cin, wf, c.stream); });
      p->tag(name, "igemm_f32", algo_fl, algo_fl, by);
      p->steps.push_back([=](const Ctx& c) { return launch_maxpool2(conv_tmp, dst, B, H, W, Cout, c.stream); });
      p->tag(name + ".pool", "maxpool2", 0, 0, 1.25 * B * H * W * Cout * 4.0);
    } else {
      q.HoPitch = HoPitch;
      p->steps.push_back([=](const Ctx& c) { return launch_igemm_f32(q, inf, Cin, wf, c.stream); });
      p->tag(name, "igemm_f32", algo_fl, algo_fl, by);
    }
  }
  return HVIT_OK;
}

// variable-length batches only: zero the pixel columns beyond each clip's own width (no-op for equal-length batches).
// A zero-padded clip then looks to the next 3x3 conv exactly like the reference's zero padding at the right border.
static void add_mask(hvit_plan* p, const std::string& name, void* buf, int H, int Hpitch, int Wmax, int pix_bytes,
                     int geo_idx) {
  const int B = p->g.B;
  p->steps.push_back([=](const Ctx& k) {
    return k.geo != nullptr ? launch_zero_cols(buf, B, H, Hpitch, Wmax, pix_bytes, k.geo, geo_idx, k.stream) : 0;
  });
  p->tag(name, "zero_cols", 0, 0, 0, 0);
}

// patch embedding step: p x p / stride p conv (+bias) + positional table -> fp32 tokens [B * Hp * Wp, D]
// (PatchEmbedding + PositionalEncoding, models/components.py:282-307,310-386); token n = h' * Wp + w'
static int add_patch_embed(hvit_plan* p, const void* in, int B, int H, int Hpitch, int W, int C, const void* pw,
                           const float* pb, const float* pos, int patch, int D, float* tokens) {
  const int bf = p->cfg.precision != HVIT_PREC_FP32 ? 1 : 0;
  const int es = bf ? 2 : 4;
  IgemmParams q = ig_zero();
  q.mode = IG_PATCH;
  q.B = B; q.H = H; q.W = W; q.Cin = C;
  q.patch = patch; q.Hq = Hpitch / patch; q.Hp = H / patch; q.Wp = W / patch;
  q.N = D; q.K = patch * patch * C;
  q.shift = pb; q.residual = pos; q.ldr = D; q.res_mod = q.Hp * q.Wp;
  q.out = tokens; q.ldc = D; q.out_f32 = 1;
  q.f16 = p->cfg.precision == HVIT_PREC_FP16;
  const long long M = static_cast<long long>(B) * q.Hp * q.Wp;
  if (bf) {
    pick_tile(q.Hp, q.Wp, &q.Wt, &q.Ht);
    q.tiles_w = (q.Wp + q.Wt - 1) / q.Wt;
    q.tiles_h = (q.Hp + q.Ht - 1) / q.Ht;
    IgemmMaps mp;
    int bn = pick_block_n(D);
    if (const char* e = getenv("HVIT_PATCH_BN")) {  // experiment: narrower tiles -> finer wave quantisation
      const int v = atoi(e);
      if ((v == 64 || v == 128 || v == 256) && D % v == 0) bn = v;
    }
    int r = tmap_patch(&mp.a, in, q.f16, B, q.Hq, W, C, patch, q.Wp, q.Wt, q.Ht);
    if (r) return r;
    r = tmap_matrix(&mp.b, pw, q.f16, D, q.K, q.K, bn / 2);
    if (r) return r;
    r = make_out_maps(q, &mp);
    if (r) return r;
    const int sms = num_sms();
    p->steps.push_back([=](const Ctx& k) { return launch_igemm_tc2(q, mp, bn, sms, k.stream); });
  } else {
    const float* inf = reinterpret_cast<const float*>(in);
    const float* wf = reinterpret_cast<const float*>(pw);
    p->steps.push_back([=](const Ctx& k) { return launch_igemm_f32(q, inf, C, wf, k.stream); });
  }
  const double fl = 2.0 * M * D * q.K;
  p->tag("patch_embed", bf ? "igemm_tc" : "igemm_f32", fl, fl,
         (static_cast<double>(M) * q.K + static_cast<double>(D) * q.K) * es + 2.0 * M * D * 4);
  return HVIT_OK;
}

static int build_steps(hvit_plan* p) {
  const hvit_model_cfg& c = p->cfg;
  const hvit_weights& w = p->w;
  const Geometry& g = p->g;
  const int bf = c.precision != HVIT_PREC_FP32 ? 1 : 0;  // 16-bit tensor-core path
  const int f16 = c.precision == HVIT_PREC_FP16 ? 1 : 0;
  const int dt = c.precision == HVIT_PREC_FP32 ? DT_F32 : (f16 ? DT_F16 : DT_BF16);
  const int B = g.B, D = c.embed_dim;
  char nm[32], nm2[32];
  int r;

  // 1. stem (encoder block 0): Conv3x3(1->C0)+BN+ReLU(+pool), CUDA cores, reads the fp32 spectrogram directly
  {
    void* out = at<void>(p, "enc0");
    const int C0 = c.enc_channels[0], pool = c.enc_pool[0], F = g.F, T = g.T;
    const float *sw = w.stem_w, *ss = w.stem_scale, *sh = w.stem_shift;
    const bool stem_tc = bf && C0 == 64 && pool == 2 && g.enc[0].pitch == g.enc[0].H && getenv("HVIT_STEM_SIMT") == nullptr;
    if (stem_tc) {
      // tensor-core stem: the position matrices are derived from the weights once, here
      void* apack = at<void>(p, "stem_a");
      // (enqueued on the caller's setup stream, after the packing kernels that produced sw / ss on that stream; the plan
      // may be used on that stream right away - see hvit_plan_create in hvit.h)
      r = launch_stem_pack(sw, ss, apack, f16, p->setup_stream);
      if (r) return r;
      const int sms = num_sms();
      const int Ho = g.enc[0].H, Wo = g.enc[0].W;
      CUtensorMap tmo;  // (64 channels, Wo, Ho, B): one pooled row of a tile = 64 pixels x 128 bytes, 128B-swizzled
      r = tmap_out4(&tmo, out, tm16(f16), 64, Wo, Ho, B, 128, static_cast<uint64_t>(Wo) * 128, static_cast<uint64_t>(Ho) * Wo * 128, 64, 1);
      if (r) return r;
      p->steps.push_back([=](const Ctx& k) { return launch_stem_tc(k.x, k.mag_max, apack, sh, tmo, f16, B, F, T, sms, k.stream); });
    } else {
      p->steps.push_back([=](const Ctx& k) { return launch_stem(k.x, k.mag_max, sw, ss, sh, out, dt, B, F, T, C0, pool, k.stream); });
    }
    {
      const double fl = 2.0 * B * F * T * C0 * 9.0;
      p->tag("encoder.0", stem_tc ? "stem_tc" : "stem", fl, fl, static_cast<double>(B) * F * T * 4 + static_cast<double>(B) * g.enc[0].H * g.enc[0].W * C0 * g.es);
    }
    add_mask(p, "encoder.0.mask", out, g.enc[0].H, g.enc[0].pitch, g.enc[0].W, C0 * g.es, GEO_ENC + 0);
  }
  // 2. encoder blocks 1.. : implicit-GEMM 3x3 conv + folded BN + ReLU (+ fused 2x2 max-pool)
  float* conv_tmp = g.bufs.count("conv_tmp") ? at<float>(p, "conv_tmp") : nullptr;
  for (int i = 1; i < c.n_enc; ++i) {
    snprintf(nm, sizeof(nm), "enc%d", i - 1);
    snprintf(nm2, sizeof(nm2), "enc%d", i);
    const EncGeo& s = g.enc[i - 1];
    const EncGeo& d = g.enc[i];
    r = add_conv(p, std::string("encoder.") + std::to_string(i), at<void>(p, nm), B, s.H, s.W, s.C, w.enc_w[i], w.enc_scale[i], w.enc_shift[i], 1,
                 c.enc_pool[i] == 2, 0, at<void>(p, nm2), d.C, d.pitch, d.C, conv_tmp);
    if (r) return r;
    add_mask(p, std::string("encoder.") + std::to_string(i) + ".mask", at<void>(p, nm2), d.H, d.pitch, d.W, d.C * g.es, GEO_ENC + i);
  }
  // 3. patch embedding (+bias +positional embedding) -> fp32 residual stream
  {
    const EncGeo& e = g.enc[c.n_enc - 1];
    snprintf(nm, sizeof(nm), "enc%d", c.n_enc - 1);
    r = add_patch_embed(p, at<void>(p, nm), B, e.H, e.pitch, e.W, e.C, w.patch_w, w.patch_b, w.pos_embed, c.patch_size, D,
                        at<float>(p, "tokens"));
    if (r) return r;
    if (g.n_samples > 0) {
      // variable-length batch: the GEMM writes the bare patch grid (bias, no positional rows), then every clip's valid
      // columns are compacted into ITS token order n = h' * Wp_b + w' and get ITS positional rows pos[n]
      // (components.py:282-307,384: the reference flattens and indexes per clip)
      hvit_plan tmp;
      tmp.cfg = p->cfg;
      float* grid = at<float>(p, "tokgrid");
      r = add_patch_embed(&tmp, at<void>(p, nm), B, e.H, e.pitch, e.W, e.C, w.patch_w, w.patch_b, nullptr, c.patch_size, D, grid);
      if (r) return r;
      const Step fixed = p->steps.back(), var = tmp.steps[0];
      float* tokens = at<float>(p, "tokens");
      const float* pos = w.pos_embed;
      const int Hp = g.Hp, Wp = g.Wp;
      p->steps.back() = [=](const Ctx& k) {
        if (k.geo == nullptr) return fixed(k);
        const int e2 = var(k);
        if (e2) return e2;
        return launch_tokens_compact(grid, pos, tokens, B, Hp, Wp, D, k.geo, k.stream);
      };
    }
  }
  // 4. transformer blocks (pre-norm), residual stream fp32
  float* tok = at<float>(p, "tokens");
  if (bf && c.num_layers > 0) {
    p->l2_pin_base = tok;
    p->l2_pin_bytes = l2_pin_window(static_cast<size_t>(g.M) * D * sizeof(float), &p->l2_pin_ratio);
    if (p->l2_pin_bytes > 0 && p->l2_pin_ratio > 0.f) {
      p->device = current_device();
      ++g_dev[p->device].pinned_plans;
    } else {
      p->l2_pin_bytes = 0;
    }
  }
  void* ln = at<void>(p, "ln");
  void* qkv = at<void>(p, "qkv");
  void* att = at<void>(p, "attn");
  void* mlp = at<void>(p, "mlp");
  const int M = g.M, Np = g.Np, heads = c.num_heads;
  const float scale = 0.125f;  // head_dim^-0.5 with head_dim = 64
  const float eps = c.ln_eps;
  CUtensorMap tq, to;
  if (bf && c.num_layers > 0) {
    r = tmap_qkv(&tq, qkv, f16, B, Np, D);
    if (r) return r;
    r = tmap_attn_out(&to, att, f16, B, Np, D);
    if (r) return r;
  }
  // LayerNorm folded into the GEMMs on either side of it (16-bit modes; kernels.h IgemmParams::ln_*):
  // proj / fc2 (the residual updates) also write the 16-bit copy of the new residual rows and their per-slot statistics,
  // qkv / fc1 / to_feature_map read that copy with gamma-scaled weights and finish the normalisation in their
  // epilogues.  Only the first block needs a stand-alone pass (rowstats) over the patch embedding's output.
  const bool fold = bf && ln_fold_enabled(c.precision, D, c.num_layers);
  float* lnstats = fold ? at<float>(p, "lnstats") : nullptr;
  const int slots = D / 128;
  uint8_t* fold_w = fold ? at<uint8_t>(p, "lnfold_w") : nullptr;
  float* fold_v = fold ? at<float>(p, "lnfold_v") : nullptr;
  // gamma / beta folded copies of a consumer's weights, derived once on the setup stream (like the stem's matrices)
  auto fold_weights = [&](const void* W, const float* gamma, const float* beta, const float* bias, int N, const void** Wp,
                          LnFold* lf) -> int {
    float* cv = fold_v;
    const int e = launch_ln_fold(W, gamma, beta, bias, fold_w, cv, dt, N, D, p->setup_stream);
    *Wp = fold_w;
    lf->stats_in = lnstats; lf->c = cv; lf->slots = slots; lf->eps = eps;
    lf->x16_out = nullptr; lf->ld16 = 0; lf->stats_out = nullptr;
    fold_w += static_cast<size_t>(N) * D * 2;
    fold_v += N;
    return e;
  };
  LnFold prod;  // producer side of proj / fc2
  prod.x16_out = ln; prod.ld16 = D; prod.stats_out = lnstats;
  for (int l = 0; l < c.num_layers; ++l) {
    const float *g1 = w.ln1_g[l], *b1 = w.ln1_b[l], *g2 = w.ln2_g[l], *b2 = w.ln2_b[l];
    const std::string L = "blocks." + std::to_string(l);
    const double ln_bytes = static_cast<double>(M) * D * (4 + g.es);
    if (fold) {
      if (l == 0) {
        p->steps.push_back([=](const Ctx& k) { return launch_rowstats(tok, ln, lnstats, dt, M, D, slots, k.stream); });
        p->tag(L + ".norm1", "rowstats", 0, 0, ln_bytes);
      }
      const void* Wq;
      LnFold cq;
      r = fold_weights(w.qkv_w[l], g1, b1, w.qkv_b[l], 3 * D, &Wq, &cq);
      if (r) return r;
      r = add_linear(p, L + ".qkv", ln, D, Wq, cq.c, ACT_NONE, nullptr, 0, 0, qkv, 3 * D, 0, M, 3 * D, D, -1.0, &cq);
      if (r) return r;
    } else {
      p->steps.push_back([=](const Ctx& k) { return launch_layernorm(tok, g1, b1, ln, dt, M, D, eps, k.stream); });
      p->tag(L + ".norm1", "layernorm", 0, 0, ln_bytes);
      r = add_linear(p, L + ".qkv", ln, D, w.qkv_w[l], w.qkv_b[l], ACT_NONE, nullptr, 0, 0, qkv, 3 * D, !bf, M, 3 * D, D);
      if (r) return r;
    }
    const size_t probs_off = static_cast<size_t>(l) * B * heads * Np * Np;
    if (bf) {
      p->steps.push_back([=](const Ctx& k) {
        // return_attentions=True: the tensor-core kernel writes the maps itself up to 1 280 tokens; beyond that the
        // CUDA-core kernel materialises them next to it
        float* pr = k.probs != nullptr ? k.probs + probs_off : nullptr;
        if (pr != nullptr && Np > 1280) {
          const int e = launch_attn_probs_16(qkv, f16, pr, B, Np, heads, D, scale, k.stream);
          if (e) return e;
          pr = nullptr;
        }
        return launch_attn_tc(tq, to, f16, B, Np, heads, D, scale, k.stream, nullptr, k.geo, pr);
      });
      p->tag(L + ".attn", "attn_tc", 4.0 * B * heads * Np * Np * 64.0, 4.0 * B * heads * Np * Np * 64.0,
             static_cast<double>(M) * 4 * D * g.es);
    } else {
      p->steps.push_back([=](const Ctx& k) {
        return launch_attn_f32(reinterpret_cast<const float*>(qkv), reinterpret_cast<float*>(att),
                               k.probs != nullptr ? k.probs + probs_off : nullptr, B, Np, heads, D, scale, k.stream, k.geo);
      });
      p->tag(L + ".attn", "attn_f32", 4.0 * B * heads * Np * Np * 64.0, 4.0 * B * heads * Np * Np * 64.0,
             static_cast<double>(M) * 4 * D * g.es);
    }
    r = add_linear(p, L + ".proj", att, D, w.proj_w[l], w.proj_b[l], ACT_NONE, tok, D, 0, tok, D, 1, M, D, D, -1.0,
                   fold ? &prod : nullptr);
    if (r) return r;
    const void* W1 = w.fc1_w[l];
    const float* c1 = w.fc1_b[l];
    LnFold cf;
    if (fold) {
      r = fold_weights(w.fc1_w[l], g2, b2, w.fc1_b[l], c.mlp_hidden, &W1, &cf);
      if (r) return r;
      c1 = cf.c;
    } else {
      p->steps.push_back([=](const Ctx& k) { return launch_layernorm(tok, g2, b2, ln, dt, M, D, eps, k.stream); });
      p->tag(L + ".norm2", "layernorm", 0, 0, ln_bytes);
    }
    // MLP in row panels: fc1(panel) is followed directly by fc2(panel), so the 16-bit hidden activation of a panel
    // (M/P x hidden) is still in L2 when fc2 reads it instead of making the round trip through HBM (130 MB per layer at
    // 64 x 4 s).  Panels are multiples of 256 rows (one CTA-pair tile).  (Measured slower than one panel: default 1.)
    {
      const int es = g.es;
      int panels = mlp_panels(M, c.mlp_hidden, es, bf);
      const int rows_per = ((M + panels - 1) / panels + 255) / 256 * 256;
      for (int r0 = 0; r0 < M; r0 += rows_per) {
        const int rows = std::min(rows_per, M - r0);
        const uint8_t* a1 = reinterpret_cast<const uint8_t*>(ln) + static_cast<size_t>(r0) * D * es;
        uint8_t* h1 = reinterpret_cast<uint8_t*>(mlp) + static_cast<size_t>(r0) * c.mlp_hidden * es;
        float* x1 = tok + static_cast<size_t>(r0) * D;
        LnFold cfp = cf, pp = prod;
        if (fold) {
          cfp.stats_in = lnstats + static_cast<size_t>(r0) * slots * 2;
          pp.stats_out = lnstats + static_cast<size_t>(r0) * slots * 2;
          pp.x16_out = reinterpret_cast<uint8_t*>(ln) + static_cast<size_t>(r0) * D * es;
        }
        r = add_linear(p, L + ".fc1", a1, D, W1, c1, ACT_GELU, nullptr, 0, 0, h1, c.mlp_hidden, !bf, rows, c.mlp_hidden, D,
                       -1.0, fold ? &cfp : nullptr);
        if (r) return r;
        r = add_linear(p, L + ".fc2", h1, c.mlp_hidden, w.fc2_w[l], w.fc2_b[l], ACT_NONE, x1, D, 0, x1, D, 1, rows, D,
                       c.mlp_hidden, -1.0, fold ? &pp : nullptr);
        if (r) return r;
      }
    }
  }
  // 5. final LayerNorm + to_feature_map, written straight into the first decoder concat buffer (NHWC == [B,N,C])
  {
    const float *gf = w.lnf_g, *bfp = w.lnf_b;
    const CatGeo& k0 = g.cat[0];
    const void* Wt = w.tofm_w;
    const float* ct = w.tofm_b;
    LnFold ctf;
    if (fold) {
      r = fold_weights(w.tofm_w, gf, bfp, w.tofm_b, k0.Cx, &Wt, &ctf);
      if (r) return r;
      ct = ctf.c;
    } else {
      p->steps.push_back([=](const Ctx& k) { return launch_layernorm(tok, gf, bfp, ln, dt, M, D, eps, k.stream); });
      p->tag("transformer.norm", "layernorm", 0, 0, static_cast<double>(M) * D * (4 + g.es));
    }
    r = add_linear(p, "to_feature_map", ln, D, Wt, ct, ACT_NONE, nullptr, 0, 0, at<void>(p, "cat0"), k0.Ccat, !bf, M, k0.Cx, D,
                   -1.0, fold ? &ctf : nullptr);
    if (r) return r;
    if (g.n_samples > 0) {
      // variable-length batch: rows are in each clip's own token order -> scatter them back onto the [Hp, Wp] grid
      hvit_plan tmp;
      tmp.cfg = p->cfg;
      void* rows = at<void>(p, "tofm_rows");
      void* cat0 = at<void>(p, "cat0");
      r = add_linear(&tmp, "to_feature_map", ln, D, Wt, ct, ACT_NONE, nullptr, 0, 0, rows, k0.Cx, !bf, M, k0.Cx, D, -1.0,
                     fold ? &ctf : nullptr);
      if (r) return r;
      const Step fixed = p->steps.back(), var = tmp.steps[0];
      const int Hp = g.Hp, Wp = g.Wp, Cx = k0.Cx, Ccat = k0.Ccat;
      p->steps.back() = [=](const Ctx& k) {
        if (k.geo == nullptr) return fixed(k);
        const int e2 = var(k);
        if (e2) return e2;
        return launch_tofm_expand(rows, dt, cat0, B, Hp, Wp, Cx, Ccat, k.geo, k.stream);
      };
    }
  }
  // 6. decoder blocks with skip connections
  for (int i = 0; i + 1 < c.n_dec; ++i) {
    const CatGeo& k = g.cat[i];
    const CatGeo& kn = g.cat[i + 1];
    snprintf(nm, sizeof(nm), "cat%d", i);
    snprintf(nm2, sizeof(nm2), "cat%d", i + 1);
    uint8_t* cat = at<uint8_t>(p, nm);
    if (k.Ccat > k.Cx) {
      const int ei = c.n_enc - 1 - i;
      const EncGeo& e = g.enc[ei];
      char en[32];
      snprintf(en, sizeof(en), "enc%d", ei);
      const void* src = at<void>(p, en);
      void* samp = at<void>(p, "samp");
      const int Hs = e.H, Hpit = e.pitch, Ws = e.W, Cs = e.C, Hd = k.H, Wd = k.W;
      p->steps.push_back([=](const Ctx& x) {
        return launch_skip_sample(src, dt, B, Hs, Hpit, Ws, Cs, Hd, Wd, samp, x.stream, x.geo, GEO_ENC + ei, GEO_CAT + i);
      });
      p->tag("skip." + std::to_string(i) + ".sample", "skip_sample", 0, 0, 5.0 * B * Hd * Wd * Cs * g.es);
      // reference graph: 1x1 conv on the full-resolution skip feature, then bilinear resize (hybrid_vit.py:377-386)
      r = add_linear(p, "skip." + std::to_string(i) + ".proj", samp, Cs, w.skip_w[i], w.skip_b[i], ACT_NONE, nullptr, 0,
                     0, cat + static_cast<size_t>(k.Cx) * g.es, k.Ccat, !bf, B * Hd * Wd, c.dec_channels[i], Cs,
                     2.0 * B * Hs * Ws * c.dec_channels[i] * Cs);
      if (r) return r;
    }
    // (variable-length batch: both halves of the concat buffer are zero beyond the clip's own width before the conv reads it)
    add_mask(p, std::string("decoder.") + std::to_string(i) + ".mask", cat, k.H, k.H, k.W, k.Ccat * g.es, GEO_CAT + i);
    r = add_conv(p, std::string("decoder.") + std::to_string(i), cat, B, k.H, k.W, k.Ccat, w.dec_w[i], w.dec_scale[i], w.dec_shift[i], 1, 0, c.dec_up[i] == 2,
                 at<void>(p, nm2), kn.Ccat, kn.H, c.dec_channels[i], nullptr);
    if (r) return r;
  }
  // 7. head conv + tanh, bilinear resize back to [F, T]
  {
    const CatGeo& k = g.cat[c.n_dec - 1];
    snprintf(nm, sizeof(nm), "cat%d", c.n_dec - 1);
    const void* in = at<void>(p, nm);
    float* logits = at<float>(p, "logits");
    float* th = at<float>(p, "tanh");
    const float* hw = w.head_w;
    const int H = k.H, W = k.W, C = k.Ccat, F = g.F, T = g.T;
    add_mask(p, "decoder." + std::to_string(c.n_dec - 1) + ".mask", at<void>(p, nm), H, H, W, C * g.es, GEO_CAT + c.n_dec - 1);
    p->steps.push_back([=](const Ctx& x) { return launch_head(in, dt, hw, B, H, W, C, p->debug ? logits : nullptr, th, x.stream); });
    p->tag("decoder." + std::to_string(c.n_dec - 1), "head", 2.0 * B * H * W * C * 9.0, 2.0 * B * H * W * C * 9.0,
           static_cast<double>(B) * H * W * (C * g.es + 4));
    p->steps.push_back([=](const Ctx& x) { return x.fused_resize ? 0 : launch_resize(th, B, H, W, x.y, F, T, x.stream); });
    p->tag("resize", "resize", 0, 0, static_cast<double>(B) * (H * W + F * T) * 4);
  }
  // enhance-only stages around the model
  if (g.n_samples > 0) {
    const int n = g.n_samples, T = g.T;
    float* max_val = at<float>(p, "max_val");
    unsigned* mag_max = at<unsigned>(p, "mag_max");
    float* mag = at<float>(p, "mag");
    float* mo = at<float>(p, "model_out");
    const double ft = static_cast<double>(B) * 257 * T;
    p->pre.push_back([=](const Ctx& x) { return launch_peak(x.wave_in, B, n, max_val, x.normalize, x.stream); });
    p->pre_meta.push_back(StepMeta{"peak_norm", "peak", 0, 0, static_cast<double>(B) * n * 4, 2});
    // the complex spectrogram is never stored: the back end recomputes the noisy phase from the waveform
    p->pre.push_back([=](const Ctx& x) { return launch_stft(x.wave_in, B, n, T, max_val, nullptr, mag, mag_max, x.stream, x.geo); });
    p->pre_meta.push_back(StepMeta{"stft", "stft", 0, 0, static_cast<double>(B) * n * 4 + ft * 4, 2});
    const CatGeo& hl = g.cat[c.n_dec - 1];
    const float* th = at<float>(p, "tanh");
    const int Hs = hl.H, Ws = hl.W;
    const int geo_ws = GEO_CAT + c.n_dec - 1;
    p->post.push_back([=](const Ctx& x) {
      return launch_enhance_istft(x.wave_in, max_val, mag_max, th, Hs, Ws, p->debug ? mo : nullptr, x.wave_out, B, n, T, x.stream,
                                  x.geo, geo_ws);
    });
    p->post_meta.push_back(StepMeta{"istft", "enhance_istft", 0, 0,
                                    2.0 * B * n * 4 + static_cast<double>(B) * Hs * Ws * 4, 1});
  }
  while (p->meta.size() < p->steps.size()) p->meta.push_back(StepMeta{"op", "op", 0, 0, 0, 1});
  p->launches_forward = 0;
  for (const StepMeta& m : p->meta) p->launches_forward += m.launches;
  // skip path on a side stream (see hvit_plan): fork before the patch embedding, join before the first step after the
  // first skip step (= the first decoder block).  HVIT_SIDE_STREAM=0 keeps everything in line.
  {
    const char* e = getenv("HVIT_SIDE_STREAM");
    const bool on = !(e != nullptr && e[0] == '0');
    p->is_side.assign(p->steps.size(), 0);
    int first_skip = -1, fork = -1;
    for (size_t i = 0; i < p->meta.size(); ++i) {
      if (p->meta[i].name.compare(0, 5, "skip.") == 0) {
        p->is_side[i] = 1;
        if (first_skip < 0) first_skip = static_cast<int>(i);
      }
      if (p->meta[i].name == "patch_embed" && fork < 0) fork = static_cast<int>(i);
    }
    if (on && first_skip >= 0 && fork >= 0 && fork < first_skip) {
      int join = first_skip;
      while (join < static_cast<int>(p->steps.size()) && p->is_side[join]) ++join;
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      if (join < static_cast<int>(p->steps.size()) &&
          cudaStreamCreateWithPriority(&p->side, cudaStreamNonBlocking, hi) == cudaSuccess &&
          cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
          cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming) == cudaSuccess) {
        p->side_fork = fork;
        p->side_join = join;
      } else {
        cudaGetLastError();
        p->side_fork = -1;
      }
    }
  }
  return HVIT_OK;
}

}  // namespace hvit

// ============================================================================================== C ABI
using namespace hvit;

extern "C" {

const char* hvit_last_error(void) { return g_err; }
int hvit_version(void) { return 100; }

int hvit_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    set_error("no usable CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    return HVIT_E_LAUNCH;
  }
  return major == 10 ? 1 : 0;
}

size_t hvit_workspace_bytes(const hvit_model_cfg* cfg, int B, int F, int T, int n_samples) {
  if (cfg == nullptr) {
    set_error("null cfg");
    return 0;
  }
  Geometry g;
  if (build_geometry(*cfg, B, F, T, n_samples, 0, g) != HVIT_OK) return 0;
  return g.total;
}

int hvit_plan_create(const hvit_model_cfg* cfg, const hvit_weights* weights, int B, int F, int T, int n_samples,
                     void* workspace_dev, size_t workspace_bytes, void* stream, hvit_plan** plan_out) {
  if (cfg == nullptr || weights == nullptr || workspace_dev == nullptr || plan_out == nullptr) {
    set_error("hvit_plan_create: null argument");
    return HVIT_E_ARG;
  }
  const int ok = hvit_device_ok();
  if (ok < 0) return ok;
  if (ok == 0) {
    set_error("this library only runs on sm_100 (B200); there is no fallback path");
    return HVIT_E_ARCH;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  hvit_plan* p = new hvit_plan();
  p->cfg = *cfg;
  p->w = *weights;
  p->ws = reinterpret_cast<uint8_t*>(workspace_dev);
  p->setup_stream = st;
  int r = build_geometry(*cfg, B, F, T, n_samples, weights->pos_len, p->g);
  if (r == HVIT_OK && (p->g.total > workspace_bytes || (reinterpret_cast<uintptr_t>(workspace_dev) & 1023) != 0)) {
    set_error("workspace too small or not 1024-byte aligned: need %zu bytes, got %zu", p->g.total, workspace_bytes);
    r = HVIT_E_ALLOC;
  }
  if (r == HVIT_OK && n_samples > 0) r = ensure_fft_tables(st);
  if (r == HVIT_OK) r = build_steps(p);
  if (r != HVIT_OK) {
    hvit_plan_destroy(p);
    return r;
  }
  *plan_out = p;
  return HVIT_OK;
}

void hvit_plan_destroy(hvit_plan* plan) {
  if (plan != nullptr && plan->device >= 0) {
    // give the residual stream's persisting lines back only when no other live plan on that device pins anything
    // (cudaCtxResetPersistingL2Cache drops every persisting line of the context), and restore the carve-out limit
    DeviceState& d = g_dev[plan->device];
    if (--d.pinned_plans <= 0) {
      d.pinned_plans = 0;
      int cur = 0;
      if (cudaGetDevice(&cur) == cudaSuccess) {
        if (cur != plan->device) cudaSetDevice(plan->device);
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, d.carve_before);
        d.carve = 0;
        if (cur != plan->device) cudaSetDevice(cur);
      }
      cudaGetLastError();
    }
  }
  if (plan != nullptr) {
    if (plan->ev_fork != nullptr) cudaEventDestroy(plan->ev_fork);
    if (plan->ev_join != nullptr) cudaEventDestroy(plan->ev_join);
    if (plan->side != nullptr) cudaStreamDestroy(plan->side);
  }
  delete plan;
}

int hvit_plan_set_debug(hvit_plan* plan, int on) {
  if (plan == nullptr) return HVIT_E_ARG;
  plan->debug = on ? 1 : 0;
  return HVIT_OK;
}

static Ctx enhance_ctx(hvit_plan* plan, const float* wave_in, float* wave_out, int normalize, void* stream) {
  Ctx c;
  memset(&c, 0, sizeof(c));
  c.x = at<float>(plan, "mag");
  c.y = at<float>(plan, "model_out");
  c.mag_max = at<unsigned>(plan, "mag_max");
  c.wave_in = wave_in; c.wave_out = wave_out; c.normalize = normalize;
  c.fused_resize = 1;
  c.stream = reinterpret_cast<cudaStream_t>(stream);
  return c;
}

static int run_steps(hvit_plan* p, const Ctx& c) {
  // the fp32 residual stream ("tokens") stays in L2 across the transformer blocks when the device allows it
  if (p->l2_pin_bytes > 0 && p->l2_pin_ratio > 0.f) l2_window_set(p->l2_pin_base, p->l2_pin_bytes, p->l2_pin_ratio);
  int r = HVIT_OK;
  if (p->side != nullptr && p->side_fork >= 0) {
    Ctx cs = c;
    cs.stream = p->side;
    for (size_t i = 0; i < p->steps.size() && r == HVIT_OK; ++i) {
      if (static_cast<int>(i) == p->side_fork) {
        if (cudaEventRecord(p->ev_fork, c.stream) != cudaSuccess || cudaStreamWaitEvent(p->side, p->ev_fork, 0) != cudaSuccess) {
          set_error("side stream fork failed: %s", cudaGetErrorString(cudaGetLastError()));
          r = HVIT_E_LAUNCH;
          break;
        }
        for (size_t j = 0; j < p->steps.size() && r == HVIT_OK; ++j)
          if (p->is_side[j]) r = p->steps[j](cs);
        if (r == HVIT_OK && cudaEventRecord(p->ev_join, p->side) != cudaSuccess) r = HVIT_E_LAUNCH;
        if (r != HVIT_OK) break;
      }
      if (static_cast<int>(i) == p->side_join && cudaStreamWaitEvent(c.stream, p->ev_join, 0) != cudaSuccess) {
        set_error("side stream join failed: %s", cudaGetErrorString(cudaGetLastError()));
        r = HVIT_E_LAUNCH;
        break;
      }
      if (!p->is_side[i]) r = p->steps[i](c);
    }
  } else {
    for (size_t i = 0; i < p->steps.size() && r == HVIT_OK; ++i) r = p->steps[i](c);
  }
  l2_window_set(nullptr, 0, 0.f);
  return r;
}

int hvit_forward(hvit_plan* plan, const float* x_dev, float* y_dev, float* attn_probs_dev, void* stream) {
  if (plan == nullptr || x_dev == nullptr || y_dev == nullptr) {
    set_error("hvit_forward: null argument");
    return HVIT_E_ARG;
  }
  Ctx c;
  memset(&c, 0, sizeof(c));
  c.x = x_dev; c.y = y_dev; c.probs = attn_probs_dev; c.mag_max = nullptr;
  c.stream = reinterpret_cast<cudaStream_t>(stream);
  return run_steps(plan, c);
}

int hvit_enhance(hvit_plan* plan, const float* wave_in_dev, float* wave_out_dev, int normalize, void* stream) {
  if (plan == nullptr || wave_in_dev == nullptr || wave_out_dev == nullptr) {
    set_error("hvit_enhance: null argument");
    return HVIT_E_ARG;
  }
  if (plan->g.n_samples <= 0) {
    set_error("hvit_enhance: plan was created without n_samples");
    return HVIT_E_ARG;
  }
  const Ctx c = enhance_ctx(plan, wave_in_dev, wave_out_dev, normalize, stream);
  for (const Step& st : plan->pre) {
    const int r = st(c);
    if (r) return r;
  }
  int r = run_steps(plan, c);
  if (r) return r;
  for (const Step& st : plan->post) {
    r = st(c);
    if (r) return r;
  }
  return HVIT_OK;
}

// smallest clip (samples) that still yields one patch column after the encoder's pooling
static int varlen_min_samples(const hvit_model_cfg& c) {
  int down = c.patch_size;
  for (int i = 0; i < c.n_enc; ++i) down *= c.enc_pool[i];
  return (down - 1) * 128;   // T = 1 + n / 128 >= down
}

int hvit_varlen_min_samples(const hvit_plan* plan) {
  if (plan == nullptr) return HVIT_E_ARG;
  return varlen_min_samples(plan->cfg);
}

int hvit_enhance_varlen(hvit_plan* plan, const float* wave_in_dev, float* wave_out_dev, const int* n_valid_dev,
                        int normalize, void* stream) {
  if (plan == nullptr || wave_in_dev == nullptr || wave_out_dev == nullptr || n_valid_dev == nullptr) {
    set_error("hvit_enhance_varlen: null argument");
    return HVIT_E_ARG;
  }
  if (plan->g.n_samples <= 0) {
    set_error("hvit_enhance_varlen: plan was created without n_samples");
    return HVIT_E_ARG;
  }
  const hvit_model_cfg& cf = plan->cfg;
  if (cf.n_enc > 8 || cf.n_dec > 8) {
    set_error("hvit_enhance_varlen: at most 8 encoder / decoder blocks");
    return HVIT_E_SHAPE;
  }
  Ctx c = enhance_ctx(plan, wave_in_dev, wave_out_dev, normalize, stream);
  VarlenCfg vc;
  memset(&vc, 0, sizeof(vc));
  vc.n_enc = cf.n_enc; vc.n_dec = cf.n_dec; vc.patch = cf.patch_size; vc.Hp = plan->g.Hp;
  for (int i = 0; i < cf.n_enc; ++i) vc.enc_pool[i] = cf.enc_pool[i];
  for (int i = 0; i < cf.n_dec; ++i) vc.dec_up[i] = cf.dec_up[i];
  vc.n_min = varlen_min_samples(cf);
  vc.n_max = plan->g.n_samples;
  int* geo = at<int>(plan, "geo");
  int r = launch_varlen_geometry(n_valid_dev, plan->g.B, vc, geo, c.stream);
  if (r) return r;
  c.geo = geo;
  for (const Step& st : plan->pre) {
    r = st(c);
    if (r) return r;
  }
  r = run_steps(plan, c);
  if (r) return r;
  for (const Step& st : plan->post) {
    r = st(c);
    if (r) return r;
  }
  return HVIT_OK;
}

int hvit_plan_num_steps(const hvit_plan* plan, int enhance) {
  if (plan == nullptr) return HVIT_E_ARG;
  return static_cast<int>(plan->steps.size() + (enhance ? plan->pre.size() + plan->post.size() : 0));
}

static const StepMeta* step_meta(const hvit_plan* plan, int enhance, int i) {
  const int npre = enhance ? static_cast<int>(plan->pre.size()) : 0;
  const int nmid = static_cast<int>(plan->steps.size());
  if (i < 0) return nullptr;
  if (i < npre) return &plan->pre_meta[i];
  if (i < npre + nmid) return &plan->meta[i - npre];
  if (enhance && i < npre + nmid + static_cast<int>(plan->post.size())) return &plan->post_meta[i - npre - nmid];
  return nullptr;
}

int hvit_plan_step_info(const hvit_plan* plan, int enhance, int i, char* name, int name_len, char* kernel,
                        int kernel_len, double* algo_flops, double* exec_flops, double* algo_bytes, int* launches) {
  if (plan == nullptr) return HVIT_E_ARG;
  const StepMeta* m = step_meta(plan, enhance, i);
  if (m == nullptr) {
    set_error("step index %d out of range", i);
    return HVIT_E_ARG;
  }
  if (name != nullptr && name_len > 0) snprintf(name, name_len, "%s", m->name.c_str());
  if (kernel != nullptr && kernel_len > 0) snprintf(kernel, kernel_len, "%s", m->kernel.c_str());
  if (algo_flops) *algo_flops = m->algo_flops;
  if (exec_flops) *exec_flops = m->exec_flops;
  if (algo_bytes) *algo_bytes = m->algo_bytes;
  if (launches) *launches = m->launches;
  return HVIT_OK;
}

int hvit_enhance_profiled(hvit_plan* plan, const float* wave_in_dev, float* wave_out_dev, int normalize, void* stream,
                          float* step_ms_host, int n_steps) {
  if (plan == nullptr || wave_in_dev == nullptr || wave_out_dev == nullptr || step_ms_host == nullptr) {
    set_error("hvit_enhance_profiled: null argument");
    return HVIT_E_ARG;
  }
  if (plan->g.n_samples <= 0 || n_steps != hvit_plan_num_steps(plan, 1)) {
    set_error("hvit_enhance_profiled: plan has no enhance stages or n_steps mismatch");
    return HVIT_E_ARG;
  }
  const Ctx c = enhance_ctx(plan, wave_in_dev, wave_out_dev, normalize, stream);
  std::vector<cudaEvent_t> ev(n_steps + 1);
  for (auto& e : ev) cudaEventCreate(&e);
  int r = HVIT_OK, i = 0;
  cudaEventRecord(ev[0], c.stream);
  auto run = [&](const std::vector<Step>& v) {
    for (const Step& st : v) {
      if (r == HVIT_OK) r = st(c);
      cudaEventRecord(ev[++i], c.stream);
    }
  };
  run(plan->pre);
  if (plan->l2_pin_bytes > 0 && plan->l2_pin_ratio > 0.f) l2_window_set(plan->l2_pin_base, plan->l2_pin_bytes, plan->l2_pin_ratio);
  run(plan->steps);
  l2_window_set(nullptr, 0, 0.f);
  run(plan->post);
  if (cudaStreamSynchronize(c.stream) != cudaSuccess && r == HVIT_OK) r = check_launch("hvit_enhance_profiled");
  for (int k = 0; k < n_steps; ++k) {
    float ms = 0.f;
    if (r == HVIT_OK) cudaEventElapsedTime(&ms, ev[k], ev[k + 1]);
    step_ms_host[k] = ms;
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return r;
}

int hvit_plan_buffer(const hvit_plan* plan, const char* name, size_t* offset, int* dims, int* elem_bytes) {
  if (plan == nullptr || name == nullptr) return HVIT_E_ARG;
  auto it = plan->g.bufs.find(name);
  if (it == plan->g.bufs.end()) {
    set_error("no such buffer: %s", name);
    return HVIT_E_ARG;
  }
  if (offset) *offset = it->second.off;
  if (dims) memcpy(dims, it->second.dims, sizeof(int) * 4);
  if (elem_bytes) *elem_bytes = it->second.es;
  return it->second.rank;
}

int hvit_plan_launch_count(const hvit_plan* plan, int enhance) {
  if (plan == nullptr) return HVIT_E_ARG;
  int n = plan->launches_forward;
  if (enhance) {
    for (const StepMeta& m : plan->pre_meta) n += m.launches;
    for (const StepMeta& m : plan->post_meta) n += m.launches;
    n -= 1;  // the stand-alone resize kernel is fused into the iSTFT frame kernel
  }
  return n;
}

int hvit_plan_tokens(const hvit_plan* plan, int* hp, int* wp) {
  if (plan == nullptr) return HVIT_E_ARG;
  if (hp) *hp = plan->g.Hp;
  if (wp) *wp = plan->g.Wp;
  return plan->g.Np;
}

// ---------------------------------------------------------------------------------------- per-kernel entry points
static int require_sm100() {
  const int ok = hvit_device_ok();
  if (ok < 0) return ok;
  if (ok == 0) {
    set_error("this library only runs on sm_100 (B200); there is no fallback path");
    return HVIT_E_ARCH;
  }
  return HVIT_OK;
}

int hvit_gemm_16(const void* a, int lda, const void* w, const float* scale, const float* shift, int act,
                 const float* residual, int ldr, void* out, int ldc, int out_f32, int M, int N, int K, int f16,
                 void* stream) {
  int r = require_sm100();
  if (r) return r;
  IgemmParams q = ig_zero();
  q.f16 = f16 ? 1 : 0;
  if (const char* e = getenv("HVIT_DBG")) q.dbg = atoi(e);
  q.mode = IG_PLAIN;
  q.M = M; q.N = N; q.K = K;
  q.scale = scale; q.shift = shift; q.act = act; q.residual = residual; q.ldr = ldr;
  q.out = out; q.ldc = ldc; q.out_f32 = out_f32;
  IgemmMaps mp;
  const int bn = pick_block_n(N);
  r = tmap_matrix(&mp.a, a, q.f16, M, K, lda, 128);
  if (r) return r;
  r = tmap_matrix(&mp.b, w, q.f16, N, K, K, bn / 2);
  if (r) return r;
  r = make_out_maps(q, &mp);
  if (r) return r;
  return launch_igemm_tc2(q, mp, bn, num_sms(), reinterpret_cast<cudaStream_t>(stream));
}

int hvit_linear_ln_producer_16(const void* a, int lda, const void* w, const float* bias, float* x, void* x16_out,
                               float* stats_out, int M, int N, int K, int f16, void* stream) {
  int r = require_sm100();
  if (r) return r;
  IgemmParams q = ig_zero();
  q.f16 = f16 ? 1 : 0;
  q.mode = IG_PLAIN;
  q.M = M; q.N = N; q.K = K;
  q.shift = bias; q.act = ACT_NONE; q.residual = x; q.ldr = N;
  q.out = x; q.ldc = N; q.out_f32 = 1;
  q.ln_x16_out = x16_out; q.ld16 = N; q.ln_stats_out = stats_out;
  IgemmMaps mp;
  const int bn = pick_block_n(N);
  r = tmap_matrix(&mp.a, a, q.f16, M, K, lda, 128);
  if (r) return r;
  r = tmap_matrix(&mp.b, w, q.f16, N, K, K, bn / 2);
  if (r) return r;
  r = make_out_maps(q, &mp);
  if (r) return r;
  return launch_igemm_tc2(q, mp, bn, num_sms(), reinterpret_cast<cudaStream_t>(stream));
}

int hvit_linear_ln_consumer_16(const void* x16, const float* stats, int slots, const void* w, const float* gamma,
                               const float* beta, const float* bias, float eps, int act, void* out, int ldc, int M, int N,
                               int K, int f16, void* w_scratch, float* gc_scratch, void* stream) {
  int r = require_sm100();
  if (r) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  r = launch_ln_fold(w, gamma, beta, bias, w_scratch, gc_scratch, f16 ? DT_F16 : DT_BF16, N, K, st);
  if (r) return r;
  IgemmParams q = ig_zero();
  q.f16 = f16 ? 1 : 0;
  q.mode = IG_PLAIN;
  q.M = M; q.N = N; q.K = K;
  q.shift = gc_scratch; q.act = act;
  q.out = out; q.ldc = ldc; q.out_f32 = 0;
  q.ln_stats_in = stats; q.ln_slots = slots; q.ln_eps = eps;
  IgemmMaps mp;
  const int bn = pick_block_n(N);
  r = tmap_matrix(&mp.a, x16, q.f16, M, K, K, 128);
  if (r) return r;
  r = tmap_matrix(&mp.b, w_scratch, q.f16, N, K, K, bn / 2);
  if (r) return r;
  r = make_out_maps(q, &mp);
  if (r) return r;
  return launch_igemm_tc2(q, mp, bn, num_sms(), st);
}

int hvit_rowstats_16(const float* x, void* x16_out, float* stats_out, int rows, int D, int slots, int f16, void* stream) {
  int r = require_sm100();
  if (r) return r;
  return launch_rowstats(x, x16_out, stats_out, f16 ? DT_F16 : DT_BF16, rows, D, slots, reinterpret_cast<cudaStream_t>(stream));
}

int hvit_gemm_f32(const float* a, int lda, const float* w, const float* scale, const float* shift, int act,
                  const float* residual, int ldr, float* out, int ldc, int M, int N, int K, void* stream) {
  int r = require_sm100();
  if (r) return r;
  IgemmParams q = ig_zero();
  q.mode = IG_PLAIN;
  q.M = M; q.N = N; q.K = K;
  q.scale = scale; q.shift = shift; q.act = act; q.residual = residual; q.ldr = ldr;
  q.out = out; q.ldc = ldc; q.out_f32 = 1;
  return launch_igemm_f32(q, a, lda, w, reinterpret_cast<cudaStream_t>(stream));
}

int hvit_conv3x3_16(const void* x, const void* w, const float* scale, const float* shift, int relu, int pool,
                    int up2, void* out, int B, int H, int W, int Cin, int Cout, int f16, void* stream) {
  int r = require_sm100();
  if (r) return r;
  if (pool && up2) {
    set_error("conv3x3: pool and up2 are exclusive");
    return HVIT_E_SHAPE;
  }
  hvit_plan tmp;
  tmp.cfg.precision = f16 ? HVIT_PREC_FP16 : HVIT_PREC_BF16;
  const int Ho = pool ? H / 2 : (up2 ? 2 * H : H);
  r = add_conv(&tmp, "conv", x, B, H, W, Cin, w, scale, shift, relu, pool, up2, out, Cout, Ho, Cout, nullptr);
  if (r) return r;
  Ctx c;
  memset(&c, 0, sizeof(c));
  c.stream = reinterpret_cast<cudaStream_t>(stream);
  return tmp.steps[0](c);
}

int hvit_conv3x3_f32(const float* x, const float* w, const float* scale, const float* shift, int relu, int up2,
                     float* out, int B, int H, int W, int Cin, int Cout, void* stream) {
  int r = require_sm100();
  if (r) return r;
  hvit_plan tmp;
  tmp.cfg.precision = HVIT_PREC_FP32;
  r = add_conv(&tmp, "conv", x, B, H, W, Cin, w, scale, shift, relu, 0, up2, out, Cout, up2 ? 2 * H : H, Cout, nullptr);
  if (r) return r;
  Ctx c;
  memset(&c, 0, sizeof(c));
  c.stream = reinterpret_cast<cudaStream_t>(stream);
  return tmp.steps[0](c);
}

int hvit_stem_16(const float* x, const void* mag_max, const float* w, const float* scale, const float* shift, void* out,
                 void* scratch, int B, int H, int W, int C, int pool, int f16, int use_tc, void* stream) {
  int r = require_sm100();
  if (r) return r;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const unsigned* mm = reinterpret_cast<const unsigned*>(mag_max);
  if (use_tc && C == 64 && pool == 2 && scratch != nullptr) {
    r = launch_stem_pack(w, scale, scratch, f16 ? 1 : 0, s);
    if (r) return r;
    const int Ho = H / 2, Wo = W / 2;
    CUtensorMap tmo;
    r = tmap_out4(&tmo, out, tm16(f16), 64, Wo, Ho, B, 128, static_cast<uint64_t>(Wo) * 128, static_cast<uint64_t>(Ho) * Wo * 128, 64, 1);
    if (r) return r;
    return launch_stem_tc(x, mm, scratch, shift, tmo, f16 ? 1 : 0, B, H, W, num_sms(), s);
  }
  return launch_stem(x, mm, w, scale, shift, out, f16 ? DT_F16 : DT_BF16, B, H, W, C, pool, s);
}

int hvit_head_16(const void* x, const float* w, float* logits, float* out_tanh, int B, int H, int W, int C, int f16,
                 void* stream) {
  int r = require_sm100();
  if (r) return r;
  return launch_head(x, f16 ? DT_F16 : DT_BF16, w, B, H, W, C, logits, out_tanh, reinterpret_cast<cudaStream_t>(stream));
}

int hvit_attention_16(const void* qkv, void* out, int B, int N, int heads, int f16, void* stream) {
  int r = require_sm100();
  if (r) return r;
  CUtensorMap tq, to;
  r = tmap_qkv(&tq, qkv, f16, B, N, heads * 64);
  if (r) return r;
  r = tmap_attn_out(&to, out, f16, B, N, heads * 64);
  if (r) return r;
  if (getenv("HVIT_PROF") != nullptr) {  // diagnostics: per-phase cycle counters of the softmax groups
    const int nc = num_sms();
    long long* d = nullptr;
    cudaMalloc(&d, sizeof(long long) * 32 * nc);
    cudaMemset(d, 0, sizeof(long long) * 32 * nc);
    r = launch_attn_tc(tq, to, f16 ? 1 : 0, B, N, heads, heads * 64, 0.125f, reinterpret_cast<cudaStream_t>(stream), d);
    cudaDeviceSynchronize();
    std::vector<long long> hst(32 * nc);
    cudaMemcpy(hst.data(), d, sizeof(long long) * 32 * nc, cudaMemcpyDeviceToHost);
    cudaFree(d);
    double sa[32] = {0};
    for (int c = 0; c < nc; ++c)
      for (int k = 0; k < 32; ++k) sa[k] += static_cast<double>(hst[c * 32 + k]) / nc;
    for (int w = 0; w < 2; ++w)
      fprintf(stderr, "[attn prof B=%d N=%d] group %c per CTA (%.1f items): wait_s %.0f ld %.0f max %.0f wait_pv/rescale %.0f exp %.0f arrive %.0f finish %.0f total %.0f cycles\n",
              B, N, 'A' + w, sa[w * 16 + 7], sa[w * 16 + 0], sa[w * 16 + 1], sa[w * 16 + 2], sa[w * 16 + 3], sa[w * 16 + 4], sa[w * 16 + 5], sa[w * 16 + 8], sa[w * 16 + 6]);
    // phase of group B relative to group A: start of the exp phase of blocks 1..6 (CTA 0 and the mean over CTAs)
    fprintf(stderr, "[attn prof] exp-phase start, B minus A, blocks 1..6: CTA0");
    for (int k = 10; k <= 15; ++k) fprintf(stderr, " %lld", hst[16 + k] - hst[k]);
    fprintf(stderr, " | A block-to-block:");
    for (int k = 11; k <= 15; ++k) fprintf(stderr, " %lld", hst[k] - hst[k - 1]);
    fprintf(stderr, "\n");
    return r;
  }
  return launch_attn_tc(tq, to, f16 ? 1 : 0, B, N, heads, heads * 64, 0.125f, reinterpret_cast<cudaStream_t>(stream));
}

int hvit_attention_f32(const float* qkv, float* out, float* probs, int B, int N, int heads, void* stream) {
  int r = require_sm100();
  if (r) return r;
  return launch_attn_f32(qkv, out, probs, B, N, heads, heads * 64, 0.125f, reinterpret_cast<cudaStream_t>(stream));
}

int hvit_layernorm(const float* x, const float* g, const float* b, void* out, int out_dtype, int rows, int D, float eps,
                   void* stream) {
  int r = require_sm100();
  if (r) return r;
  if (out_dtype < 0 || out_dtype > 2) {
    set_error("layernorm: out_dtype must be 0 (fp32), 1 (bf16) or 2 (fp16)");
    return HVIT_E_ARG;
  }
  return launch_layernorm(x, g, b, out, out_dtype, rows, D, eps, reinterpret_cast<cudaStream_t>(stream));
}

int hvit_stft(const float* wave, int B, int n, int normalize, void* max_val, void* spec, float* mag, void* mag_max,
              void* stream) {
  int r = require_sm100();
  if (r) return r;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  r = ensure_fft_tables(s);
  if (r) return r;
  r = launch_peak(wave, B, n, reinterpret_cast<float*>(max_val), normalize, s);
  if (r) return r;
  return launch_stft(wave, B, n, 1 + n / 128, reinterpret_cast<const float*>(max_val), reinterpret_cast<float2*>(spec),
                     mag, reinterpret_cast<unsigned*>(mag_max), s);
}

int hvit_istft(const float* mag_norm, const void* spec, const void* mag_max, const void* max_val, float* frames,
               float* wave_out, int B, int n, void* stream) {
  int r = require_sm100();
  if (r) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  r = ensure_fft_tables(st);
  if (r) return r;
  r = launch_istft_frames(const_cast<float*>(mag_norm), nullptr, 0, 0, reinterpret_cast<const float2*>(spec),
                          reinterpret_cast<const unsigned*>(mag_max), frames, B, 1 + n / 128, st);
  if (r) return r;
  return launch_istft_ola(frames, reinterpret_cast<const float*>(max_val), wave_out, B, n, 1 + n / 128, st);
}

int hvit_patch_embed_16(const void* x, int B, int H, int W, int C, const void* w, const float* bias, const float* pos,
                        int patch, int D, float* tokens, int f16, void* stream) {
  int r = require_sm100();
  if (r) return r;
  if (x == nullptr || w == nullptr || tokens == nullptr || patch < 1 || patch > 16 || H % patch != 0 || H / patch < 1 ||
      W / patch < 1 || C % 64 != 0 || D % 64 != 0) {
    set_error("hvit_patch_embed_16: bad argument (needs H %% p == 0, C and D multiples of 64)");
    return HVIT_E_ARG;
  }
  hvit_plan tmp;
  tmp.cfg.precision = f16 ? HVIT_PREC_FP16 : HVIT_PREC_BF16;
  r = add_patch_embed(&tmp, x, B, H, H, W, C, w, bias, pos, patch, D, tokens);
  if (r) return r;
  Ctx c;
  memset(&c, 0, sizeof(c));
  c.stream = reinterpret_cast<cudaStream_t>(stream);
  return tmp.steps[0](c);
}

int hvit_skip_concat_16(const void* src, int B, int Hs, int Ws, int Cs, const void* w, const float* bias, int Cdec,
                        void* cat, int Hd, int Wd, int Ccat, int c_off, void* scratch, int f16, void* stream) {
  int r = require_sm100();
  if (r) return r;
  if (src == nullptr || w == nullptr || cat == nullptr || scratch == nullptr || Cs % 64 != 0 || Cdec % 64 != 0 ||
      c_off % 64 != 0 || c_off + Cdec > Ccat || Ccat % 8 != 0) {
    set_error("hvit_skip_concat_16: bad argument (channel counts / offset must be multiples of 64 and fit the concat buffer)");
    return HVIT_E_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // bilinear sample first (align_corners=False), then the 1x1 projection on the low-resolution grid: exact, because the
  // interpolation weights sum to 1 and the projection is affine per pixel
  r = launch_skip_sample(src, f16 ? DT_F16 : DT_BF16, B, Hs, Hs, Ws, Cs, Hd, Wd, scratch, st);
  if (r) return r;
  hvit_plan tmp;
  tmp.cfg.precision = f16 ? HVIT_PREC_FP16 : HVIT_PREC_BF16;
  r = add_linear(&tmp, "skip", scratch, Cs, w, bias, ACT_NONE, nullptr, 0, 0,
                 reinterpret_cast<uint8_t*>(cat) + static_cast<size_t>(c_off) * 2, Ccat, 0, B * Hd * Wd, Cdec, Cs);
  if (r) return r;
  Ctx c;
  memset(&c, 0, sizeof(c));
  c.stream = st;
  return tmp.steps[0](c);
}

}  // extern "C"
