// Device-side weight packing behind hvit_pack_weights (include/hvit.h): turns the reference state_dict (fp32, torch
// layouts) into the packed buffers the launch plan reads.  Replaces the parameter half of HybridViT.__init__ /
// load_state_dict for this path (reference models/hybrid_vit.py:172-284): BatchNorm(eval) folding
// (models/components.py:62-69), K-major conv re-layout, the four pre-summed 2x2 parity kernels of a
// "nearest x2 upsample + 3x3 conv" block (models/components.py:139-167) and the conversion to the operand type.
// Runs once per weight version; plain grid-stride kernels, no torch.
#include <cstring>

#include "common.cuh"
#include "hvit.h"
#include "kernels.h"

namespace hvit {
namespace {

constexpr float BN_EPS = 1e-5f;  // nn.BatchNorm2d default used by the reference (components.py:67,162)

__device__ __forceinline__ void store_act(void* dst, size_t i, float v, int prec) {
  if (prec == HVIT_PREC_FP32) reinterpret_cast<float*>(dst)[i] = v;
  else if (prec == HVIT_PREC_FP16) reinterpret_cast<__half*>(dst)[i] = __float2half_rn(v);
  else reinterpret_cast<bf16*>(dst)[i] = __float2bfloat16_rn(v);
}

// scale = w / sqrt(var + eps), shift = b - mean * scale
__global__ void bn_fold_kernel(const float* w, const float* b, const float* mean, const float* var, float* scale,
                               float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s = w[c] / sqrtf(var[c] + BN_EPS);
  scale[c] = s;
  shift[c] = b[c] - mean[c] * s;
}

// [Cout][Cin][k][k] fp32 -> [Cout][k][k][Cin] act, optionally multiplied by scale[Cout] (in fp32, one rounding)
__global__ void conv_pack_kernel(const float* w, const float* scale, void* out, int Cout, int Cin, int k, int prec) {
  const size_t total = static_cast<size_t>(Cout) * Cin * k * k;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % Cin);
    size_t r = i / Cin;
    const int kx = static_cast<int>(r % k);
    r /= k;
    const int ky = static_cast<int>(r % k);
    const int co = static_cast<int>(r / k);
    float v = w[((static_cast<size_t>(co) * Cin + ci) * k + ky) * k + kx];
    if (scale != nullptr) v *= scale[co];
    store_act(out, i, v, prec);
  }
}

// "nearest x2 upsample + 3x3 / pad 1 conv" as four 2x2 convolutions on the low-resolution input, one per output parity
// (py, px): out[2y+py, 2x+px] = sum_{a,b} K[py*2+px][:, a, b, :] . in[y+a+py-1, x+b+px-1].  Kernel rows collapse as
//   py = 0: a = 0 <- {ky 0}, a = 1 <- {ky 1, 2};   py = 1: a = 0 <- {ky 0, 1}, a = 1 <- {ky 2}      (same for columns)
// [Cout][Cin][3][3] fp32 -> [4][Cout][2][2][Cin] act, weights pre-multiplied by scale[Cout].
__global__ void up2_pack_kernel(const float* w, const float* scale, void* out, int Cout, int Cin, int prec) {
  const size_t total = static_cast<size_t>(4) * Cout * 4 * Cin;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % Cin);
    size_t r = i / Cin;
    const int b = static_cast<int>(r & 1), a = static_cast<int>((r >> 1) & 1);
    r >>= 2;
    const int co = static_cast<int>(r % Cout);
    const int par = static_cast<int>(r / Cout);
    const int py = par >> 1, px = par & 1;
    const float* wk = w + (static_cast<size_t>(co) * Cin + ci) * 9;
    const float s = scale != nullptr ? scale[co] : 1.0f;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const bool ry = py == 0 ? (a == 0 ? ky == 0 : ky >= 1) : (a == 0 ? ky <= 1 : ky == 2);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const bool rx = px == 0 ? (b == 0 ? kx == 0 : kx >= 1) : (b == 0 ? kx <= 1 : kx == 2);
        if (ry && rx) acc += wk[ky * 3 + kx] * s;
      }
    }
    store_act(out, i, acc, prec);
  }
}

__global__ void cast_kernel(const float* src, void* dst, size_t n, int prec) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    store_act(dst, i, src[i], prec);
}

// stem / head weights: w[c][0][ky][kx] (stem, [C][1][3][3]) or w[0][c][ky][kx] (head, [1][C][3][3]) -> [3][3][C] fp32;
// both are the same index map of a [C][9] matrix: out[t * C + c] = w[c * 9 + t]
__global__ void tap_major_kernel(const float* w, float* out, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * C) return;
  const int c = i % C, t = i / C;
  out[i] = w[c * 9 + t];
}

unsigned grid_for(size_t n) {
  const size_t b = (n + 255) / 256;
  return static_cast<unsigned>(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

// bump allocator over the caller's packed buffer (sizes only when base == nullptr)
struct Arena {
  uint8_t* base;
  size_t off;
  void* take(size_t bytes) {
    void* p = base != nullptr ? base + off : nullptr;
    off = (off + bytes + 255) / 256 * 256;
    return p;
  }
};

struct Dims {
  int cin_enc[HVIT_MAX_STAGES];
  int cin_dec[HVIT_MAX_STAGES];
  int cskip_in[HVIT_MAX_STAGES];
  bool has_skip[HVIT_MAX_STAGES];
};

int derive_dims(const hvit_model_cfg& c, Dims& d) {
  if (c.n_enc < 1 || c.n_enc > HVIT_MAX_STAGES || c.n_dec < 2 || c.n_dec > HVIT_MAX_STAGES || c.num_layers < 0 ||
      c.num_layers > HVIT_MAX_LAYERS || c.embed_dim < 1 || c.mlp_hidden < 1 || c.patch_size < 1) {
    set_error("hvit_pack_weights: bad model configuration");
    return HVIT_E_SHAPE;
  }
  if (c.precision != HVIT_PREC_FP32 && c.precision != HVIT_PREC_BF16 && c.precision != HVIT_PREC_FP16) {
    set_error("unknown precision %d", c.precision);
    return HVIT_E_SHAPE;
  }
  for (int i = 0; i < c.n_enc; ++i) d.cin_enc[i] = i == 0 ? 1 : c.enc_channels[i - 1];
  for (int i = 0; i < c.n_dec; ++i) {
    const bool final_blk = i == c.n_dec - 1;
    d.has_skip[i] = c.use_skip && !final_blk && i < c.n_enc;
    const int prev = i == 0 ? c.dec_channels[0] : c.dec_channels[i - 1];
    d.cin_dec[i] = prev + (d.has_skip[i] ? c.dec_channels[i] : 0);
    d.cskip_in[i] = d.has_skip[i] ? c.enc_channels[c.n_enc - 1 - i] : 0;
  }
  return HVIT_OK;
}

// One pass over the weight set: with ref == nullptr it only measures (arena.base == nullptr), otherwise it also
// launches the packing kernels on `s` and fills `out`.
int pack_all(const hvit_model_cfg& c, const hvit_ref_weights* ref, int pos_len, Arena& ar, hvit_weights* out,
             cudaStream_t s) {
  Dims d;
  int r = derive_dims(c, d);
  if (r) return r;
  const int prec = c.precision;
  const bool lowp = prec != HVIT_PREC_FP32;
  const size_t es = lowp ? 2 : 4;
  const bool run = ref != nullptr;
  hvit_weights w;
  memset(&w, 0, sizeof(w));

  auto f32copy = [&](const float* src, size_t n) -> const float* {
    float* dst = reinterpret_cast<float*>(ar.take(n * 4));
    if (run) {
      if (src == nullptr) return nullptr;
      cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, s);
    }
    return dst;
  };
  auto cast = [&](const float* src, size_t n) -> const void* {
    void* dst = ar.take(n * es);
    if (run) {
      if (src == nullptr) return nullptr;
      cast_kernel<<<grid_for(n), 256, 0, s>>>(src, dst, n, prec);
    }
    return dst;
  };
  auto need = [&](const void* p, const char* what, int i) -> bool {
    if (run && p == nullptr) {
      set_error("hvit_pack_weights: missing reference tensor %s[%d]", what, i);
      return false;
    }
    return true;
  };

  // ---- encoder (block 0 = stem: fp32 [3][3][C0] + scale / shift; the rest implicit-GEMM weights)
  for (int i = 0; i < c.n_enc; ++i) {
    const int C = c.enc_channels[i], Cin = d.cin_enc[i];
    if (run && !(need(ref->enc_conv_w[i], "enc_conv_w", i) && need(ref->enc_bn_w[i], "enc_bn_w", i) &&
                 need(ref->enc_bn_b[i], "enc_bn_b", i) && need(ref->enc_bn_mean[i], "enc_bn_mean", i) &&
                 need(ref->enc_bn_var[i], "enc_bn_var", i)))
      return HVIT_E_ARG;
    float* scale = reinterpret_cast<float*>(ar.take(static_cast<size_t>(C) * 4));
    float* shift = reinterpret_cast<float*>(ar.take(static_cast<size_t>(C) * 4));
    if (run)
      bn_fold_kernel<<<(C + 255) / 256, 256, 0, s>>>(ref->enc_bn_w[i], ref->enc_bn_b[i], ref->enc_bn_mean[i],
                                                     ref->enc_bn_var[i], scale, shift, C);
    if (i == 0) {
      float* sw = reinterpret_cast<float*>(ar.take(static_cast<size_t>(9) * C * 4));
      if (run) tap_major_kernel<<<(9 * C + 255) / 256, 256, 0, s>>>(ref->enc_conv_w[0], sw, C);
      w.stem_w = sw; w.stem_scale = scale; w.stem_shift = shift;
    } else {
      const size_t n = static_cast<size_t>(C) * 9 * Cin;
      void* pw = ar.take(n * es);
      // 16-bit modes: the BN scale is folded into the weights in fp32 (one rounding), the epilogue only adds the shift
      if (run) conv_pack_kernel<<<grid_for(n), 256, 0, s>>>(ref->enc_conv_w[i], lowp ? scale : nullptr, pw, C, Cin, 3, prec);
      w.enc_w[i] = pw;
      w.enc_scale[i] = lowp ? nullptr : scale;
      w.enc_shift[i] = shift;
    }
  }
  // ---- patch embedding + positional table
  {
    const int C = c.enc_channels[c.n_enc - 1], D = c.embed_dim, p = c.patch_size;
    if (run && !(need(ref->patch_w, "patch_w", 0) && need(ref->patch_b, "patch_b", 0) && need(ref->pos_embed, "pos_embed", 0)))
      return HVIT_E_ARG;
    const size_t n = static_cast<size_t>(D) * p * p * C;
    void* pw = ar.take(n * es);
    if (run) conv_pack_kernel<<<grid_for(n), 256, 0, s>>>(ref->patch_w, nullptr, pw, D, C, p, prec);
    w.patch_w = pw;
    w.patch_b = f32copy(run ? ref->patch_b : nullptr, D);
    w.pos_embed = f32copy(run ? ref->pos_embed : nullptr, static_cast<size_t>(pos_len) * D);
    w.pos_len = pos_len;
  }
  // ---- transformer
  {
    const size_t D = c.embed_dim, Hd = c.mlp_hidden;
    for (int l = 0; l < c.num_layers; ++l) {
      if (run && !(need(ref->ln1_w[l], "ln1_w", l) && need(ref->ln1_b[l], "ln1_b", l) && need(ref->ln2_w[l], "ln2_w", l) &&
                   need(ref->ln2_b[l], "ln2_b", l) && need(ref->qkv_w[l], "qkv_w", l) && need(ref->qkv_b[l], "qkv_b", l) &&
                   need(ref->proj_w[l], "proj_w", l) && need(ref->proj_b[l], "proj_b", l) && need(ref->fc1_w[l], "fc1_w", l) &&
                   need(ref->fc1_b[l], "fc1_b", l) && need(ref->fc2_w[l], "fc2_w", l) && need(ref->fc2_b[l], "fc2_b", l)))
        return HVIT_E_ARG;
      w.ln1_g[l] = f32copy(run ? ref->ln1_w[l] : nullptr, D);
      w.ln1_b[l] = f32copy(run ? ref->ln1_b[l] : nullptr, D);
      w.ln2_g[l] = f32copy(run ? ref->ln2_w[l] : nullptr, D);
      w.ln2_b[l] = f32copy(run ? ref->ln2_b[l] : nullptr, D);
      w.qkv_w[l] = cast(run ? ref->qkv_w[l] : nullptr, 3 * D * D);
      w.qkv_b[l] = f32copy(run ? ref->qkv_b[l] : nullptr, 3 * D);
      w.proj_w[l] = cast(run ? ref->proj_w[l] : nullptr, D * D);
      w.proj_b[l] = f32copy(run ? ref->proj_b[l] : nullptr, D);
      w.fc1_w[l] = cast(run ? ref->fc1_w[l] : nullptr, Hd * D);
      w.fc1_b[l] = f32copy(run ? ref->fc1_b[l] : nullptr, Hd);
      w.fc2_w[l] = cast(run ? ref->fc2_w[l] : nullptr, D * Hd);
      w.fc2_b[l] = f32copy(run ? ref->fc2_b[l] : nullptr, D);
    }
    if (run && !(need(ref->lnf_w, "lnf_w", 0) && need(ref->lnf_b, "lnf_b", 0) && need(ref->tofm_w, "tofm_w", 0) &&
                 need(ref->tofm_b, "tofm_b", 0)))
      return HVIT_E_ARG;
    const size_t Cl = c.enc_channels[c.n_enc - 1];
    w.lnf_g = f32copy(run ? ref->lnf_w : nullptr, D);
    w.lnf_b = f32copy(run ? ref->lnf_b : nullptr, D);
    w.tofm_w = cast(run ? ref->tofm_w : nullptr, Cl * D);
    w.tofm_b = f32copy(run ? ref->tofm_b : nullptr, Cl);
  }
  // ---- decoder (+ skip projections); the last block is the fp32 head [3][3][C]
  for (int i = 0; i < c.n_dec; ++i) {
    const int Cin = d.cin_dec[i], C = c.dec_channels[i];
    if (run && !need(ref->dec_conv_w[i], "dec_conv_w", i)) return HVIT_E_ARG;
    if (i == c.n_dec - 1) {
      float* hw = reinterpret_cast<float*>(ar.take(static_cast<size_t>(9) * Cin * 4));
      if (run) tap_major_kernel<<<(9 * Cin + 255) / 256, 256, 0, s>>>(ref->dec_conv_w[i], hw, Cin);
      w.head_w = hw;
      continue;
    }
    if (run && !(need(ref->dec_bn_w[i], "dec_bn_w", i) && need(ref->dec_bn_b[i], "dec_bn_b", i) &&
                 need(ref->dec_bn_mean[i], "dec_bn_mean", i) && need(ref->dec_bn_var[i], "dec_bn_var", i)))
      return HVIT_E_ARG;
    float* scale = reinterpret_cast<float*>(ar.take(static_cast<size_t>(C) * 4));
    float* shift = reinterpret_cast<float*>(ar.take(static_cast<size_t>(C) * 4));
    if (run)
      bn_fold_kernel<<<(C + 255) / 256, 256, 0, s>>>(ref->dec_bn_w[i], ref->dec_bn_b[i], ref->dec_bn_mean[i],
                                                     ref->dec_bn_var[i], scale, shift, C);
    const bool up2 = c.dec_up[i] == 2;
    if (lowp && up2) {
      const size_t n = static_cast<size_t>(16) * C * Cin;
      void* pw = ar.take(n * es);
      if (run) up2_pack_kernel<<<grid_for(n), 256, 0, s>>>(ref->dec_conv_w[i], scale, pw, C, Cin, prec);
      w.dec_w[i] = pw;
    } else {
      const size_t n = static_cast<size_t>(C) * 9 * Cin;
      void* pw = ar.take(n * es);
      if (run) conv_pack_kernel<<<grid_for(n), 256, 0, s>>>(ref->dec_conv_w[i], lowp ? scale : nullptr, pw, C, Cin, 3, prec);
      w.dec_w[i] = pw;
    }
    w.dec_scale[i] = lowp ? nullptr : scale;
    w.dec_shift[i] = shift;
    if (d.has_skip[i]) {
      if (run && !(need(ref->skip_w[i], "skip_w", i) && need(ref->skip_b[i], "skip_b", i))) return HVIT_E_ARG;
      w.skip_w[i] = cast(run ? ref->skip_w[i] : nullptr, static_cast<size_t>(C) * d.cskip_in[i]);
      w.skip_b[i] = f32copy(run ? ref->skip_b[i] : nullptr, C);
    }
  }
  if (out != nullptr) *out = w;
  if (run) return check_launch("hvit_pack_weights");
  return HVIT_OK;
}

}  // namespace
}  // namespace hvit

using namespace hvit;

extern "C" {

size_t hvit_packed_weights_bytes(const hvit_model_cfg* cfg, int pos_len) {
  if (cfg == nullptr || pos_len < 1) {
    set_error("hvit_packed_weights_bytes: null cfg or bad pos_len");
    return 0;
  }
  Arena ar{nullptr, 0};
  if (pack_all(*cfg, nullptr, pos_len, ar, nullptr, nullptr) != HVIT_OK) return 0;
  return ar.off;
}

int hvit_pack_weights(const hvit_model_cfg* cfg, const hvit_ref_weights* ref, void* packed_dev, size_t packed_bytes,
                      hvit_weights* out, void* stream) {
  if (cfg == nullptr || ref == nullptr || packed_dev == nullptr || out == nullptr) {
    set_error("hvit_pack_weights: null argument");
    return HVIT_E_ARG;
  }
  if (ref->pos_len < 1) {
    set_error("hvit_pack_weights: pos_len must be >= 1");
    return HVIT_E_ARG;
  }
  const size_t need_bytes = hvit_packed_weights_bytes(cfg, ref->pos_len);
  if (need_bytes == 0) return HVIT_E_SHAPE;
  if (packed_bytes < need_bytes || (reinterpret_cast<uintptr_t>(packed_dev) & 255) != 0) {
    set_error("hvit_pack_weights: packed buffer too small or not 256-byte aligned (need %zu bytes, got %zu)", need_bytes,
              packed_bytes);
    return HVIT_E_ALLOC;
  }
  Arena ar{reinterpret_cast<uint8_t*>(packed_dev), 0};
  return pack_all(*cfg, ref, ref->pos_len, ar, out, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
