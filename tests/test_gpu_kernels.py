"""-m gpu: every CUDA kernel, called through the C ABI, against a plain PyTorch fp32 reference of the same op
(and the tcgen05 kernels additionally against the fp32 CUDA-core kernels)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_util as U
    from hvit_b200 import _lib
    from hvit_b200.models.packing import conv_khwc, up2_parity_kernels

DEV = "cuda"
DT16 = {"bf16": torch.bfloat16, "fp16": torch.float16}
OUT_TOL = {"bf16": 6e-3, "fp16": 8e-4}   # rounding of a 16-bit output (2^-8 / 2^-11) with margin


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def test_device_is_sm100():
    assert U.lib().hvit_device_ok() == 1
    assert torch.cuda.get_device_capability()[0] == 10


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (200, 128, 128), (496, 1536, 512), (1000, 512, 2048),
                                   (64 * 48, 384, 128), (37, 256, 512)])
def test_gemm_f32(M, N, K):
    a, w = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=0.05)
    bias = _rand(N, seed=3)
    out = U.gemm_f32(a, w, shift=bias)
    ref = a.double() @ w.double().t() + bias.double()
    assert U.rel_err(out, ref) < 1e-5


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 256, 64), (200, 128, 128), (496, 1536, 512),
                                   (1000, 512, 2048), (64 * 48, 384, 128), (37, 256, 512), (31744, 512, 512)])
@pytest.mark.parametrize("kind", ["bf16", "fp16"])
def test_gemm_16_tcgen05(M, N, K, kind):
    a = _rand(M, K, seed=1).to(DT16[kind])
    w = _rand(N, K, seed=2, scale=0.05).to(DT16[kind])
    out = U.gemm_16(a, w, out_f32=True)
    ref = a.double() @ w.double().t()
    assert U.rel_err(out, ref) < 2e-5          # fp32 accumulation of exact 16-bit products
    out16 = U.gemm_16(a, w, out_f32=False)
    assert out16.dtype == DT16[kind]
    assert U.rel_err(out16.float(), ref) < OUT_TOL[kind]


@pytest.mark.parametrize("kind", ["bf16", "fp16"])
def test_gemm_16_epilogues(kind):
    M, N, K = 300, 512, 256
    a = _rand(M, K, seed=4).to(DT16[kind])
    w = _rand(N, K, seed=5, scale=0.05).to(DT16[kind])
    scale, shift, res = _rand(N, seed=6), _rand(N, seed=7), _rand(M, N, seed=8)
    base = a.double() @ w.double().t()
    out = U.gemm_16(a, w, scale=scale, shift=shift, act=_lib.ACT_RELU, out_f32=True)
    assert U.rel_err(out, torch.relu(base * scale.double() + shift.double())) < 2e-5
    out = U.gemm_16(a, w, shift=shift, act=_lib.ACT_GELU, out_f32=True)
    assert U.rel_err(out, F.gelu(base + shift.double())) < 2e-5
    # residual add, in place on the fp32 residual stream (x += proj(attn))
    x = res.clone()
    U.gemm_16(a, w, shift=shift, residual=x, out_f32=True, out=x)
    assert U.rel_err(x, base + shift.double() + res.double()) < 2e-5
    # strided output: write a channel slice of a wider (concat) buffer
    wide = torch.zeros((M, N + 128), dtype=DT16[kind], device=DEV)
    U.gemm_16(a, w, shift=shift, out=wide[:, 128:], ldc=N + 128)
    U.sync()
    assert U.rel_err(wide[:, 128:].float(), base + shift.double()) < OUT_TOL[kind]
    assert float(wide[:, :128].abs().max()) == 0.0


def test_fp16_conversion_saturates():
    """fp16 outputs saturate to the largest finite value instead of overflowing to inf."""
    a = torch.full((128, 64), 200.0, dtype=torch.float16, device=DEV)
    w = torch.full((64, 64), 200.0, dtype=torch.float16, device=DEV)
    out = U.gemm_16(a, w)          # 64 * 4e4 = 2.56e6 > 65504
    U.sync()
    assert torch.isfinite(out.float()).all() and float(out.float().max()) == 65504.0


def test_gemm_16_matches_simt_fp32():
    M, N, K = 512, 256, 1152
    a = _rand(M, K, seed=9).bfloat16()
    w = _rand(N, K, seed=10, scale=0.03).bfloat16()
    tc = U.gemm_16(a, w, out_f32=True)
    simt = U.gemm_f32(a.float(), w.float())
    assert U.rel_err(tc, simt) < 2e-5


# ------------------------------------------------------------------------------------------------ conv
def _conv_ref(x_nhwc, w, scale, shift, relu, pool, up2):
    x = x_nhwc.permute(0, 3, 1, 2).double()
    if up2:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    y = F.conv2d(x, w.double(), padding=1) * scale.double()[None, :, None, None] + shift.double()[None, :, None, None]
    if relu:
        y = torch.relu(y)
    if pool:
        y = F.max_pool2d(y, 2)
    return y.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("B,H,W,Cin,Cout,pool,up2", [
    (2, 16, 31, 128, 64, 0, 0), (1, 64, 125, 128, 256, 0, 0), (2, 37, 50, 64, 128, 1, 0), (1, 128, 250, 64, 128, 1, 0),
    (2, 16, 31, 384, 128, 0, 1), (1, 32, 62, 192, 64, 0, 1), (3, 5, 7, 64, 64, 0, 0), (2, 9, 3, 64, 64, 0, 1)])
@pytest.mark.parametrize("kind", ["bf16", "fp16"])
def test_conv3x3_16_tcgen05(B, H, W, Cin, Cout, pool, up2, kind):
    dt = DT16[kind]
    x = _rand(B, H, W, Cin, seed=11).to(dt)
    w = _rand(Cout, Cin, 3, 3, seed=12, scale=(2.0 / (9 * Cin)) ** 0.5)
    scale, shift = _rand(Cout, seed=13).abs() + 0.5, _rand(Cout, seed=14, scale=0.1)
    if up2:
        wp = up2_parity_kernels(w).to(dt)
    else:
        wp = conv_khwc(w).to(dt)
    Ho, Wo = (H // 2, W // 2) if pool else ((2 * H, 2 * W) if up2 else (H, W))
    out = torch.full((B, Ho, Wo, Cout), float("nan"), dtype=dt, device=DEV)
    _lib.check(U.lib().hvit_conv3x3_16(U.P(x), U.P(wp), U.P(scale), U.P(shift), 1, pool, up2, U.P(out), B, H, W, Cin,
                                       Cout, 1 if kind == "fp16" else 0, U.stream()), "hvit_conv3x3_16")
    U.sync()
    if up2:   # reference assembled from the same bf16-rounded parity kernels
        ref = _up2_from_parity(x.float().permute(0, 3, 1, 2).double(), wp.double())
        ref = torch.relu(ref * scale.double()[None, :, None, None] + shift.double()[None, :, None, None])
        ref = ref.permute(0, 2, 3, 1)
    else:
        ref = _conv_ref(x.float(), wp.float().permute(0, 3, 1, 2), scale, shift, True, pool, up2)
    assert torch.isfinite(out.float()).all()
    assert U.rel_err(out.float(), ref) < OUT_TOL[kind]


@pytest.mark.parametrize("B,H,W,Cin,Cout,up2", [(2, 16, 31, 64, 64, 0), (1, 9, 14, 32, 48, 1)])
def test_conv3x3_f32(B, H, W, Cin, Cout, up2):
    x = _rand(B, H, W, Cin, seed=15)
    w = _rand(Cout, Cin, 3, 3, seed=16, scale=(2.0 / (9 * Cin)) ** 0.5)
    scale, shift = _rand(Cout, seed=17).abs() + 0.5, _rand(Cout, seed=18, scale=0.1)
    Ho, Wo = (2 * H, 2 * W) if up2 else (H, W)
    out = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32, device=DEV)
    _lib.check(U.lib().hvit_conv3x3_f32(U.P(x), U.P(conv_khwc(w)), U.P(scale), U.P(shift), 1, up2, U.P(out), B, H, W, Cin,
                                        Cout, U.stream()), "hvit_conv3x3_f32")
    U.sync()
    assert U.rel_err(out, _conv_ref(x, w, scale, shift, True, False, up2)) < 1e-5


def _up2_from_parity(x, k):
    """x [B,Cin,H,W] fp64, k [4,Cout,2,2,Cin] -> [B,Cout,2H,2W]: the four 2x2 parity convolutions."""
    B, _, H, W = x.shape
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros((B, k.shape[1], 2 * H, 2 * W), dtype=x.dtype, device=x.device)
    for py in range(2):
        for px in range(2):
            kk = k[py * 2 + px].permute(0, 3, 1, 2)  # [Cout, Cin, 2, 2]
            out[:, :, py::2, px::2] = F.conv2d(xp[:, :, py:py + H + 1, px:px + W + 1], kk)
    return out


def test_up2_parity_kernels_identity():
    """nearest-x2 + 3x3 conv == four 2x2 parity convs with pre-summed kernels (exact algebra)."""
    w = _rand(8, 4, 3, 3, seed=19).double()
    x = _rand(1, 4, 5, 6, seed=20).double()
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)
    out = _up2_from_parity(x, up2_parity_kernels(w.float()).double())
    assert float((out - ref).abs().max()) < 1e-5


# ------------------------------------------------------------------------------------------------ attention
def _attn_ref(qkv, B, N, h):
    D = h * 64
    q, k, v = qkv.double().reshape(B, N, 3, h, 64).permute(2, 0, 3, 1, 4)
    a = ((q @ k.transpose(-2, -1)) * 0.125).softmax(-1)
    return (a @ v).transpose(1, 2).reshape(B * N, D), a


@pytest.mark.parametrize("B,N,h", [(2, 48, 2), (1, 112, 8), (2, 496, 8), (1, 640, 2), (1, 1248, 1)])
def test_attention_f32(B, N, h):
    qkv = _rand(B * N, 3 * h * 64, seed=21)
    out = torch.empty((B * N, h * 64), dtype=torch.float32, device=DEV)
    probs = torch.empty((B, h, N, N), dtype=torch.float32, device=DEV) if N <= 496 else None
    _lib.check(U.lib().hvit_attention_f32(U.P(qkv), U.P(out), U.P(probs), B, N, h, U.stream()), "attention_f32")
    U.sync()
    ref, a = _attn_ref(qkv, B, N, h)
    assert U.rel_err(out, ref) < 1e-5
    if probs is not None:
        assert float((probs.double() - a).abs().max()) < 1e-6


@pytest.mark.parametrize("B,N,h", [(2, 48, 2), (1, 112, 8), (2, 496, 8), (1, 640, 2), (1, 1248, 1), (64, 496, 8)])
@pytest.mark.parametrize("kind", ["bf16", "fp16"])
def test_attention_16_tcgen05(B, N, h, kind):
    qkv = _rand(B * N, 3 * h * 64, seed=22).to(DT16[kind])
    out = torch.full((B * N, h * 64), float("nan"), dtype=DT16[kind], device=DEV)
    _lib.check(U.lib().hvit_attention_16(U.P(qkv), U.P(out), B, N, h, 1 if kind == "fp16" else 0, U.stream()),
               "attention_16")
    U.sync()
    sl = slice(0, min(B, 2) * N)
    ref, _ = _attn_ref(qkv[sl].float(), min(B, 2), N, h)
    assert torch.isfinite(out.float()).all()
    assert U.rel_err(out[sl].float(), ref) < (1.5e-2 if kind == "bf16" else 2e-3)   # P and the output are 16-bit


# ------------------------------------------------------------------------------------------------ glue
@pytest.mark.parametrize("rows,D", [(1000, 512), (37, 128), (64, 768)])
def test_layernorm(rows, D):
    x, g, b = _rand(rows, D, seed=23, scale=3.0) + 1.0, _rand(D, seed=24), _rand(D, seed=25)
    out = torch.empty((rows, D), dtype=torch.float32, device=DEV)
    _lib.check(U.lib().hvit_layernorm(U.P(x), U.P(g), U.P(b), U.P(out), 0, rows, D, 1e-5, U.stream()), "layernorm")
    ref = F.layer_norm(x.double(), (D,), g.double(), b.double(), 1e-5)
    assert U.rel_err(out, ref) < 1e-5
    for code, kind in ((1, "bf16"), (2, "fp16")):
        out16 = torch.empty((rows, D), dtype=DT16[kind], device=DEV)
        _lib.check(U.lib().hvit_layernorm(U.P(x), U.P(g), U.P(b), U.P(out16), code, rows, D, 1e-5, U.stream()), "layernorm")
        assert U.rel_err(out16.float(), ref) < OUT_TOL[kind]


def _ln_merge(st, D):
    """(mean, var) of every row from the per-slot (mean, M2) partials, as the consumer epilogue merges them."""
    S = st.shape[1]
    mean = st[..., 0].double().mean(1)
    m2 = st[..., 1].double().sum(1) + (D / S) * ((st[..., 0].double() - mean[:, None]) ** 2).sum(1)
    return mean, m2 / D


@pytest.mark.parametrize("kind", ["bf16", "fp16"])
@pytest.mark.parametrize("M,D,K,N2,act", [(1000, 512, 512, 1536, 0), (300, 512, 2048, 2048, 2), (257, 768, 768, 256, 0),
                                          (31744, 512, 512, 1536, 0)])
def test_layernorm_folded_into_gemms(M, D, K, N2, act, kind):
    """x += a W^T + b (producer: fp32 x, 16-bit copy, row statistics) followed by act(LN(x) W2^T + b2) (consumer)
    against fp64 nn.LayerNorm / nn.Linear: attention.py:258-262 (x = x + attn(norm1(x)); x = x + mlp(norm2(x)))."""
    f16 = 1 if kind == "fp16" else 0
    a = _rand(M, K, seed=40).to(DT16[kind])
    w = _rand(D, K, seed=41, scale=K ** -0.5).to(DT16[kind])
    bias = _rand(D, seed=42)
    x0 = _rand(M, D, seed=43, scale=2.0) + 0.7          # non-zero row means
    x = x0.clone()
    x16 = torch.full((M, D), float("nan"), dtype=DT16[kind], device=DEV)
    S = D // 128
    stats = torch.full((M, S, 2), float("nan"), dtype=torch.float32, device=DEV)
    _lib.check(U.lib().hvit_linear_ln_producer_16(U.P(a), K, U.P(w), U.P(bias), U.P(x), U.P(x16), U.P(stats), M, D, K, f16,
                                                  U.stream()), "ln_producer")
    U.sync()
    xref = x0.double() + a.double() @ w.double().T + bias.double()
    assert U.rel_err(x, xref) < 1e-5
    assert torch.equal(x16, x.to(DT16[kind]))                       # the 16-bit copy is the rounded fp32 result
    mean, var = _ln_merge(stats, D)
    assert (mean - x.double().mean(1)).abs().max() < 1e-5 * (1 + x.double().mean(1).abs().max())
    assert ((var - x.double().var(1, unbiased=False)).abs() / var).max() < 1e-4
    # rowstats of the same matrix must merge to the same statistics
    x16b = torch.empty_like(x16); stats_b = torch.empty_like(stats)
    _lib.check(U.lib().hvit_rowstats_16(U.P(x), U.P(x16b), U.P(stats_b), M, D, S, f16, U.stream()), "rowstats")
    U.sync()
    mean_b, var_b = _ln_merge(stats_b, D)
    assert torch.equal(x16b, x16)
    assert (mean_b - mean).abs().max() < 1e-5 and ((var_b - var).abs() / var).max() < 1e-4
    # consumer
    g, be = 1.0 + 0.3 * _rand(D, seed=44), 0.2 * _rand(D, seed=45)
    w2 = _rand(N2, D, seed=46, scale=D ** -0.5).to(DT16[kind])
    b2 = _rand(N2, seed=47)
    out = torch.full((M, N2), float("nan"), dtype=DT16[kind], device=DEV)
    wsc = torch.empty((N2, D), dtype=DT16[kind], device=DEV)
    gc = torch.empty(2 * N2, dtype=torch.float32, device=DEV)
    _lib.check(U.lib().hvit_linear_ln_consumer_16(U.P(x16), U.P(stats), S, U.P(w2), U.P(g), U.P(be), U.P(b2), 1e-5, act,
                                                  U.P(out), N2, M, N2, D, f16, U.P(wsc), U.P(gc), U.stream()), "ln_consumer")
    U.sync()
    ln = F.layer_norm(x.double(), (D,), g.double(), be.double(), 1e-5)
    ref = ln @ w2.double().T + b2.double()
    if act == 2:
        ref = F.gelu(ref)
    assert torch.isfinite(out.float()).all()
    assert U.rel_err(out.float(), ref) < OUT_TOL[kind]


@pytest.mark.parametrize("n", [64000, 16000, 9001, 2048])
def test_stft_istft_against_oracle(oracle, n):
    _, noisy = oracle.synth_clip(seed=n, n_samples=n)
    from hvit_b200.utils.audio_processing import compute_stft, compute_istft
    s = compute_stft(noisy)
    ref = oracle.stft(noisy)
    assert s.shape == ref.shape and s.dtype == np.complex64
    assert np.abs(s - ref).max() <= 2e-5 * np.abs(ref).max()
    y = compute_istft(ref, length=n)
    yref = oracle.istft(ref, length=n)
    assert np.abs(y - yref).max() <= 1e-5 * max(np.abs(yref).max(), 1e-6)
    assert np.abs(y - noisy).max() <= 1e-4 * np.abs(noisy).max()     # round trip


def test_stft_batched_normalised(oracle):
    B, n = 3, 8000
    clips = np.stack([oracle.synth_clip(seed=s, n_samples=n)[1] * (s + 1) for s in range(B)])
    x = torch.from_numpy(clips).cuda()
    T = 1 + n // 128
    spec = torch.empty((B, 257, T), dtype=torch.complex64, device=DEV)
    mag = torch.empty((B, 257, T), dtype=torch.float32, device=DEV)
    mv = torch.empty(B, dtype=torch.float32, device=DEV)
    mm = torch.empty(B, dtype=torch.float32, device=DEV)
    _lib.check(U.lib().hvit_stft(U.P(x), B, n, 1, U.P(mv), U.P(spec), U.P(mag), U.P(mm), U.stream()), "hvit_stft")
    U.sync()
    for b in range(B):
        peak = np.abs(clips[b]).max()
        assert abs(float(mv[b]) - peak) <= 1e-7 * peak
        ref = oracle.stft(clips[b] / peak)
        assert np.abs(spec[b].cpu().numpy() - ref).max() <= 2e-5 * np.abs(ref).max()
        assert abs(float(mm[b]) - np.abs(ref).max()) <= 2e-5 * np.abs(ref).max()
        assert np.abs(mag[b].cpu().numpy() - np.abs(ref)).max() <= 2e-5 * np.abs(ref).max()


@pytest.mark.parametrize("B,H,W,use_tc", [(2, 257, 501, 1), (1, 64, 70, 1), (3, 33, 129, 1), (2, 40, 51, 0)])
@pytest.mark.parametrize("kind", ["fp16", "bf16"])
def test_stem_16(B, H, W, use_tc, kind):
    """Encoder block 0: conv3x3(1->64)+BN+ReLU+pool, tensor-core window-GEMM kernel and CUDA-core kernel against torch
    fp32 (incl. the /mag_max input scaling, odd sizes, partial tiles)."""
    dt = torch.float16 if kind == "fp16" else torch.bfloat16
    g = torch.Generator(device="cpu").manual_seed(H * W + B)
    x = torch.rand((B, H, W), generator=g).to(DEV) * 3.0
    w = (torch.randn((64, 1, 3, 3), generator=g) * 0.4).to(DEV)
    scale = (torch.rand(64, generator=g) + 0.5).to(DEV)
    shift = (torch.randn(64, generator=g) * 0.2).to(DEV)
    mm = (torch.rand(B, generator=g) * 2 + 1.5).to(DEV)
    out = torch.empty((B, H // 2, W // 2, 64), dtype=dt, device=DEV)
    scratch = torch.empty(64 * 1024, dtype=torch.uint8, device=DEV)
    w9c = w[:, 0].permute(1, 2, 0).contiguous()
    _lib.check(U.lib().hvit_stem_16(U.P(x), U.P(mm.view(torch.int32)), U.P(w9c), U.P(scale), U.P(shift), U.P(out),
                                    U.P(scratch), B, H, W, 64, 2, 1 if kind == "fp16" else 0, use_tc, U.stream()), "stem")
    U.sync()
    xin = (x / mm[:, None, None])[:, None].double()
    ref = torch.nn.functional.conv2d(xin, w.double(), padding=1) * scale[None, :, None, None] + shift[None, :, None, None]
    ref = torch.nn.functional.max_pool2d(torch.relu(ref), 2).permute(0, 2, 3, 1)
    assert torch.isfinite(out.float()).all()
    assert U.rel_err(out.float(), ref) < OUT_TOL[kind]


@pytest.mark.parametrize("B,H,W", [(2, 64, 124), (1, 5, 7), (3, 17, 33), (1, 9, 140), (1, 6, 150)])
@pytest.mark.parametrize("kind", ["fp16", "bf16"])
def test_head_16(B, H, W, kind):
    """Last decoder block: conv3x3(64->1) + tanh with fp32 accumulation (row-streaming kernel) against torch."""
    dt = torch.float16 if kind == "fp16" else torch.bfloat16
    g = torch.Generator(device="cpu").manual_seed(B * H + W)
    x = torch.randn((B, H, W, 64), generator=g).to(DEV).to(dt)
    w = (torch.randn((1, 64, 3, 3), generator=g) * 0.05).to(DEV)
    logits = torch.empty((B, H, W), dtype=torch.float32, device=DEV)
    th = torch.empty_like(logits)
    w9c = w[0].permute(1, 2, 0).contiguous()
    _lib.check(U.lib().hvit_head_16(U.P(x), U.P(w9c), U.P(logits), U.P(th), B, H, W, 64, 1 if kind == "fp16" else 0,
                                    U.stream()), "head")
    U.sync()
    ref = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.double(), padding=1)[:, 0]  # (fp64: no TF32)
    assert U.rel_err(logits, ref) < 1e-5
    assert (th.double() - torch.tanh(ref)).abs().max() < 1e-5


# ------------------------------------------------------------------------------------------------ round-2 entry points
@pytest.mark.parametrize("kind", ["fp16", "bf16"])
@pytest.mark.parametrize("B,H,W,C,D,p", [(2, 64, 125, 256, 512, 4), (1, 16, 18, 128, 128, 4), (3, 8, 35, 64, 192, 4)])
def test_patch_embed_16(B, H, W, C, D, p, kind):
    """hvit_patch_embed_16 (IG_PATCH A-operand mode, 5-D TMA map) vs F.conv2d(stride=p) + flatten + pos_embed
    (reference components.py:282-307, 310-386); W % p != 0 drops the trailing columns like the reference conv."""
    x = _rand(B, H, W, C, seed=1).to(DT16[kind])
    w = _rand(D, C, p, p, seed=2, scale=0.03)
    bias, pos = _rand(D, seed=3, scale=0.1), _rand(2000, D, seed=4, scale=0.1)
    wk = conv_khwc(w).to(DT16[kind]).contiguous()
    Hp, Wp = H // p, W // p
    tokens = torch.empty(B * Hp * Wp, D, dtype=torch.float32, device=DEV)
    _lib.check(U.lib().hvit_patch_embed_16(U.P(x), B, H, W, C, U.P(wk), U.P(bias), U.P(pos), p, D, U.P(tokens),
                                           1 if kind == "fp16" else 0, U.stream()), "hvit_patch_embed_16")
    U.sync()
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), wk.double().permute(0, 3, 1, 2), bias.double(), stride=p)
    ref = ref.flatten(2).transpose(1, 2) + pos.double()[None, :Hp * Wp]
    assert U.rel_err(tokens.view(B, Hp * Wp, D), ref) < 2e-5


@pytest.mark.parametrize("kind", ["fp16", "bf16"])
@pytest.mark.parametrize("B,Hs,Ws,Cs,Cdec,Hd,Wd,Ccat,c_off", [(2, 64, 125, 256, 256, 16, 31, 512, 256),
                                                               (1, 128, 250, 64, 64, 32, 62, 192, 128),
                                                               (2, 16, 18, 128, 128, 16, 18, 256, 128)])
def test_skip_concat_16(B, Hs, Ws, Cs, Cdec, Hd, Wd, Ccat, c_off, kind):
    """hvit_skip_concat_16 vs the reference order of operations (1x1 conv on the full-resolution feature, THEN bilinear
    resize, then torch.cat - hybrid_vit.py:367-389); the other channels of the concat buffer stay untouched."""
    src = _rand(B, Hs, Ws, Cs, seed=1).to(DT16[kind])
    w = _rand(Cdec, Cs, seed=2, scale=0.05).to(DT16[kind])
    bias = _rand(Cdec, seed=3, scale=0.1)
    cat = torch.full((B, Hd, Wd, Ccat), 7.0, dtype=DT16[kind], device=DEV)
    scratch = torch.empty(B * Hd * Wd * Cs, dtype=DT16[kind], device=DEV)
    _lib.check(U.lib().hvit_skip_concat_16(U.P(src), B, Hs, Ws, Cs, U.P(w), U.P(bias), Cdec, U.P(cat), Hd, Wd, Ccat, c_off,
                                           U.P(scratch), 1 if kind == "fp16" else 0, U.stream()), "hvit_skip_concat_16")
    U.sync()
    proj = F.conv2d(src.double().permute(0, 3, 1, 2), w.double()[:, :, None, None], bias.double())
    if (Hs, Ws) != (Hd, Wd):
        proj = F.interpolate(proj, size=(Hd, Wd), mode="bilinear", align_corners=False)
    got = cat[..., c_off:c_off + Cdec].double().permute(0, 3, 1, 2)
    assert U.rel_err(got, proj) < OUT_TOL[kind] * 2      # two 16-bit roundings (sampled feature, output)
    rest = torch.cat([cat[..., :c_off], cat[..., c_off + Cdec:]], dim=-1)
    assert bool((rest == 7.0).all())


@pytest.mark.parametrize("precision", ["fp16", "bf16", "fp32"])
def test_pack_weights_matches_torch_reexpression(oracle, precision):
    """hvit_pack_weights (csrc/pack.cu, device kernels) against the same re-expressions written with torch ops
    (models/packing.py: fold_bn / conv_khwc / up2_parity_kernels) on seeded weights with non-trivial BatchNorm stats."""
    from hvit_b200.models import HybridViT
    from hvit_b200.models.packing import PackedWeights, fold_bn
    cfg = oracle.full_cfg(dict(encoder_channels=[64, 64, 128], embed_dim=128, num_heads=2, num_layers=2,
                               decoder_channels=[128, 64, 64, 1]))
    sd = oracle.make_state_dict(cfg, seed=9)
    m = HybridViT(precision=precision, encoder_channels=cfg["encoder_channels"], embed_dim=128, num_heads=2, num_layers=2,
                  decoder_channels=cfg["decoder_channels"])
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    prec = {"fp16": _lib.PREC_FP16, "bf16": _lib.PREC_BF16, "fp32": _lib.PREC_FP32}[precision]
    act = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[precision]
    pw = PackedWeights(m, prec)
    U.sync()
    base = pw.buffer.data_ptr()

    def view(ptr, shape, dtype):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        off = ptr - base
        assert 0 <= off and off + n <= pw.nbytes
        return pw.buffer[off:off + n].view(dtype).view(*shape)

    sdc = {k: v.cuda() for k, v in sd.items()}
    lowp = precision != "fp32"
    # stem
    sc, sh = fold_bn(sdc, "encoder.0.block.1")
    assert torch.equal(view(pw.c.stem_w, (3, 3, 64), torch.float32), sdc["encoder.0.block.0.weight"][:, 0].permute(1, 2, 0))
    assert torch.allclose(view(pw.c.stem_scale, (64,), torch.float32), sc, rtol=1e-6, atol=0)
    assert torch.allclose(view(pw.c.stem_shift, (64,), torch.float32), sh, rtol=1e-5, atol=1e-7)
    # encoder.1 (BN scale folded into the weights in the 16-bit modes)
    w = sdc["encoder.1.block.0.weight"]
    sc, sh = fold_bn(sdc, "encoder.1.block.1")
    ref = conv_khwc(w * sc[:, None, None, None] if lowp else w)
    got = view(pw.c.enc_w[1], (64, 3, 3, 64), act).float()
    ulp = {"bf16": 2.0 ** -8, "fp16": 2.0 ** -11, "fp32": 0.0}[precision]   # one unit in the last place of the type
    assert float((got - ref.to(act).float()).abs().max()) <= float(ref.abs().max()) * ulp
    assert (pw.c.enc_scale[1] is None) == lowp
    # decoder.1: nearest x2 + 3x3 -> four 2x2 parity kernels (16-bit modes), original layout in fp32
    w = sdc["decoder.1.block.1.weight"]
    sc, sh = fold_bn(sdc, "decoder.1.block.2")
    if lowp:
        ref = up2_parity_kernels(w * sc[:, None, None, None])
        got = view(pw.c.dec_w[1], tuple(ref.shape), act).float()
        assert float((got - ref.to(act).float()).abs().max()) <= float(ref.abs().max()) * ulp
    else:
        assert torch.equal(view(pw.c.dec_w[1], (64, 3, 3, w.shape[1]), act), conv_khwc(w))
        assert torch.allclose(view(pw.c.dec_scale[1], (64,), torch.float32), sc, rtol=1e-6, atol=0)
    # linear weights / biases / positional table / head
    assert torch.equal(view(pw.c.qkv_w[1], (384, 128), act), sdc["transformer.blocks.1.attn.qkv.weight"].to(act))
    assert torch.equal(view(pw.c.fc2_b[0], (128,), torch.float32), sdc["transformer.blocks.0.mlp.net.3.bias"])
    assert torch.equal(view(pw.c.pos_embed, (10000, 128), torch.float32), sdc["pos_encoding.pos_embed"][0])
    assert torch.equal(view(pw.c.head_w, (3, 3, 64), torch.float32), sdc["decoder.3.block.0.weight"][0].permute(1, 2, 0))
    assert torch.equal(view(pw.c.skip_w[2], (64, 64), act), sdc["skip_projections.2.weight"].reshape(64, 64).to(act))


@pytest.mark.parametrize("n", [64000, 9001, 2048, 1920, 3333, 16000 + 127])
def test_fused_enhance_backend_matches_two_kernel_istft(oracle, n):
    """The fused back end of the enhance path (phase recomputed from the waveform, overlap-add in shared memory, block
    borders every 13 hops) against the stand-alone STFT -> iSTFT entry points: with a model output of all ones
    (mag_norm == 1) both must reproduce |S|-free reconstruction; checked through a tiny model's enhance() vs oracle in
    test_gpu_model.py, here at the kernel level through clip lengths that put the block borders and the ragged tail in
    every position."""
    from hvit_b200.utils.audio_processing import compute_stft, compute_istft
    _, noisy = oracle.synth_clip(seed=n, n_samples=n)
    s = compute_stft(noisy)
    ref = oracle.stft(noisy)
    assert np.abs(s - ref).max() <= 3e-6 * np.abs(ref).max()
    y = compute_istft(s, length=n)
    assert np.abs(y - noisy).max() <= 2e-5


def test_second_device_has_its_own_tables():
    """Per-device one-time state (FFT tables, function attributes, SM count): a second GPU in the same process works."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from hvit_b200.utils.audio_processing import compute_stft
    x = np.sin(np.arange(4096) * 0.01).astype(np.float32)
    with torch.cuda.device(0):
        a = compute_stft(x)
    with torch.cuda.device(1):
        b = compute_stft(x)
    assert np.array_equal(a, b)
