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
