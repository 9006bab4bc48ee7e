"""CPU oracle for the HybridViT speech-enhancement inference path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline.  The product path (``hvit_b200``) never imports this file
and fails loudly when its CUDA library is missing.

What it is: a functional, state-dict driven restatement, in fp32 on the CPU, of
the reference's arithmetic for

  * ``AudioEnhancer.enhance``           (reference inference/enhancer.py:55-135)
  * ``HybridViT.forward``               (reference models/hybrid_vit.py:396-469)
  * the librosa STFT / iSTFT front end  (reference utils/audio_processing.py:67-193)

Pinning status
  * model forward: PINNED.  ``tests/golden/make_golden.py`` runs the reference's
    own ``models.HybridViT`` (imported from /root/reference in the build
    container) on seeded weights/inputs and stores its outputs under
    ``tests/golden/``; ``tests/test_oracle.py`` checks this file against them.
  * enhance pipeline: pinned the same way, by executing the reference's
    ``inference/enhancer.py`` verbatim with a shim ``librosa`` module.
  * STFT / iSTFT arithmetic itself: PARITY UNPINNED against librosa.  librosa
    (``librosa>=0.10.0``, reference requirements.txt:9, no lock file) is a
    third-party dependency that is neither vendored in the reference nor
    installed here, and the reference ships no tests or golden vectors.  The
    restatement below follows librosa 0.10's published algorithm (centered
    zero padding, periodic Hann from ``scipy.signal.get_window``, float64
    window times float32 frames, ``scipy.fft.rfft``; inverse: ``irfft`` times
    window, overlap-add, trim, divide by the window sum-square envelope) and is
    cross-checked against ``torch.stft`` / ``torch.istft`` in the tests.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.fft
import scipy.signal
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# Model configuration (defaults = reference models/hybrid_vit.py:36-67)
# ----------------------------------------------------------------------------

DEFAULT_CFG = dict(
    input_channels=1,
    output_channels=1,
    encoder_channels=[64, 128, 256],
    encoder_kernel_sizes=[3, 3, 3],
    encoder_pool_sizes=[2, 2, 1],
    embed_dim=512,
    num_heads=8,
    num_layers=6,
    mlp_ratio=4.0,
    patch_size=4,
    decoder_channels=[256, 128, 64, 1],
    decoder_kernel_sizes=[3, 3, 3, 3],
    decoder_upsample_factors=[1, 2, 2, 1],
    use_skip_connections=True,
)

BN_EPS = 1e-5  # nn.BatchNorm2d default, reference components.py:67
LN_EPS = 1e-5  # nn.LayerNorm default, reference attention.py:152


def full_cfg(cfg: Optional[dict] = None) -> dict:
    out = dict(DEFAULT_CFG)
    if cfg:
        out.update(cfg)
    return out


# ----------------------------------------------------------------------------
# STFT front end (librosa >= 0.10 semantics; reference enhancer.py:82-93,122-129)
# ----------------------------------------------------------------------------

def _padded_window(win_length: int, n_fft: int, window: str = "hann") -> np.ndarray:
    """Periodic window, float64, centre-padded to n_fft (librosa.filters.get_window
    + util.pad_center)."""
    w = scipy.signal.get_window(window, win_length, fftbins=True).astype(np.float64)
    if win_length < n_fft:
        lpad = (n_fft - win_length) // 2
        w = np.pad(w, (lpad, n_fft - win_length - lpad))
    return w


def stft(y: np.ndarray, n_fft: int = 512, hop_length: int = 128,
         win_length: Optional[int] = None, window: str = "hann",
         center: bool = True) -> np.ndarray:
    """librosa.stft restatement.  float32 in -> complex64 [1+n_fft/2, 1+n//hop]."""
    win_length = win_length or n_fft
    y = np.asarray(y)
    w = _padded_window(win_length, n_fft, window)
    if center:
        y = np.pad(y, (n_fft // 2, n_fft // 2), mode="constant")
    n_frames = 1 + (len(y) - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_frames)[None, :]
    frames = y[idx]                                   # [n_fft, T], input dtype
    spec = scipy.fft.rfft(w[:, None] * frames, axis=0)  # float64 product -> complex128
    out_dtype = np.complex64 if y.dtype == np.float32 else np.complex128
    return spec.astype(out_dtype)


def window_sumsquare(n_frames: int, n_fft: int, hop_length: int, win_length: int,
                     window: str = "hann", dtype=np.float32) -> np.ndarray:
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=dtype)
    wsq = _padded_window(win_length, n_fft, window) ** 2
    for i in range(n_frames):
        s = i * hop_length
        x[s:min(n, s + n_fft)] += wsq[:max(0, min(n_fft, n - s))]
    return x


def istft(spec: np.ndarray, hop_length: int = 128, win_length: Optional[int] = None,
          window: str = "hann", center: bool = True,
          length: Optional[int] = None) -> np.ndarray:
    """librosa.istft restatement.  complex64 [F, T] -> float32 [length]."""
    n_fft = 2 * (spec.shape[0] - 1)
    win_length = win_length or n_fft
    n_frames = spec.shape[1]
    dtype = np.float32 if spec.dtype == np.complex64 else np.float64
    w = _padded_window(win_length, n_fft, window)
    full_len = n_fft + hop_length * (n_frames - 1)
    ytmp = w[:, None] * scipy.fft.irfft(spec, n=n_fft, axis=0)   # float64 product
    y_full = np.zeros(full_len, dtype=dtype)
    for t in range(n_frames):
        y_full[t * hop_length:t * hop_length + n_fft] += ytmp[:, t].astype(dtype)
    start = n_fft // 2 if center else 0
    if length is None:
        length = full_len - 2 * start
    y = np.zeros(length, dtype=dtype)
    avail = min(length, full_len - start)
    y[:avail] = y_full[start:start + avail]
    wss = window_sumsquare(n_frames, n_fft, hop_length, win_length, window, dtype=dtype)
    wss_c = np.zeros(length, dtype=dtype)
    wss_c[:avail] = wss[start:start + avail]
    nz = wss_c > np.finfo(dtype).tiny
    y[nz] /= wss_c[nz]
    return y


# ----------------------------------------------------------------------------
# HybridViT forward, functional (reference hybrid_vit.py:286-469)
# ----------------------------------------------------------------------------

def _conv_block(sd, prefix: str, x, pool: int, stages, name):
    """ConvBlock in eval mode (reference components.py:15-99): Conv3x3(no bias) ->
    BN(running stats) -> ReLU -> [Dropout2d = id] -> [MaxPool(pool)]."""
    w = sd[f"{prefix}.block.0.weight"]
    x = F.conv2d(x, w, None, stride=1, padding=w.shape[-1] // 2)
    x = F.batch_norm(x, sd[f"{prefix}.block.1.running_mean"], sd[f"{prefix}.block.1.running_var"],
                     sd[f"{prefix}.block.1.weight"], sd[f"{prefix}.block.1.bias"],
                     training=False, eps=BN_EPS)
    x = F.relu(x)
    if pool > 1:
        x = F.max_pool2d(x, pool)
    if stages is not None:
        stages[name] = x
    return x


def _layer_norm(x, sd, prefix):
    return F.layer_norm(x, (x.shape[-1],), sd[f"{prefix}.weight"], sd[f"{prefix}.bias"], LN_EPS)


def _attention(sd, prefix, x, num_heads: int, attns: Optional[list]):
    """MultiHeadSelfAttention, eval mode, no mask (reference attention.py:64-115)."""
    B, N, C = x.shape
    hd = C // num_heads
    qkv = F.linear(x, sd[f"{prefix}.qkv.weight"], sd[f"{prefix}.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    a = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    a = a.softmax(dim=-1)
    if attns is not None:
        attns.append(a)
    o = (a @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(o, sd[f"{prefix}.proj.weight"], sd[f"{prefix}.proj.bias"])


def hybrid_vit_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, cfg: Optional[dict] = None,
                       return_attentions: bool = False,
                       stages: Optional[Dict[str, torch.Tensor]] = None):
    """fp32 CPU forward of HybridViT in eval mode from a reference-keyed state_dict.

    ``stages`` (optional dict) receives the intermediate tensors listed in
    SURVEY.md section 8(a) (NCHW / [B,N,D] as in the reference)."""
    cfg = full_cfg(cfg)
    assert not cfg.get("use_cls_token", False)
    in_hw = x.shape[2:]
    # 1. CNN encoder; post-pool outputs are the skip features (hybrid_vit.py:286-307)
    skips = []
    for i, pool in enumerate(cfg["encoder_pool_sizes"]):
        x = _conv_block(sd, f"encoder.{i}", x, pool, stages, f"encoder.{i}")
        skips.append(x)
    # 2. patch embedding: Conv2d(k=p, s=p, bias) -> [B, N, D] (components.py:282-307)
    p = cfg["patch_size"]
    x = F.conv2d(x, sd["patch_embed.projection.weight"], sd["patch_embed.projection.bias"], stride=p)
    B, D, Hp, Wp = x.shape
    x = x.flatten(2).transpose(1, 2)
    if stages is not None:
        stages["patch_embed"] = x
    # 3. learnable positional encoding + transformer (hybrid_vit.py:309-350, attention.py:273-304)
    N = x.shape[1]
    x = x + sd["pos_encoding.pos_embed"][:, :N, :]
    attns = [] if return_attentions else None
    for l in range(cfg["num_layers"]):
        pre = f"transformer.blocks.{l}"
        x = x + _attention(sd, f"{pre}.attn", _layer_norm(x, sd, f"{pre}.norm1"), cfg["num_heads"], attns)
        h = F.linear(_layer_norm(x, sd, f"{pre}.norm2"), sd[f"{pre}.mlp.net.0.weight"], sd[f"{pre}.mlp.net.0.bias"])
        h = F.gelu(h)  # exact erf GELU (nn.GELU default, components.py:225)
        x = x + F.linear(h, sd[f"{pre}.mlp.net.3.weight"], sd[f"{pre}.mlp.net.3.bias"])
        if stages is not None:
            stages[f"transformer.blocks.{l}"] = x
    x = _layer_norm(x, sd, "transformer.norm")
    if stages is not None:
        stages["transformer"] = x
    x = F.linear(x, sd["to_feature_map.weight"], sd["to_feature_map.bias"])
    x = x.transpose(1, 2).reshape(B, x.shape[-1], Hp, Wp)
    if stages is not None:
        stages["to_feature_map"] = x
    # 4. decoder with skip projections (hybrid_vit.py:352-394)
    rskips = skips[::-1]
    n_dec = len(cfg["decoder_channels"])
    for i in range(n_dec):
        final = i == n_dec - 1
        if cfg["use_skip_connections"] and not final and i < len(rskips):
            s = F.conv2d(rskips[i], sd[f"skip_projections.{i}.weight"], sd[f"skip_projections.{i}.bias"])
            if s.shape[2:] != x.shape[2:]:
                s = F.interpolate(s, size=x.shape[2:], mode="bilinear", align_corners=False)
            x = torch.cat([x, s], dim=1)
        up = cfg["decoder_upsample_factors"][i]
        ci = 0
        if up > 1:
            x = F.interpolate(x, scale_factor=up, mode="nearest")
            ci = 1
        w = sd[f"decoder.{i}.block.{ci}.weight"]
        x = F.conv2d(x, w, None, padding=w.shape[-1] // 2)
        if final:
            if stages is not None:
                stages[f"decoder.{i}.pre_tanh"] = x
            x = torch.tanh(x)
        else:
            bn = f"decoder.{i}.block.{ci + 1}"
            x = F.batch_norm(x, sd[f"{bn}.running_mean"], sd[f"{bn}.running_var"], sd[f"{bn}.weight"],
                             sd[f"{bn}.bias"], training=False, eps=BN_EPS)
            x = F.relu(x)
        if stages is not None:
            stages[f"decoder.{i}"] = x
    # 5. resize back to the input resolution (hybrid_vit.py:458-465)
    if x.shape[2:] != in_hw:
        x = F.interpolate(x, size=in_hw, mode="bilinear", align_corners=False)
    if return_attentions:
        return x, attns
    return x


# ----------------------------------------------------------------------------
# AudioEnhancer.enhance (reference inference/enhancer.py:55-135)
# ----------------------------------------------------------------------------

def enhance(sd: Dict[str, torch.Tensor], noisy_audio: np.ndarray, cfg: Optional[dict] = None,
            normalize: bool = True, n_fft: int = 512, hop_length: int = 128,
            win_length: int = 512, window: str = "hann",
            debug: Optional[dict] = None) -> np.ndarray:
    """One clip, host numpy in/out, float32.  ``debug`` receives intermediate
    arrays (noisy magnitude, enhanced magnitude, scalars)."""
    x = np.asarray(noisy_audio, dtype=np.float32)
    if normalize:
        max_val = np.abs(x).max() if x.size else np.float32(0)
        if max_val > 1e-8:
            x = x / max_val
        else:
            max_val = 1.0
    else:
        max_val = 1.0
    spec = stft(x, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window, center=True)
    mag = np.abs(spec)
    phase = np.angle(spec)
    mag_max = mag.max()
    if mag_max > 1e-8:
        mag_n = mag / mag_max
    else:
        mag_n = mag
        mag_max = 1.0
    t = torch.from_numpy(mag_n).float()[None, None]
    with torch.no_grad():
        out = hybrid_vit_forward(sd, t, cfg)
    enh_mag = out.squeeze().numpy() * mag_max
    enh_spec = enh_mag * np.exp(1j * phase)
    y = istft(enh_spec.astype(np.complex64), hop_length=hop_length, win_length=win_length,
              window=window, center=True, length=len(x))
    if normalize:
        y = y * max_val
    if debug is not None:
        debug.update(noisy_mag=mag, noisy_mag_norm=mag_n, mag_max=float(mag_max), max_val=float(max_val),
                     enhanced_mag=enh_mag, model_out=out.squeeze().numpy())
    return y.astype(np.float32)


def enhance_batch(sd: Dict[str, torch.Tensor], noisy_batch: np.ndarray, cfg: Optional[dict] = None,
                  normalize: bool = True, chunk: int = 16) -> np.ndarray:
    """Equal-length clips [B, n] -> [B, n]: the same per-clip arithmetic as ``enhance`` (reference enhancer.py:55-135
    applied clip by clip), with the model forward batched over ``chunk`` clips at a time (the reference's forward
    accepts a batch, hybrid_vit.py:396-411; chunking only bounds the CPU memory of the fp32 activations).  Used by
    bench.py as the same-configuration CPU arm (batch 64)."""
    x = np.asarray(noisy_batch, dtype=np.float32)
    B, n = x.shape
    out = np.zeros_like(x)
    for c0 in range(0, B, chunk):
        xs = x[c0:c0 + chunk]
        max_vals, specs, mags, mag_maxs = [], [], [], []
        for xi in xs:
            mv = np.abs(xi).max() if (normalize and xi.size) else np.float32(1.0)
            if not (normalize and mv > 1e-8):
                mv = 1.0
            sp = stft(xi / mv if mv != 1.0 else xi)
            mg = np.abs(sp)
            mm = mg.max()
            if not mm > 1e-8:
                mm = 1.0
            max_vals.append(mv); specs.append(sp); mags.append(mg / mm if mm != 1.0 else mg); mag_maxs.append(mm)
        t = torch.from_numpy(np.stack(mags)).float()[:, None]
        with torch.no_grad():
            y = hybrid_vit_forward(sd, t, cfg)[:, 0].numpy()
        for i, (mv, sp, mm) in enumerate(zip(max_vals, specs, mag_maxs)):
            enh_spec = (y[i] * mm) * np.exp(1j * np.angle(sp))
            w = istft(enh_spec.astype(np.complex64), length=n)
            out[c0 + i] = w * mv if normalize else w
    return out


# ----------------------------------------------------------------------------
# Parity metrics (SI-SDR follows reference evaluation/metrics.py:100-145)
# ----------------------------------------------------------------------------

def si_sdr(clean: np.ndarray, enhanced: np.ndarray, eps: float = 1e-8) -> float:
    n = min(len(clean), len(enhanced))
    c = np.asarray(clean[:n], dtype=np.float64)
    e = np.asarray(enhanced[:n], dtype=np.float64)
    c = c - c.mean()
    e = e - e.mean()
    alpha = np.dot(e, c) / (np.dot(c, c) + eps)
    cs = alpha * c
    return float(10 * np.log10(np.sum(cs ** 2) / (np.sum((e - cs) ** 2) + eps)))


def max_rel_err(ours: np.ndarray, ref: np.ndarray) -> float:
    """max|ours-ref| / max|ref|  (the north-star 'max relative spectrogram error')."""
    ref = np.asarray(ref, dtype=np.float64)
    d = np.abs(np.asarray(ours, dtype=np.float64) - ref).max()
    return float(d / max(np.abs(ref).max(), 1e-30))


def rms_rel_err(ours: np.ndarray, ref: np.ndarray) -> float:
    ref = np.asarray(ref, dtype=np.float64)
    d = np.asarray(ours, dtype=np.float64) - ref
    return float(np.sqrt((d ** 2).mean()) / max(np.sqrt((ref ** 2).mean()), 1e-30))


# ----------------------------------------------------------------------------
# Deterministic weights and clips shared by tests, golden generator and bench
# ----------------------------------------------------------------------------

def state_dict_spec(cfg: Optional[dict] = None, max_len: int = 10000) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) for every entry of the reference state_dict, in the
    reference's registration order (keys measured in SURVEY.md section 8 a18)."""
    cfg = full_cfg(cfg)
    spec: List[Tuple[str, Tuple[int, ...], str]] = []

    def bn(prefix, c):
        spec.extend([(f"{prefix}.weight", (c,), "bn_w"), (f"{prefix}.bias", (c,), "bn_b"),
                     (f"{prefix}.running_mean", (c,), "bn_mean"), (f"{prefix}.running_var", (c,), "bn_var"),
                     (f"{prefix}.num_batches_tracked", (), "count")])

    cin = cfg["input_channels"]
    for i, (c, k) in enumerate(zip(cfg["encoder_channels"], cfg["encoder_kernel_sizes"])):
        spec.append((f"encoder.{i}.block.0.weight", (c, cin, k, k), "conv"))
        bn(f"encoder.{i}.block.1", c)
        cin = c
    D, p = cfg["embed_dim"], cfg["patch_size"]
    spec.append(("patch_embed.projection.weight", (D, cin, p, p), "conv"))
    spec.append(("patch_embed.projection.bias", (D,), "bias"))
    spec.append(("pos_encoding.pos_embed", (1, max_len, D), "pos"))
    hid = int(D * cfg["mlp_ratio"])
    for l in range(cfg["num_layers"]):
        pre = f"transformer.blocks.{l}"
        spec += [(f"{pre}.norm1.weight", (D,), "ln_w"), (f"{pre}.norm1.bias", (D,), "ln_b"),
                 (f"{pre}.norm2.weight", (D,), "ln_w"), (f"{pre}.norm2.bias", (D,), "ln_b"),
                 (f"{pre}.attn.qkv.weight", (3 * D, D), "linear"), (f"{pre}.attn.qkv.bias", (3 * D,), "bias"),
                 (f"{pre}.attn.proj.weight", (D, D), "linear"), (f"{pre}.attn.proj.bias", (D,), "bias"),
                 (f"{pre}.mlp.net.0.weight", (hid, D), "linear"), (f"{pre}.mlp.net.0.bias", (hid,), "bias"),
                 (f"{pre}.mlp.net.3.weight", (D, hid), "linear"), (f"{pre}.mlp.net.3.bias", (D,), "bias")]
    spec += [("transformer.norm.weight", (D,), "ln_w"), ("transformer.norm.bias", (D,), "ln_b")]
    spec += [("to_feature_map.weight", (cin, D), "linear"), ("to_feature_map.bias", (cin,), "bias")]
    dch = cfg["decoder_channels"]
    for i, (c, k, up) in enumerate(zip(dch, cfg["decoder_kernel_sizes"], cfg["decoder_upsample_factors"])):
        ic = dch[0] if i == 0 else dch[i - 1]
        final = i == len(dch) - 1
        if cfg["use_skip_connections"] and not final:
            ic += c
        ci = 1 if up > 1 else 0
        spec.append((f"decoder.{i}.block.{ci}.weight", (c, ic, k, k), "head" if final else "conv"))
        if not final:
            bn(f"decoder.{i}.block.{ci + 1}", c)
    if cfg["use_skip_connections"]:
        for i, (ec, dc) in enumerate(zip(cfg["encoder_channels"][::-1], dch[:-1])):
            spec.append((f"skip_projections.{i}.weight", (dc, ec, 1, 1), "conv"))
            spec.append((f"skip_projections.{i}.bias", (dc,), "bias"))
    return spec


def make_state_dict(cfg: Optional[dict] = None, seed: int = 0, head_scale: float = 0.05,
                    max_len: int = 10000) -> Dict[str, torch.Tensor]:
    """Seeded, construction-order independent weights that exercise every
    parameter: randomised BatchNorm statistics/affine (a BN-folding bug is
    invisible with the default 0/1 stats), non-zero biases, non-trivial
    LayerNorm affine, and a down-scaled output head so tanh is not saturated
    (SURVEY.md section 7 'hard parts')."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def randn(shape, std=1.0, mean=0.0):
        return torch.randn(shape, generator=g, dtype=torch.float32) * std + mean

    def rand(shape, lo, hi):
        return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    for key, shape, kind in state_dict_spec(cfg, max_len):
        if kind in ("conv", "head"):
            fan_out = shape[0] * shape[2] * shape[3]
            t = randn(shape, math.sqrt(2.0 / fan_out))
            if kind == "head":
                t = t * head_scale
        elif kind == "linear":
            t = randn(shape, 0.02).clamp_(-0.04, 0.04)
        elif kind == "bias":
            t = randn(shape, 0.02)
        elif kind == "pos":
            t = randn(shape, 0.02)
        elif kind == "bn_w":
            t = rand(shape, 0.6, 1.4)
        elif kind == "bn_b":
            t = randn(shape, 0.1)
        elif kind == "bn_mean":
            t = randn(shape, 0.1)
        elif kind == "bn_var":
            t = rand(shape, 0.5, 1.5)
        elif kind == "ln_w":
            t = randn(shape, 0.1, 1.0)
        elif kind == "ln_b":
            t = randn(shape, 0.05)
        elif kind == "count":
            t = torch.zeros((), dtype=torch.long)
        else:
            raise KeyError(kind)
        sd[key] = t
    return sd


def state_dict_digest(sd: Dict[str, torch.Tensor]) -> str:
    import hashlib
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def synth_clip(seconds: float = 4.0, sr: int = 16000, snr_db: float = 5.0, seed: int = 0,
               n_samples: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
    """(clean, noisy) float32: harmonic stack x slow envelope + white noise at
    ``snr_db`` (SURVEY.md section 8 c)."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr)) if n_samples is None else n_samples
    t = np.arange(n) / sr
    f0 = rng.uniform(90.0, 250.0)
    clean = np.zeros(n)
    for h in range(1, 9):
        clean += (1.0 / h) * np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 2 * np.pi))
    env = 0.55 + 0.45 * np.sin(2 * np.pi * rng.uniform(1.5, 4.0) * t + rng.uniform(0, 2 * np.pi))
    clean = clean * env
    clean = 0.3 * clean / max(np.abs(clean).max(), 1e-9)
    noise = rng.standard_normal(n)
    ps = np.mean(clean ** 2) + 1e-12
    pn = np.mean(noise ** 2) + 1e-12
    noise = noise * math.sqrt(ps / (pn * 10 ** (snr_db / 10)))
    return clean.astype(np.float32), (clean + noise).astype(np.float32)
