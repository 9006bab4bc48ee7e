"""Speech-quality metrics with the reference's names, signatures and conventions
(reference evaluation/metrics.py:16-365).  SI-SDR / SNR / segmental SNR / log-spectral distance are plain numpy
(float64 accumulation like numpy on the reference's float inputs); PESQ and STOI are third-party packages in the
reference too - they are used when importable and report 0.0 otherwise, exactly like the reference does."""
from __future__ import annotations

from typing import Dict

import numpy as np


def _same_length(a: np.ndarray, b: np.ndarray):
    n = min(len(a), len(b))
    return a[:n], b[:n]


def compute_pesq(clean: np.ndarray, enhanced: np.ndarray, sr: int = 16000, mode: str = "wb") -> float:
    """reference metrics.py:16-55"""
    try:
        from pesq import pesq
        clean, enhanced = _same_length(clean, enhanced)
        return float(pesq(sr, clean, enhanced, mode))
    except ImportError:
        print("Warning: pesq library not installed. Install with: pip install pesq")
        return 0.0
    except Exception as e:  # noqa: BLE001  (the reference swallows metric errors the same way)
        print(f"Error computing PESQ: {e}")
        return 0.0


def compute_stoi(clean: np.ndarray, enhanced: np.ndarray, sr: int = 16000, extended: bool = False) -> float:
    """reference metrics.py:58-97"""
    try:
        from pystoi import stoi
        clean, enhanced = _same_length(clean, enhanced)
        return float(stoi(clean, enhanced, sr, extended=extended))
    except ImportError:
        print("Warning: pystoi library not installed. Install with: pip install pystoi")
        return 0.0
    except Exception as e:  # noqa: BLE001
        print(f"Error computing STOI: {e}")
        return 0.0


def compute_sisdr(clean: np.ndarray, enhanced: np.ndarray, eps: float = 1e-8) -> float:
    """Scale-invariant SDR in dB (reference metrics.py:100-145): zero-mean both signals, project the estimate on the
    reference, 10 log10(|target|^2 / (|residual|^2 + eps))."""
    clean, enhanced = _same_length(np.asarray(clean), np.asarray(enhanced))
    clean = clean - np.mean(clean)
    enhanced = enhanced - np.mean(enhanced)
    alpha = np.dot(enhanced, clean) / (np.dot(clean, clean) + eps)
    target = alpha * clean
    return float(10 * np.log10(np.sum(target ** 2) / (np.sum((enhanced - target) ** 2) + eps)))


def compute_snr(clean: np.ndarray, noisy_or_enhanced: np.ndarray, eps: float = 1e-8) -> float:
    """reference metrics.py:148-184"""
    clean, other = _same_length(np.asarray(clean), np.asarray(noisy_or_enhanced))
    return float(10 * np.log10(np.mean(clean ** 2) / (np.mean((other - clean) ** 2) + eps)))


def compute_segsnr(clean: np.ndarray, enhanced: np.ndarray, frame_length: int = 512, hop_length: int = 256,
                   eps: float = 1e-8) -> float:
    """Mean of the per-frame SNRs clipped to [-10, 35] dB over frames with signal and noise power > eps
    (reference metrics.py:187-243; note its frame loop stops at len - frame_length, exclusive)."""
    clean, enhanced = _same_length(np.asarray(clean), np.asarray(enhanced))
    vals = []
    for i in range(0, len(clean) - frame_length, hop_length):
        c = clean[i:i + frame_length]
        n = enhanced[i:i + frame_length] - c
        ps, pn = np.mean(c ** 2), np.mean(n ** 2)
        if ps > eps and pn > eps:
            vals.append(np.clip(10 * np.log10(ps / pn), -10, 35))
    return float(np.mean(vals)) if vals else 0.0


def _stft_mag(x: np.ndarray, n_fft: int, hop_length: int) -> np.ndarray:
    """|librosa.stft(x, n_fft, hop_length)|: centred frames, zero padding, periodic Hann, [1 + n_fft/2, 1 + n/hop]."""
    x = np.asarray(x)
    xp = np.pad(x, n_fft // 2)
    win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)).astype(x.dtype if x.dtype.kind == "f" else np.float64)
    t = 1 + len(x) // hop_length
    idx = np.arange(n_fft)[None, :] + hop_length * np.arange(t)[:, None]
    from scipy import fft as _fft  # (keeps float32 input in single precision, like librosa)
    return np.abs(_fft.rfft(xp[idx] * win[None, :], axis=1)).T


def compute_lsd(clean: np.ndarray, enhanced: np.ndarray, sr: int = 16000, n_fft: int = 512, hop_length: int = 128,
                eps: float = 1e-10) -> float:
    """Log-spectral distance (reference metrics.py:246-296): mean over frames of the rms over bins of the natural-log
    magnitude difference."""
    clean, enhanced = _same_length(np.asarray(clean), np.asarray(enhanced))
    a, b = _stft_mag(clean, n_fft, hop_length), _stft_mag(enhanced, n_fft, hop_length)
    return float(np.mean(np.sqrt(np.mean((np.log(a + eps) - np.log(b + eps)) ** 2, axis=0))))


def compute_metrics_device(clean, enhanced, lengths=None) -> Dict[str, np.ndarray]:
    """SI-SDR / SNR / segmental SNR / LSD of a whole batch ON THE GPU (``hvit_metrics``, csrc/metrics.cu): ``clean`` and
    ``enhanced`` are float32 CUDA tensors [B, n] (zero-padded to a common length, ``lengths`` = true sample counts), so
    scoring the enhancer's device output needs no device -> host -> numpy trip.  Returns per-clip numpy float64 arrays
    under the reference's keys.  Same arithmetic as the host functions above (which stay the reference-pinned checker)."""
    import ctypes as C
    import torch
    from .. import _lib
    if clean.shape != enhanced.shape or clean.dim() != 2 or not clean.is_cuda or not enhanced.is_cuda:
        raise ValueError("compute_metrics_device expects two float32 CUDA tensors [B, n] of the same shape")
    clean, enhanced = clean.contiguous().float(), enhanced.contiguous().float()
    B, n = clean.shape
    lib = _lib.load()
    with torch.cuda.device(clean.device):
        nbytes = lib.hvit_metrics_scratch_bytes(B, n)
        scratch = torch.empty(nbytes + 256, dtype=torch.uint8, device=clean.device)
        off = (-scratch.data_ptr()) % 256
        out = torch.empty((B, 4), dtype=torch.float64, device=clean.device)
        nv = None
        if lengths is not None:
            nv = torch.as_tensor(list(lengths), dtype=torch.int32).to(clean.device)
        _lib.check(lib.hvit_metrics(clean.data_ptr(), enhanced.data_ptr(), B, n, _lib.ptr(nv), scratch.data_ptr() + off,
                                    nbytes, out.data_ptr(), _lib.current_stream_ptr()), "hvit_metrics")
        res = out.cpu().numpy()
    return {"sisdr": res[:, 0], "snr": res[:, 1], "segsnr": res[:, 2], "lsd": res[:, 3]}


def compute_all_metrics(clean: np.ndarray, enhanced: np.ndarray, noisy: np.ndarray = None, sr: int = 16000) -> Dict[str, float]:
    """reference metrics.py:299-349: same keys (pesq, stoi, sisdr, snr, segsnr, lsd and the *_improvement entries when
    the noisy input is given)."""
    m = {"pesq": compute_pesq(clean, enhanced, sr), "stoi": compute_stoi(clean, enhanced, sr),
         "sisdr": compute_sisdr(clean, enhanced), "snr": compute_snr(clean, enhanced),
         "segsnr": compute_segsnr(clean, enhanced), "lsd": compute_lsd(clean, enhanced, sr)}
    if noisy is not None:
        m["pesq_improvement"] = m["pesq"] - compute_pesq(clean, noisy, sr)
        m["stoi_improvement"] = m["stoi"] - compute_stoi(clean, noisy, sr)
        m["sisdr_improvement"] = m["sisdr"] - compute_sisdr(clean, noisy)
        m["snr_improvement"] = m["snr"] - compute_snr(clean, noisy)
    return m


def print_metrics(metrics: Dict[str, float], title: str = "Metrics") -> None:
    """reference metrics.py:352-365"""
    print(f"\n{title}")
    print("=" * 50)
    for k, v in metrics.items():
        print(f"{k:25s}: {v:8.4f}")
    print("=" * 50)
