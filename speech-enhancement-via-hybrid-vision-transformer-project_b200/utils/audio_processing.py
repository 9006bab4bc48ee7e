"""STFT front-end wrappers with the reference's signatures (reference utils/audio_processing.py:67-193), backed by
the CUDA STFT / iSTFT kernels (hvit_stft / hvit_istft in include/hvit.h).  numpy in, numpy out, float32/complex64.

Also ``load_audio`` / ``save_audio`` with the reference's signatures (reference audio_processing.py:15-64) on a
dependency-free RIFF/WAVE reader / writer (PCM 8/16/24/32-bit and IEEE float) standing in for librosa.load /
soundfile.write, which are not installed in this image.
"""
from __future__ import annotations

import struct
from pathlib import Path
from typing import Optional, Tuple

import numpy as np
import torch

from .. import _lib

_ONE_BITS = 0x3F800000


def _check_cfg(n_fft, hop_length, win_length, window, center):
    if (n_fft, hop_length, win_length, window, center) != (512, 128, 512, "hann", True):
        raise NotImplementedError("CUDA STFT supports n_fft=512, hop_length=128, win_length=512, window='hann', center=True")


def compute_stft(audio: np.ndarray, n_fft: int = 512, hop_length: int = 128, win_length: int = 512,
                 window: str = "hann", center: bool = True) -> np.ndarray:
    """Complex STFT [257, 1 + n//128], complex64 (reference audio_processing.py:67-98)."""
    _check_cfg(n_fft, hop_length, win_length, window, center)
    lib = _lib.load()
    x = torch.from_numpy(np.ascontiguousarray(np.asarray(audio, dtype=np.float32))).cuda()
    n = x.numel()
    T = 1 + n // 128
    spec = torch.empty((1, 257, T), dtype=torch.complex64, device="cuda")
    mag = torch.empty((1, 257, T), dtype=torch.float32, device="cuda")
    scal = torch.empty(2, dtype=torch.int32, device="cuda")
    _lib.check(lib.hvit_stft(x.data_ptr(), 1, n, 0, scal[0:1].data_ptr(), spec.data_ptr(), mag.data_ptr(),
                             scal[1:2].data_ptr(), _lib.current_stream_ptr()), "hvit_stft")
    return spec[0].cpu().numpy()


def compute_istft(stft: np.ndarray, hop_length: int = 128, win_length: int = 512, window: str = "hann",
                  center: bool = True, length: Optional[int] = None) -> np.ndarray:
    """Inverse STFT (reference audio_processing.py:101-132)."""
    _check_cfg(2 * (stft.shape[0] - 1), hop_length, win_length, window, center)
    lib = _lib.load()
    T = stft.shape[1]
    n = int(length) if length is not None else 128 * (T - 1)
    if 1 + n // 128 != T:
        raise ValueError(f"length={n} is inconsistent with {T} frames at hop 128")
    spec = torch.from_numpy(np.ascontiguousarray(stft.astype(np.complex64)))[None].cuda()
    mag = spec.abs().contiguous()
    ones = torch.full((2,), _ONE_BITS, dtype=torch.int32, device="cuda")
    frames = torch.empty((1, T, 512), dtype=torch.float32, device="cuda")
    out = torch.empty((1, n), dtype=torch.float32, device="cuda")
    _lib.check(lib.hvit_istft(mag.data_ptr(), spec.data_ptr(), ones[0:1].data_ptr(), ones[1:2].data_ptr(),
                              frames.data_ptr(), out.data_ptr(), 1, n, _lib.current_stream_ptr()), "hvit_istft")
    return out[0].cpu().numpy()


def normalize_audio(audio: np.ndarray, target_level: float = 1.0, eps: float = 1e-8) -> np.ndarray:
    """Peak normalisation (reference audio_processing.py:135-156); the fused path does this inside hvit_stft."""
    max_val = np.abs(audio).max()
    return audio * (target_level / max_val) if max_val > eps else audio


def compute_magnitude_phase(stft: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """reference audio_processing.py:159-174"""
    return np.abs(stft), np.angle(stft)


def reconstruct_from_magnitude_phase(magnitude: np.ndarray, phase: np.ndarray) -> np.ndarray:
    """reference audio_processing.py:177-193"""
    return magnitude * np.exp(1j * phase)


def _read_wav(path) -> Tuple[np.ndarray, int]:
    """RIFF/WAVE -> (float32 [frames, channels] in [-1, 1), sample rate).  Integer PCM of 8 / 16 / 24 / 32 bits is scaled
    by 1 / 2^(bits-1) and IEEE float (32 / 64 bits) is taken as is - the conversions libsndfile (behind librosa.load /
    soundfile) applies; WAVE_FORMAT_EXTENSIBLE is resolved through its sub-format.  (Python's ``wave`` module rejects
    float files, hence the hand-written chunk walk.)"""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, raw = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, rate, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == 0xFFFE and len(body) >= 26:      # WAVE_FORMAT_EXTENSIBLE: the real tag leads the sub-format GUID
                tag = struct.unpack("<H", body[24:26])[0]
            fmt = (tag, ch, rate, bits)
        elif cid == b"data":
            raw = body
        pos += 8 + size + (size & 1)
    if fmt is None or raw is None:
        raise ValueError(f"{path}: missing fmt / data chunk")
    tag, ch, rate, bits = fmt
    if tag == 1:
        if bits == 8:
            x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(raw[:len(raw) // 2 * 2], dtype="<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(raw[:len(raw) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = (v - ((v & 0x800000) << 1)).astype(np.float32) / 8388608.0
        elif bits == 32:
            x = (np.frombuffer(raw[:len(raw) // 4 * 4], dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            raise ValueError(f"{path}: unsupported PCM width {bits}")
    elif tag == 3:
        if bits == 32:
            x = np.frombuffer(raw[:len(raw) // 4 * 4], dtype="<f4").astype(np.float32)
        elif bits == 64:
            x = np.frombuffer(raw[:len(raw) // 8 * 8], dtype="<f8").astype(np.float32)
        else:
            raise ValueError(f"{path}: unsupported float width {bits}")
    else:
        raise ValueError(f"{path}: unsupported WAVE format tag {tag} (PCM and IEEE float are supported)")
    return x[:len(x) // ch * ch].reshape(-1, ch), rate


def load_audio(file_path, sr: int = 16000, mono: bool = True, offset: float = 0.0, duration: Optional[float] = None,
               res_type: Optional[str] = None) -> Tuple[np.ndarray, int]:
    """reference audio_processing.py:15-43 (= ``librosa.load(file_path, sr=sr, mono=mono, offset=..., duration=...)``):
    float32 waveform (mono: channel mean) and the sample rate.  WAV only (PCM 8/16/24/32-bit, IEEE float 32/64-bit).

    Resampling: librosa's default resampler (``soxr_hq``) is a third-party library that is not available here, and a
    different resampler would silently change the audio.  A file whose rate differs from ``sr`` therefore raises unless
    ``res_type`` names the resampler to use: ``"polyphase"`` = ``scipy.signal.resample_poly`` with the gcd-reduced ratio
    (exactly librosa's ``res_type="polyphase"``), ``"linear"`` = linear interpolation."""
    x, rate = _read_wav(file_path)
    if offset:
        x = x[int(round(offset * rate)):]
    if duration is not None:
        x = x[:int(round(duration * rate))]
    x = x.mean(axis=1) if mono else np.ascontiguousarray(x.T)
    if sr is not None and rate != sr:
        if res_type == "polyphase":
            from math import gcd
            from scipy.signal import resample_poly
            g = gcd(int(sr), int(rate))
            x = resample_poly(x, int(sr) // g, int(rate) // g, axis=-1)
        elif res_type == "linear":
            n_out = int(np.ceil(x.shape[-1] * sr / rate))
            t = np.arange(n_out) * (rate / sr)
            x = np.interp(t, np.arange(x.shape[-1]), x) if x.ndim == 1 else \
                np.stack([np.interp(t, np.arange(x.shape[-1]), c) for c in x])
        else:
            raise NotImplementedError(
                f"{file_path}: sample rate {rate} != {sr}. librosa's default resampler (soxr_hq) is not available; pass "
                "res_type='polyphase' (scipy.signal.resample_poly, librosa's 'polyphase') or 'linear', or resample offline")
        rate = sr
    return np.ascontiguousarray(x, dtype=np.float32), rate


_SUBTYPES = {"PCM_16": (1, 16), "PCM_24": (1, 24), "PCM_32": (1, 32), "FLOAT": (3, 32)}


def save_audio(file_path, audio: np.ndarray, sr: int = 16000, subtype: str = "PCM_16") -> None:
    """reference audio_processing.py:46-64 (= ``soundfile.write(file_path, audio, sr, subtype=subtype)``): mono [n] or
    [n, channels] float audio -> WAV.  Integer subtypes are scaled by 2^(bits-1) - 1 and rounded to nearest like
    libsndfile's float -> int conversion; samples outside [-1, 1] are clipped (libsndfile would wrap them)."""
    if subtype not in _SUBTYPES:
        raise ValueError(f"unsupported subtype {subtype!r}; supported: {sorted(_SUBTYPES)}")
    path = Path(file_path)
    path.parent.mkdir(parents=True, exist_ok=True)
    x = np.asarray(audio)
    if x.ndim == 1:
        x = x[:, None]
    ch = x.shape[1]
    tag, bits = _SUBTYPES[subtype]
    if tag == 3:
        payload = np.ascontiguousarray(x, dtype="<f4").tobytes()
    else:
        full = float(2 ** (bits - 1) - 1)
        q = np.rint(np.clip(x.astype(np.float64), -1.0, 1.0) * full)
        if bits == 16:
            payload = q.astype("<i2").tobytes()
        elif bits == 32:
            payload = q.astype("<i4").tobytes()
        else:
            v = q.astype(np.int32).reshape(-1)
            b = np.empty((v.size, 3), dtype=np.uint8)
            b[:, 0], b[:, 1], b[:, 2] = v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF
            payload = b.tobytes()
    block = ch * bits // 8
    fmt = struct.pack("<HHIIHH", tag, ch, int(sr), int(sr) * block, block, bits)
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt
    if tag == 3:
        chunks += b"fact" + struct.pack("<II", 4, x.shape[0])
    chunks += b"data" + struct.pack("<I", len(payload)) + payload + (b"\x00" if len(payload) & 1 else b"")
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks)
