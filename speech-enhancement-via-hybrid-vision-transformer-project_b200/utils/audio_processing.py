"""STFT front-end wrappers with the reference's signatures (reference utils/audio_processing.py:67-193), backed by
the CUDA STFT / iSTFT kernels (hvit_stft / hvit_istft in include/hvit.h).  numpy in, numpy out, float32/complex64.

Also a dependency-free PCM WAV reader/writer standing in for librosa.load / soundfile.write
(reference audio_processing.py:15-64), which are not installed in this image.
"""
from __future__ import annotations

import wave as _wave
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


def load_audio(file_path, sr: int = 16000, mono: bool = True) -> Tuple[np.ndarray, int]:
    """PCM WAV -> float32 mono in [-1, 1]; linear resampling when the file rate differs."""
    with _wave.open(str(file_path), "rb") as f:
        ch, width, rate, frames = f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()
        raw = f.readframes(frames)
    if width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"unsupported WAV sample width {width}")
    x = x.reshape(-1, ch)
    x = x.mean(axis=1) if mono else x.T
    if rate != sr:
        n_out = int(round(x.shape[-1] * sr / rate))
        x = np.interp(np.arange(n_out) * (rate / sr), np.arange(x.shape[-1]), x).astype(np.float32)
    return x.astype(np.float32), sr


def save_audio(audio: np.ndarray, file_path, sr: int = 16000, subtype: str = "PCM_16") -> None:
    path = Path(file_path)
    path.parent.mkdir(parents=True, exist_ok=True)
    pcm = (np.clip(np.asarray(audio, dtype=np.float32), -1.0, 1.0) * 32767.0).astype("<i2")
    with _wave.open(str(path), "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(sr)
        f.writeframes(pcm.tobytes())
