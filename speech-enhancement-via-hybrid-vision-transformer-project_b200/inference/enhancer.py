"""AudioEnhancer with the reference's constructor and ``enhance`` signature
(reference inference/enhancer.py:18-135), running entirely on the GPU.

Reference data flow (host numpy, one H2D/D2H of the spectrogram around the model):
    peak-normalise -> librosa.stft -> abs/angle -> max-normalise -> model -> * mag_max
    -> * exp(1j*phase) -> librosa.istft -> * max_val
Here only the waveform crosses PCIe; every stage above is a CUDA kernel inside
``hvit_enhance`` (include/hvit.h), enqueued on one stream.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Optional, Sequence, Union

import numpy as np
import torch
import torch.nn as nn

from .. import _lib

HOP = 128
N_FFT = 512


class AudioEnhancer:
    """Audio enhancement inference engine (reference enhancer.py:18-53)."""

    def __init__(self, model: nn.Module, device: str = "cuda", sample_rate: int = 16000, n_fft: int = 512,
                 hop_length: int = 128, win_length: int = 512, window: str = "hann"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("hvit_b200.AudioEnhancer runs on a CUDA (sm_100) device only; there is no CPU path")
        if (n_fft, hop_length, win_length, window) != (N_FFT, HOP, N_FFT, "hann"):
            raise NotImplementedError(
                "the CUDA STFT/iSTFT kernels implement the configuration of config/model_config.yaml "
                "(n_fft=512, hop_length=128, win_length=512, window='hann')")
        if not hasattr(model, "plan_for"):
            raise TypeError("model must be an hvit_b200.HybridViT")
        self.model = model.to(dev).eval()
        self.device = device
        self._dev = dev
        self.sample_rate = sample_rate
        self.n_fft, self.hop_length, self.win_length, self.window = n_fft, hop_length, win_length, window
        self._pinned = {}
        self._pipes = {}

    # ------------------------------------------------------------------ helpers
    def _staging(self, B: int, n: int):
        key = (B, n)
        buf = self._pinned.get(key)
        if buf is None:
            if len(self._pinned) >= 8:
                self._pinned.pop(next(iter(self._pinned)))
            buf = (torch.empty((B, n), dtype=torch.float32).pin_memory(),
                   torch.empty((B, n), dtype=torch.float32).pin_memory(),
                   torch.empty((B, n), dtype=torch.float32, device=self._dev),
                   torch.empty((B, n), dtype=torch.float32, device=self._dev))
            self._pinned[key] = buf
        return buf

    def enhance_device(self, wave: torch.Tensor, normalize: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Device-resident batch API (extension): ``wave`` fp32 CUDA [B, n] -> enhanced fp32 CUDA [B, n]."""
        if wave.dim() != 2 or not wave.is_cuda or wave.dtype != torch.float32:
            raise ValueError("enhance_device expects a float32 CUDA tensor [B, n]")
        wave = wave.contiguous()
        B, n = wave.shape
        if n < 1:
            raise ValueError("empty audio")
        T = 1 + n // HOP
        plan = self.model.plan_for(B, N_FFT // 2 + 1, T, n_samples=n)
        if out is None:
            out = torch.empty_like(wave)
        with torch.cuda.device(self._dev):
            _lib.check(plan.lib.hvit_enhance(plan.handle, wave.data_ptr(), out.data_ptr(), 1 if normalize else 0,
                                             _lib.current_stream_ptr()), "hvit_enhance")
        return out

    def _pipeline(self, B: int, n: int):
        """Two-slot copy/compute pipeline for the host-to-host path: device staging buffers, one H2D and one D2H
        stream and the events that order them against the compute (current) stream."""
        key = (B, n)
        pipe = self._pipes.get(key)
        if pipe is None:
            if len(self._pipes) >= 4:
                self._pipes.pop(next(iter(self._pipes)))
            with torch.cuda.device(self._dev):
                pipe = dict(
                    slot=0,
                    d_in=[torch.empty((B, n), dtype=torch.float32, device=self._dev) for _ in range(2)],
                    d_out=[torch.empty((B, n), dtype=torch.float32, device=self._dev) for _ in range(2)],
                    s_in=torch.cuda.Stream(device=self._dev), s_out=torch.cuda.Stream(device=self._dev),
                    ev_in=[torch.cuda.Event() for _ in range(2)],      # H2D of the slot landed
                    ev_comp=[torch.cuda.Event() for _ in range(2)],    # enhance of the slot finished
                    ev_out=[torch.cuda.Event() for _ in range(2)])     # D2H of the slot finished
            self._pipes[key] = pipe
        return pipe

    def enhance_pinned(self, pinned_in: torch.Tensor, pinned_out: torch.Tensor, normalize: bool = True,
                       synchronize: bool = True) -> torch.Tensor:
        """Host-to-host batch API on caller-owned pinned buffers [B, n] fp32 (what ``enhance_batch`` does after
        staging the numpy input).  The H2D copy, the kernels and the D2H copy run on three streams over two device
        slots, so consecutive calls overlap: copy-in of batch i+1 and copy-out of batch i-1 hide behind the compute
        of batch i.  With ``synchronize=False`` the call only enqueues; call :meth:`join` (or pass
        ``synchronize=True`` on the last batch) before reading ``pinned_out``."""
        B, n = pinned_in.shape
        pipe = self._pipeline(B, n)
        k = pipe["slot"]
        pipe["slot"] = k ^ 1
        with torch.cuda.device(self._dev):
            cur = torch.cuda.current_stream()
            s_in, s_out = pipe["s_in"], pipe["s_out"]
            s_in.wait_event(pipe["ev_comp"][k])          # the kernels that last read this input slot are done
            with torch.cuda.stream(s_in):
                pipe["d_in"][k].copy_(pinned_in, non_blocking=True)
                pipe["ev_in"][k].record(s_in)
            cur.wait_event(pipe["ev_in"][k])
            cur.wait_event(pipe["ev_out"][k])            # the D2H that last read this output slot is done
            self._enhance_slot(pipe, k, normalize)
            pipe["ev_comp"][k].record(cur)
            s_out.wait_event(pipe["ev_comp"][k])
            with torch.cuda.stream(s_out):
                pinned_out.copy_(pipe["d_out"][k], non_blocking=True)
                pipe["ev_out"][k].record(s_out)
            if synchronize:
                pipe["ev_out"][k].synchronize()
        return pinned_out

    def _enhance_slot(self, pipe, k: int, normalize: bool) -> None:
        """Kernels of one pipeline slot.  The slot's device buffers are fixed, so after two eager calls the ~64 launches
        of ``hvit_enhance`` are replayed as one CUDA graph (the plan only enqueues kernels - no allocation, no sync -
        and is capturable, PDL edges included): single-clip latency is launch-bound otherwise.
        HVIT_NO_GRAPH=1 keeps the eager path."""
        B, n = pipe["d_in"][k].shape
        # the graph bakes in raw pointers into the plan's workspace and packed weights: it is keyed by everything that
        # selects a plan (precision included), holds a strong reference to the plan it captured, and is only replayed
        # while that very plan is still the model's current one for the shape (weights unchanged, not evicted + rebuilt)
        plan = self.model.plan_for(B, N_FFT // 2 + 1, 1 + n // HOP, n_samples=n)
        key = (k, bool(normalize), self.model.precision)
        graphs = pipe.setdefault("graphs", {})
        calls = pipe.setdefault("calls", {})
        g = graphs.get(key)
        if g is not None and g[1] is not plan:
            del graphs[key]
            calls.pop(key, None)
            g = None
        if g is None:
            calls[key] = calls.get(key, 0) + 1
            # graphs only pay off while the step is launch-bound (measured: batch 1 x 4 s 0.75 -> 0.69 ms p50, but a
            # GPU-bound batch of 64 loses 7 % to the graph launch); large batches stay eager
            small = pipe["d_in"][k].numel() <= 8 * 160000
            if calls[key] <= 2 or not small or os.environ.get("HVIT_NO_GRAPH") == "1" or \
                    torch.cuda.is_current_stream_capturing():
                self.enhance_device(pipe["d_in"][k], normalize=normalize, out=pipe["d_out"][k])
                return
            cur = torch.cuda.current_stream()
            cap = torch.cuda.Stream(device=self._dev)
            cap.wait_stream(cur)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=cap):
                self.enhance_device(pipe["d_in"][k], normalize=normalize, out=pipe["d_out"][k])
            cur.wait_stream(cap)
            g = graphs[key] = (graph, plan)
        g[0].replay()

    def join(self, block: bool = False) -> None:
        """Order the current stream after every outstanding D2H copy of :meth:`enhance_pinned`
        (``block=True`` additionally waits on the host)."""
        with torch.cuda.device(self._dev):
            cur = torch.cuda.current_stream()
            for pipe in self._pipes.values():
                for ev in pipe["ev_out"]:
                    cur.wait_event(ev)
                    if block:
                        ev.synchronize()

    @torch.no_grad()
    def enhance_batch(self, noisy_audio: Union[np.ndarray, Sequence[np.ndarray]], normalize: bool = True) -> np.ndarray:
        """Batched extension of :meth:`enhance`: equal-length clips [B, n] (numpy) -> [B, n] float32."""
        x = np.ascontiguousarray(np.asarray(noisy_audio, dtype=np.float32))
        if x.ndim != 2:
            raise ValueError("enhance_batch expects [B, n]")
        B, n = x.shape
        pin_in, pin_out, d_in, d_out = self._staging(B, n)
        pin_in.numpy()[...] = x
        self.enhance_pinned(pin_in, pin_out, normalize=normalize, synchronize=True)
        return pin_out.numpy().copy()

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def enhance(self, noisy_audio: np.ndarray, normalize: bool = True) -> np.ndarray:
        """Enhance one noisy waveform (reference enhancer.py:55-135).  The result has the input's floating dtype like the
        reference's (float32 in -> float32 out, float64 in -> float64 out); the arithmetic itself is the plan's
        (fp32 / fp16 on the device) - a float64 input is rounded to float32 first."""
        x = np.asarray(noisy_audio)
        if x.ndim != 1:
            raise ValueError(f"expected a 1-D waveform, got shape {x.shape}")
        if x.size == 0:
            raise ValueError("zero-size array to reduction operation maximum which has no identity")
        y = self.enhance_batch(x[None, :], normalize=normalize)[0]
        return y.astype(np.float64) if x.dtype == np.float64 else y

    def enhance_file(self, input_path, output_path, normalize: bool = True) -> None:
        """reference enhancer.py:137-162 - 16-bit / float PCM WAV I/O without librosa/soundfile."""
        from ..utils.audio_processing import load_audio, save_audio
        audio, _ = load_audio(input_path, sr=self.sample_rate)
        out = self.enhance(audio, normalize=normalize)
        save_audio(output_path, out, self.sample_rate)
        print(f"Enhanced audio saved to {output_path}")

    def enhance_varlen_device(self, wave: torch.Tensor, n_valid: torch.Tensor, normalize: bool = True,
                              out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Device-resident mixed-length batch: ``wave`` fp32 CUDA [B, n_max] zero-padded clips, ``n_valid`` int32 CUDA [B]
        true lengths -> enhanced fp32 CUDA [B, n_max] (zero beyond each clip's length).  See :meth:`enhance_varlen`."""
        if wave.dim() != 2 or not wave.is_cuda or wave.dtype != torch.float32:
            raise ValueError("enhance_varlen_device expects a float32 CUDA tensor [B, n]")
        if n_valid.dtype != torch.int32 or not n_valid.is_cuda or n_valid.numel() != wave.shape[0]:
            raise ValueError("n_valid must be an int32 CUDA tensor with one length per clip")
        wave = wave.contiguous()
        B, n = wave.shape
        plan = self.model.plan_for(B, N_FFT // 2 + 1, 1 + n // HOP, n_samples=n)
        if out is None:
            out = torch.empty_like(wave)
        with torch.cuda.device(self._dev):
            _lib.check(plan.lib.hvit_enhance_varlen(plan.handle, wave.data_ptr(), out.data_ptr(), n_valid.data_ptr(),
                                                    1 if normalize else 0, _lib.current_stream_ptr()), "hvit_enhance_varlen")
        return out

    @torch.no_grad()
    def enhance_varlen(self, clips: Sequence[np.ndarray], normalize: bool = True, pad_multiple: int = 8000,
                       return_device: bool = False):
        """Mixed-length batch: clips of different lengths share ONE batch and every result
        equals ``enhance(clip)`` of that clip alone.  The clips are zero-padded to a common length (rounded up to
        ``pad_multiple`` samples so that a stream of batches reuses a small set of plans) and ``hvit_enhance_varlen`` is
        given their true lengths: per-clip frame counts, right-border zero padding in every convolution, per-clip token
        order / positional rows / key mask in attention, per-clip bilinear resizes and iSTFT length - all on the device
        (reference plumbing that this completes: attention.py:94-98 mask, dataset.py:297-347 zero-pad collate)."""
        xs = [np.ascontiguousarray(np.asarray(c, dtype=np.float32)) for c in clips]
        if not xs:
            return []
        if any(x.ndim != 1 for x in xs):
            raise ValueError("enhance_varlen expects a sequence of 1-D waveforms")
        lens = [int(x.shape[0]) for x in xs]
        B = len(xs)
        n_max = max(1, -(-max(lens) // pad_multiple) * pad_multiple)
        plan = self.model.plan_for(B, N_FFT // 2 + 1, 1 + n_max // HOP, n_samples=n_max)
        n_min = plan.lib.hvit_varlen_min_samples(plan.handle)
        if min(lens) < n_min:
            raise ValueError(f"clips must have at least {n_min} samples (one 4x4 patch after the encoder), got {min(lens)}")
        pin_in, pin_out, d_in, d_out = self._staging(B, n_max)
        buf = pin_in.numpy()
        buf[...] = 0.0
        for i, x in enumerate(xs):
            buf[i, :lens[i]] = x
        with torch.cuda.device(self._dev):
            n_valid = torch.tensor(lens, dtype=torch.int32).to(self._dev, non_blocking=False)
            d_in.copy_(pin_in, non_blocking=True)
            _lib.check(plan.lib.hvit_enhance_varlen(plan.handle, d_in.data_ptr(), d_out.data_ptr(), n_valid.data_ptr(),
                                                    1 if normalize else 0, _lib.current_stream_ptr()), "hvit_enhance_varlen")
            pin_out.copy_(d_out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        out = pin_out.numpy()
        host = [out[i, :lens[i]].copy() for i in range(B)]
        if return_device:   # (+ the padded device batch [B, n_max] - valid until the next call with this shape - and lengths)
            return host, d_out, lens
        return host

    def enhance_files(self, input_paths: Sequence, output_paths: Sequence, normalize: bool = True,
                      batch_size: int = 64, pad_multiple: int = 8000) -> None:
        """Batched file enhancement.  The reference enhances one file at a time; here the
        decoded clips are sorted by length and consecutive runs of up to ``batch_size`` clips share a batch through
        :meth:`enhance_varlen` (mixed lengths, each clip processed exactly as if alone), so a directory of arbitrary
        lengths runs in ceil(files / batch_size) launches of the plan instead of one per file - and the number of
        distinct plans is bounded by the number of distinct (batch size, padded length) pairs.  Results are written as
        16-bit PCM like ``enhance_file``."""
        from ..utils.audio_processing import load_audio, save_audio
        if len(input_paths) != len(output_paths):
            raise ValueError("input_paths and output_paths must have the same length")
        items = []
        for i, path in enumerate(input_paths):
            audio, _ = load_audio(path, sr=self.sample_rate)
            if audio.size == 0:
                raise ValueError(f"{path}: empty audio")
            items.append((audio.shape[0], i, audio))
        items.sort(key=lambda t: (t[0], t[1]))
        for start in range(0, len(items), batch_size):
            chunk = items[start:start + batch_size]
            outs = self.enhance_varlen([a for _, _, a in chunk], normalize=normalize, pad_multiple=pad_multiple)
            for (_, i, _), y in zip(chunk, outs):
                save_audio(output_paths[i], y, self.sample_rate)

    def enhance_directory(self, input_dir, output_dir, extension: str = ".wav", normalize: bool = True,
                          batch_size: int = 64) -> None:
        """reference enhancer.py:164-194 (same messages, same outputs), batched through :meth:`enhance_files`."""
        src, dst = Path(input_dir), Path(output_dir)
        dst.mkdir(parents=True, exist_ok=True)
        files = sorted(src.glob(f"*{extension}"))
        print(f"Found {len(files)} audio files to enhance")
        self.enhance_files(files, [dst / f.name for f in files], normalize=normalize, batch_size=batch_size)
        for f in files:
            print(f"Enhanced audio saved to {dst / f.name}")
        print(f"All files enhanced and saved to {dst}")


def enhance_audio(noisy_audio: np.ndarray, model: nn.Module, device: str = "cuda", sample_rate: int = 16000,
                  n_fft: int = 512, hop_length: int = 128) -> np.ndarray:
    """Convenience wrapper (reference enhancer.py:197-229)."""
    return AudioEnhancer(model=model, device=device, sample_rate=sample_rate, n_fft=n_fft,
                         hop_length=hop_length).enhance(noisy_audio)


def load_model_for_inference(checkpoint_path, model: nn.Module, device: str = "cuda") -> nn.Module:
    """reference enhancer.py:258-290"""
    ckpt = torch.load(checkpoint_path, map_location=device)
    model.load_state_dict(ckpt["model_state_dict"] if "model_state_dict" in ckpt else ckpt)
    return model.to(device).eval()
