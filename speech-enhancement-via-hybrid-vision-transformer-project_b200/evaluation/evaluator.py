"""Evaluator with the reference's constructor and methods (reference evaluation/evaluator.py:20-290).

``enhance_audio`` is the same data flow as ``AudioEnhancer.enhance`` (the reference duplicates that code,
evaluator.py:54-117) and runs through the same CUDA plan; ``evaluate_dataset`` additionally batches the directory:
all noisy files are enhanced first in mixed-length batches (``AudioEnhancer.enhance_varlen``, in memory), then scored
(SI-SDR / SNR / segmental SNR / LSD).  PCM WAV decoding replaces librosa.load / soundfile (neither is a dependency here)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, Optional

import numpy as np
import torch.nn as nn

from ..inference.enhancer import AudioEnhancer
from ..utils.audio_processing import load_audio, save_audio
from .metrics import compute_metrics_device, compute_pesq, compute_stoi


class Evaluator:
    def __init__(self, model: nn.Module, device: str = "cuda", sample_rate: int = 16000, n_fft: int = 512,
                 hop_length: int = 128, win_length: int = 512):
        self._enh = AudioEnhancer(model, device=device, sample_rate=sample_rate, n_fft=n_fft, hop_length=hop_length,
                                  win_length=win_length)
        self.model = self._enh.model
        self.device = device
        self.sample_rate = sample_rate
        self.n_fft, self.hop_length, self.win_length = n_fft, hop_length, win_length

    def enhance_audio(self, noisy_audio: np.ndarray) -> np.ndarray:
        """reference evaluator.py:54-117"""
        return self._enh.enhance(noisy_audio, normalize=True)

    def _score_batch(self, noisy, clean, enhanced_host, enhanced_dev, lens):
        """Metric dictionaries (the reference's keys, metrics.py:299-349) for one enhanced batch.  SI-SDR / SNR / segmental
        SNR / LSD are reduced ON THE DEVICE from the enhancer's output buffer (``hvit_metrics``); only the clean and noisy
        references are uploaded, and 4 doubles per clip come back.  PESQ / STOI are third-party host packages in the
        reference too (0.0 when not installed)."""
        import torch
        dev = enhanced_dev.device
        B, n_max = enhanced_dev.shape

        def padded(xs):
            buf = np.zeros((B, n_max), dtype=np.float32)
            for i, x in enumerate(xs):
                buf[i, :lens[i]] = x[:lens[i]]
            return torch.from_numpy(buf).to(dev)

        d_clean = padded(clean)
        enh = compute_metrics_device(d_clean, enhanced_dev, lens)
        noi = compute_metrics_device(d_clean, padded(noisy), lens)
        out = []
        for i in range(B):
            c, e, x = clean[i][:lens[i]], enhanced_host[i][:lens[i]], noisy[i][:lens[i]]
            m = {"pesq": compute_pesq(c, e, self.sample_rate), "stoi": compute_stoi(c, e, self.sample_rate),
                 "sisdr": float(enh["sisdr"][i]), "snr": float(enh["snr"][i]), "segsnr": float(enh["segsnr"][i]),
                 "lsd": float(enh["lsd"][i])}
            m["pesq_improvement"] = m["pesq"] - compute_pesq(c, x, self.sample_rate)
            m["stoi_improvement"] = m["stoi"] - compute_stoi(c, x, self.sample_rate)
            m["sisdr_improvement"] = m["sisdr"] - float(noi["sisdr"][i])
            m["snr_improvement"] = m["snr"] - float(noi["snr"][i])
            out.append(m)
        return out

    def evaluate_pair(self, noisy_path: Path, clean_path: Path) -> Dict[str, float]:
        """reference evaluator.py:119-155"""
        noisy, _ = load_audio(noisy_path, sr=self.sample_rate, mono=True)
        clean, _ = load_audio(clean_path, sr=self.sample_rate, mono=True)
        n = min(len(noisy), len(clean))
        noisy, clean = noisy[:n], clean[:n]
        host, dev, lens = self._enh.enhance_varlen([noisy], return_device=True)
        return self._score_batch([noisy], [clean], host, dev, lens)[0]

    def evaluate_dataset(self, noisy_dir: Path, clean_dir: Path, output_dir: Optional[Path] = None,
                         save_enhanced: bool = False, batch_size: int = 64) -> Dict[str, object]:
        """reference evaluator.py:157-231 (same result dictionary), with the enhancement run in mixed-length batches and
        the signal metrics computed on the device from the enhancer's output buffer."""
        noisy_files = sorted(Path(noisy_dir).glob("*.wav"))
        if len(noisy_files) == 0:
            raise ValueError(f"No .wav files found in {noisy_dir}")
        print(f"Evaluating {len(noisy_files)} audio files...")
        if save_enhanced and output_dir:
            output_dir = Path(output_dir)
            output_dir.mkdir(parents=True, exist_ok=True)
        pairs = []
        for noisy_path in noisy_files:
            clean_path = Path(clean_dir) / noisy_path.name
            if not clean_path.exists():
                print(f"Warning: No clean file found for {noisy_path.name}")
                continue
            noisy, _ = load_audio(noisy_path, sr=self.sample_rate, mono=True)
            clean, _ = load_audio(clean_path, sr=self.sample_rate, mono=True)
            n = min(len(noisy), len(clean))
            pairs.append((noisy_path.name, noisy[:n], clean[:n]))
        # mixed-length batches (clips sorted by length; every clip is processed exactly as if alone)
        per_file = {}
        order = sorted(range(len(pairs)), key=lambda i: (len(pairs[i][1]), i))
        for s in range(0, len(order), batch_size):
            idx = order[s:s + batch_size]
            noisy_b, clean_b = [pairs[i][1] for i in idx], [pairs[i][2] for i in idx]
            host, dev, lens = self._enh.enhance_varlen(noisy_b, return_device=True)
            for i, m, y in zip(idx, self._score_batch(noisy_b, clean_b, host, dev, lens), host):
                per_file[pairs[i][0]] = m
                if save_enhanced and output_dir:
                    save_audio(output_dir / pairs[i][0], y, self.sample_rate)
        all_metrics = [per_file[name] for name, _, _ in pairs]
        per_file = {name: per_file[name] for name, _, _ in pairs}
        average = {}
        if all_metrics:
            for k in all_metrics[0].keys():
                vals = [m[k] for m in all_metrics]
                average[k] = np.mean(vals)
                average[f"{k}_std"] = np.std(vals)
        return {"per_file_metrics": per_file, "average_metrics": average, "num_files": len(all_metrics)}

    def save_results(self, results: Dict, output_path: Path) -> None:
        """reference evaluator.py:233-263"""
        output_path = Path(output_path)
        output_path.parent.mkdir(parents=True, exist_ok=True)
        ser = {"num_files": results["num_files"],
               "average_metrics": {k: float(v) for k, v in results["average_metrics"].items()},
               "per_file_metrics": {f: {k: float(v) for k, v in m.items()} for f, m in results["per_file_metrics"].items()}}
        with open(output_path, "w") as f:
            json.dump(ser, f, indent=2)
        print(f"Results saved to {output_path}")

    def print_results(self, results: Dict) -> None:
        """reference evaluator.py:265-290"""
        print(f"\nEvaluation Results ({results['num_files']} files)")
        print("=" * 70)
        avg = results["average_metrics"]
        print("\nMain Metrics:")
        print("-" * 70)
        for metric in ("pesq", "stoi", "sisdr", "snr"):
            if metric in avg:
                print(f"{metric.upper():15s}: {avg[metric]:7.4f} ± {avg.get(f'{metric}_std', 0.0):6.4f}")
        print("\nImprovements over the noisy input:")
        print("-" * 70)
        for metric in ("pesq_improvement", "stoi_improvement", "sisdr_improvement", "snr_improvement"):
            if metric in avg:
                print(f"{metric.upper():20s}: {avg[metric]:7.4f} ± {avg.get(f'{metric}_std', 0.0):6.4f}")
        print("=" * 70)
