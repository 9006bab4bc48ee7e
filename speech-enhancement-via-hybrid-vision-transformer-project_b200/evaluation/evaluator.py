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
from .metrics import compute_all_metrics


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

    def _score(self, noisy_audio: np.ndarray, clean_audio: np.ndarray, enhanced_audio: np.ndarray) -> Dict[str, float]:
        n = min(len(noisy_audio), len(clean_audio))
        return compute_all_metrics(clean=clean_audio[:n], enhanced=enhanced_audio[:n], noisy=noisy_audio[:n],
                                   sr=self.sample_rate)

    def evaluate_pair(self, noisy_path: Path, clean_path: Path) -> Dict[str, float]:
        """reference evaluator.py:119-155"""
        noisy, _ = load_audio(noisy_path, sr=self.sample_rate, mono=True)
        clean, _ = load_audio(clean_path, sr=self.sample_rate, mono=True)
        n = min(len(noisy), len(clean))
        noisy, clean = noisy[:n], clean[:n]
        return self._score(noisy, clean, self.enhance_audio(noisy))

    def evaluate_dataset(self, noisy_dir: Path, clean_dir: Path, output_dir: Optional[Path] = None,
                         save_enhanced: bool = False, batch_size: int = 64) -> Dict[str, object]:
        """reference evaluator.py:157-231 (same result dictionary), with the enhancement run in mixed-length batches."""
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
        # enhance in mixed-length batches (clips sorted by length; every clip is processed exactly as if alone)
        enhanced = {}
        order = sorted(range(len(pairs)), key=lambda i: (len(pairs[i][1]), i))
        for s in range(0, len(order), batch_size):
            idx = order[s:s + batch_size]
            outs = self._enh.enhance_varlen([pairs[i][1] for i in idx])
            for i, y in zip(idx, outs):
                enhanced[pairs[i][0]] = y
        all_metrics, per_file = [], {}
        for name, noisy, clean in pairs:
            m = self._score(noisy, clean, enhanced[name])
            per_file[name] = m
            all_metrics.append(m)
            if save_enhanced and output_dir:
                save_audio(output_dir / name, enhanced[name], self.sample_rate)
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
