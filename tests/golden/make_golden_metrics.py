"""Generates tests/golden/metrics_v1.json by running the REFERENCE's own evaluation/metrics.py (loaded by file path;
librosa is shimmed with torch.stft as in make_golden.py) on seeded signals.  Run in the build container only:
    python tests/golden/make_golden_metrics.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def shim_librosa():
    lib = types.ModuleType("librosa")

    def stft(y, n_fft=512, hop_length=128, win_length=None, **kw):
        t = torch.from_numpy(np.asarray(y, dtype=np.float32))
        s = torch.stft(t, n_fft=n_fft, hop_length=hop_length, win_length=win_length or n_fft,
                       window=torch.hann_window(n_fft, periodic=True), center=True, pad_mode="constant",
                       return_complex=True)
        return s.numpy()
    lib.stft = stft
    sys.modules["librosa"] = lib


def main():
    shim_librosa()
    spec = importlib.util.spec_from_file_location("ref_metrics", "/root/reference/evaluation/metrics.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(123)
    cases = []
    for i, n in enumerate((16000, 12345, 40000)):
        t = np.arange(n) / 16000.0
        clean = (0.3 * np.sin(2 * np.pi * 220 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 2 * t))
                 + 0.01 * rng.standard_normal(n)).astype(np.float32)   # broadband floor: no bin at the fp32 noise level
        noisy = (clean + 0.1 * rng.standard_normal(n)).astype(np.float32)
        enh = (0.9 * clean + 0.02 * rng.standard_normal(n)).astype(np.float32)
        case = dict(n=n, seed=123, index=i,
                    sisdr=ref.compute_sisdr(clean, enh), snr=ref.compute_snr(clean, enh),
                    segsnr=ref.compute_segsnr(clean, enh), lsd=ref.compute_lsd(clean, enh),
                    sisdr_noisy=ref.compute_sisdr(clean, noisy), snr_noisy=ref.compute_snr(clean, noisy))
        cases.append(case)
    with open(os.path.join(HERE, "metrics_v1.json"), "w") as f:
        json.dump(dict(generator="tests/golden/make_golden_metrics.py", cases=cases), f, indent=1)
    print(json.dumps(cases, indent=1))


if __name__ == "__main__":
    main()
