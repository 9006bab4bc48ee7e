"""Second caller of the enhance path: the reference's Evaluator re-hosted on the CUDA
pipeline, plus the waveform metrics that need no third-party package."""
from .evaluator import Evaluator
from .metrics import (compute_all_metrics, compute_lsd, compute_pesq, compute_segsnr, compute_sisdr, compute_snr,
                      compute_stoi, print_metrics)

__all__ = ["Evaluator", "compute_all_metrics", "compute_lsd", "compute_pesq", "compute_segsnr", "compute_sisdr",
           "compute_snr", "compute_stoi", "print_metrics"]
