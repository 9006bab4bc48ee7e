"""Spectrogram losses with the reference's constructors (reference training/losses.py:15-85 SpectrogramLoss, :88-141
STOILoss, :286-387 CombinedLoss, :390-410 create_loss_function), evaluated for validation: every term is a reduction of
(pred, target), done in ONE pass on the device by ``hvit_spec_loss`` (fp64 sums per sample) - no torch elementwise
temporaries of the [B, 1, 257, T] spectrograms.  Forward-only: these modules return detached scalars (there is no
backward on this path; the reference's PerceptualLoss placeholder is its L1 term, losses.py:279-283)."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch
import torch.nn as nn

from .. import _lib


def _sums(pred: torch.Tensor, target: torch.Tensor, use_log: bool) -> torch.Tensor:
    """[B, 5] fp64 on the device: sum |p'-t'|, sum (p'-t')^2, sum p^2, sum t^2, sum p t (p' = ln(p + 1e-8) if use_log)."""
    if pred.shape != target.shape:
        raise ValueError(f"pred {tuple(pred.shape)} and target {tuple(target.shape)} differ")
    if not pred.is_cuda:
        raise RuntimeError("the spectrogram losses run on the CUDA device (no CPU fallback): move pred / target to cuda")
    p = pred.detach().to(torch.float32).contiguous()
    t = target.detach().to(device=p.device, dtype=torch.float32).contiguous()
    B = p.shape[0]
    n_per = p.numel() // B
    out = torch.empty((B, 5), dtype=torch.float64, device=p.device)
    lib = _lib.load()
    _lib.check(lib.hvit_spec_loss(p.data_ptr(), t.data_ptr(), B, n_per, 1 if use_log else 0, out.data_ptr(),
                                  _lib.current_stream_ptr()), "hvit_spec_loss")
    return out


def _reduce(per_sample_sum: torch.Tensor, n_per: int, reduction: str) -> torch.Tensor:
    if reduction == "mean":
        return per_sample_sum.sum() / (per_sample_sum.numel() * n_per)
    if reduction == "sum":
        return per_sample_sum.sum()
    raise ValueError("reduction must be 'mean' or 'sum' on the device path (elementwise 'none' is not a reduction)")


def _stoi_proxy(s: torch.Tensor) -> torch.Tensor:
    """1 - cosine similarity per sample; F.normalize clamps each norm at 1e-12 (losses.py:126-135)."""
    den = s[:, 2].sqrt().clamp_min(1e-12) * s[:, 3].sqrt().clamp_min(1e-12)
    return 1.0 - s[:, 4] / den


class SpectrogramLoss(nn.Module):
    def __init__(self, loss_type: str = "l1", reduction: str = "mean", use_log_compression: bool = False):
        super().__init__()
        if loss_type not in ("l1", "mse", "l1+mse"):
            raise ValueError(f"Unknown loss type: {loss_type}")
        self.loss_type, self.reduction, self.use_log_compression = loss_type, reduction, use_log_compression

    def forward(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        s = _sums(pred, target, self.use_log_compression)
        n_per = pred.numel() // pred.shape[0]
        loss = 0.0
        if "l1" in self.loss_type:
            loss = loss + _reduce(s[:, 0], n_per, self.reduction)
        if "mse" in self.loss_type:
            loss = loss + _reduce(s[:, 1], n_per, self.reduction)
        return loss.to(torch.float32)


class STOILoss(nn.Module):
    """The reference's differentiable proxy: one minus the cosine similarity of the flattened spectrograms."""

    def __init__(self, reduction: str = "mean"):
        super().__init__()
        self.reduction = reduction

    def forward(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        loss = _stoi_proxy(_sums(pred, target, False))
        if self.reduction == "mean":
            return loss.mean().to(torch.float32)
        if self.reduction == "sum":
            return loss.sum().to(torch.float32)
        return loss.to(torch.float32)


class CombinedLoss(nn.Module):
    def __init__(self, l1_weight: float = 1.0, mse_weight: float = 0.0, stoi_weight: float = 0.1,
                 perceptual_weight: float = 0.0, use_log_compression: bool = False):
        super().__init__()
        self.l1_weight, self.mse_weight = l1_weight, mse_weight
        self.stoi_weight, self.perceptual_weight = stoi_weight, perceptual_weight
        self.use_log_compression = use_log_compression

    def forward(self, pred: torch.Tensor, target: torch.Tensor, return_components: bool = False):
        n = pred.numel()
        s = _sums(pred, target, self.use_log_compression)
        # STOI / perceptual terms see the RAW spectrograms (losses.py:362-371); their sums do not depend on use_log
        raw = s if not (self.use_log_compression and self.perceptual_weight > 0) else _sums(pred, target, False)
        losses: Dict[str, float] = {}
        total = torch.zeros((), dtype=torch.float64, device=pred.device)
        if self.l1_weight > 0:
            l1 = s[:, 0].sum() / n
            losses["l1"] = float(l1)
            total = total + self.l1_weight * l1
        if self.mse_weight > 0:
            mse = s[:, 1].sum() / n
            losses["mse"] = float(mse)
            total = total + self.mse_weight * mse
        if self.stoi_weight > 0:
            stoi = _stoi_proxy(s).mean()
            losses["stoi"] = float(stoi)
            total = total + self.stoi_weight * stoi
        if self.perceptual_weight > 0:   # the reference's placeholder: plain L1 of the raw spectrograms
            perceptual = raw[:, 0].sum() / n
            losses["perceptual"] = float(perceptual)
            total = total + self.perceptual_weight * perceptual
        total = total.to(torch.float32)
        losses["total"] = float(total)
        return (total, losses) if return_components else total


def create_loss_function(config: Dict) -> nn.Module:
    c = config.get("loss", {})
    return CombinedLoss(l1_weight=c.get("l1_weight", 1.0), mse_weight=c.get("mse_weight", 0.0),
                        stoi_weight=c.get("stoi_weight", 0.1), perceptual_weight=c.get("perceptual_weight", 0.0),
                        use_log_compression=c.get("use_log_compression", False))
