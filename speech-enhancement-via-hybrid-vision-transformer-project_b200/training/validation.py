"""``Trainer.validate`` re-hosted on the inference plan (reference training/trainer.py:207-251): for every batch of the
validation loader, ``enhanced_spec = model(noisy_spec)`` (``HybridViT.forward`` -> ``hvit_forward``), the criterion on
(enhanced_spec, clean_spec), and the mean of the per-batch losses.  The reference wraps this in ``torch.no_grad`` and
``model.eval()``; the CUDA plan is inference-only, so both hold by construction.  Batches are dictionaries with
``noisy_spec`` / ``clean_spec`` [B, 1, F, T] (data/dataset.py:297-347 collate)."""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import torch
import torch.nn as nn

from .losses import CombinedLoss


class Validator:
    def __init__(self, model: nn.Module, val_loader: Optional[Iterable], criterion: Optional[nn.Module] = None,
                 device: str = "cuda"):
        self.model = model.to(device).eval()
        self.val_loader = val_loader
        self.criterion = criterion if criterion is not None else CombinedLoss()
        self.device = device

    @torch.no_grad()
    def validate(self) -> Dict[str, float]:
        if self.val_loader is None:
            return {}
        total_loss, num_batches = 0.0, 0
        for batch in self.val_loader:
            noisy_spec = batch["noisy_spec"].to(self.device, non_blocking=True)
            clean_spec = batch["clean_spec"].to(self.device, non_blocking=True)
            enhanced_spec = self.model(noisy_spec)
            total_loss += float(self.criterion(enhanced_spec, clean_spec))
            num_batches += 1
        if num_batches == 0:
            raise ZeroDivisionError("validation loader yielded no batches")  # the reference divides by zero here too
        return {"loss": total_loss / num_batches}


def validate(model: nn.Module, val_loader: Iterable, criterion: Optional[nn.Module] = None, device: str = "cuda"):
    """Functional form of ``Validator(...).validate()``."""
    return Validator(model, val_loader, criterion, device).validate()
