"""Weights enter the hot path through ``load_state_dict`` exactly as in the reference
(reference utils/checkpoint.py:127-161): a checkpoint is either ``{'model_state_dict': ...}`` or a bare state_dict."""
from __future__ import annotations

from pathlib import Path

import torch
import torch.nn as nn


def load_model_weights(filepath: str, model: nn.Module, device: str = "cpu", strict: bool = True) -> nn.Module:
    path = Path(filepath)
    if not path.exists():
        raise FileNotFoundError(f"Checkpoint not found: {path}")
    ckpt = torch.load(path, map_location=device)
    state = ckpt["model_state_dict"] if isinstance(ckpt, dict) and "model_state_dict" in ckpt else ckpt
    model.load_state_dict(state, strict=strict)
    return model.to(device)
