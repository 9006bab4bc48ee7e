"""Parameter containers for the CNN / embedding blocks of HybridViT.

These classes exist so that ``HybridViT.state_dict()`` has exactly the reference's keys
(reference models/components.py:15-386) and so that the reference's literal random
initialisation is reproduced under the same ``torch.manual_seed``.  They do NOT compute:
the arithmetic of every block is fused into the CUDA launch plan driven by
``HybridViT.forward`` (csrc/).  Calling a block directly raises.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn


class _FusedBlock(nn.Module):
    def forward(self, *args, **kwargs):  # pragma: no cover - guard
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container; its arithmetic runs inside the fused CUDA plan of "
            "HybridViT.forward (hvit_b200 has no eager / CPU path).")


def _activation(name: str) -> nn.Module:
    table = {"relu": lambda: nn.ReLU(inplace=True), "gelu": nn.GELU, "leaky_relu": lambda: nn.LeakyReLU(0.2, inplace=True)}
    if name not in table:
        raise ValueError(f"Unknown activation: {name}")  # same error as reference components.py:77
    return table[name]()


class ConvBlock(_FusedBlock):
    """Conv2d -> BatchNorm2d -> activation -> [Dropout2d] -> [MaxPool2d]  (reference components.py:15-99)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, stride: int = 1, padding: int = 1,
                 pool_size: Optional[int] = 2, activation: str = "relu", use_batchnorm: bool = True,
                 dropout: float = 0.0):
        super().__init__()
        seq = [nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                         bias=not use_batchnorm)]
        if use_batchnorm:
            seq.append(nn.BatchNorm2d(out_channels))
        seq.append(_activation(activation))
        if dropout > 0:
            seq.append(nn.Dropout2d(dropout))
        if pool_size is not None and pool_size > 1:
            seq.append(nn.MaxPool2d(kernel_size=pool_size))
        self.block = nn.Sequential(*seq)
        self.meta = dict(kernel_size=kernel_size, stride=stride, padding=padding, pool=pool_size or 1,
                         activation=activation, batchnorm=use_batchnorm)


class TransposeConvBlock(_FusedBlock):
    """[Upsample(nearest)] -> Conv2d -> [BatchNorm2d] -> ReLU | Tanh -> [Dropout2d]  (reference components.py:102-192)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, stride: int = 1, padding: int = 1,
                 output_padding: int = 0, upsample_factor: Optional[int] = 2, activation: str = "relu",
                 use_batchnorm: bool = True, dropout: float = 0.0, final_layer: bool = False):
        super().__init__()
        seq = []
        if upsample_factor is not None and upsample_factor > 1:
            seq.append(nn.Upsample(scale_factor=upsample_factor, mode="nearest"))
        seq.append(nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                             bias=not use_batchnorm))
        if use_batchnorm and not final_layer:
            seq.append(nn.BatchNorm2d(out_channels))
        if final_layer:
            seq.append(nn.Tanh())
        elif activation in ("relu", "gelu", "leaky_relu"):
            seq.append(_activation(activation))
        if dropout > 0 and not final_layer:
            seq.append(nn.Dropout2d(dropout))
        self.block = nn.Sequential(*seq)
        self.meta = dict(kernel_size=kernel_size, stride=stride, padding=padding, up=upsample_factor or 1,
                         activation=activation, batchnorm=use_batchnorm, final=final_layer)


class FeedForward(_FusedBlock):
    """Linear -> GELU(erf) -> Dropout -> Linear -> Dropout  (reference components.py:195-241)."""

    def __init__(self, dim: int, hidden_dim: Optional[int] = None, dropout: float = 0.0):
        super().__init__()
        hidden_dim = hidden_dim or 4 * dim
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))


class PatchEmbedding(_FusedBlock):
    """p x p / stride p Conv2d projection to tokens  (reference components.py:244-307)."""

    def __init__(self, in_channels: int, embed_dim: int, patch_size: int = 4, flatten: bool = True):
        super().__init__()
        self.patch_size = patch_size
        self.flatten = flatten
        self.projection = nn.Conv2d(in_channels, embed_dim, kernel_size=patch_size, stride=patch_size)


class PositionalEncoding(_FusedBlock):
    """Learnable (or sinusoidal) positional table added to the tokens  (reference components.py:310-386)."""

    def __init__(self, embed_dim: int, max_len: int = 5000, learnable: bool = True, dropout: float = 0.1):
        super().__init__()
        self.embed_dim = embed_dim
        self.learnable = learnable
        self.dropout = nn.Dropout(dropout)
        if learnable:
            self.pos_embed = nn.Parameter(torch.zeros(1, max_len, embed_dim))
            nn.init.trunc_normal_(self.pos_embed, std=0.02)
        else:
            pos = torch.arange(max_len).unsqueeze(1)
            freq = torch.exp(torch.arange(0, embed_dim, 2) * (-math.log(10000.0) / embed_dim))
            table = torch.zeros(1, max_len, embed_dim)
            table[0, :, 0::2] = torch.sin(pos * freq)
            table[0, :, 1::2] = torch.cos(pos * freq)
            self.register_buffer("pos_embed", table)


class DropPath(nn.Module):
    """Stochastic depth; identity in eval mode, which is the only mode this package runs
    (reference components.py:389-427)."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        raise RuntimeError("hvit_b200 is inference-only: DropPath in training mode is not implemented")
