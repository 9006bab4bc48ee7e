"""Parameter containers for the Vision Transformer part of HybridViT
(reference models/attention.py:17-304).  See components.py for why these do not compute."""
from __future__ import annotations

import torch
import torch.nn as nn

from .components import DropPath, FeedForward, _FusedBlock


class MultiHeadSelfAttention(_FusedBlock):
    def __init__(self, embed_dim: int, num_heads: int = 8, qkv_bias: bool = True, attn_dropout: float = 0.0,
                 proj_dropout: float = 0.0):
        super().__init__()
        assert embed_dim % num_heads == 0, \
            f"embed_dim ({embed_dim}) must be divisible by num_heads ({num_heads})"  # reference attention.py:46
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.head_dim = embed_dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(embed_dim, embed_dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(embed_dim, embed_dim)
        self.attn_dropout = nn.Dropout(attn_dropout)
        self.proj_dropout = nn.Dropout(proj_dropout)


class TransformerEncoderBlock(_FusedBlock):
    def __init__(self, embed_dim: int, num_heads: int = 8, mlp_ratio: float = 4.0, qkv_bias: bool = True,
                 dropout: float = 0.0, attn_dropout: float = 0.0, drop_path: float = 0.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(embed_dim)
        self.norm2 = nn.LayerNorm(embed_dim)
        self.attn = MultiHeadSelfAttention(embed_dim, num_heads, qkv_bias, attn_dropout, dropout)
        self.mlp = FeedForward(embed_dim, int(embed_dim * mlp_ratio), dropout)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()


class VisionTransformer(_FusedBlock):
    def __init__(self, embed_dim: int, num_layers: int = 6, num_heads: int = 8, mlp_ratio: float = 4.0,
                 qkv_bias: bool = True, dropout: float = 0.0, attn_dropout: float = 0.0, drop_path_rate: float = 0.0):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_layers = num_layers
        rates = [r.item() for r in torch.linspace(0, drop_path_rate, num_layers)]
        self.blocks = nn.ModuleList([
            TransformerEncoderBlock(embed_dim, num_heads, mlp_ratio, qkv_bias, dropout, attn_dropout, rates[i])
            for i in range(num_layers)])
        self.norm = nn.LayerNorm(embed_dim)
