"""Mirror of the reference ``models`` package interface (reference models/__init__.py:8-27)."""
from .components import (ConvBlock, TransposeConvBlock, FeedForward, PatchEmbedding, PositionalEncoding,
                         DropPath)
from .attention import MultiHeadSelfAttention, TransformerEncoderBlock, VisionTransformer
from .hybrid_vit import HybridViT, create_hybrid_vit

__all__ = ["ConvBlock", "TransposeConvBlock", "FeedForward", "PatchEmbedding", "PositionalEncoding", "DropPath",
           "MultiHeadSelfAttention", "TransformerEncoderBlock", "VisionTransformer", "HybridViT",
           "create_hybrid_vit"]
