"""hvit_b200 - B200-native (sm_100a) drop-in for the HybridViT speech-enhancement inference path.

Mirrors the reference's public interface for that path only:
  models.HybridViT / models.create_hybrid_vit      (reference models/hybrid_vit.py)
  inference.AudioEnhancer / inference.enhance_audio (reference inference/enhancer.py)
  utils.audio_processing.compute_stft / compute_istft / ...  (reference utils/audio_processing.py)
  utils.config.load_all_configs, utils.checkpoint.load_model_weights

All arithmetic runs in hand-written CUDA kernels behind the C ABI in include/hvit.h
(csrc/libhvit_sm100.so); there is no CPU or PyTorch-eager fallback.
"""
from . import _lib  # noqa: F401
from .models import HybridViT, create_hybrid_vit  # noqa: F401
from .inference import AudioEnhancer, enhance_audio  # noqa: F401

__all__ = ["HybridViT", "create_hybrid_vit", "AudioEnhancer", "enhance_audio"]
