"""Mirror of the reference ``inference`` package interface (reference inference/__init__.py)."""
from .enhancer import AudioEnhancer, enhance_audio, load_model_for_inference

__all__ = ["AudioEnhancer", "enhance_audio", "load_model_for_inference"]
