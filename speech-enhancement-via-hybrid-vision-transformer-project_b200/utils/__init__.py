"""Mirror of the hot-path part of the reference ``utils`` package (config, checkpoint, STFT wrappers)."""
from .audio_processing import (compute_stft, compute_istft, normalize_audio, compute_magnitude_phase,
                               reconstruct_from_magnitude_phase, load_audio, save_audio)
from .config import load_config, merge_configs, load_all_configs
from .checkpoint import load_model_weights

__all__ = ["compute_stft", "compute_istft", "normalize_audio", "compute_magnitude_phase",
           "reconstruct_from_magnitude_phase", "load_audio", "save_audio", "load_config", "merge_configs",
           "load_all_configs", "load_model_weights"]
