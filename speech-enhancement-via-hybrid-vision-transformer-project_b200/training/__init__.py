"""Validation side of the reference's training package: ``Trainer.validate`` is a plain
``HybridViT.forward`` over a loader plus a spectrogram loss, so it runs on the inference plan.  Training itself
(optimiser, backward, schedulers, checkpoints) is out of scope."""
from .losses import CombinedLoss, SpectrogramLoss, STOILoss, create_loss_function
from .validation import Validator, validate

__all__ = ["CombinedLoss", "SpectrogramLoss", "STOILoss", "create_loss_function", "Validator", "validate"]
