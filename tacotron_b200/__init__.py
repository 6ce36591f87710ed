"""tacotron_b200 -- B200-native (sm_100a) Tacotron mel/linear-spectrogram hot path.

Python host over PyTorch tensors calling hand-written CUDA through a ctypes C-ABI
(include/taco_b200.h).  Mirrors the reference's `models.ops` / `models.tacotron` surface.
"""
from . import _lib  # noqa: F401
from .models.tacotron import Config, Tacotron  # noqa: F401
from .models import ops  # noqa: F401

__all__ = ["Config", "Tacotron", "ops"]
