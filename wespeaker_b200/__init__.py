"""wespeaker_b200 — B200-native speaker-embedding extraction + PLDA scoring engine (drop-in for the
WeSpeaker hot path).  Python host code over a C-ABI/ctypes layer (lib.py) loading hand-written sm_100a
kernels (csrc/)."""
from .lib import B200Error, LIB_PATH  # noqa: F401

__all__ = ["B200Error", "LIB_PATH"]
__version__ = "0.1.0"
