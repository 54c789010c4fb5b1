"""tmglow_b200 -- B200-native (sm_100a) implementation of the TM-Glow flow hot path.

    import sys; sys.path.insert(0, "deep-turbulence_b200")
    from tmglow_b200 import TMGlow            # drop-in for tmglow/nn/tmGlow.py:TMGlow

Host side in Python/PyTorch (device memory, streams), compute in libtmglow_b200.so through the
C ABI of include/tmglow_b200.h.  No CPU fallback.
"""
from . import _lib  # noqa: F401
from .nn.tmGlow import TMGlow  # noqa: F401
from .optim import FlatAdam  # noqa: F401

__all__ = ["TMGlow", "FlatAdam"]
