"""nefii_b200 -- B200-native (sm_100a) implementation of NeFII's per-ray-batch rendering hot path.

Host code is Python/PyTorch and mirrors the reference's module API (`nefii_b200.model.*` has the
same class / function names, arguments and return dicts as the reference's `code/model/*`);
all device work goes through the C ABI in include/nefii_b200.h (libnefii_b200.so).
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)

__all__ = ["_lib"]
