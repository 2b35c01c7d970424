"""nefii_b200 -- B200-native (sm_100a) implementation of NeFII's per-ray-batch rendering hot path.

Host code is Python/PyTorch and mirrors the reference's module API (`nefii_b200.model.*` has the
same class / function names, arguments and return dicts as the reference's `code/model/*`);
all device work goes through the C ABI in include/nefii_b200.h (libnefii_b200.so, loaded by
`nefii_b200._lib`, which raises ImportError when the library has not been built -- there is no
CPU or PyTorch fallback).  Build with `python nefii_b200/build.py`.
"""
__version__ = "0.1.0"
