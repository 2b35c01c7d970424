"""Host emulation of the device math headers -- test infrastructure only (see sg_emu.cpp)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.normpath(os.path.join(_HERE, "..", "..", "nefii_b200", "csrc"))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "_hostemu.so")
    srcs = [os.path.join(_HERE, f) for f in sorted(os.listdir(_HERE)) if f.endswith(".cpp")]
    deps = srcs + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith("_math.cuh")]
    if force or not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
               "-I", _CSRC, "-x", "c++"] + srcs + ["-o", so]
        subprocess.check_call(cmd)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB
