"""In-tree build of libnefii_b200.so (nvcc, sm_100a only).

``python -m nefii_b200.build`` or ``nefii_b200.build.build()``.  The .so lands next to this file
(git-ignored, but shipped to the GPU box by gpurun).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libnefii_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
BASE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]
# Translation units whose results feed comparisons / ill-conditioned cancellations replicate the
# reference's per-op float32 rounding: no FMA contraction, IEEE div/sqrt (see sg_math.cuh).
EXACT_FP32 = {"sg_render.cu", "tracer.cu", "mis.cu", "sample_network.cu"}
EXACT_FLAGS = ["-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false"]
# Instruction-bound gradient kernels whose results only have to meet the rel 1e-3 gradient tolerance
FAST_FP32 = {"sg_render_bwd.cu"}
FAST_FLAGS = ["-prec-div=false", "-prec-sqrt=false"]


def _needs(src, obj, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src] + deps)


def _compile(src):
    name = os.path.basename(src)
    obj = os.path.join(OBJ, name.replace(".cu", ".o"))
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "nefii_b200.h"))
    if not _needs(src, obj, headers + [os.path.abspath(__file__)]):
        return obj, ""
    flags = list(BASE_FLAGS) + (EXACT_FLAGS if name in EXACT_FP32 else []) + (FAST_FLAGS if name in FAST_FP32 else [])
    flags += os.environ.get("NEFII_NVCC_EXTRA", "").split()      # development: A/B builds (-D switches)
    cmd = [NVCC] + flags + ["-I", CSRC, "-c", src, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (name, p.stdout, p.stderr))
    with open(obj + ".ptxas.log", "w") as f:
        f.write(p.stderr)
    return obj, p.stderr


def build(verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(_compile, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    if (not os.path.exists(LIB)) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
               "--cudart", "shared", "-o", LIB] + objs
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
