"""CPU: the C-ABI library loads and exports every symbol include/nefii_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nefii_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nefii_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from nefii_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert "nefii_sg_render_fwd" in names and len(names) >= 5
    for n in names:
        assert hasattr(lib, n), "missing export: " + n


def test_python_signature_table_covers_header():
    from nefii_b200 import _lib
    declared = set(_declared_symbols()) - {"nefii_last_error"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_abi_version_and_error_text():
    from nefii_b200 import _lib
    assert _lib.raw().nefii_abi_version() >= 1
    assert isinstance(_lib.raw().nefii_last_error(), bytes)


def test_no_cpu_fallback():
    import pytest
    import torch
    from nefii_b200 import _lib
    from nefii_b200.model.sg_render import render_with_sg
    with pytest.raises(_lib.NefiiError):
        render_with_sg(torch.zeros(4, 7), torch.zeros(1, 3), torch.ones(1, 1), torch.zeros(5, 3), torch.zeros(5, 3),
                       torch.zeros(5, 3))


def test_argument_errors_return_codes_not_crashes():
    """The ABI's error convention (include/nefii_b200.h: negative int + nefii_last_error text, nothing throws or exits):
    malformed sizes are rejected before any CUDA call, so this runs without a GPU."""
    from nefii_b200 import _lib
    lib = _lib.raw()
    null = None
    one = ctypes.c_void_p(16)          # a non-null pointer that is never dereferenced (argument checks come first)
    cases = [
        lambda: lib.nefii_sg_render_fwd(null, 8, 0, 1, one, one, one, one, one, one, null, one, one, one),       # no light SGs
        lambda: lib.nefii_sg_render_fwd(null, 8, 128, 99, one, one, one, one, one, one, null, one, one, one),    # too many materials
        lambda: lib.nefii_sg_render_fwd(null, 8, 128, 1, null, one, one, one, one, one, null, one, one, one),    # null input
        lambda: lib.nefii_background_sg_fwd(null, 8, 0, one, one, one),
        lambda: lib.nefii_idr_loss_fwd(null, 10, 4, one, one, one, one, one, one, one, 0, 1, ctypes.c_float(50.0), one),   # 10 pixels, patches of 4
        lambda: lib.nefii_idr_loss_fwd(null, 8, 4, one, one, one, one, one, one, one, 7, 1, ctypes.c_float(50.0), one),    # unknown loss kind
        lambda: lib.nefii_idr_loss_fwd(null, 8, 4, one, one, one, one, one, one, one, 0, 1, ctypes.c_float(-1.0), one),    # alpha <= 0
        lambda: lib.nefii_gemm_set_cluster(3),
        lambda: lib.nefii_gemm_set_k_flush(0),
    ]
    for i, call in enumerate(cases):
        rc = call()
        assert rc < 0, (i, rc)
        assert len(lib.nefii_last_error()) > 0
    # empty problems are fine and touch nothing
    assert lib.nefii_sg_render_fwd(null, 0, 128, 1, null, null, null, null, null, null, null, null, null, null) == 0
    assert lib.nefii_idr_loss_bwd(null, 0, 4, null, null, null, null, null, null, null, 0, 1, ctypes.c_float(50.0), null, null,
                                  null, null, null, null) == 0


def test_ctypes_struct_mirrors_match_the_header_layout(tmp_path):
    """Every descriptor struct of include/nefii_b200.h against its ctypes mirror: size and the offset of every field, as gcc lays
    the header out (a drifted field order would silently mis-read pointers)."""
    import subprocess
    from nefii_b200 import mlp, ops
    from nefii_b200.model import ray_tracing
    mirrors = {"nefii_gemm_desc": ops.GemmDesc, "nefii_sdf_config": ops.SdfConfig, "nefii_trace_config": ray_tracing.TraceConfig,
               "nefii_dense_stack_desc": mlp.DenseStackDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include <stdint.h>', '#include "%s"' % os.path.join(ROOT, "include", "nefii_b200.h"),
             'int main(void) {']
    for cname, mirror in mirrors.items():
        lines.append('  printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in mirror._fields_:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for row in filter(None, out):
        cname, field, value = row.split()
        mirror = mirrors[cname]
        want = ctypes.sizeof(mirror) if field == "sizeof" else getattr(mirror, field).offset
        assert int(value) == want, (cname, field, int(value), want)
        seen += 1
    assert seen == sum(len(m._fields_) + 1 for m in mirrors.values())
