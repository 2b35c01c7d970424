"""ctypes binding of libnefii_b200.so (the C ABI declared in include/nefii_b200.h).

There is no CPU fallback: if the shared library is missing the import fails loudly, and every
call on a non-CUDA tensor raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NEFII_LIB_PATH") or os.path.join(_HERE, "libnefii_b200.so")   # the override is for A/B builds

c_void_p, c_int, c_float, c_longlong = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong

# symbol -> argtypes; kept in sync with include/nefii_b200.h (tests/test_abi.py checks both ways)
SIGNATURES = {
    "nefii_abi_version": [],
    "nefii_launch_count": [],
    "nefii_sg_render_fwd": [c_void_p, c_int, c_int, c_int] + [c_void_p] * 10,
    "nefii_sg_render_bwd": [c_void_p, c_int, c_int, c_int] + [c_void_p] * 18,
    "nefii_background_sg_fwd": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "nefii_gemm_split_bf16": [c_void_p, c_void_p],
    "nefii_probe_fp32": [c_void_p, c_int, c_int, c_void_p],
    "nefii_camera_rays": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    "nefii_sample_network_fwd": [c_void_p, c_int] + [c_void_p] * 7,
    "nefii_sample_network_bwd": [c_void_p, c_int] + [c_void_p] * 12,
    "nefii_gemm_profile_enable": [c_int],
    "nefii_gemm_set_cluster": [c_int],
    "nefii_gemm_set_debug": [c_int],
    "nefii_gemm_set_pdl": [c_int],
    "nefii_gemm_set_grid_cap": [c_int],
    "nefii_gemm_set_k_flush": [c_int],
    "nefii_gemm_set_k_flush_head": [c_int],
    "nefii_gemm_set_trunc_comp": [c_int, c_float],
    "nefii_gemm_profile_fetch": [c_void_p],
    "nefii_assemble_input": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int],
    "nefii_transpose_planes": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "nefii_last_layer_bwd": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                             c_void_p, c_int, c_void_p, c_void_p],
    "nefii_reduce_splits": [c_void_p, c_void_p, c_int, c_longlong, c_int, c_int, c_int, c_void_p],
    "nefii_mis_sample": [c_void_p, c_int, c_int] + [c_void_p] * 9,
    "nefii_mis_shade_fwd": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int] + [c_void_p] * 13,
    "nefii_mis_shade_bwd": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int] + [c_void_p] * 19,
    "nefii_background_sg_bwd": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "nefii_sg_param_grad": [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int],
    "nefii_trace_workspace_bytes": [c_int, c_void_p, c_int, c_int],
    "nefii_ray_trace": [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                        c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p],
    "nefii_trace_set_tiers": [c_int, c_int],
    "nefii_trace_set_graph_mode": [c_int],
    "nefii_trace_graph_mode": [],
    "nefii_trace_graph_captures": [],
    "nefii_trace_set_quad_rows": [c_int],
    "nefii_trace_set_bisect_depth": [c_int],
    "nefii_trace_graph_clear": [],
    "nefii_analytic_sdf_eval": [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p],
    "nefii_sdf_create": [c_void_p, c_void_p],
    "nefii_sdf_destroy": [c_void_p],
    "nefii_sdf_set_weights": [c_void_p, c_void_p, c_void_p, c_void_p],
    "nefii_sdf_workspace_bytes": [c_void_p, c_int, c_int],
    "nefii_sdf_eval": [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_int],
    "nefii_split_to_planes": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_int, c_int],
    "nefii_split_to_planes_fmt": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_int, c_int, c_int],
    "nefii_sdf_set_format": [c_void_p, c_int],
    "nefii_sdf_get_format": [c_void_p],
    "nefii_sdf_set_pe_prologue": [c_int],
    "nefii_gemm_set_trunc_comp_fmt": [c_int, c_int, c_float],
    "nefii_dense_stack_workspace_bytes": [c_void_p],
    "nefii_dense_stack_fwd": [c_void_p, c_void_p],
    "nefii_dense_stack_bwd": [c_void_p, c_void_p],
    "nefii_idr_loss_fwd": [c_void_p, c_int, c_int] + [c_void_p] * 7 + [c_int, c_int, c_float, c_void_p],
    "nefii_idr_loss_bwd": [c_void_p, c_int, c_int] + [c_void_p] * 7 + [c_int, c_int, c_float] + [c_void_p] * 6,
}


class NefiiError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "nefii_b200: %s not found -- build it with `python -m nefii_b200.build` "
            "(there is no CPU / PyTorch fallback path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.nefii_last_error.restype = ctypes.c_char_p
    lib.nefii_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    lib.nefii_sdf_workspace_bytes.restype = c_longlong
    lib.nefii_dense_stack_workspace_bytes.restype = c_longlong
    lib.nefii_launch_count.restype = c_longlong
    lib.nefii_trace_graph_captures.restype = c_longlong
    lib.nefii_trace_workspace_bytes.restype = c_longlong
    return lib


_lib = _load()


def raw():
    return _lib


def check(rc):
    if rc != 0:
        raise NefiiError("nefii_b200 error %d: %s" % (rc, _lib.nefii_last_error().decode("utf-8", "replace")))


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dptr(t, dtype=torch.float32, allow_none=False):
    """Device pointer of a contiguous CUDA tensor (raises instead of silently copying)."""
    if t is None:
        if allow_none:
            return None
        raise NefiiError("nefii_b200: required tensor is None")
    if not t.is_cuda:
        raise NefiiError("nefii_b200: expected a CUDA tensor (no CPU path exists)")
    if t.dtype != dtype:
        raise NefiiError("nefii_b200: expected dtype %s, got %s" % (dtype, t.dtype))
    if not t.is_contiguous():
        raise NefiiError("nefii_b200: expected a contiguous tensor")
    return c_void_p(t.data_ptr())


def f32c(t):
    """float32 contiguous view/copy of a CUDA tensor (expanded views are materialised)."""
    if not t.is_cuda:
        raise NefiiError("nefii_b200: expected a CUDA tensor (no CPU path exists)")
    return t.detach().to(torch.float32).contiguous()
