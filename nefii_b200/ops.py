"""Thin Python wrappers over the C ABI building blocks (tensors in, tensors out; no math here)."""
import ctypes

import torch

from . import _lib

c_void_p, c_int, c_float = ctypes.c_void_p, ctypes.c_int32, ctypes.c_float

ACT_NONE, ACT_SOFTPLUS100, ACT_RELU, ACT_ELU = 0, 1, 2, 3


class GemmDesc(ctypes.Structure):
    """Mirror of ``nefii_gemm_desc`` in include/nefii_b200.h (field order matters)."""
    _fields_ = [
        ("a_hi", c_void_p), ("a_lo", c_void_p), ("a_ld", c_int), ("rows_cap", c_int),
        ("b_hi", c_void_p), ("b_lo", c_void_p), ("b_ld", c_int), ("n_pad", c_int),
        ("k_pad", c_int),
        ("count", c_void_p),
        ("mode", c_int), ("act", c_int), ("n_valid", c_int),
        ("bias", c_void_p),
        ("out_scale", c_float),
        ("dst_hi", c_void_p), ("dst_lo", c_void_p), ("dst_ld", c_int), ("dst_col0", c_int), ("dst_ncols", c_int),
        ("dst_f32", c_void_p), ("f32_ld", c_int), ("f32_begin", c_int), ("f32_end", c_int),
        ("w_last", c_void_p), ("b_last", c_void_p), ("n_last", c_int), ("w_last_ld", c_int), ("dst_last", c_void_p),
        ("seed_hi", c_void_p), ("seed_lo", c_void_p), ("seed_ld", c_int),
        ("sav_hi", c_void_p), ("sav_lo", c_void_p), ("sav_ld", c_int), ("sav_ncols", c_int), ("sav_scale", c_float),
    ]


def _p(t):
    return None if t is None else t.data_ptr()


def round_up(x, m):
    return (x + m - 1) // m * m


def split_to_planes(src, rows_pad=None, cols_pad=None, transpose=False, scale=1.0):
    """fp32 [rows, cols] -> (hi, lo) bf16 planes, zero padded to [rows_pad, cols_pad]."""
    assert src.is_cuda and src.dtype == torch.float32 and src.dim() == 2 and src.stride(1) == 1
    rows, cols = src.shape
    orow, ocol = (cols, rows) if transpose else (rows, cols)
    rows_pad = rows_pad or orow
    cols_pad = cols_pad or round_up(ocol, 64)
    hi = torch.empty(rows_pad, cols_pad, device=src.device, dtype=torch.bfloat16)
    lo = torch.empty_like(hi)
    _lib.check(_lib.raw().nefii_split_to_planes(
        _lib.stream_ptr(src.device), _lib.dptr(src) if src.is_contiguous() else c_void_p(src.data_ptr()),
        rows, cols, src.stride(0), 1 if transpose else 0, float(scale),
        c_void_p(hi.data_ptr()), c_void_p(lo.data_ptr()), rows_pad, cols_pad))
    return hi, lo


def gemm_split_bf16(a, b, k_pad, n_valid, *, mode=0, act=ACT_NONE, bias=None, out_scale=1.0, count=None,
                    dst=None, dst_col0=0, dst_ncols=0, dst_f32=None, f32_begin=0, f32_end=0,
                    w_last=None, b_last=None, dst_last=None, seed=None, sav=None, sav_ncols=0, sav_scale=1.0):
    """a=(hi,lo) activations planes [rows_cap, a_ld], b=(hi,lo) weight planes [n_pad, b_ld]."""
    d = GemmDesc()
    d.a_hi, d.a_lo, d.a_ld, d.rows_cap = _p(a[0]), _p(a[1]), a[0].stride(0), a[0].shape[0]
    d.b_hi, d.b_lo, d.b_ld, d.n_pad = _p(b[0]), _p(b[1]), b[0].stride(0), b[0].shape[0]
    d.k_pad = k_pad
    d.count = _p(count)
    d.mode, d.act, d.n_valid = mode, act, n_valid
    d.bias = _p(bias)
    d.out_scale = out_scale
    if dst is not None:
        d.dst_hi, d.dst_lo, d.dst_ld = _p(dst[0]), _p(dst[1]), dst[0].stride(0)
        d.dst_col0, d.dst_ncols = dst_col0, dst_ncols
    if dst_f32 is not None:
        d.dst_f32, d.f32_ld, d.f32_begin, d.f32_end = _p(dst_f32), dst_f32.stride(0), f32_begin, f32_end
    if w_last is not None:
        d.w_last, d.b_last, d.n_last, d.w_last_ld = _p(w_last), _p(b_last), w_last.shape[0], w_last.stride(0)
        d.dst_last = _p(dst_last)
    if seed is not None:
        d.seed_hi, d.seed_lo, d.seed_ld = _p(seed[0]), _p(seed[1]), seed[0].stride(0)
    if sav is not None:
        d.sav_hi, d.sav_lo, d.sav_ld = _p(sav[0]), _p(sav[1]), sav[0].stride(0)
        d.sav_ncols, d.sav_scale = sav_ncols, sav_scale
    _lib.check(_lib.raw().nefii_gemm_split_bf16(_lib.stream_ptr(a[0].device), ctypes.byref(d)))
