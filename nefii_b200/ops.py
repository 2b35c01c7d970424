"""Thin Python wrappers over the C ABI building blocks (tensors in, tensors out; no math here)."""
import ctypes

import torch

from . import _lib

c_void_p, c_int, c_float = ctypes.c_void_p, ctypes.c_int32, ctypes.c_float

ACT_NONE, ACT_SOFTPLUS100, ACT_RELU, ACT_ELU = 0, 1, 2, 3
# plane formats (include/nefii_b200.h): bf16 split (trainable stacks, backward passes), fp16 split (SDF inference chain)
PLANES_BF16, PLANES_FP16 = 0, 1


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
        ("dst_hi", c_void_p), ("dst_lo", c_void_p), ("dst_ld", c_int), ("dst_col0", c_int), ("dst_ncols", c_int), ("dst_zero_to", c_int),
        ("dst_f32", c_void_p), ("f32_ld", c_int), ("f32_begin", c_int), ("f32_end", c_int),
        ("w_last", c_void_p), ("b_last", c_void_p), ("n_last", c_int), ("w_last_ld", c_int), ("dst_last", c_void_p),
        ("seed_hi", c_void_p), ("seed_lo", c_void_p), ("seed_ld", c_int),
        ("sav_hi", c_void_p), ("sav_lo", c_void_p), ("sav_ld", c_int), ("sav_ncols", c_int), ("sav_scale", c_float),
        ("k_splits", c_int), ("f32_split_stride", ctypes.c_int64), ("k_splits_used", c_int),
        ("k_flush", c_int), ("dst_pad_ok", c_int), ("fmt", c_int),
        ("pe_x", c_void_p), ("pe_n_freqs", c_int), ("pe_side_hi", c_void_p), ("pe_side_lo", c_void_p), ("pe_side_ld", c_int),
        ("pe_side_col0", c_int), ("pe_side_scale", c_float),
    ]


def _p(t):
    return None if t is None else t.data_ptr()


def round_up(x, m):
    return (x + m - 1) // m * m


def split_to_planes(src, rows_pad=None, cols_pad=None, transpose=False, scale=1.0, fmt=PLANES_BF16):
    """fp32 [rows, cols] -> (hi, lo) 16-bit planes of format `fmt`, zero padded to [rows_pad, cols_pad]."""
    assert src.is_cuda and src.dtype == torch.float32 and src.dim() == 2
    rows, cols = src.shape
    if cols > 1 and src.stride(1) != 1:
        src = src.contiguous()
    # (a size-1 dimension may carry any stride: only its index 0 is ever addressed)
    orow, ocol = (cols, rows) if transpose else (rows, cols)
    rows_pad = rows_pad or orow
    cols_pad = cols_pad or round_up(ocol, 64)
    hi = torch.empty(rows_pad, cols_pad, device=src.device, dtype=torch.float16 if fmt == PLANES_FP16 else torch.bfloat16)
    lo = torch.empty_like(hi)
    _lib.check(_lib.raw().nefii_split_to_planes_fmt(
        _lib.stream_ptr(src.device), _lib.dptr(src) if src.is_contiguous() else c_void_p(src.data_ptr()),
        rows, cols, src.stride(0), 1 if transpose else 0, float(scale),
        c_void_p(hi.data_ptr()), c_void_p(lo.data_ptr()), rows_pad, cols_pad, int(fmt)))
    return hi, lo


def gemm_split_bf16(a, b, k_pad, n_valid, *, mode=0, act=ACT_NONE, bias=None, out_scale=1.0, count=None,
                    dst=None, dst_col0=0, dst_ncols=0, dst_zero_to=0, dst_f32=None, f32_begin=0, f32_end=0, f32_ld=None,
                    w_last=None, b_last=None, dst_last=None, seed=None, sav=None, sav_ncols=0, sav_scale=1.0,
                    k_splits=1, f32_split_stride=0, rows_cap=None, k_flush=0, fmt=None, dst_pad_ok=0,
                    pe_x=None, pe_n_freqs=0, pe_side=None, pe_side_col0=0, pe_side_scale=1.0):
    """a=(hi,lo) activations planes [rows_cap, a_ld], b=(hi,lo) weight planes [n_pad, b_ld].  fmt: plane format of every
    plane of the launch (default: from the dtype of a)."""
    d = GemmDesc()
    if pe_x is not None:
        # PE prologue: a is None, the rows of A are the positional encoding of pe_x [rows, 3], computed inside the kernel
        d.fmt = (PLANES_FP16 if b[0].dtype == torch.float16 else PLANES_BF16) if fmt is None else fmt
        d.pe_x, d.pe_n_freqs = _p(pe_x), pe_n_freqs
        d.rows_cap = pe_x.shape[0] if rows_cap is None else rows_cap
        if pe_side is not None:
            d.pe_side_hi, d.pe_side_lo, d.pe_side_ld = _p(pe_side[0]), _p(pe_side[1]), pe_side[0].stride(0)
            d.pe_side_col0, d.pe_side_scale = pe_side_col0, pe_side_scale
    else:
        d.fmt = (PLANES_FP16 if a[0].dtype == torch.float16 else PLANES_BF16) if fmt is None else fmt
        d.a_hi, d.a_lo, d.a_ld, d.rows_cap = _p(a[0]), _p(a[1]), a[0].stride(0), (a[0].shape[0] if rows_cap is None else rows_cap)
    d.b_hi, d.b_lo, d.b_ld, d.n_pad = _p(b[0]), _p(b[1]), b[0].stride(0), b[0].shape[0]
    d.k_pad = k_pad
    d.count = _p(count)
    d.mode, d.act, d.n_valid = mode, act, n_valid
    d.bias = _p(bias)
    d.out_scale = out_scale
    if dst is not None:
        d.dst_hi, d.dst_lo, d.dst_ld = _p(dst[0]), _p(dst[1]), dst[0].stride(0)
        d.dst_col0, d.dst_ncols, d.dst_zero_to = dst_col0, dst_ncols, dst_zero_to
    if dst_f32 is not None:
        d.dst_f32, d.f32_begin, d.f32_end = _p(dst_f32), f32_begin, f32_end
        d.f32_ld = dst_f32.stride(0) if f32_ld is None else f32_ld
    if w_last is not None:
        d.w_last, d.b_last, d.n_last, d.w_last_ld = _p(w_last), _p(b_last), w_last.shape[0], w_last.stride(0)
        d.dst_last = _p(dst_last)
    if seed is not None:
        d.seed_hi, d.seed_lo, d.seed_ld = _p(seed[0]), _p(seed[1]), seed[0].stride(0)
    if sav is not None:
        d.sav_hi, d.sav_lo, d.sav_ld = _p(sav[0]), _p(sav[1]), sav[0].stride(0)
        d.sav_ncols, d.sav_scale = sav_ncols, sav_scale
    d.k_splits, d.f32_split_stride = k_splits, f32_split_stride
    d.k_flush = k_flush
    d.dst_pad_ok = dst_pad_ok
    _lib.check(_lib.raw().nefii_gemm_split_bf16(_lib.stream_ptr(b[0].device), ctypes.byref(d)))
    return d.k_splits_used


class SdfConfig(ctypes.Structure):
    """Mirror of ``nefii_sdf_config``."""
    _fields_ = [("d_in", c_int), ("n_freqs", c_int), ("width", c_int), ("n_hidden", c_int),
                ("skip_layer", c_int), ("d_out", c_int), ("d_feat", c_int)]


class SdfMlp:
    """Owner of a ``nefii_sdf_*`` handle: packed weights live in the library, workspaces are torch tensors."""

    def __init__(self, n_freqs=6, width=512, n_hidden=8, skip_layer=4, device=None, d_feat=0, fmt=None):
        """d_feat = 0: feature vector = input of the last layer (use_last_as_f); > 0: rows 1.. of the last Linear.
        fmt: plane format of the inference chain (None = library default: fp16 split, NEFII_SDF_FORMAT overrides)."""
        self.device = torch.device(device if device is not None else "cuda")
        self.cfg = SdfConfig(3, n_freqs, width, n_hidden, skip_layer, 1, d_feat)
        self.width, self.n_hidden = width, n_hidden
        self.feat_width = d_feat if d_feat > 0 else width
        h = c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.raw().nefii_sdf_create(ctypes.byref(h), ctypes.byref(self.cfg)))
        self._h = h
        self._ws = None
        self._keep = None
        if fmt is not None:
            self.set_format(fmt)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.raw().nefii_sdf_destroy(h)

    @property
    def handle(self):
        return self._h

    def set_format(self, fmt):
        """Plane format of the inference chain; the packed weights are dropped (call set_weights again)."""
        _lib.check(_lib.raw().nefii_sdf_set_format(self._h, int(fmt)))

    @property
    def format(self):
        return int(_lib.raw().nefii_sdf_get_format(self._h))

    def set_weights(self, weights, biases):
        """weights[l]: effective fp32 [out_l, in_l] CUDA tensors (weight norm folded)."""
        n = self.n_hidden + 1
        assert len(weights) == n and len(biases) == n
        ws = [_lib.f32c(w) for w in weights]
        bs = [_lib.f32c(b) for b in biases]
        wp = (c_void_p * n)(*[w.data_ptr() for w in ws])
        bp = (c_void_p * n)(*[b.data_ptr() for b in bs])
        with torch.cuda.device(self.device):
            _lib.check(_lib.raw().nefii_sdf_set_weights(self._h, _lib.stream_ptr(self.device), wp, bp))
        self._keep = (ws, bs)   # keep sources alive until the async copies have been ordered on the stream

    def workspace(self, rows_cap, with_grad):
        need = int(_lib.raw().nefii_sdf_workspace_bytes(self._h, int(rows_cap), 1 if with_grad else 0))
        # one workspace per stream: evaluations enqueued on different streams (IDRNetwork.prefetch_trace) may overlap
        sid = torch.cuda.current_stream(self.device).cuda_stream
        if self._ws is None:
            self._ws = {}
        ws = self._ws.get(sid)
        if ws is None or ws.numel() < need:
            self._ws.pop(sid, None)
            ws = self._ws[sid] = torch.empty(need, dtype=torch.uint8, device=self.device)
        return ws

    def eval(self, x, want_feat=False, want_grad=False, count=None, k_flush=0):
        """x [N,3] -> (sdf [N], feat [N,width] | None, grad [N,3] | None).  k_flush: accuracy tier (1 = most accurate,
        0 = library default)."""
        x = _lib.f32c(x).reshape(-1, 3)
        n = x.shape[0]
        sdf = torch.empty(n, device=x.device, dtype=torch.float32)
        feat = torch.empty(n, self.feat_width, device=x.device, dtype=torch.float32) if want_feat else None
        grad = torch.empty(n, 3, device=x.device, dtype=torch.float32) if want_grad else None
        if n == 0:
            return sdf, feat, grad
        ws = self.workspace(n, want_grad)
        with torch.cuda.device(self.device):
            _lib.check(_lib.raw().nefii_sdf_eval(
                self._h, _lib.stream_ptr(self.device), n, _p(count), x.data_ptr(), ws.data_ptr(), ws.numel(),
                sdf.data_ptr(), _p(feat), _p(grad), int(k_flush)))
        return sdf, feat, grad
