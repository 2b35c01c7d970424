"""Trainable dense stacks (radiance net, material net) on the tcgen05 layer GEMM, with autograd.

``dense_mlp(segments, weights, biases, act)`` computes what the reference computes with
``cat -> [Linear + act] * L -> Linear`` (RenderingNetwork.forward, implicit_differentiable_renderer.py:196-241;
EnvmapMaterialNetwork.diffuse_albedo_layers, sg_envmap_material.py:357-366) and returns the raw output of the
last Linear.  Gradients flow to the weights and biases only -- in the reference's step-2 training the inputs
(points, normals, view directions, frozen geometry features) carry no gradient.

Forward : input assembly (PE + concat) -> hidden layers (GEMM + bias + act fused) -> the last hidden layer's
          epilogue also applies the tiny output layer in fp32.
Backward: output layer backward (elementwise + reductions), then per hidden layer a weight-gradient GEMM
          (K = points, split-K over the grid, operands transposed into K-major planes), the bias gradient as
          column sums, and the data-gradient GEMM with the activation derivative fused into its epilogue.
"""
import ctypes
import os

import torch

from . import _lib, ops

c_void_p = ctypes.c_void_p
# accuracy tier of the forward layer GEMMs of the trainable stacks (K blocks per TMEM partial; 1 = shortest partial sums,
# 0 = library default).  With the truncation compensation of the layer GEMM (csrc/mlp_gemm.cu) the default is as accurate as 1.
FWD_FLUSH = 0


# one native call per direction for the plain stacks (csrc/dense_stack.cu) instead of one ctypes call per kernel; 0 = the
# per-kernel sequence below (kept: it is the only one that knows skip connections and want_hidden)
NATIVE_STACK = int(os.environ.get("NEFII_DENSE_NATIVE", "1"))


class DenseStackDesc(ctypes.Structure):
    """Mirror of ``nefii_dense_stack_desc``."""
    _fields_ = [("rows", ctypes.c_int32), ("n_hidden", ctypes.c_int32), ("act", ctypes.c_int32), ("n_seg", ctypes.c_int32),
                ("seg_src", c_void_p * 4), ("seg_width", ctypes.c_int32 * 4), ("seg_freqs", ctypes.c_int32 * 4),
                ("weights", c_void_p), ("biases", c_void_p), ("dim_in", c_void_p), ("dim_out", c_void_p),
                ("need_grad", ctypes.c_int32), ("workspace", c_void_p), ("workspace_bytes", ctypes.c_int64),
                ("y", c_void_p), ("gy", c_void_p), ("grad_w", c_void_p), ("grad_b", c_void_p)]


class _NativeStack:
    """Descriptor + the host arrays it points to + the tensors the device pointers belong to (kept alive together)."""

    def __init__(self, act, srcs, seg_freqs, Ws, bs, need_grad):
        n_lin = len(Ws)
        self.n_lin = n_lin
        self.keep = (srcs, Ws, bs)
        self.w_ptrs = (c_void_p * n_lin)(*[w.data_ptr() for w in Ws])
        self.b_ptrs = (c_void_p * n_lin)(*[b.data_ptr() for b in bs])
        self.dim_in = (ctypes.c_int32 * n_lin)(*[w.shape[1] for w in Ws])
        self.dim_out = (ctypes.c_int32 * n_lin)(*[w.shape[0] for w in Ws])
        d = DenseStackDesc()
        d.rows, d.n_hidden, d.act, d.n_seg = srcs[0].shape[0], n_lin - 1, act, len(srcs)
        for i, (t, f) in enumerate(zip(srcs, seg_freqs)):
            d.seg_src[i], d.seg_width[i], d.seg_freqs[i] = t.data_ptr(), t.shape[1], f
        d.weights, d.biases = ctypes.addressof(self.w_ptrs), ctypes.addressof(self.b_ptrs)
        d.dim_in, d.dim_out = ctypes.addressof(self.dim_in), ctypes.addressof(self.dim_out)
        d.need_grad = 1 if need_grad else 0
        self.desc = d
        need = int(_lib.raw().nefii_dense_stack_workspace_bytes(ctypes.byref(d)))
        if need < 0:
            _lib.check(1)
        self.ws = torch.empty(need, dtype=torch.uint8, device=srcs[0].device)
        d.workspace, d.workspace_bytes = self.ws.data_ptr(), need

    def forward(self, y):
        self.desc.y = y.data_ptr()
        _lib.check(_lib.raw().nefii_dense_stack_fwd(_lib.stream_ptr(y.device), ctypes.byref(self.desc)))

    def backward(self, gy):
        """-> [gW_0, gb_0, gW_1, ...] as views of one flat buffer."""
        Ws, bs = self.keep[1], self.keep[2]
        sizes = []
        for w, b in zip(Ws, bs):
            sizes += [w.numel(), b.numel()]
        flat = torch.empty(sum(sizes), device=gy.device, dtype=torch.float32)
        parts = list(torch.split(flat, sizes))
        base = flat.data_ptr()
        offs = [0]
        for sz in sizes:
            offs.append(offs[-1] + 4 * sz)
        gw = (c_void_p * self.n_lin)(*[base + offs[2 * l] for l in range(self.n_lin)])
        gb = (c_void_p * self.n_lin)(*[base + offs[2 * l + 1] for l in range(self.n_lin)])
        self.desc.gy = gy.data_ptr()
        self.desc.grad_w, self.desc.grad_b = ctypes.addressof(gw), ctypes.addressof(gb)
        _lib.check(_lib.raw().nefii_dense_stack_bwd(_lib.stream_ptr(gy.device), ctypes.byref(self.desc)))
        return [p.view(t.shape) for p, t in zip(parts, [x for wb in zip(Ws, bs) for x in wb])]


def _num_sms(device):
    return torch.cuda.get_device_properties(device).multi_processor_count


def _planes(rows, cols, device):
    return (torch.empty(rows, cols, device=device, dtype=torch.bfloat16),
            torch.empty(rows, cols, device=device, dtype=torch.bfloat16))


def assemble_input(segments, k_pad):
    """segments: list of (tensor [N, w], n_freqs) with n_freqs = -1 for a raw copy.  -> planes [N, k_pad]."""
    n = segments[0][0].shape[0]
    dev = segments[0][0].device
    srcs = [_lib.f32c(t).reshape(n, -1) for t, _ in segments]
    dst = _planes(max(n, 1), k_pad, dev)
    if n == 0:
        return dst, srcs
    ns = len(segments)
    ptrs = (c_void_p * ns)(*[s.data_ptr() for s in srcs])
    widths = (ctypes.c_int32 * ns)(*[s.shape[1] for s in srcs])
    freqs = (ctypes.c_int32 * ns)(*[f for _, f in segments])
    _lib.check(_lib.raw().nefii_assemble_input(_lib.stream_ptr(dev), n, ns, ptrs, widths, freqs, dst[0].data_ptr(),
                                               dst[1].data_ptr(), k_pad, k_pad))
    return dst, srcs


def transpose_planes(src, rows, cols, rows_pad, cols_pad, col_sum=None):
    dst = _planes(cols_pad, rows_pad, src[0].device)
    _lib.check(_lib.raw().nefii_transpose_planes(
        _lib.stream_ptr(src[0].device), src[0].data_ptr(), src[1].data_ptr(), src[0].stride(0), rows, cols,
        dst[0].data_ptr(), dst[1].data_ptr(), rows_pad, rows_pad, cols_pad, col_sum.data_ptr() if col_sum is not None else None))
    return dst


def segment_width(segments):
    return sum((3 + 6 * f) if f >= 0 else t.shape[-1] for t, f in segments)


def encode_segments(segments):
    """fp32 [N, d_in] of what assemble_input writes as planes: per segment the raw tensor or its positional encoding
    [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...] (reference embedder.py:22-36)."""
    cols = []
    for t, f in segments:
        t = t.detach().float()
        cols.append(t)
        for k in range(max(f, 0)):
            cols += [torch.sin(t * float(2 ** k)), torch.cos(t * float(2 ** k))]
    return torch.cat(cols, dim=-1)


class _DenseMlp(torch.autograd.Function):
    """skip = 0: plain stack.  skip = s > 0: the input of hidden layer s is cat([h_{s-1}, input]) * skip_scale
    (ImplicitNetwork.forward, implicit_differentiable_renderer.py:97-101: `x = torch.cat([x, input], 1) / np.sqrt(2)`);
    layer s-1 then has width - d_in outputs."""

    @staticmethod
    def forward(ctx, act, seg_freqs, n_hidden, skip, skip_scale, want_hidden, need_grad, *tensors):
        n_seg = len(seg_freqs)
        seg_src = tensors[:n_seg]
        params = tensors[n_seg:]
        Ws = [params[2 * l] for l in range(n_hidden + 1)]
        bs = [params[2 * l + 1] for l in range(n_hidden + 1)]
        dev = seg_src[0].device
        n = seg_src[0].shape[0]
        segments = list(zip(seg_src, seg_freqs))
        d_in = segment_width(segments)
        assert d_in == Ws[0].shape[1], (d_in, Ws[0].shape)
        n_out = Ws[-1].shape[0]
        assert n_out <= 4
        y = torch.empty(n, n_out, device=dev, dtype=torch.float32)
        ctx.n_seg = n_seg
        hidden = torch.empty(n if want_hidden else 0, Ws[-1].shape[1], device=dev, dtype=torch.float32)
        ctx.mark_non_differentiable(hidden)
        if n == 0:
            ctx.empty = True
            ctx.shapes = [p.shape for p in params]
            return y, hidden
        if NATIVE_STACK and not skip and not want_hidden:
            srcs = [_lib.f32c(t).reshape(n, -1) for t in seg_src]
            st = _NativeStack(act, srcs, seg_freqs, [_lib.f32c(w) for w in Ws], [_lib.f32c(b) for b in bs], need_grad)
            st.forward(y)
            ctx.empty = False
            ctx.native = st if need_grad else None
            ctx.acts = True
            if need_grad:
                ctx.needs = [ctx.needs_input_grad[7 + n_seg + i] for i in range(len(params))]
            return y, hidden
        ctx.native = None
        k0 = ops.round_up(d_in, 64)
        in0, _keep = assemble_input(segments, k0)
        w_last = _lib.f32c(Ws[-1])
        b_last = _lib.f32c(bs[-1])
        acts = [in0]
        packed = []
        for l in range(n_hidden):
            W = _lib.f32c(Ws[l])
            out_dim, in_dim = W.shape
            k_pad = ops.round_up(in_dim, 64)
            wp = ops.split_to_planes(W, rows_pad=ops.round_up(out_dim, 256), cols_pad=k_pad)
            packed.append(W)
            bias = _lib.f32c(bs[l])
            a = acts[-1]
            if skip and l == skip - 1:
                # [h * scale | input * scale | 0]: the GEMM writes the first out_dim columns, the encoding goes next to them
                assert l < n_hidden - 1
                wide = ops.round_up(out_dim + d_in, 64)
                dst = (torch.zeros(n, wide, device=dev, dtype=torch.bfloat16), torch.zeros(n, wide, device=dev, dtype=torch.bfloat16))
                enc = encode_segments(segments) * skip_scale
                enc_hi = enc.to(torch.bfloat16)
                dst[0][:, out_dim:out_dim + d_in] = enc_hi
                dst[1][:, out_dim:out_dim + d_in] = (enc - enc_hi.float()).to(torch.bfloat16)
                ops.gemm_split_bf16(a, wp, k_pad, out_dim, act=act, bias=bias, out_scale=skip_scale, dst=dst, dst_ncols=out_dim,
                                    k_flush=FWD_FLUSH)
            elif l < n_hidden - 1:
                dst = _planes(n, ops.round_up(out_dim, 64), dev)
                ops.gemm_split_bf16(a, wp, k_pad, out_dim, act=act, bias=bias, dst=dst, dst_ncols=out_dim,
                                    dst_zero_to=ops.round_up(out_dim, 64), k_flush=FWD_FLUSH)
            else:
                dst = _planes(n, ops.round_up(out_dim, 64), dev) if (need_grad or want_hidden) else None
                ops.gemm_split_bf16(a, wp, k_pad, out_dim, act=act, bias=bias, dst=dst, dst_ncols=out_dim if dst else 0,
                                    dst_zero_to=ops.round_up(out_dim, 64) if dst else 0,
                                    w_last=w_last, b_last=b_last, dst_last=y, k_flush=FWD_FLUSH)
            if need_grad or l < n_hidden - 1:
                if not need_grad:
                    acts = [dst]          # ping-pong: drop what is no longer needed
                else:
                    acts.append(dst)
        ctx.empty = False
        if need_grad:
            ctx.act, ctx.n_hidden, ctx.n = act, n_hidden, n
            ctx.skip, ctx.skip_scale = skip, skip_scale
            ctx.acts = acts                # in0, h_1 .. h_L (planes)
            ctx.weights = packed           # effective fp32 hidden weights
            ctx.w_last = w_last
            ctx.needs = [ctx.needs_input_grad[7 + n_seg + i] for i in range(len(params))]
        if want_hidden:
            # the last hidden activation (ImplicitNetwork's feature vector), rebuilt from its two bf16 planes; no gradient
            torch.add(dst[0][:, :hidden.shape[1]].float(), dst[1][:, :hidden.shape[1]].float(), out=hidden)
        return y, hidden

    @staticmethod
    def backward(ctx, gy, _g_hidden=None):
        n_lead = 7
        if ctx.empty:
            return (None,) * (n_lead + ctx.n_seg) + tuple(torch.zeros(s, device=gy.device) for s in ctx.shapes)
        if ctx.acts is None:
            raise _lib.NefiiError("dense_mlp: backward called a second time; the activation planes are released after the first "
                                  "pass (retain_graph is not supported on this path)")
        if ctx.native is not None:
            grads = ctx.native.backward(_lib.f32c(gy))
            ctx.native = ctx.acts = None
            return (None,) * (n_lead + ctx.n_seg) + tuple(g if need else None for g, need in zip(grads, ctx.needs))
        act, L, n = ctx.act, ctx.n_hidden, ctx.n
        dev = gy.device
        gy = _lib.f32c(gy)
        lib = _lib.raw()
        sp = _lib.stream_ptr(dev)
        acts = ctx.acts
        width = ctx.w_last.shape[1]
        n_out = ctx.w_last.shape[0]
        gw_last = torch.zeros_like(ctx.w_last)
        gb_last = torch.zeros(n_out, device=dev)
        hL = acts[L]
        G = _planes(n, hL[0].shape[1], dev)
        _lib.check(lib.nefii_last_layer_bwd(sp, act, n, width, n_out, gy.data_ptr(), ctx.w_last.data_ptr(),
                                            hL[0].data_ptr(), hL[1].data_ptr(), hL[0].stride(0),
                                            G[0].data_ptr(), G[1].data_ptr(), G[0].stride(0),
                                            gw_last.data_ptr(), gb_last.data_ptr()))
        n_pad = ops.round_up(n, 64)
        grads_w, grads_b = [None] * L, [None] * L
        for l in range(L - 1, -1, -1):
            W = ctx.weights[l]
            out_dim, in_dim = W.shape
            gb = torch.zeros(out_dim, device=dev)
            GT = transpose_planes(G, n, out_dim, n_pad, ops.round_up(out_dim, 128), col_sum=gb)
            h_prev = acts[l]
            HT = transpose_planes(h_prev, n, in_dim, n_pad, ops.round_up(in_dim, 256))
            n_sms = _num_sms(dev)
            splits = max(1, min(n_pad // 128, (n_sms + (GT[0].shape[0] // 128) - 1) // (GT[0].shape[0] // 128)))
            partial = torch.empty(splits, out_dim, in_dim, device=dev, dtype=torch.float32)
            used = ops.gemm_split_bf16(GT, HT, n_pad, in_dim, dst_f32=partial, f32_begin=0, f32_end=in_dim,
                                       f32_ld=in_dim, k_splits=splits, f32_split_stride=out_dim * in_dim,
                                       rows_cap=out_dim)
            gW = torch.empty(out_dim, in_dim, device=dev, dtype=torch.float32)
            _lib.check(lib.nefii_reduce_splits(sp, partial.data_ptr(), used, out_dim * in_dim, out_dim, in_dim, in_dim,
                                               gW.data_ptr()))
            grads_w[l], grads_b[l] = gW, gb
            if l > 0:
                wt = ops.split_to_planes(W, rows_pad=ops.round_up(in_dim, 256), cols_pad=ops.round_up(out_dim, 64), transpose=True)
                if ctx.skip and l == ctx.skip:
                    # only the h part of this layer's input carries on; it was stored scaled, and d(h * scale)/dh = scale
                    prev_dim = ctx.weights[l - 1].shape[0]
                    G_prev = _planes(n, ops.round_up(prev_dim, 64), dev)
                    ops.gemm_split_bf16(G, wt, ops.round_up(out_dim, 64), prev_dim, mode=1, act=act, out_scale=ctx.skip_scale,
                                        dst=G_prev, dst_ncols=prev_dim, dst_zero_to=ops.round_up(prev_dim, 64), sav=h_prev,
                                        sav_ncols=prev_dim, sav_scale=1.0 / ctx.skip_scale)
                else:
                    G_prev = _planes(n, ops.round_up(in_dim, 64), dev)
                    ops.gemm_split_bf16(G, wt, ops.round_up(out_dim, 64), in_dim, mode=1, act=act, dst=G_prev, dst_ncols=in_dim,
                                        dst_zero_to=ops.round_up(in_dim, 64), sav=h_prev, sav_ncols=in_dim)
                G = G_prev
        out = [None] * (n_lead + ctx.n_seg)
        for l in range(L):
            out += [grads_w[l] if ctx.needs[2 * l] else None, grads_b[l] if ctx.needs[2 * l + 1] else None]
        out += [gw_last if ctx.needs[2 * L] else None, gb_last if ctx.needs[2 * L + 1] else None]
        ctx.acts = None
        return tuple(out)


def dense_mlp(segments, weights, biases, act, skip=0, skip_scale=1.0, want_hidden=False):
    """segments: [(tensor [N,w], n_freqs | -1)], weights/biases: hidden layers then the output layer.
    Returns the output layer's raw result [N, n_out] (n_out <= 4); with want_hidden also the last hidden activation
    [N, width] (fp32, detached).  skip / skip_scale: see _DenseMlp."""
    seg_src = [t for t, _ in segments]
    if torch.is_grad_enabled() and any(t.requires_grad for t in seg_src):
        # inputs that carry a graph (a trainable geometry: points from SampleNetwork, normals = d sdf/dx, features): the same
        # stack as a composition of differentiable pieces -- every Linear on the tcgen05 layer GEMM through gemm_nt (defined
        # below), encoding / bias / activation as torch ops.  Slower than the fused path; step 2 (frozen geometry) never takes it.
        return _dense_mlp_autograd(segments, weights, biases, act, skip, skip_scale, want_hidden)
    seg_freqs = tuple(f for _, f in segments)
    params = []
    for w, b in zip(weights, biases):
        params += [w, b]
    # bias Parameters keep requires_grad under torch.no_grad(): the activation planes are only kept when a backward can follow
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    y, hidden = _DenseMlp.apply(act, seg_freqs, len(weights) - 1, int(skip), float(skip_scale), bool(want_hidden), need_grad,
                                *seg_src, *params)
    return (y, hidden) if want_hidden else y


# --------------------------------------------------------------------------------------------------------------------------
# A matrix product that can be differentiated any number of times (the eikonal / un-frozen-geometry path)
# --------------------------------------------------------------------------------------------------------------------------
def _gemm_nt_raw(a, b):
    """a [M,K] @ b[N,K]^T -> fp32 [M,N] on the tcgen05 layer GEMM (operands split into bf16 hi/lo planes here)."""
    a, b = _lib.f32c(a), _lib.f32c(b)
    m, k = a.shape
    n = b.shape[0]
    out = torch.empty(m, n, device=a.device, dtype=torch.float32)
    if m == 0 or n == 0:
        return out
    if k == 0:
        return out.zero_()
    k_pad = ops.round_up(k, 64)
    ap = ops.split_to_planes(a, rows_pad=m, cols_pad=k_pad)
    bp = ops.split_to_planes(b, rows_pad=ops.round_up(n, 256), cols_pad=k_pad)
    ops.gemm_split_bf16(ap, bp, k_pad, n, dst_f32=out, f32_begin=0, f32_end=n, f32_ld=n, rows_cap=m)
    return out


class _GemmNT(torch.autograd.Function):
    """y = a @ b^T.  Its backward is two more products of the same kind, built from the same Function, so autograd can
    differentiate it again: this is what `ImplicitNetwork.gradient(create_graph=True)` (the eikonal term and the normals of a
    trainable geometry, reference implicit_differentiable_renderer.py:110-123) runs on."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return _gemm_nt_raw(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ga = gemm_nt(g, b.t()) if ctx.needs_input_grad[0] else None          # [M,N] @ [N,K]
        gb = gemm_nt(g.t(), a.t()) if ctx.needs_input_grad[1] else None      # [N,M] @ [M,K]
        return ga, gb


def gemm_nt(a, b):
    """a [M,K], b [N,K] -> a @ b^T [M,N]; differentiable to any order w.r.t. both operands."""
    return _GemmNT.apply(a, b)


def _dense_mlp_autograd(segments, weights, biases, act, skip=0, skip_scale=1.0, want_hidden=False):
    cols = []
    for t, f in segments:
        t = t.float()
        cols.append(t)
        for k in range(max(f, 0)):
            cols += [torch.sin(t * float(2 ** k)), torch.cos(t * float(2 ** k))]
    x0 = torch.cat(cols, dim=-1)
    fn = {ops.ACT_NONE: lambda z: z, ops.ACT_SOFTPLUS100: lambda z: torch.nn.functional.softplus(z, beta=100),
          ops.ACT_RELU: torch.relu, ops.ACT_ELU: torch.nn.functional.elu}[act]
    h = x0
    n_lin = len(weights)
    hidden = None
    for l, (w, b) in enumerate(zip(weights, biases)):
        if l == n_lin - 1:
            hidden = h
        if skip and l == skip:
            h = torch.cat([h, x0], dim=-1) * skip_scale
        h = gemm_nt(h, w) + b
        if l < n_lin - 1:
            h = fn(h)
    return (h, hidden.detach()) if want_hidden else h
