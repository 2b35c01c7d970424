"""Parity of the tcgen05 split layer GEMM (C ABI nefii_gemm_split_bf16; plane formats: bf16 split and fp16 split) against
float64 torch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _planes_to_f32(p):
    return p[0].float() + p[1].float()


@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("rows,k,n,act", [
    (128, 64, 256, 0), (1000, 512, 512, 1), (4096, 512, 473, 1), (300, 640, 512, 2), (257, 576, 512, 3),
])
def test_forward_layer(cuda_device, rows, k, n, act, fmt):
    from nefii_b200 import ops
    torch.manual_seed(rows + k + n)
    dev = cuda_device
    x = torch.randn(rows, k, device=dev) * 0.5
    w = torch.randn(n, k, device=dev) * (1.0 / k ** 0.5)
    bias = torch.randn(n, device=dev) * 0.1
    k_pad, n_pad = ops.round_up(k, 64), ops.round_up(n, 256)
    a = ops.split_to_planes(x, cols_pad=k_pad, fmt=fmt)
    b = ops.split_to_planes(w, rows_pad=n_pad, cols_pad=k_pad, fmt=fmt)
    pdt = torch.float16 if fmt == ops.PLANES_FP16 else torch.bfloat16
    dst = (torch.zeros(rows, n_pad, device=dev, dtype=pdt), torch.zeros(rows, n_pad, device=dev, dtype=pdt))
    f32 = torch.zeros(rows, n, device=dev)
    ops.gemm_split_bf16(a, b, k_pad, n, act=act, bias=bias, dst=dst, dst_ncols=n, dst_f32=f32, f32_begin=0, f32_end=n)
    torch.cuda.synchronize()
    z = x.double() @ w.double().t() + bias.double()
    if act == 1:
        ref = torch.nn.functional.softplus(z, beta=100)
    elif act == 2:
        ref = torch.relu(z)
    elif act == 3:
        ref = torch.nn.functional.elu(z)
    else:
        ref = z
    err = (f32.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = 2e-5 if fmt == ops.PLANES_BF16 else 2e-6      # 16 against 22 significant operand bits
    assert err < tol * max(scale, 1.0), (err, scale)
    # planes reproduce the fp32 result to ~2^-16 (bf16 split) / ~2^-22 (fp16 split) relative
    perr = (_planes_to_f32(dst)[:, :n].double() - ref).abs().max().item()
    assert perr < 2 * tol * max(scale, 1.0), perr
    # padding columns of the destination stay untouched
    assert _planes_to_f32(dst)[:, n:].abs().max().item() == 0 if n_pad > n else True


def test_count_limits_rows_and_fused_last(cuda_device):
    from nefii_b200 import ops
    torch.manual_seed(3)
    dev = cuda_device
    rows, k, n = 700, 512, 512
    x = torch.randn(rows, k, device=dev) * 0.3
    w = torch.randn(n, k, device=dev) / k ** 0.5
    bias = torch.zeros(n, device=dev)
    w_last = torch.randn(3, n, device=dev) / n ** 0.5
    b_last = torch.randn(3, device=dev)
    a = ops.split_to_planes(x)
    b = ops.split_to_planes(w)
    count = torch.tensor([333], device=dev, dtype=torch.int32)
    y = torch.full((rows, 3), -7.0, device=dev)
    feat = torch.full((rows, n), -7.0, device=dev)
    seed = (torch.zeros(rows, n, device=dev, dtype=torch.bfloat16), torch.zeros(rows, n, device=dev, dtype=torch.bfloat16))
    ops.gemm_split_bf16(a, b, k, n, act=1, bias=bias, count=count, dst_f32=feat, f32_begin=0, f32_end=n,
                        w_last=w_last, b_last=b_last, dst_last=y, seed=seed)
    torch.cuda.synchronize()
    z = x.double() @ w.double().t()
    h = torch.nn.functional.softplus(z, beta=100)
    yref = h @ w_last.double().t() + b_last.double()
    assert (y[:333].double() - yref[:333]).abs().max().item() < 3e-5
    assert (y[333:] == -7.0).all() and (feat[333:] == -7.0).all()
    sref = w_last[0].double() * torch.sigmoid(100 * z)
    assert ((seed[0].float() + seed[1].float())[:333].double() - sref[:333]).abs().max().item() < 3e-5


@pytest.mark.parametrize("fmt", [0, 1])
def test_backward_layer(cuda_device, fmt):
    from nefii_b200 import ops
    torch.manual_seed(5)
    dev = cuda_device
    rows, k, n = 513, 512, 512          # g [rows, k] @ W[k(out), n(in)]  (B = W^T planes)
    g = torch.randn(rows, k, device=dev)
    w = torch.randn(k, n, device=dev) / k ** 0.5       # layer weight [out=k, in=n]
    h_saved = torch.rand(rows, n, device=dev) * 0.05   # forward activations of the layer input
    a = ops.split_to_planes(g, fmt=fmt)
    bt = ops.split_to_planes(w, transpose=True, fmt=fmt)        # [n, k]
    sav = ops.split_to_planes(h_saved * 0.70710678, fmt=fmt)
    pdt = torch.float16 if fmt == ops.PLANES_FP16 else torch.bfloat16
    dst = (torch.zeros(rows, n, device=dev, dtype=pdt), torch.zeros(rows, n, device=dev, dtype=pdt))
    pe = torch.zeros(rows, 39, device=dev)
    ops.gemm_split_bf16(a, bt, k, n, mode=1, act=1, out_scale=0.70710678, dst=dst, dst_ncols=473,
                        dst_f32=pe, f32_begin=473, f32_end=512, sav=sav, sav_ncols=473, sav_scale=1.41421356)
    torch.cuda.synchronize()
    full = g.double() @ w.double()
    sig = 1 - torch.exp(-100 * (sav[0].float() + sav[1].float()).double() * 1.41421356)
    ref = full.clone()
    ref[:, :473] = full[:, :473] * sig[:, :473] * 0.70710678
    got = (dst[0].float() + dst[1].float()).double()
    assert (got[:, :473] - ref[:, :473]).abs().max().item() < (1e-4 if fmt == ops.PLANES_BF16 else 1e-5)
    assert (got[:, 473:] == 0).all()
    # fp32 side output holds the raw product (no out_scale) for the PE columns
    assert (pe.double() - full[:, 473:]).abs().max().item() < 1e-4


@pytest.mark.parametrize("cluster", [1, 2])
@pytest.mark.parametrize("k_flush", [1, 2, 4])
def test_both_kernel_instantiations_and_partial_lengths(cuda_device, cluster, k_flush):
    """The single-CTA kernel (cta_group::1) and the CTA-pair kernel (cta_group::2) are separate template instantiations, and
    the partial-sum length is a runtime knob: every combination must reproduce the fp64 product (hidden layer with bulk
    stores, ragged last tile, odd number of row tiles) -- whatever the library's defaults are."""
    from nefii_b200 import _lib, ops
    dev = cuda_device
    lib = _lib.raw()
    torch.manual_seed(11)
    rows, k, n = 128 * 5 + 37, 512, 512
    x = torch.randn(rows, k, device=dev) * 0.4
    w = torch.randn(n, k, device=dev) / k ** 0.5
    bias = torch.randn(n, device=dev) * 0.1
    a, b = ops.split_to_planes(x), ops.split_to_planes(w)
    dst = (torch.zeros(rows, n, device=dev, dtype=torch.bfloat16), torch.zeros(rows, n, device=dev, dtype=torch.bfloat16))
    ref = torch.nn.functional.softplus(x.double() @ w.double().t() + bias.double(), beta=100)
    try:
        _lib.check(lib.nefii_gemm_set_cluster(cluster))
        _lib.check(lib.nefii_gemm_set_k_flush(k_flush))
        ops.gemm_split_bf16(a, b, k, n, act=1, bias=bias, dst=dst, dst_ncols=n)
        torch.cuda.synchronize()
    finally:
        _lib.check(lib.nefii_gemm_set_cluster(2))
        _lib.check(lib.nefii_gemm_set_k_flush(4))
    err = (_planes_to_f32(dst).double() - ref).abs().max().item()
    assert err < 4e-5 * max(ref.abs().max().item(), 1.0), err


@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("rows", [1, 127, 300, 5000])
def test_pe_prologue_equals_encoded_planes(cuda_device, rows, fmt):
    """The "PE prologue" (positional encoding computed inside the kernel, straight into the swizzled operand tile) gives the
    same bits as the product on planes that hold the encoding; its side copy lands in the requested plane columns only."""
    from nefii_b200 import mlp, ops
    dev = cuda_device
    torch.manual_seed(rows)
    x = torch.rand(rows, 3, device=dev) * 2 - 1
    w = torch.randn(512, 39, device=dev) / 6
    enc = mlp.encode_segments([(x, 6)])
    a = ops.split_to_planes(enc, cols_pad=64, fmt=fmt)
    b = ops.split_to_planes(w, rows_pad=512, cols_pad=64, fmt=fmt)
    o1 = torch.zeros(rows, 512, device=dev)
    o2 = torch.zeros(rows, 512, device=dev)
    ops.gemm_split_bf16(a, b, 64, 512, dst_f32=o1, f32_begin=0, f32_end=512)
    pdt = torch.float16 if fmt else torch.bfloat16
    side = (torch.zeros(rows, 512, device=dev, dtype=pdt), torch.zeros(rows, 512, device=dev, dtype=pdt))
    ops.gemm_split_bf16(None, b, 64, 512, dst_f32=o2, f32_begin=0, f32_end=512, pe_x=x, pe_n_freqs=6, pe_side=side, pe_side_col0=473,
                        pe_side_scale=0.5)
    assert torch.equal(o1, o2)
    got = (side[0].float() + side[1].float())
    assert (got[:, 473:] - enc * 0.5).abs().max().item() < (1e-7 if fmt else 4e-6)
    assert got[:, :473].abs().max().item() == 0
