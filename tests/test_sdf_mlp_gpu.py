"""GPU: the tcgen05 SDF/feature MLP (forward + closed-form input gradient) against the oracle.

The CUDA path computes each fp32 product as three 16-bit MMAs (hi*hi + hi*lo + lo*hi) on one of two plane formats
(include/nefii_b200.h): the fp16 split (the default of the SDF network: 22 significant bits, SDF within 5e-6 of the f64
value, 4e-7 on average -- torch's own fp32 evaluation is 1.5e-7 on average) and the bf16 split (16 bits, fp32's range;
1e-5 / 2e-6).  The bounds asserted here are what DESIGN.md reports."""
import pytest
import torch

from oracle import mlp

pytestmark = pytest.mark.gpu


def _net(dev, params, fmt=None):
    from nefii_b200 import ops
    net = ops.SdfMlp(n_freqs=params.n_freqs, width=params.W[1].shape[0], n_hidden=params.n_layers - 1,
                     skip_layer=params.skip_layer, device=dev, fmt=fmt)
    net.set_weights([w.to(dev) for w in params.W], [b.to(dev) for b in params.b])
    return net


# (format, max |sdf err|, mean |sdf err|, p99 of the normal's direction error)
FORMATS = {"fp16": (1, 6e-6, 8e-7, 4e-5), "bf16": (0, 5e-5, 4e-6, 1.5e-4)}


def test_default_format_is_the_fp16_split(cuda_device):
    from nefii_b200 import ops
    net = _net(cuda_device, mlp.sdf_init(seed=1, bumps=0.3))
    assert net.format == ops.PLANES_FP16


@pytest.mark.parametrize("fmt", ["fp16", "bf16"])
@pytest.mark.parametrize("n", [1, 127, 4096, 20001])
def test_forward_features_gradient(cuda_device, n, fmt):
    dev = cuda_device
    fmt_id, tol_max, tol_mean, tol_dir = FORMATS[fmt]
    params = mlp.sdf_init(seed=1, bumps=0.3)
    net = _net(dev, params, fmt_id)
    assert net.format == fmt_id
    g = torch.Generator().manual_seed(n)
    x = (torch.rand(n, 3, generator=g) * 1.8 - 0.9).to(dev)
    sdf, feat, grad = net.eval(x, want_feat=True, want_grad=True)
    p64 = params.to(dev, torch.float64)
    ref = mlp.sdf_forward(p64, x.double())
    gref = mlp.sdf_gradient(p64, x.double())
    err = (sdf.double() - ref[:, 0]).abs()
    assert err.max().item() < tol_max, err.max().item()
    if n >= 4096:
        assert err.mean().item() < tol_mean, err.mean().item()
    assert (feat.double() - ref[:, 1:]).abs().max().item() < tol_max
    # direction error of the normal; points where |grad| is tiny (critical points of the random field)
    # are ill-conditioned for the fp32 reference as well, hence the floor on the denominator
    rel = (grad.double() - gref).norm(dim=-1) / gref.norm(dim=-1).clamp_min(0.5)
    assert rel.max().item() < 1e-3, rel.max().item()
    assert rel.kthvalue(max(1, int(0.99 * n)))[0].item() < tol_dir
    # forward-only path (ping-pong workspace) returns the same SDF bit for bit
    sdf2, _, _ = net.eval(x)
    assert torch.equal(sdf, sdf2)


def test_count_and_reference_width_256(cuda_device):
    dev = cuda_device
    params = mlp.sdf_init(seed=2, width=256, bumps=0.2)
    net = _net(dev, params)
    x = (torch.rand(1000, 3) * 1.6 - 0.8).to(dev)
    count = torch.tensor([600], device=dev, dtype=torch.int32)
    sdf, feat, grad = net.eval(x, want_feat=True, want_grad=True, count=count)
    ref = mlp.sdf_forward(params.to(dev, torch.float64), x.double())
    assert (sdf[:600].double() - ref[:600, 0]).abs().max().item() < 5e-5
    gref = mlp.sdf_gradient(params.to(dev, torch.float64), x.double())
    assert ((grad[:600].double() - gref[:600]).norm(dim=-1) / gref[:600].norm(dim=-1).clamp_min(0.5)).max().item() < 1e-3


def test_neus_layout_width_256_features_from_last_linear(cuda_device):
    """confs_sg/conf_neus.conf: 8 x 256, use_last_as_f = False -- the feature vector is rows 1.. of the last Linear
    (reference implicit_differentiable_renderer.py:39-42,85-108).  ImplicitNetwork.forward / .gradient of nefii_b200 with
    that layout against the oracle in float64."""
    from nefii_b200.model.implicit_differentiable_renderer import ImplicitNetwork
    from oracle import mlp
    dev = cuda_device
    params = mlp.sdf_init(seed=5, width=256, bias=0.5, bumps=0.05, d_feat=256)
    net = ImplicitNetwork(256, d_in=3, d_out=1, dims=[256] * 8, geometric_init=True, bias=0.5, skip_in=[4], weight_norm=True,
                          multires=6, use_last_as_f=False).to(dev)
    net.load_state_dict(params.state_dict(""))
    for p in net.parameters():
        p.requires_grad_(False)
    x = (torch.rand(5000, 3, generator=torch.Generator().manual_seed(0)) * 1.6 - 0.8).to(dev)
    p64 = params.to(dev, torch.float64)
    with torch.no_grad():
        out = net(x)
        ref = mlp.sdf_forward(p64, x.double())
        g = net.gradient(x, no_grad=True)[:, 0]
        g_ref = mlp.sdf_gradient(p64, x.double())
    assert out.shape == (5000, 257)
    assert (out[:, 0].double() - ref[:, 0]).abs().max().item() < 5e-5
    assert (out[:, 1:].double() - ref[:, 1:]).abs().max().item() < 1e-4
    rel = (g.double() - g_ref).norm(dim=-1) / g_ref.norm(dim=-1)
    assert rel.kthvalue(int(0.99 * rel.numel()))[0].item() < 5e-4


@pytest.mark.parametrize("n", [127, 4096, 20001])
def test_pe_prologue_mode_is_bit_identical(cuda_device, n):
    """nefii_sdf_set_pe_prologue(1): the encoding computed inside layer 0's GEMM instead of by the encode kernel -- same SDF,
    features and gradient bits (the option is kept for its measurements; the default is the faster encode kernel)."""
    from nefii_b200 import _lib
    dev = cuda_device
    params = mlp.sdf_init(seed=1, bumps=0.3)
    net = _net(dev, params)
    x = (torch.rand(n, 3, generator=torch.Generator().manual_seed(n)) * 1.8 - 0.9).to(dev)
    a = net.eval(x, want_feat=True, want_grad=True)
    a0 = net.eval(x)[0]
    _lib.check(_lib.raw().nefii_sdf_set_pe_prologue(1))
    try:
        b = net.eval(x, want_feat=True, want_grad=True)
        b0 = net.eval(x)[0]
    finally:
        _lib.check(_lib.raw().nefii_sdf_set_pe_prologue(0))
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    assert torch.equal(a0, b0)
