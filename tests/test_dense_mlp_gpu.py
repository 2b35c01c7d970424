"""GPU: the trainable dense stacks (radiance / material nets) -- forward and weight gradients against torch fp32."""
import pytest
import torch

from oracle import mlp as omlp

pytestmark = pytest.mark.gpu


def _check_grad(got, want):
    """rel 1e-3 in the Frobenius sense; single entries may move more where a ReLU/ELU pre-activation sits within
    float rounding of its kink (the mask flips for any forward that is not bit-identical to cuBLAS)."""
    rel = (got - want).norm().item() / (want.norm().item() + 1e-20)
    assert rel < 1e-3, rel
    assert (got - want).abs().max().item() <= 5e-3 * (want.abs().max().item() + 1e-12)


def _grads(params):
    return [p.grad.clone() for p in params]


@pytest.mark.parametrize("n", [1, 500, 6000])
def test_radiance_net_forward_backward(cuda_device, n):
    from nefii_b200 import mlp, ops
    dev = cuda_device
    g = torch.Generator().manual_seed(n)
    pts = (torch.rand(n, 3, generator=g) - 0.5).to(dev)
    nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(dev)
    view = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(dev)
    feat = torch.rand(n, 512, generator=g).to(dev)
    p = omlp.radiance_init(seed=2).to(dev)
    p.requires_grad_(True)
    gy = torch.rand(n, 3, generator=g).to(dev)
    ref = omlp.radiance_forward(p, pts, nrm, view, feat)
    (ref * gy).sum().backward()
    ref_grads = _grads(p.tensors())
    for t in p.tensors():
        t.grad = None
    raw = mlp.dense_mlp([(pts, 10), (view, 4), (nrm, -1), (feat, -1)], p.W, p.b, ops.ACT_RELU)
    out = raw ** 2
    assert torch.allclose(out, ref.detach(), rtol=2e-4, atol=2e-5)
    (out * gy).sum().backward()
    for got, want in zip(_grads(p.tensors()), ref_grads):
        _check_grad(got, want)


def test_material_net_forward_backward(cuda_device):
    from nefii_b200 import mlp, ops
    dev = cuda_device
    n = 3000
    g = torch.Generator().manual_seed(11)
    pts = (torch.rand(n, 3, generator=g) - 0.5).to(dev)
    feat = torch.rand(n, 512, generator=g).to(dev)
    p = omlp.material_init(seed=3).to(dev)
    p.requires_grad_(True)
    ga, gr = torch.rand(n, 3, generator=g).to(dev), torch.rand(n, 1, generator=g).to(dev)
    a_ref, r_ref = omlp.material_forward(p, pts, feat)
    ((a_ref * ga).sum() + (r_ref * gr).sum()).backward()
    ref_grads = _grads(p.tensors())
    for t in p.tensors():
        t.grad = None
    raw = mlp.dense_mlp([(pts, 10), (feat, -1)], p.W, p.b, ops.ACT_ELU)
    albedo = torch.sigmoid(raw[:, :3])
    rough = (1 - 0.089) * torch.sigmoid(raw[:, 3:4]) + 0.089
    assert torch.allclose(albedo, a_ref.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(rough, r_ref.detach(), rtol=1e-4, atol=1e-5)
    ((albedo * ga).sum() + (rough * gr).sum()).backward()
    for got, want in zip(_grads(p.tensors()), ref_grads):
        _check_grad(got, want)


def test_no_grad_path_and_empty(cuda_device):
    from nefii_b200 import mlp, ops
    dev = cuda_device
    p = omlp.material_init(seed=4).to(dev)
    pts, feat = torch.rand(200, 3, device=dev), torch.rand(200, 512, device=dev)
    with torch.no_grad():
        raw = mlp.dense_mlp([(pts, 10), (feat, -1)], p.W, p.b, ops.ACT_ELU)
    a_ref, _ = omlp.material_forward(p, pts, feat)
    assert torch.allclose(torch.sigmoid(raw[:, :3]), a_ref, rtol=1e-4, atol=1e-5)
    empty = mlp.dense_mlp([(pts[:0], 10), (feat[:0], -1)], p.W, p.b, ops.ACT_ELU)
    assert empty.shape == (0, 4)


@pytest.mark.parametrize("n,need_grad", [(1, True), (777, True), (5000, True), (5000, False)])
def test_native_sequencer_matches_per_kernel_sequence(cuda_device, n, need_grad, monkeypatch):
    """csrc/dense_stack.cu (one host call per direction) runs the same kernels in the same order as the per-kernel ctypes
    sequence of mlp.py: outputs and weight gradients bit-identical, bias gradients (atomic column sums) to 1e-6."""
    from nefii_b200 import mlp, ops
    dev = cuda_device
    g = torch.Generator().manual_seed(100 + n)
    pts = (torch.rand(n, 3, generator=g) - 0.5).to(dev)
    nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(dev)
    view = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(dev)
    feat = torch.rand(n, 512, generator=g).to(dev)
    gy = torch.rand(n, 3, generator=g).to(dev)
    p = omlp.radiance_init(seed=5).to(dev)
    p.requires_grad_(need_grad)
    res = {}
    for native in (1, 0):
        monkeypatch.setattr(mlp, "NATIVE_STACK", native)
        for t in p.tensors():
            t.grad = None
        with torch.set_grad_enabled(need_grad):
            raw = mlp.dense_mlp([(pts, 10), (view, 4), (nrm, -1), (feat, -1)], p.W, p.b, ops.ACT_RELU)
        grads = []
        if need_grad:
            (raw * gy).sum().backward()
            grads = _grads(p.tensors())
        res[native] = (raw.detach().clone(), grads)
    assert torch.equal(res[1][0], res[0][0])
    names = ["W%d" % l for l in range(len(p.W))] + ["b%d" % l for l in range(len(p.b))]      # DenseParams.tensors() order
    assert len(res[1][1]) == len(res[0][1])
    for name, a, b in zip(names, res[1][1], res[0][1]):
        if name.startswith("W") and name != "W%d" % (len(p.W) - 1):
            assert torch.equal(a, b), name
        else:
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6 * (b.abs().max().item() + 1e-12)), name


def test_native_sequencer_second_backward_raises(cuda_device):
    from nefii_b200 import mlp, ops, _lib
    dev = cuda_device
    p = omlp.material_init(seed=6).to(dev)
    p.requires_grad_(True)
    pts, feat = torch.rand(64, 3, device=dev), torch.rand(64, 512, device=dev)
    raw = mlp.dense_mlp([(pts, 10), (feat, -1)], p.W, p.b, ops.ACT_ELU)
    raw.sum().backward(retain_graph=True)
    with pytest.raises(_lib.NefiiError):
        raw.sum().backward()
