"""GPU: ImplicitNetwork.gradient(x, no_grad=False) -- `create_graph=True`, reference implicit_differentiable_renderer.py:110-123:
the normals carry a graph to the SDF parameters (eikonal term, normals of a trainable geometry).  Every Linear runs on the
tcgen05 layer GEMM through mlp.gemm_nt, whose backward is built from itself; checked against autograd through the oracle MLP in
float64 on the same device."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _net(dev, width, n_hidden, skip):
    from nefii_b200.model.implicit_differentiable_renderer import ImplicitNetwork
    torch.manual_seed(0)
    net = ImplicitNetwork(width, d_in=3, d_out=1, dims=[width] * n_hidden, geometric_init=True, bias=0.6, skip_in=[skip],
                          weight_norm=True, multires=6, use_last_as_f=True).to(dev)
    with torch.no_grad():                                   # a bumpy blob instead of the exact sphere
        net.lin0.weight_v[:, 3:] = torch.randn_like(net.lin0.weight_v[:, 3:]) * 0.004
        net.lin0.weight_g.copy_(net.lin0.weight_v.norm(dim=1, keepdim=True))
    return net


def _oracle_params(net, dtype):
    from nefii_b200.model.implicit_differentiable_renderer import _effective_weight
    from oracle import mlp as omlp
    ws = [_effective_weight(l).detach().to(dtype).requires_grad_(True) for l in net._layers()]
    bs = [l.bias.detach().to(dtype).requires_grad_(True) for l in net._layers()]
    return omlp.SdfParams(ws, bs, n_freqs=net.multires, skip_layer=net.skip_in[0]), ws, bs


@pytest.mark.parametrize("width,n_hidden,skip,n", [(512, 8, 4, 4096), (256, 4, 2, 777)])
def test_eikonal_gradients_reach_the_parameters(cuda_device, width, n_hidden, skip, n):
    from nefii_b200.model.implicit_differentiable_renderer import _effective_weight
    from oracle import mlp as omlp
    dev = cuda_device
    net = _net(dev, width, n_hidden, skip)
    x = (torch.rand(n, 3, generator=torch.Generator().manual_seed(1)) * 1.4 - 0.7).to(dev)
    g = net.gradient(x.clone(), no_grad=False)
    assert g.shape == (n, 1, 3) and g.requires_grad
    # value: the same normals as the fused inference chain
    with torch.no_grad():
        g_fused = net.gradient(x.clone(), no_grad=True)
    assert (g.detach() - g_fused).norm(dim=-1).max().item() < 1e-3 * g_fused.norm(dim=-1).max().item()
    eik = ((g.norm(2, dim=-1) - 1) ** 2).mean()              # loss.py:188-193
    # gradients w.r.t. the EFFECTIVE weights (what the oracle holds): hook them through torch's weight-norm fold
    layers = net._layers()
    grads = torch.autograd.grad(eik, [l.weight_v for l in layers] + [l.bias for l in layers], allow_unused=True)
    params, ws, bs = _oracle_params(net, torch.float64)
    xx = x.double().requires_grad_(True)
    y = omlp.sdf_forward(params, xx)[:, :1]
    gg = torch.autograd.grad(y, xx, torch.ones_like(y), create_graph=True)[0]
    eik_ref = ((gg.unsqueeze(1).norm(2, dim=-1) - 1) ** 2).mean()
    assert abs(eik.item() - eik_ref.item()) < 1e-4 * max(1e-3, abs(eik_ref.item())) + 1e-7
    ref = torch.autograd.grad(eik_ref, ws + bs, allow_unused=True)
    L = len(layers)
    for l in range(L):
        gW = ref[l]
        if gW is None:
            assert grads[l] is None or grads[l].abs().max().item() == 0
            continue
        # d/d weight_v of g * v / |v| at g == |v|: the tangential part of dL/dW
        v = _effective_weight(layers[l]).detach().double()
        nrm = v.norm(dim=1, keepdim=True)
        gv = gW - (gW * v).sum(1, keepdim=True) * v / (nrm * nrm)
        rel = (grads[l].double() - gv).norm().item() / (gv.norm().item() + 1e-30)
        assert rel < 1e-3, (l, rel)                          # north_star: gradients rel 1e-3
    for l in range(L):
        gb = ref[L + l]
        if gb is None or gb.norm().item() == 0:
            continue
        rel = (grads[L + l].double() - gb).norm().item() / gb.norm().item()
        assert rel < 1e-3, ("bias", l, rel)


def test_forward_with_an_input_that_requires_grad(cuda_device):
    """points that depend on parameters (SampleNetwork's output): forward keeps the graph through the input"""
    from oracle import mlp as omlp
    dev = cuda_device
    net = _net(dev, 512, 8, 4)
    for p in net.parameters():
        p.requires_grad_(False)
    x = (torch.rand(1000, 3, generator=torch.Generator().manual_seed(2)) * 1.4 - 0.7).to(dev).requires_grad_(True)
    out = net(x)
    assert out.shape == (1000, 513) and out.requires_grad
    (gx,) = torch.autograd.grad(out[:, 0].sum() + out[:, 1:].pow(2).sum() * 1e-3, x)
    params, _, _ = _oracle_params(net, torch.float64)
    xx = x.detach().double().requires_grad_(True)
    o = omlp.sdf_forward(params, xx)
    (gref,) = torch.autograd.grad(o[:, 0].sum() + o[:, 1:].pow(2).sum() * 1e-3, xx)
    assert (out.detach().double() - o.detach()).abs().max().item() < 1e-4
    rel = (gx.double() - gref).norm(dim=-1) / gref.norm(dim=-1)
    assert rel.max().item() < 1e-3
