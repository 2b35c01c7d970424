"""GPU: step-1 geometry fitting (reference training/geometry_train.py:366-376: L1 between ImplicitNetwork(points)[:, 0:1] and
sampled SDF values at batch 16384) on the trainable tcgen05 stack: forward values and the gradients of every weight_g /
weight_v / bias against autograd through the oracle's restatement of ImplicitNetwork.forward with the same parameters
(rel 1e-3, north_star's gradient tolerance), and a few optimizer steps that actually reduce the loss."""
import pytest
import torch

from oracle import mlp as omlp

pytestmark = pytest.mark.gpu


def _net(dev, seed=0):
    from nefii_b200.model.implicit_differentiable_renderer import ImplicitNetwork
    torch.manual_seed(seed)
    return ImplicitNetwork(feature_vector_size=512, d_in=3, d_out=1, dims=[512] * 8, geometric_init=True, bias=0.6, skip_in=[4],
                           weight_norm=True, multires=6, use_last_as_f=True).to(dev)


def _oracle_forward(net, x):
    layers = [getattr(net, "lin%d" % l) for l in range(net.num_layers - 1)]
    W = [omlp.fold_weight_norm(l.weight_g, l.weight_v) for l in layers]
    p = omlp.SdfParams(W, [l.bias for l in layers], n_freqs=6, skip_layer=4)
    return omlp.sdf_forward(p, x)


def _target(x):
    """a shape that is not the initial sphere: sphere of radius 0.5 with bumps"""
    return (x.norm(dim=-1, keepdim=True) - 0.5) + 0.05 * torch.sin(6 * x[:, :1]) * torch.cos(5 * x[:, 1:2])


def test_forward_and_weight_gradients_match_oracle(cuda_device):
    dev = cuda_device
    net = _net(dev)
    with torch.no_grad():                       # move off the exactly-zero initial biases / PE columns
        for p in net.parameters():
            p.add_(0.01 * torch.randn_like(p))
    ref = _net(dev)
    ref.load_state_dict(net.state_dict())
    x = (torch.rand(16384, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 2 - 1) * 0.9
    gt = _target(x)
    net.train()
    out = net(x)
    assert out.shape == (16384, 513)
    loss = torch.nn.functional.l1_loss(out[:, 0:1], gt)
    loss.backward()
    out_ref = _oracle_forward(ref, x)
    loss_ref = torch.nn.functional.l1_loss(out_ref[:, 0:1], gt)
    loss_ref.backward()
    assert (out[:, 0] - out_ref[:, 0]).abs().max().item() < 5e-5
    assert abs(loss.item() - loss_ref.item()) < 1e-5
    assert (out[:, 1:] - out_ref[:, 1:]).abs().max().item() < 2e-3          # feature columns: rebuilt from two bf16 planes
    worst = 0.0
    for (name, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, name
        rel = (p.grad - q.grad).norm().item() / (q.grad.norm().item() + 1e-12)
        worst = max(worst, rel)
        assert rel < 1e-3, (name, rel)
    assert worst > 0          # gradients are not trivially identical objects


def test_a_few_adam_steps_reduce_the_loss(cuda_device):
    dev = cuda_device
    net = _net(dev, seed=1)
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    g = torch.Generator(device=dev).manual_seed(2)
    losses = []
    for it in range(40):
        x = (torch.rand(16384, 3, device=dev, generator=g) * 2 - 1) * 0.9
        loss = torch.nn.functional.l1_loss(net(x)[:, 0:1], _target(x))
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert min(losses[-5:]) < 0.8 * losses[0], losses
    # and the inference path (tracer / renderer) sees the updated weights
    with torch.no_grad():
        x = (torch.rand(4096, 3, device=dev, generator=g) * 2 - 1) * 0.9
        sdf_inf = net(x)[:, 0]
    net_ref_out = _oracle_forward(net, x)[:, 0]
    assert (sdf_inf - net_ref_out).abs().max().item() < 5e-5


def test_eikonal_path_returns_a_graph(cuda_device):
    """gradient(x, no_grad=False) -- create_graph=True in the reference (:110-123) -- carries a graph to the parameters
    (round 1 refused it; the parity of its gradients is in tests/test_second_order_gpu.py)"""
    net = _net(cuda_device)
    net.train()
    g = net.gradient(torch.rand(8, 3, device=cuda_device), no_grad=False)
    assert g.shape == (8, 1, 3) and g.requires_grad
    ((g.norm(2, dim=-1) - 1) ** 2).mean().backward()
    assert net.lin3.weight_v.grad is not None and torch.isfinite(net.lin3.weight_v.grad).all()
