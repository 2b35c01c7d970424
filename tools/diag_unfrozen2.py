"""Where does the idr_rgb loss term's gradient into a TRAINABLE geometry differ from the oracle's?  Compares d L / d (points,
normals, features) at the radiance network's inputs, then the SDF parameter gradients with single paths cut.  Diagnostic."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mlp as omlp, pipeline, ref_harness as rh  # noqa: E402


def rel(a, b):
    return (a - b).norm().item() / (b.norm().item() + 1e-30)


def main():
    from nefii_b200.model.implicit_differentiable_renderer import IDRNetwork
    from nefii_b200.utils.conf import default_model_conf
    dev = torch.device("cuda:0")
    om = rh.small_model(seed=0)
    torch.manual_seed(0)
    net = IDRNetwork(default_model_conf()).to(dev)
    rh.load_oracle_weights(net, om)
    net.unfreeze_geometry()
    net.train()
    om = om.to(dev)
    uv, pose, K = rh.camera_batch(20, 2, seed=1)
    S = uv.shape[1]
    obj = torch.ones(1, S, dtype=torch.bool)
    obj[0, ::5] = False
    g = torch.Generator().manual_seed(7)
    U = torch.rand(4096, 7, generator=g).to(dev)
    vecs = [torch.rand(100, generator=g) for _ in range(2)]
    eik = (torch.rand(S * 2 // 2, 3, generator=g) * 2 - 1).to(dev)
    gt = torch.rand(S, 3, generator=g).to(dev)
    inp = dict(uv=uv.to(dev), pose=pose.to(dev), intrinsics=K.to(dev), object_mask=obj.to(dev))
    for t in om.sdf.W + om.sdf.b + om.radiance.tensors() + om.material.tensors() + [om.lgtSGs]:
        t.requires_grad_(True)

    cap_mine, cap_ref = {}, {}
    rn_fwd = net.rendering_network.forward

    def my_rn(points, normals, view_dirs, feature_vectors=None):
        if 'points' not in cap_mine and points.requires_grad:
            for k, v in (('points', points), ('normals', normals), ('feats', feature_vectors)):
                v.retain_grad()
                cap_mine[k] = v
        out = rn_fwd(points, normals, view_dirs, feature_vectors)
        if 'out' not in cap_mine and points.requires_grad:
            cap_mine['out'] = out
        return out
    net.rendering_network.forward = my_rn
    orf = omlp.radiance_forward

    def ref_rn(p, points, normals, view_dirs, features, *a, **k):
        if 'points' not in cap_ref and points.requires_grad:
            for kk, v in (('points', points), ('normals', normals), ('feats', features)):
                v.retain_grad()
                cap_ref[kk] = v
        out = orf(p, points, normals, view_dirs, features, *a, **k)
        if 'out' not in cap_ref and points.requires_grad:
            cap_ref['out'] = out
        return out
    omlp.radiance_forward = ref_rn

    def idr_term(o):
        m = o['network_object_mask'] & o['object_mask']
        return (o['idr_rgb_values'][m] - gt[m]).abs().mean()

    mine = net.forward_with_uv(inp, uniforms=U, trace_uniforms=vecs[0], eikonal_points=eik)
    idr_term(mine).backward()
    ref = pipeline.forward_with_uv_trainable(om, inp['uv'], inp['pose'], inp['intrinsics'], inp['object_mask'], lambda n: U[:n], eik,
                                             vecs[0], vecs[1])
    idr_term(ref).backward()
    print("n_hit mine %d ref %d" % (cap_mine['points'].shape[0], cap_ref['points'].shape[0]))
    for k in ('points', 'normals', 'feats', 'out'):
        print("value %-8s rel %.2e" % (k, rel(cap_mine[k].detach(), cap_ref[k].detach())))
    for k in ('points', 'normals', 'feats'):
        a, b = cap_mine[k].grad, cap_ref[k].grad
        per = (a - b).norm(dim=-1) / (b.norm(dim=-1) + 1e-30)
        print("grad  %-8s rel %.2e  |ref| %.3e  per-point median %.2e p95 %.2e max %.2e" % (
            k, rel(a, b), b.norm().item(), per.median().item(), per.kthvalue(int(0.95 * per.numel()))[0].item(), per.max().item()))
    for l in (0, 4, 7, 8):
        lin = getattr(net.implicit_network, "lin%d" % l)
        print("L%d bias grad rel %.2e" % (l, rel(lin.bias.grad, om.sdf.b[l].grad)))
    # the radiance network alone, same inputs (the oracle's), gradients w.r.t. inputs: isolates mlp._dense_mlp_autograd
    P, N, F = (cap_ref[k].detach().clone().requires_grad_(True) for k in ('points', 'normals', 'feats'))
    view = pipeline.unit(torch.randn(P.shape[0], 3, generator=torch.Generator().manual_seed(5)).to(dev))
    o1 = rn_fwd(P, N, view, F)
    w = torch.randn(o1.shape, generator=torch.Generator().manual_seed(3)).to(dev)
    (o1 * w).sum().backward()
    P2, N2, F2 = (cap_ref[k].detach().clone().requires_grad_(True) for k in ('points', 'normals', 'feats'))
    o2 = orf(om.radiance, P2, N2, view, F2)
    (o2 * w).sum().backward()
    print("radiance net alone: out rel %.2e; grad points %.2e normals %.2e feats %.2e" % (
        rel(o1.detach(), o2.detach()), rel(P.grad, P2.grad), rel(N.grad, N2.grad), rel(F.grad, F2.grad)))
    per = (P.grad - P2.grad).norm(dim=-1) / (P2.grad.norm(dim=-1) + 1e-30)
    print("   per-point grad points: median %.2e p95 %.2e max %.2e" % (per.median().item(), per.kthvalue(int(0.95 * per.numel()))[0].item(),
                                                                     per.max().item()))


if __name__ == "__main__":
    main()
