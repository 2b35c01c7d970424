"""Which loss term's gradient into the SDF network differs between nefii_b200 (trainable geometry) and the oracle?  Diagnostic."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pipeline, ref_harness as rh  # noqa: E402


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
    terms = {
        "eikonal": lambda o, m: ((o['grad_theta'].norm(2, dim=1) - 1) ** 2).mean(),
        "mask": lambda o, m: torch.nn.functional.binary_cross_entropy_with_logits(-50 * o['sdf_output'][~m].reshape(-1), o['object_mask'][~m].float()) / 50,
        "idr_rgb": lambda o, m: (o['idr_rgb_values'][m] - gt[m]).abs().mean(),
        "sg_rgb": lambda o, m: (o['sg_rgb_values'][m] - gt[m]).abs().mean(),
        "sg_diffuse": lambda o, m: (o['sg_diffuse_rgb_values'][m] - gt[m]).abs().mean(),
        "sg_specular": lambda o, m: (o['sg_specular_rgb_values'][m] - gt[m]).abs().mean(),
        "roughness": lambda o, m: (o['sg_roughness_values'][m]).mean(),
        "albedo": lambda o, m: (o['sg_diffuse_albedo_values'][m] - gt[m]).abs().mean(),
        "normal": lambda o, m: (o['normal_values'][m] * gt[m]).sum(-1).mean(),
    }
    for name, fn in terms.items():
        for p in net.parameters():
            p.grad = None
        for t in om.sdf.W + om.sdf.b:
            t.grad = None
        mine = net.forward_with_uv(inp, uniforms=U, trace_uniforms=vecs[0], eikonal_points=eik)
        m = mine['network_object_mask'] & mine['object_mask']
        fn(mine, m).backward()
        ref = pipeline.forward_with_uv_trainable(om, inp['uv'], inp['pose'], inp['intrinsics'], inp['object_mask'], lambda n: U[:n], eik,
                                                 vecs[0], vecs[1])
        mr = ref['network_object_mask'] & ref['object_mask']
        fn(ref, mr).backward()
        out = []
        for l in (0, 4, 8):
            lin = getattr(net.implicit_network, "lin%d" % l)
            a, b = lin.bias.grad, om.sdf.b[l].grad
            if a is None or b is None:
                out.append("L%d bias None/None" % l)
                continue
            out.append("L%d bias rel %.2e (|ref| %.2e)" % (l, (a - b).norm().item() / (b.norm().item() + 1e-30), b.norm().item()))
        print("%-12s %s" % (name, " | ".join(out)))


if __name__ == "__main__":
    main()
