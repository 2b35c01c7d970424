"""One launch of every shading / integrator kernel at bench-like sizes, for `ncu --set full -k regex:"sg_render|mis_|background_sg"`
(test infrastructure: inputs come from oracle/inputs.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import inputs  # noqa: E402


def main():
    from nefii_b200 import integrator
    from nefii_b200.model.sg_render import render_with_sg
    dev = torch.device("cuda:0")
    lgt = inputs.synthetic_light_sgs(128, seed=2).to(dev).requires_grad_(True)
    # render_with_sg: 1 Mi rays, 128 SGs (BASELINE configs[0] scaled up), forward + backward
    n = 1 << 20
    normal, view, albedo = [x.to(dev) for x in inputs.shading_inputs(n, seed=0)]
    spec = torch.full((1, 3), 0.04, device=dev)
    rough = torch.tensor([[0.3]], device=dev, requires_grad=True)
    out = render_with_sg(lgt, spec, rough, albedo.requires_grad_(True), normal, view)
    out['sg_rgb'].sum().backward()
    # integrator: 131 072 surface points x 3 secondary directions (one bench step's worth)
    n = 131072
    normal, view, albedo = [x.to(dev) for x in inputs.shading_inputs(n, seed=1)]
    g = torch.Generator().manual_seed(1)
    r = (torch.rand(n, 1, generator=g) * 0.9 + 0.089).to(dev).requires_grad_(True)
    u = torch.rand(n, 7, generator=g).to(dev)
    wi, pdf, weight, _ = integrator.mis_sample(lgt.detach(), r.detach(), normal, view, u)
    hit = (torch.rand(3, n, generator=g) < 0.3).to(dev)
    indirect = torch.rand(3, n, 3, generator=g).to(dev).requires_grad_(True)
    sr = torch.full((1, 3), 0.04, device=dev)
    rgb = integrator.mis_shade(lgt, sr, r, albedo.requires_grad_(True), normal, view, wi, pdf, weight, hit, indirect)
    rgb["sg_rgb"].sum().backward()
    bg = integrator.background_sg(lgt, view)
    bg.sum().backward()
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
