"""Generate tests/golden/*.npz by running the REAL reference (imported from /root/reference).

Run in the build container only:  python -m oracle.make_golden
The fixtures travel to the GPU box; the reference does not.
"""
import os
import sys

import numpy as np
import torch

from . import inputs, ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def golden_sg_render():
    from model.sg_render import render_with_sg   # the reference's own function
    normal, view, albedo = inputs.shading_inputs(1024, seed=0)
    sunrise = torch.from_numpy(np.load(os.path.join(ref_shim.REF_CODE, "envmaps", "sunrise", "sg_128.npy"))).float()
    env1 = torch.from_numpy(np.load(os.path.join(ref_shim.REF_CODE, "envmaps", "envmap1_sg_fit", "tmp_lgtSGs_100.npy"))).float()
    sets = {"sunrise": sunrise, "synthetic": inputs.synthetic_light_sgs(128, seed=0), "envmap1": env1}
    spec = torch.tensor([[0.04, 0.04, 0.04]])
    out = dict(normal=normal.numpy(), view=view.numpy(), albedo=albedo.numpy(), spec=spec.numpy())
    for name, lgt in sets.items():
        out["lgt_" + name] = lgt.numpy()
        for r in inputs.ROUGHNESS_SWEEP:
            rough = torch.tensor([[r]])
            res32 = render_with_sg(lgt, spec, rough, albedo, normal, view)
            res64 = render_with_sg(lgt.double(), spec.double(), rough.double(), albedo.double(), normal.double(), view.double())
            for k in ("sg_rgb", "sg_specular_rgb", "sg_diffuse_rgb"):
                out["%s_r%g_%s_f32" % (name, r, k)] = res32[k].numpy()
                out["%s_r%g_%s_f64" % (name, r, k)] = res64[k].numpy().astype(np.float64)
    # K = 2 base materials with blending weights (exercises the K axis quirk of the diffuse term)
    g = torch.Generator().manual_seed(1)
    rough2 = torch.tensor([[0.3], [0.6]])
    spec2 = torch.tensor([[0.04, 0.04, 0.04], [0.1, 0.2, 0.3]])
    bw = torch.softmax(torch.randn(1024, 2, generator=g), -1)
    res = render_with_sg(sunrise, spec2, rough2, albedo, normal, view, blending_weights=bw)
    out.update(k2_rough=rough2.numpy(), k2_spec=spec2.numpy(), k2_blend=bw.numpy())
    for k in ("sg_rgb", "sg_specular_rgb", "sg_diffuse_rgb"):
        out["k2_" + k] = res[k].numpy()
    np.savez_compressed(os.path.join(OUT, "sg_render_cfg1.npz"), **out)


def main():
    if not ref_shim.available():
        sys.exit("reference checkout not present; golden vectors can only be regenerated in the build container")
    ref_shim.install()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    golden_sg_render()
    for extra in EXTRA:
        extra()
    print("wrote", sorted(os.listdir(OUT)))


EXTRA = []

if __name__ == "__main__":
    main()
