"""Which piece carries the tail of the sg_rgb error at configs[2] size?  Swaps single pieces of the CUDA path for the oracle's
fp32 torch arithmetic (diagnostic; test infrastructure)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mlp as omlp  # noqa: E402
from tests import parity_util  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    from nefii_b200.model import implicit_differentiable_renderer as idr
    base = parity_util.fullsize_compare(dev, 0.08, 2048, 64, True, grads=False, verbose=False)
    print("baseline: sg_rgb lanes %.4f p95 %.2e p99 %.2e | idr lanes %.4f" % (
        base['keys']['sg_rgb_values']['frac_1e4'], base['keys']['sg_rgb_values']['q'][1], base['keys']['sg_rgb_values']['q'][2],
        base['keys']['idr_rgb_values']['frac_1e4']))
    # 1. radiance network forward in torch fp32 (same weights)
    orig_fwd = idr.RenderingNetwork.forward

    def torch_radiance(self, points, normals, view_dirs, feature_vectors=None):
        layers = [getattr(self, "lin%d" % l) for l in range(self.num_layers - 1)]
        p = omlp.DenseParams([idr._effective_weight(l).detach() for l in layers], [l.bias.detach() for l in layers])
        x = omlp.radiance_forward(p, points, normals, view_dirs, feature_vectors)
        return x
    idr.RenderingNetwork.forward = torch_radiance
    try:
        r = parity_util.fullsize_compare(dev, 0.08, 2048, 64, True, grads=False, verbose=False)
    finally:
        idr.RenderingNetwork.forward = orig_fwd
    print("radiance MLP in torch fp32: sg_rgb lanes %.4f p95 %.2e p99 %.2e | idr lanes %.4f" % (
        r['keys']['sg_rgb_values']['frac_1e4'], r['keys']['sg_rgb_values']['q'][1], r['keys']['sg_rgb_values']['q'][2],
        r['keys']['idr_rgb_values']['frac_1e4']))


def per_ray():
    """single-ray pixels: what differs on the rays whose sg_rgb is off by more than 1e-4?"""
    dev = torch.device("cuda:0")
    r = parity_util.fullsize_compare(dev, 0.08, 32768, 0, True, grads=False, verbose=False)
    mine, ref = r['_outputs']
    hit = mine['network_object_mask'] & ref['network_object_mask']
    idx = torch.nonzero(hit).squeeze(1)
    rel = ((mine['sg_rgb_values'] - ref['sg_rgb_values']).abs() / (ref['sg_rgb_values'].abs() + 1e-6))[idx].amax(-1)
    dp = (mine['points'] - ref['points'])[idx].abs().amax(-1)
    dn = (mine['normal_values'] - ref['normal_values'])[idx].abs().amax(-1)
    # secondary rays are stored per HIT ray in hit order: [3, n_hit, ...]
    sm, sr = mine['secondary_mask'], ref['secondary_mask']
    n_hit_m, n_hit_r = sm.shape[1], sr.shape[1]
    print("hits: common %d, ours %d, oracle %d" % (idx.numel(), n_hit_m, n_hit_r))
    if n_hit_m == n_hit_r == idx.numel():
        mism = (sm != sr).reshape(3, -1).any(0)
        both = (sm & sr).reshape(3, -1)
        d2 = (mine['secondary_points'] - ref['secondary_points']).abs().amax(-1)
        d2 = torch.where(both, d2, torch.zeros_like(d2)).amax(0)
        dirs = (mine['secondary_dir'] - ref['secondary_dir']).abs().amax(-1).amax(0)
        bad = rel > 1e-4
        print("rays with sg_rgb rel err > 1e-4: %d of %d" % (int(bad.sum()), bad.numel()))
        for name, x, thr in (("depth |dp| > 1e-6", dp, 1e-6), ("normal |dn| > 1e-4", dn, 1e-4), ("secondary depth diff > 1e-4", d2, 1e-4),
                             ("secondary depth diff > 1e-5", d2, 1e-5), ("secondary dir diff > 1e-4", dirs, 1e-4)):
            print("   %-30s among bad rays %.3f | among good rays %.4f" % (name, (x[bad] > thr).float().mean().item(), (x[~bad] > thr).float().mean().item()))
        print("   secondary mask mismatch        among bad rays %.3f | among good rays %.4f" % (mism[bad].float().mean().item(), mism[~bad].float().mean().item()))
        explained = (dp > 1e-6) | (dn > 1e-4) | (d2 > 1e-5) | mism
        print("   bad rays with none of the above: %d" % int((bad & ~explained).sum()))
        # the worst few
        worst = rel.argsort(descending=True)[:8]
        for w in worst.tolist():
            print("   ray %6d rel %.2e dp %.2e dn %.2e sec depth %.2e sec dir %.2e mask mism %d" % (idx[w].item(), rel[w].item(), dp[w].item(), dn[w].item(), d2[w].item(), dirs[w].item(), int(mism[w])))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "per_ray":
        per_ray()
    else:
        main()
