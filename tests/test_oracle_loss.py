"""CPU: the loss oracle (oracle/loss.py) against the golden vectors produced by the REAL reference IDRLoss
(tests/golden/idr_loss.npz, oracle/make_golden.py) and, where /root/reference exists, against the reference run live."""
import os

import numpy as np
import pytest
import torch

from oracle import loss as oloss, ref_shim

GOLD = os.path.join(os.path.dirname(__file__), "golden", "idr_loss.npz")
CASES = {"conf": dict(loss_type='L1', env_loss_type='L2'), "l2": dict(loss_type='L2', env_loss_type='L1'),
         "smooth": dict(loss_type='L1_smooth', env_loss_type='L2')}
TERMS = ('idr_rgb_loss', 'sg_rgb_loss', 'mask_loss', 'normalsmooth_loss', 'background_rgb_loss')


def _inputs(g, tag, grad=False):
    d = {k: torch.from_numpy(g["%s_in_%s" % (tag, k)]) for k in ("idr_rgb", "sg_rgb", "rgb_gt", "normal", "sdf_output", "net_mask", "obj_mask")}
    if grad:
        for k in ("idr_rgb", "sg_rgb", "normal", "sdf_output"):
            d[k].requires_grad_(True)
    return d


@pytest.mark.parametrize("tag", sorted(CASES))
def test_oracle_matches_reference_golden(tag):
    g = np.load(GOLD)
    d = _inputs(g, tag, grad=True)
    terms = oloss.idr_loss_terms(alpha=50.0, r_patch=1, **CASES[tag], **d)
    for k in TERMS:
        assert abs(terms[k].item() - float(g["%s_%s" % (tag, k)])) <= 1e-6 * max(1.0, abs(float(g["%s_%s" % (tag, k)]))), k
    total = oloss.idr_loss(terms)
    assert abs(total.item() - float(g[tag + "_loss"])) < 1e-5
    total.backward()
    for k in ("idr_rgb", "sg_rgb", "normal", "sdf_output"):
        ref = torch.from_numpy(g["%s_grad_%s" % (tag, k)])
        assert torch.allclose(d[k].grad, ref, rtol=1e-5, atol=1e-8), k


def test_empty_masks_give_zero_terms():
    d = oloss.loss_inputs(64, seed=3)
    d['net_mask'][:] = True
    d['obj_mask'][:] = True
    t = oloss.idr_loss_terms(alpha=50.0, **d)
    assert t['mask_loss'].item() == 0 and t['background_rgb_loss'].item() == 0 and t['idr_rgb_loss'].item() > 0
    d['net_mask'][:] = False
    t = oloss.idr_loss_terms(alpha=50.0, **d)
    assert t['idr_rgb_loss'].item() == 0 and t['normalsmooth_loss'].item() == 0 and t['mask_loss'].item() > 0


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout only exists in the build container")
def test_oracle_matches_live_reference():
    ref_shim.install()
    from oracle import make_golden
    inp = oloss.loss_inputs(n_pixels=256, seed=11, hit_frac=0.4)
    res, grads = make_golden.reference_idr_loss(inp, **make_golden.LOSS_CONF)
    d = {k: v.clone() for k, v in inp.items()}
    for k in ("idr_rgb", "sg_rgb", "normal", "sdf_output"):
        d[k].requires_grad_(True)
    terms = oloss.idr_loss_terms(alpha=50.0, r_patch=1, loss_type='L1', env_loss_type='L2', **d)
    for k in TERMS:
        assert abs(terms[k].item() - res[k].item()) < 1e-6, k
    oloss.idr_loss(terms).backward()
    for k in grads:
        assert torch.allclose(d[k].grad, grads[k], rtol=1e-5, atol=1e-8), k
