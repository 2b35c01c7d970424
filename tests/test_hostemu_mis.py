"""CPU: csrc/mis_math.cuh compiled for the host against the oracle (sampling, shading, and the
hand-derived shading backward against autograd through the oracle, all in float64)."""
import ctypes

import pytest
import torch

from oracle import inputs, mis
from tests.util import hostemu


@pytest.fixture(scope="module")
def emu():
    return hostemu().lib()


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _setup(n, dtype, seed=0):
    normal, view, albedo = [x.to(dtype) for x in inputs.shading_inputs(n, seed=seed)]
    g = torch.Generator().manual_seed(seed + 1)
    rough = (torch.rand(n, 1, generator=g) * 0.9 + 0.089).to(dtype)
    lgt = inputs.synthetic_light_sgs(48, seed=seed + 2).to(dtype)
    u = torch.rand(n, 7, generator=g).to(dtype)
    return lgt, rough, normal, view, albedo, u, g


def test_sampling_matches_oracle_f64(emu):
    n, dt = 400, torch.float64
    lgt, rough, normal, view, albedo, u, _ = _setup(n, dt)
    wi = torch.empty(3, n, 3, dtype=dt); pdf = torch.empty(3, n, dtype=dt); wt = torch.empty(3, n, dtype=dt)
    mat = torch.empty(3, 3, n, dtype=dt)
    emu.emu_mis_sample_f64(n, lgt.shape[0], _p(lgt), _p(rough), _p(normal), _p(view), _p(u), _p(wi), _p(pdf), _p(wt), _p(mat))
    o_wi, o_pdf, o_mat = mis.sample_directions(lgt, rough, normal, view, u)
    assert torch.allclose(wi, o_wi, rtol=1e-9, atol=1e-11)
    assert torch.allclose(pdf, o_pdf[..., 0], rtol=1e-9, atol=1e-12)
    assert torch.allclose(mat, o_mat[..., 0], rtol=1e-9, atol=1e-12)
    tot = (o_mat[..., 0] ** 2).sum(1)
    o_wt = torch.stack([o_mat[i, i, :, 0] ** 2 for i in range(3)]) / tot.clamp_min(1e-6)
    assert torch.allclose(wt, o_wt, rtol=1e-9, atol=1e-12)


def test_sampling_f32_lobe_choice(emu):
    """float32: same lobe / direction as the oracle except where a uniform falls within rounding of a CDF step."""
    n, dt = 2000, torch.float32
    lgt, rough, normal, view, albedo, u, _ = _setup(n, dt, seed=3)
    wi = torch.empty(3, n, 3); pdf = torch.empty(3, n); wt = torch.empty(3, n); mat = torch.empty(3, 3, n)
    emu.emu_mis_sample_f32(n, lgt.shape[0], _p(lgt), _p(rough), _p(normal), _p(view), _p(u), _p(wi), _p(pdf), _p(wt), _p(mat))
    o_wi, _, _ = mis.sample_directions(lgt, rough, normal, view, u)
    close = ((wi - o_wi).abs().amax(-1) < 1e-3).float().mean(-1)
    assert close.min().item() > 0.995


def test_shading_forward_and_backward_f64(emu):
    n, dt = 300, torch.float64
    lgt, rough, normal, view, albedo, u, g = _setup(n, dt, seed=5)
    wi, pdf, mat = mis.sample_directions(lgt, rough, normal, view, u)
    tot = (mat[..., 0] ** 2).sum(1)
    weight = (torch.stack([mat[i, i, :, 0] ** 2 for i in range(3)]) / tot.clamp_min(1e-6)).contiguous()
    vis = (torch.rand(3, n, generator=g) > 0.4).to(dt)
    indirect = torch.rand(3, n, 3, generator=g).to(dt)
    spec = torch.rand(n, 3, generator=g).to(dt) * 0.3
    lgt_r, rough_r, alb_r, spec_r, ind_r, n_r = [t.clone().requires_grad_(True) for t in (lgt, rough, albedo, spec, indirect, normal)]
    ref = mis.shade(lgt_r, spec_r, rough_r, alb_r, n_r, view, wi, pdf, mat, vis.unsqueeze(-1), ind_r)
    g_rgb = torch.rand(n, 3, generator=g).to(dt)
    # the emulation feeds the same upstream gradient into the specular and the diffuse estimate
    loss = ((ref["sg_specular_rgb"] + ref["sg_diffuse_rgb"]) * g_rgb).sum()
    loss.backward()
    out = [torch.empty(n, 3, dtype=dt) for _ in range(3)]
    g_rough = torch.empty(n, dtype=dt); g_alb = torch.empty(n, 3, dtype=dt); g_sr = torch.empty(n, 3, dtype=dt)
    g_ind = torch.empty(3, n, 3, dtype=dt); acc = torch.zeros(lgt.shape[0], 7, dtype=dt); g_n = torch.empty(n, 3, dtype=dt)
    wi_c, pdf_c = wi.contiguous(), pdf[..., 0].contiguous()
    emu.emu_mis_shade_f64(n, lgt.shape[0], _p(lgt), _p(spec), 3, _p(rough), _p(albedo), _p(normal), _p(view), _p(wi_c),
                          _p(pdf_c), _p(weight), _p(vis), _p(indirect), _p(out[0]), _p(out[1]), _p(out[2]), _p(g_rgb),
                          _p(g_rough), _p(g_alb), _p(g_sr), _p(g_ind), _p(acc), _p(g_n))
    assert torch.allclose(out[0], ref["sg_rgb"].detach(), rtol=1e-9, atol=1e-12)
    assert torch.allclose(out[1], ref["sg_specular_rgb"].detach(), rtol=1e-9, atol=1e-12)
    assert torch.allclose(out[2], ref["sg_diffuse_rgb"].detach(), rtol=1e-9, atol=1e-12)
    assert torch.allclose(g_rough, rough_r.grad[:, 0], rtol=1e-7, atol=1e-10)
    assert torch.allclose(g_alb, alb_r.grad, rtol=1e-7, atol=1e-10)
    assert torch.allclose(g_sr, spec_r.grad, rtol=1e-7, atol=1e-10)
    assert torch.allclose(g_ind, ind_r.grad, rtol=1e-7, atol=1e-10)
    # d / d normal (only needed when the geometry trains: the normal is d sdf/dx with a graph to the SDF parameters)
    assert torch.allclose(g_n, n_r.grad, rtol=1e-6, atol=1e-9), (g_n - n_r.grad).abs().max()
    # unit-parametrisation accumulator -> raw parameter gradient (what nefii_sg_param_grad does)
    raw = lgt
    ln = raw[:, :3].norm(dim=-1, keepdim=True)
    d = ln + 1e-6
    g_axis = acc[:, :3] / d - raw[:, :3] * (raw[:, :3] * acc[:, :3]).sum(-1, keepdim=True) / (ln * d * d)
    g_raw = torch.cat([g_axis, acc[:, 3:4] * torch.sign(raw[:, 3:4]), acc[:, 4:] * torch.sign(raw[:, 4:])], dim=-1)
    assert torch.allclose(g_raw, lgt_r.grad, rtol=1e-6, atol=1e-9)
