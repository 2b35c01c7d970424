"""CPU: the device math header (csrc/sg_math.cuh) compiled for the host, against the oracle.

In float64 the two agree to round-off, which pins the *logic* of the CUDA kernel before it ever
runs on a GPU; float32 parity (operation order, libdevice) is checked on the GPU box."""
import ctypes

import pytest
import torch

from oracle import inputs, sg
from tests.util import hostemu


@pytest.fixture(scope="module")
def emu():
    return hostemu().lib()


def _run(emu, dtype, lgt, spec, rough, albedo, normal, view, blend=None):
    fn = emu.emu_sg_render_fwd_f64 if dtype == torch.float64 else emu.emu_sg_render_fwd_f32
    args = [t.to(dtype).contiguous() for t in (lgt, spec, rough, albedo, normal, view)]
    bl = blend.to(dtype).contiguous() if blend is not None else None
    n = normal.shape[0]
    outs = [torch.empty(n, 3, dtype=dtype) for _ in range(3)]
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    fn(n, lgt.shape[0], spec.shape[0], *[p(a) for a in args], p(bl), *[p(o) for o in outs])
    ref = sg.render_with_sg(args[0], args[1], args[2], args[3], args[4], args[5], blending_weights=bl)
    return outs, [ref["sg_rgb"], ref["sg_specular_rgb"], ref["sg_diffuse_rgb"]]


@pytest.mark.parametrize("rough", inputs.ROUGHNESS_SWEEP)
def test_f64_matches_oracle(emu, rough):
    normal, view, albedo = inputs.shading_inputs(300, seed=1)
    lgt = inputs.synthetic_light_sgs(128, seed=2)
    outs, refs = _run(emu, torch.float64, lgt, torch.full((1, 3), 0.04), torch.tensor([[rough]]), albedo, normal, view)
    for o, r in zip(outs, refs):
        assert torch.allclose(o, r, rtol=1e-7, atol=1e-9)


def test_f64_multi_material(emu):
    normal, view, albedo = inputs.shading_inputs(200, seed=4)
    lgt = inputs.synthetic_light_sgs(32, seed=4)
    spec = torch.tensor([[0.04] * 3, [0.1, 0.2, 0.3], [0.5, 0.4, 0.3]])
    rough = torch.tensor([[0.2], [0.5], [0.9]])
    blend = torch.softmax(torch.randn(200, 3), -1)
    for bl in (None, blend):
        outs, refs = _run(emu, torch.float64, lgt, spec, rough, albedo, normal, view, bl)
        for o, r in zip(outs, refs):
            assert torch.allclose(o, r, rtol=1e-7, atol=1e-9)


def test_f32_close_to_oracle(emu):
    normal, view, albedo = inputs.shading_inputs(300, seed=1)
    lgt = inputs.synthetic_light_sgs(128, seed=2)
    outs, refs = _run(emu, torch.float32, lgt, torch.full((1, 3), 0.04), torch.tensor([[0.5]]), albedo, normal, view)
    for o, r in zip(outs, refs):
        assert torch.allclose(o, r, rtol=2e-3, atol=1e-4)


def test_backward_f64_matches_autograd_of_oracle(emu):
    """render_with_sg backward by the hand-derived adjoint (csrc/sg_adjoint_math.cuh) == autograd through the oracle, the
    gradient w.r.t. the normals included."""
    dt = torch.float64
    n, M, K = 120, 24, 2
    normal, view, albedo = [x.to(dt) for x in inputs.shading_inputs(n, seed=8)]
    lgt = inputs.synthetic_light_sgs(M, seed=9).to(dt)
    lgt[3, 3] = -lgt[3, 3]          # abs() branches of the raw parameters
    lgt[5, 4:] = -lgt[5, 4:]
    spec = torch.tensor([[0.04, 0.05, 0.06], [0.1, 0.2, 0.3]], dtype=dt)
    rough = torch.tensor([[0.35], [0.7]], dtype=dt)
    g = torch.Generator().manual_seed(3)
    g_spec = torch.rand(n, 3, generator=g, dtype=dt)
    g_diff = torch.rand(n, 3, generator=g, dtype=dt)
    leaves = [t.clone().requires_grad_(True) for t in (lgt, spec, rough, albedo, normal)]
    ref = sg.render_with_sg(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], view)
    ((ref["sg_specular_rgb"] * g_spec).sum() + (ref["sg_diffuse_rgb"] * g_diff).sum()).backward()
    acc = torch.zeros(M, 7, dtype=dt)
    g_rough = torch.zeros(K, dtype=dt)
    g_sr = torch.zeros(K, 3, dtype=dt)
    g_alb = torch.zeros(n, 3, dtype=dt)
    g_nrm = torch.zeros(n, 3, dtype=dt)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    emu.emu_sg_render_bwd_f64(n, M, K, p(lgt), p(spec), p(rough), p(albedo.contiguous()), p(normal.contiguous()), p(view.contiguous()),
                              p(g_spec), p(g_diff), p(acc), p(g_rough), p(g_sr), p(g_alb), p(g_nrm), None, None)
    ln = lgt[:, :3].norm(dim=-1, keepdim=True)
    d = ln + 1e-6
    g_axis = acc[:, :3] / d - lgt[:, :3] * (lgt[:, :3] * acc[:, :3]).sum(-1, keepdim=True) / (ln * d * d)
    g_raw = torch.cat([g_axis, acc[:, 3:4] * torch.sign(lgt[:, 3:4]), acc[:, 4:] * torch.sign(lgt[:, 4:])], dim=-1)
    assert torch.allclose(g_raw, leaves[0].grad, rtol=1e-6, atol=1e-9)
    assert torch.allclose(g_sr, leaves[1].grad, rtol=1e-6, atol=1e-9)
    assert torch.allclose(g_rough, leaves[2].grad[:, 0], rtol=1e-6, atol=1e-9)
    assert torch.allclose(g_alb, leaves[3].grad, rtol=1e-6, atol=1e-9)
    assert torch.allclose(g_nrm, leaves[4].grad, rtol=1e-6, atol=1e-8), (g_nrm - leaves[4].grad).abs().max()


def test_backward_f64_with_blending_weights(emu):
    """K = 3 base materials with per-point blending weights (sg_render.py:254-256): gradients incl. d / d blending weight."""
    dt = torch.float64
    n, M, K = 90, 16, 3
    normal, view, albedo = [x.to(dt) for x in inputs.shading_inputs(n, seed=18)]
    lgt = inputs.synthetic_light_sgs(M, seed=19).to(dt)
    spec = torch.tensor([[0.04, 0.05, 0.06], [0.1, 0.2, 0.3], [0.5, 0.4, 0.3]], dtype=dt)
    rough = torch.tensor([[0.35], [0.7], [0.2]], dtype=dt)
    g = torch.Generator().manual_seed(5)
    blend = torch.softmax(torch.randn(n, K, generator=g, dtype=dt), -1)
    g_spec = torch.rand(n, 3, generator=g, dtype=dt)
    g_diff = torch.rand(n, 3, generator=g, dtype=dt)
    leaves = [t.clone().requires_grad_(True) for t in (lgt, spec, rough, albedo, normal, blend)]
    ref = sg.render_with_sg(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], view, blending_weights=leaves[5])
    ((ref["sg_specular_rgb"] * g_spec).sum() + (ref["sg_diffuse_rgb"] * g_diff).sum()).backward()
    acc = torch.zeros(M, 7, dtype=dt)
    g_rough = torch.zeros(K, dtype=dt)
    g_sr = torch.zeros(K, 3, dtype=dt)
    g_alb = torch.zeros(n, 3, dtype=dt)
    g_nrm = torch.zeros(n, 3, dtype=dt)
    g_bl = torch.zeros(n, K, dtype=dt)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    emu.emu_sg_render_bwd_f64(n, M, K, p(lgt), p(spec), p(rough), p(albedo.contiguous()), p(normal.contiguous()), p(view.contiguous()),
                              p(g_spec), p(g_diff), p(acc), p(g_rough), p(g_sr), p(g_alb), p(g_nrm), p(blend.contiguous()), p(g_bl))
    assert torch.allclose(g_sr, leaves[1].grad, rtol=1e-6, atol=1e-9)
    assert torch.allclose(g_rough, leaves[2].grad[:, 0], rtol=1e-6, atol=1e-9)
    assert torch.allclose(g_alb, leaves[3].grad, rtol=1e-6, atol=1e-9)
    assert torch.allclose(g_nrm, leaves[4].grad, rtol=1e-6, atol=1e-8)
    assert torch.allclose(g_bl, leaves[5].grad, rtol=1e-6, atol=1e-9)
