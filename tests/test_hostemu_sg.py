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
