"""CPU: the per-ray march state machine of csrc/tracer_math.cuh (compiled for the host) against the lock-step
sphere_tracing of oracle/tracer.py -- bit-exact accumulated distances, unfinished masks and evaluation counts.

The CUDA tracer lets every ray run through its own iterations / line-search steps (a ray that needs no line search never
waits for one that does); the reference marches all rays together.  This test pins that the two are the same function."""
import ctypes
import unittest.mock as mock

import numpy as np
import pytest
import torch

from oracle import tracer as otr
from tests.util import hostemu


@pytest.fixture(scope="module")
def emu():
    return hostemu().lib()


@pytest.fixture(autouse=True)
def ieee_sqrt():
    """torch's CPU sqrt is a vectorised approximation that is 1 ulp off on ~0.5 % of inputs (CUDA's is IEEE); the bit-exact
    comparison needs the correctly rounded one on both sides, so the oracle runs with numpy's."""
    def sqrt(t):
        return torch.from_numpy(np.sqrt(t.detach().numpy()))
    with mock.patch.object(torch, "sqrt", sqrt):
        yield


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _rays(n_side, seed, f, cam=(0.0, 0.0, -3.0)):
    g = torch.Generator().manual_seed(seed)
    K = torch.eye(4); K[0, 0] = K[1, 1] = f; K[0, 2] = n_side / 2; K[1, 2] = n_side / 2
    pose = torch.eye(4); pose[:3, 3] = torch.tensor(cam)
    ii, jj = torch.meshgrid(torch.arange(n_side).float(), torch.arange(n_side).float(), indexing="xy")
    uv = torch.stack([ii, jj], -1).reshape(1, -1, 2) + torch.rand(1, n_side * n_side, 2, generator=g)
    dirs, loc = otr.camera_rays(uv, pose[None], K[None])
    n = n_side * n_side
    return loc.expand(n, 3).contiguous(), dirs.reshape(n, 3).contiguous()


def _run(emu, o, d, prims, cfg):
    n = o.shape[0]
    acc_s, acc_e, mn, mx = [torch.empty(n) for _ in range(4)]
    flags = torch.empty(n, dtype=torch.uint8)
    stats = (ctypes.c_longlong * 2)()
    emu.emu_sphere_trace(n, _p(o), _p(d), _p(prims), prims.shape[0], ctypes.c_float(cfg.object_bounding_sphere),
                         ctypes.c_float(cfg.sdf_threshold), ctypes.c_float(cfg.line_search_step), cfg.line_step_iters,
                         cfg.sphere_tracing_iters, _p(acc_s), _p(acc_e), _p(mn), _p(mx), _p(flags), stats)
    return acc_s, acc_e, mn, mx, flags, stats[0], stats[1]


def test_analytic_evaluator_bit_exact(emu):
    prims = otr.robot_scene()
    x = torch.rand(5000, 3, generator=torch.Generator().manual_seed(0)) * 2 - 1
    out = torch.empty(5000)
    emu.emu_analytic_sdf(5000, _p(x), _p(prims), prims.shape[0], _p(out))
    assert torch.equal(out, otr.analytic_sdf(prims)(x))


@pytest.mark.parametrize("n_side,f,over", [
    (64, 100.0, {}),
    (96, 250.0, {}),
    (48, 100.0, dict(sphere_tracing_iters=2, line_step_iters=1)),
    (48, 100.0, dict(sphere_tracing_iters=0)),
    (48, 100.0, dict(line_step_iters=0)),
    (48, 80.0, dict(line_search_step=0.25, sphere_tracing_iters=6)),
])
def test_per_ray_march_equals_lock_step_reference(emu, n_side, f, over):
    cfg = otr.TraceConfig(**over)
    prims = otr.robot_scene()
    o, d = _rays(n_side, seed=n_side, f=f)
    acc_s, acc_e, mn, mx, flags, evals, rounds = _run(emu, o, d, prims, cfg)
    t_sph, hits = otr.sphere_intersection(o, d, cfg.object_bounding_sphere)
    st = otr.sphere_tracing(otr.analytic_sdf(prims), o, d, hits, t_sph, cfg)
    assert torch.equal((flags & 1).bool(), hits)
    assert torch.equal(acc_s, st["acc_start"]), (acc_s - st["acc_start"]).abs().max()
    assert torch.equal(acc_e, st["acc_end"]), (acc_e - st["acc_end"]).abs().max()
    assert torch.equal((flags & 2).bool(), st["unfinished"])
    assert torch.equal(mn, st["min_dis"]) and torch.equal(mx, st["max_dis"])
    assert evals == st["n_evals"]
    # rounds of the device-driven schedule: never more than the lock-step schedule's worst case
    assert rounds <= 1 + cfg.sphere_tracing_iters * (1 + cfg.line_step_iters)
    assert int(hits.sum()) > 100


def test_secondary_style_rays(emu):
    """origins inside the sphere, one ray per origin (how the integrator calls the tracer)"""
    g = torch.Generator().manual_seed(3)
    n = 4000
    o = ((torch.rand(n, 3, generator=g) - 0.5) * 1.2).contiguous()
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).contiguous()
    cfg = otr.TraceConfig()
    prims = otr.robot_scene()
    acc_s, acc_e, mn, mx, flags, evals, rounds = _run(emu, o, d, prims, cfg)
    t_sph, hits = otr.sphere_intersection(o, d, cfg.object_bounding_sphere)
    st = otr.sphere_tracing(otr.analytic_sdf(prims), o, d, hits, t_sph, cfg)
    assert torch.equal(acc_s, st["acc_start"]) and torch.equal(acc_e, st["acc_end"])
    assert torch.equal((flags & 2).bool(), st["unfinished"])
    assert evals == st["n_evals"]
