"""GPU: the CUDA tracer against the oracle.

With the analytic scene both sides evaluate the SDF with the same rounded operations, so hit masks,
the sampler / bisection control flow and the depths must agree BIT FOR BIT (BASELINE.json: "hit masks
and pixel indices bit-exact after fixing the march order")."""
import pytest
import torch

from oracle import tracer as otr

pytestmark = pytest.mark.gpu


def _camera(dev, n_side=64, seed=0, cam=(0.0, 0.0, -3.0), f=100.0):
    g = torch.Generator().manual_seed(seed)
    K = torch.eye(4); K[0, 0] = K[1, 1] = f; K[0, 2] = n_side / 2; K[1, 2] = n_side / 2
    pose = torch.eye(4); pose[:3, 3] = torch.tensor(cam)
    ii, jj = torch.meshgrid(torch.arange(n_side).float(), torch.arange(n_side).float(), indexing="xy")
    uv = torch.stack([ii, jj], -1).reshape(1, -1, 2) + torch.rand(1, n_side * n_side, 2, generator=g)
    dirs, loc = otr.camera_rays(uv, pose[None], K[None])
    return dirs.to(dev), loc.to(dev)


def _module(training, **over):
    from nefii_b200.model.ray_tracing import RayTracing
    cfg = otr.TraceConfig(**over)
    rt = RayTracing(**cfg.as_kwargs())
    rt.train(training)
    rt.collect_stats = True
    return rt, cfg


@pytest.mark.parametrize("training", [False, True])
@pytest.mark.parametrize("n_side,f", [(64, 100.0), (128, 250.0)])
def test_analytic_scene_bit_exact(cuda_device, training, n_side, f):
    from nefii_b200.model.ray_tracing import AnalyticSDF
    dev = cuda_device
    dirs, loc = _camera(dev, n_side, seed=n_side, f=f)
    n = n_side * n_side
    obj = torch.ones(n, dtype=torch.bool, device=dev)
    obj[::7] = False
    prims = otr.robot_scene()
    sdf_dev = AnalyticSDF(prims, dev)
    rt, cfg = _module(training)
    u = torch.rand(100, generator=torch.Generator().manual_seed(5))
    pts, mask, dist = rt(sdf_dev, loc, obj, dirs, uniforms=u if training else None)
    o_pts, o_mask, o_dist, st = otr.ray_trace(otr.analytic_sdf(prims), loc, obj, dirs, cfg, training=training, uniforms=u)
    # the analytic evaluators themselves agree bit for bit
    probe = torch.rand(5000, 3, device=dev) * 2 - 1
    assert torch.equal(sdf_dev(probe), otr.analytic_sdf(prims)(probe))
    assert torch.equal(mask, o_mask)
    assert rt.last_stats["n_sampler"] == st["n_sampler"]
    assert torch.equal(dist, o_dist), (dist - o_dist).abs().max().item()
    live = st["sphere_hits"]
    assert torch.equal(pts[live], o_pts[live])
    # with one bisection iteration per round the compacted trace evaluates exactly the reference's points; the default (up to
    # four iterations per round while few rays are refined) evaluates the 2^D - 1 candidates of D iterations -- same results
    from nefii_b200 import _lib
    lib = _lib.raw()
    _lib.check(lib.nefii_trace_set_quad_rows(0))
    try:
        p1, m1, d1 = rt(sdf_dev, loc, obj, dirs, uniforms=u if training else None)
        assert rt.last_stats["n_evals"] <= st["n_evals"]
        assert torch.equal(m1, mask) and torch.equal(d1, dist) and torch.equal(p1[live], pts[live])
    finally:
        _lib.check(lib.nefii_trace_set_quad_rows(12288))
    # every depth, with a row budget that makes the device pick it, and one (1 << 30) that always takes the deepest tree
    for depth, rows in ((2, 12288), (3, 12288), (4, 1 << 30), (3, 1 << 30)):
        _lib.check(lib.nefii_trace_set_bisect_depth(depth))
        _lib.check(lib.nefii_trace_set_quad_rows(rows))
        try:
            p1, m1, d1 = rt(sdf_dev, loc, obj, dirs, uniforms=u if training else None)
            assert torch.equal(m1, mask) and torch.equal(d1, dist) and torch.equal(p1[live], pts[live]), (depth, rows)
        finally:
            _lib.check(lib.nefii_trace_set_bisect_depth(4))
            _lib.check(lib.nefii_trace_set_quad_rows(12288))


def test_secondary_style_rays_inside_the_sphere(cuda_device):
    """Origins on / near the surface, one ray per 'image' (how the integrator calls the tracer)."""
    from nefii_b200.model.ray_tracing import AnalyticSDF
    dev = cuda_device
    g = torch.Generator().manual_seed(3)
    n = 5000
    origin = (torch.rand(n, 3, generator=g) - 0.5).to(dev) * 1.2
    dirs = torch.nn.functional.normalize(torch.randn(n, 1, 3, generator=g), dim=-1).to(dev)
    obj = torch.ones(n, dtype=torch.bool, device=dev)
    prims = otr.robot_scene()
    for training in (False, True):
        rt, cfg = _module(training)
        u = torch.rand(100, generator=g)
        pts, mask, dist = rt(AnalyticSDF(prims, dev), origin, obj, dirs, uniforms=u)
        o_pts, o_mask, o_dist, st = otr.ray_trace(otr.analytic_sdf(prims), origin, obj, dirs, cfg, training=training, uniforms=u)
        assert torch.equal(mask, o_mask)
        assert torch.equal(dist, o_dist)
        assert torch.equal(pts, o_pts)
        if training:   # skipping min-SDF sampling changes only lanes that are not hits
            rt.skip_min_sdf = True
            p2, m2, d2 = rt(AnalyticSDF(prims, dev), origin, obj, dirs)
            assert torch.equal(m2, mask) and torch.equal(d2[mask], dist[mask]) and torch.equal(p2[mask], pts[mask])


def test_edge_cases(cuda_device):
    from nefii_b200.model.ray_tracing import AnalyticSDF
    dev = cuda_device
    prims = otr.robot_scene()
    rt, cfg = _module(False)
    sdf_dev = AnalyticSDF(prims, dev)
    # empty batch
    p, m, d = rt(sdf_dev, torch.zeros(1, 3, device=dev), torch.zeros(0, dtype=torch.bool, device=dev), torch.zeros(1, 0, 3, device=dev))
    assert p.shape == (0, 3) and m.shape == (0,)
    # every ray misses the bounding sphere
    dirs = torch.nn.functional.normalize(torch.tensor([[[0.0, 1.0, 0.0]] * 33]), dim=-1).to(dev)
    loc = torch.tensor([[0.0, 0.0, -3.0]], device=dev)
    p, m, d = rt(sdf_dev, loc, torch.ones(33, dtype=torch.bool, device=dev), dirs)
    assert not m.any() and (d == 0).all()
    # shallow marching budget (forces the sampler on almost every ray) still matches the oracle
    rt2, cfg2 = _module(False, sphere_tracing_iters=2, line_step_iters=1, n_rootfind_steps=8)
    dirs, loc = _camera(dev, 48, seed=9)
    obj = torch.ones(48 * 48, dtype=torch.bool, device=dev)
    p, m, d = rt2(sdf_dev, loc, obj, dirs)
    op, om, od, st = otr.ray_trace(otr.analytic_sdf(prims), loc, obj, dirs, cfg2, training=False)
    assert torch.equal(m, om) and torch.equal(d, od)
    assert st["n_sampler"] > 100


def test_mlp_scene_statistics(cuda_device):
    """With the MLP the SDF values differ by ~1e-5 from the fp32 oracle, so masks are compared on rays that
    are not within that margin of a decision; mismatches are counted, depths on agreeing rays abs 1e-4."""
    from nefii_b200 import ops
    from oracle import mlp
    dev = cuda_device
    params = mlp.sdf_init(seed=1, bumps=0.03)
    net = ops.SdfMlp(device=dev)
    net.set_weights([w.to(dev) for w in params.W], [b.to(dev) for b in params.b])

    class Src:
        def nefii_sdf_source(self):
            return 0, net.handle.value, 0, net

    p32 = params.to(dev)
    oracle_sdf = lambda x: mlp.sdf_forward(p32, x)[:, 0]
    dirs, loc = _camera(dev, 64, seed=2, f=2.4 * 64)     # ~50 % of the rays hit the blob
    obj = torch.ones(64 * 64, dtype=torch.bool, device=dev)
    rt, cfg = _module(False)
    pts, mask, dist = rt(Src(), loc, obj, dirs)
    o_pts, o_mask, o_dist, st = otr.ray_trace(oracle_sdf, loc, obj, dirs, cfg, training=False)
    agree = mask == o_mask
    assert agree.float().mean().item() > 0.995, agree.float().mean().item()
    both = mask & o_mask
    assert both.sum() > 500
    err = (dist - o_dist)[both].abs()
    assert err.median().item() < 1e-5
    assert (err < 1e-4).float().mean().item() > 0.98


@pytest.mark.parametrize("training", [False, True])
def test_fixed_schedule_equals_graph_mode(cuda_device, training):
    """nefii_trace_set_graph_mode(0): every loop unrolled to its worst case (empty rounds exit at once) instead of a CUDA graph
    with conditional WHILE nodes -- the same kernels, the same results, with the speculative rounds on and off."""
    from nefii_b200 import _lib
    from nefii_b200.model.ray_tracing import AnalyticSDF
    dev = cuda_device
    lib = _lib.raw()
    dirs, loc = _camera(dev, 96, seed=4, f=160.0)
    n = 96 * 96
    obj = torch.ones(n, dtype=torch.bool, device=dev)
    obj[::5] = False
    sdf_dev = AnalyticSDF(otr.robot_scene(), dev)
    rt, cfg = _module(training)
    u = torch.rand(100, generator=torch.Generator().manual_seed(9))
    ref = None
    try:
        for mode in (1, 0):
            for rows in (12288, 0):
                _lib.check(lib.nefii_trace_set_graph_mode(mode))
                _lib.check(lib.nefii_trace_set_quad_rows(rows))
                out = rt(sdf_dev, loc, obj, dirs, uniforms=u if training else None)
                if ref is None:
                    ref = out
                else:
                    for a, b in zip(out, ref):
                        assert torch.equal(a, b), (mode, rows)
    finally:
        _lib.check(lib.nefii_trace_set_graph_mode(1))
        _lib.check(lib.nefii_trace_set_quad_rows(12288))
