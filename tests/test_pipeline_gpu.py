"""GPU: the whole per-ray-batch path (IDRNetwork.forward of nefii_b200) against the restated reference pipeline
(oracle/pipeline.py) on identical weights, rays and random numbers -- forward outputs and parameter gradients."""
import pytest
import torch

from oracle import pipeline, ref_harness as rh

pytestmark = pytest.mark.gpu


def _build(dev, seed=0, n_sg=128):
    from nefii_b200.model.implicit_differentiable_renderer import IDRNetwork
    from nefii_b200.utils.conf import default_model_conf
    om = rh.small_model(seed=seed, n_sg=n_sg)
    torch.manual_seed(0)
    net = IDRNetwork(default_model_conf(num_lgt_sgs=n_sg)).to(dev)
    rh.load_oracle_weights(net, om)
    return net, om.to(dev)


def _inputs(dev, n_side, rays, seed):
    uv, pose, K = rh.camera_batch(n_side, rays, seed=seed)
    S = uv.shape[1]
    obj = torch.ones(1, S, dtype=torch.bool)
    obj[0, ::9] = False
    g = torch.Generator().manual_seed(seed + 100)
    U = torch.rand(S * max(rays, 1), 7, generator=g).to(dev)
    vecs = [torch.rand(100, generator=g) for _ in range(2)]
    return dict(uv=uv.to(dev), pose=pose.to(dev), intrinsics=K.to(dev), object_mask=obj.to(dev)), U, vecs


KEYS = ['points', 'idr_rgb_values', 'sg_rgb_values', 'normal_values', 'sdf_output', 'sg_diffuse_rgb_values',
        'sg_diffuse_albedo_values', 'sg_specular_rgb_values', 'sg_roughness_values', 'sg_specular_reflection_values']


@pytest.mark.parametrize("training,rays", [(False, 0), (True, 4)])
def test_forward_matches_oracle_pipeline(cuda_device, training, rays):
    dev = cuda_device
    net, om = _build(dev)
    net.train(training)
    inp, U, vecs = _inputs(dev, 32, rays, seed=3)
    with torch.no_grad():
        mine = net.forward_with_uv(inp, uniforms=U, trace_uniforms=vecs[0])
        ref = pipeline.forward_with_uv(om, inp['uv'], inp['pose'], inp['intrinsics'], inp['object_mask'], lambda n: U[:n],
                                       training, vecs[0], vecs[1])
    a, b = mine['network_object_mask'], ref['network_object_mask']
    agree = a == b
    assert int((~agree).sum()) <= 1, int((~agree).sum())      # measured: 0 mismatches (tools/diag_gpu.py pipeline / fullsize)
    assert int((a & b).sum()) > 100
    assert torch.equal(mine['object_mask'], ref['object_mask'])
    sel = agree
    # depth abs 1e-4 (north_star) on the surface pixels whose rays took the same branches.  Non-hit lanes in train mode hold
    # the arg-min over 100 random depths (minimal_sdf_points): the arg-min may jump between two near-equal samples, so
    # there the minimum VALUE (sdf_output) is compared instead of its position.
    hit = sel & a
    assert ((mine['points'] - ref['points'])[hit].abs().max(-1)[0] < 1e-4).float().mean().item() > 0.99      # measured 0.9995 / 1.0
    miss = sel & ~a
    if bool(miss.any()):
        assert (mine['sdf_output'] - ref['sdf_output'])[miss].abs().max().item() < 1e-4
    for k in KEYS:
        if k == 'points':
            continue
        x, y = mine[k][sel].float(), ref[k][sel].float()
        if k == 'sdf_output':      # |sdf| ~ 1e-6 on the surface: absolute, against the fp32 noise floor of the 8-layer MLP
            assert (x - y).abs().flatten().kthvalue(max(1, int(0.95 * x.numel())))[0].item() < 1e-5
            continue
        err = (x - y).abs() / (y.abs() + 1e-3)
        p95 = err.flatten().kthvalue(max(1, int(0.95 * err.numel())))[0].item()
        # idr_rgb = (raw MLP output)^2 of a random-init net: tiny values, the relative error of the square is amplified
        print('pipeline train=%d %-28s p95 rel err %.2e' % (training, k, p95))
        assert p95 < (1.5e-3 if k == 'idr_rgb_values' else 1e-4), (k, p95)        # measured 3.5e-4 / <= 1.3e-5 (fp16-split planes)
    # secondary rays: same directions (bit-exact sampler) wherever the primary hit point agrees
    if mine['secondary_dir'] is not None and mine['secondary_dir'].shape == ref['secondary_dir'].shape:
        d = (mine['secondary_dir'] - ref['secondary_dir']).abs().amax(-1)
        assert (d < 1e-3).float().mean().item() > 0.97


def test_gradients_match_oracle_pipeline(cuda_device):
    dev = cuda_device
    net, om = _build(dev, seed=1)
    net.train(True)
    inp, U, vecs = _inputs(dev, 32, 2, seed=5)
    S = inp['uv'].shape[1]
    gt = torch.rand(S, 3, generator=torch.Generator().manual_seed(9)).to(dev)

    def loss_of(out):
        m = out['network_object_mask'] & out['object_mask']
        bg = (~out['network_object_mask']) & (~out['object_mask'])
        l = (out['sg_rgb_values'][m] - gt[m]).abs().mean() + (out['idr_rgb_values'][m] - gt[m]).abs().mean()
        if bool(bg.any()):
            l = l + ((out['sg_rgb_values'][bg] - gt[bg]) ** 2).mean()
        return l

    mine = net.forward_with_uv(inp, uniforms=U, trace_uniforms=vecs[0])
    loss_of(mine).backward()
    g_lgt = net.envmap_material_network.lgtSGs.grad.clone()
    g_mat = [l.weight.grad.clone() for l in net.envmap_material_network.diffuse_albedo_layers if hasattr(l, "weight")]
    g_rad_v = [getattr(net.rendering_network, "lin%d" % i).weight_v.grad.clone() for i in range(5)]

    om.lgtSGs.requires_grad_(True)
    om.material.requires_grad_(True)
    om.radiance.requires_grad_(True)
    ref = pipeline.forward_with_uv(om, inp['uv'], inp['pose'], inp['intrinsics'], inp['object_mask'], lambda n: U[:n], True,
                                   vecs[0], vecs[1])
    assert (mine['network_object_mask'] == ref['network_object_mask']).float().mean().item() > 0.99
    loss_of(ref).backward()

    def rel(a, b):
        return (a - b).norm().item() / (b.norm().item() + 1e-20)

    # north_star: gradients rel 1e-3 -- met at BASELINE size (tests/test_parity_fullsize_gpu.py); this 1024-pixel x 2-ray batch
    # has so few samples per weight that single ReLU / ELU sign flips of near-zero pre-activations show (measured: lgtSGs 1e-6,
    # material <= 1.4e-5, radiance 1e-5 .. 1.9e-3): 1e-2 for the radiance network here, rel 1e-3 for the others
    print('grad rel: lgtSGs %.2e material %s' % (rel(g_lgt, om.lgtSGs.grad), ' '.join('%.1e' % rel(a, w.grad) for a, w in zip(g_mat, om.material.W))))
    assert rel(g_lgt, om.lgtSGs.grad) < 1e-3, rel(g_lgt, om.lgtSGs.grad)       # measured 7e-7
    for a, b in zip(g_mat, [w.grad for w in om.material.W]):
        assert rel(a, b) < 1e-3, rel(a, b)                                     # measured <= 1.4e-5
    # radiance net: the oracle holds effective weights W = g v/|v| with |v| = g at init, so dL/dW == dL/dv + radial part;
    # compare the tangential gradient (what weight_v receives) after projecting the oracle's gradient the same way
    for a, w in zip(g_rad_v, om.radiance.W):
        gW = w.grad
        v = w.detach()
        nrm = v.norm(dim=1, keepdim=True)
        gv = gW - (gW * v).sum(1, keepdim=True) * v / (nrm * nrm)      # d/dv of g * v/|v| at g == |v|
        print('grad rel radiance %.2e' % rel(a, gv))
        assert rel(a, gv) < 1e-2, rel(a, gv)


def test_forward_with_point_entry(cuda_device):
    """train_with_secondary's entry (idr_train.py:804-852 -> forward(input, with_point=True),
    implicit_differentiable_renderer.py:503-527): same kernels, points + directions instead of uv."""
    dev = cuda_device
    net, om = _build(dev, seed=2)
    net.train(True)
    g = torch.Generator().manual_seed(4)
    n, R = 96, 4
    # points on the blob's surface: trace a few rays first
    inp, U, vecs = _inputs(dev, 24, 0, seed=7)
    with torch.no_grad():
        out = net.forward_with_uv(inp, uniforms=U)
    hit = out['network_object_mask']
    pts = out['points'][hit][:n]
    n = pts.shape[0]
    dirs = torch.nn.functional.normalize(torch.randn(n, R, 3, generator=g), dim=-1).to(dev)
    points = pts.unsqueeze(1).expand(n, R, 3).contiguous()
    res = net({'points': points, 'ray_dirs': dirs}, with_point=True)
    assert res['idr_rgb_values'].shape == (n, 3) and res['sg_rgb_values'].shape == (n, 3)
    assert torch.isfinite(res['sg_rgb_values']).all() and torch.isfinite(res['idr_rgb_values']).all()
    loss = (res['sg_rgb_values'] - res['idr_rgb_values'].detach()).abs().mean()
    loss.backward()
    assert net.envmap_material_network.lgtSGs.grad is not None
    # idr branch equals the oracle's radiance net on the same points / normals
    from oracle import mlp as omlp
    with torch.no_grad():
        feats = omlp.sdf_forward(om.sdf, points.reshape(-1, 3))[:, 1:]
        nrm = pipeline.unit(omlp.sdf_gradient(om.sdf, points.reshape(-1, 3)))
        ref = omlp.radiance_forward(om.radiance, points.reshape(-1, 3), nrm, pipeline.unit(-dirs.reshape(-1, 3)), feats)
    assert torch.allclose(res['idr_rgb_values'], ref.reshape(n, R, 3).mean(1), rtol=2e-3, atol=2e-4)
    # the SG branch of the same entry: get_rbg_value with injected uniforms against the oracle's get_rgb_value (the entry
    # itself draws its random numbers like the reference, so it is compared through the function it wraps)
    U2 = torch.rand(n * R, 7, generator=g).to(dev)
    with torch.no_grad():
        mine = net.get_rbg_value(points.reshape(-1, 3), -dirs.reshape(-1, 3), uniforms=U2)
        oref = pipeline.get_rgb_value(om, points.reshape(-1, 3), -dirs.reshape(-1, 3), U2, True, vecs[1])
    assert torch.equal(mine['secondary_dir'].reshape(3, -1, 3), oref['secondary_dir'].reshape(3, -1, 3)) or \
        (mine['secondary_dir'].reshape(3, -1, 3) - oref['secondary_dir'].reshape(3, -1, 3)).abs().max().item() < 1e-4
    same = (mine['secondary_mask'].reshape(3, -1) == oref['secondary_mask'].reshape(3, -1)).all(0)
    assert same.float().mean().item() > 0.995
    err = ((mine['sg_rgb'] - oref['sg_rgb']).abs() / (oref['sg_rgb'].abs() + 1e-6))[same]
    frac = (err <= 1e-4).float().mean().item()
    print('forward_with_point sg_rgb: lanes within rel 1e-4: %.4f, p95 %.2e' % (frac, err.flatten().kthvalue(int(0.95 * err.numel()))[0].item()))
    assert frac > 0.85 and err.flatten().kthvalue(int(0.95 * err.numel()))[0].item() < 1e-3
    for k in ('sg_diffuse_rgb', 'sg_specular_rgb'):
        assert torch.allclose(mine[k][same], oref[k][same], rtol=5e-3, atol=5e-5), k


def test_degenerate_batches(cuda_device):
    """All rays miss (background only, gradient reaches lgtSGs through the environment lookup) and a one-ray batch."""
    dev = cuda_device
    net, om = _build(dev, seed=3)
    net.train(True)
    uv, pose, K = rh.camera_batch(8, 2, seed=1, cam=(0.0, 0.0, -3.0))
    pose = pose.clone()
    pose[0, :3, :3] = torch.diag(torch.tensor([1.0, 1.0, -1.0]))          # look away from the object
    S = uv.shape[1]
    inp = dict(uv=uv.to(dev), pose=pose.to(dev), intrinsics=K.to(dev), object_mask=torch.zeros(1, S, dtype=torch.bool, device=dev))
    out = net(inp)
    assert not out['network_object_mask'].any()
    assert out['secondary_points'] is None
    assert torch.isfinite(out['sg_rgb_values']).all()
    out['sg_rgb_values'].sum().backward()
    g = net.envmap_material_network.lgtSGs.grad
    assert g is not None and torch.isfinite(g).all() and g.abs().sum() > 0
    # a single pixel, single ray, eval mode
    net.eval()
    one = dict(uv=torch.tensor([[[4.0, 4.0]]], device=dev), pose=rh.camera_batch(8, 0)[1].to(dev), intrinsics=K.to(dev),
               object_mask=torch.ones(1, 1, dtype=torch.bool, device=dev))
    with torch.no_grad():
        o = net(one)
    assert o['sg_rgb_values'].shape == (1, 3) and torch.isfinite(o['sg_rgb_values']).all()


def test_render_frame_chunked_matches_single_call(cuda_device):
    """utils/general.render_frame (reference scripts/render.py:283-360 + utils/general.py:24-82): a 48 x 48 frame rendered in
    512-pixel chunks equals the same frame rendered in one call (eval mode: every ray is independent)."""
    from nefii_b200.utils import general
    dev = cuda_device
    net, _ = _build(dev)
    net.eval()
    n = 48
    ii, jj = torch.meshgrid(torch.arange(n, device=dev).float(), torch.arange(n, device=dev).float(), indexing="xy")
    f = 2.0 * n
    K = torch.tensor([[f, 0, n / 2, 0], [0, f, n / 2, 0], [0, 0, 1, 0], [0, 0, 0, 1]], device=dev)[None]
    pose = torch.eye(4, device=dev)[None].clone()
    pose[0, 2, 3] = -3.0
    inp = {'uv': torch.stack([ii, jj], -1).reshape(1, -1, 2) + 0.5, 'object_mask': torch.ones(1, n * n, dtype=torch.bool, device=dev),
           'pose': pose, 'intrinsics': K}
    torch.manual_seed(0)
    whole = general.render_frame(net, inp, n * n, memory_capacity_level=20)
    torch.manual_seed(0)
    parts = general.render_frame(net, inp, n * n, memory_capacity_level=9)
    assert int(whole['network_object_mask'].sum()) > 100
    assert torch.equal(whole['network_object_mask'], parts['network_object_mask'])
    m = whole['network_object_mask']
    for name in ('points', 'normal_values', 'idr_rgb_values', 'sg_diffuse_albedo_values', 'sg_roughness_values'):
        # geometry and material do not depend on the sampled secondary rays; bisection's batch-wide stop moves depths < 2e-5
        assert (whole[name] - parts[name])[m].abs().max().item() < 2e-3, name
    assert set(whole) == {k for k, _ in general.FRAME_PLANES}
    # ... and the ORACLE's frame (eval mode, one call): geometry and material planes do not depend on the sampled directions
    om = rh.small_model(seed=0).to(dev)
    Uo = torch.rand(n * n, 7, generator=torch.Generator().manual_seed(1)).to(dev)
    with torch.no_grad():
        ref = pipeline.forward_with_uv(om, inp['uv'], inp['pose'], inp['intrinsics'], inp['object_mask'], lambda k: Uo[:k], False)
    assert int((parts['network_object_mask'] != ref['network_object_mask']).sum()) <= 1
    both = parts['network_object_mask'] & ref['network_object_mask']
    assert ((parts['points'] - ref['points'])[both].abs().amax(-1) < 1e-4).float().mean().item() > 0.99
    assert (parts['normal_values'] - ref['normal_values'])[both].abs().flatten().kthvalue(int(0.99 * int(both.sum()) * 3))[0].item() < 5e-4
    assert torch.allclose(parts['sg_diffuse_albedo_values'][both], ref['sg_diffuse_albedo_values'][both], rtol=1e-3, atol=1e-4)
    assert torch.allclose(parts['sg_roughness_values'][both], ref['sg_roughness_values'][both], rtol=1e-3, atol=1e-4)


def test_render_frame_multi_ray_chunks(cuda_device):
    """BASELINE configs[4]-style rendering (several jittered rays per pixel): chunks hold 2**level / num_rays pixels
    (utils/general.py:29-30) and the per-pixel averages of the chunked frame equal the single-call frame."""
    from nefii_b200.utils import general
    dev = cuda_device
    net, _ = _build(dev)
    net.eval()
    n, R = 32, 4
    ii, jj = torch.meshgrid(torch.arange(n, device=dev).float(), torch.arange(n, device=dev).float(), indexing="xy")
    g = torch.Generator().manual_seed(0)
    jitter = (torch.rand(R, 2, generator=g) - 0.5).to(dev)
    uv = (torch.stack([ii, jj], -1).reshape(-1, 1, 2) + 0.5 + jitter[None]).unsqueeze(0)        # [1, n*n, R, 2]
    f = 2.0 * n
    K = torch.tensor([[f, 0, n / 2, 0], [0, f, n / 2, 0], [0, 0, 1, 0], [0, 0, 0, 1]], device=dev)[None]
    pose = torch.eye(4, device=dev)[None].clone()
    pose[0, 2, 3] = -3.0
    inp = {'uv': uv, 'object_mask': torch.ones(1, n * n, dtype=torch.bool, device=dev), 'pose': pose, 'intrinsics': K}
    assert general.split_input(inp, n * n, num_rays=R, memory_capacity_level=10)[0]['uv'].shape == (1, 256, R, 2)
    whole = general.render_frame(net, inp, n * n, num_rays=R, memory_capacity_level=20)
    parts = general.render_frame(net, inp, n * n, num_rays=R, memory_capacity_level=10)
    assert whole['points'].shape == (n * n, 3) and parts['sg_rgb_values'].shape == (n * n, 3)
    assert torch.equal(whole['network_object_mask'], parts['network_object_mask'])
    m = whole['network_object_mask']
    assert int(m.sum()) > 50
    for name in ('points', 'normal_values', 'idr_rgb_values', 'sg_diffuse_albedo_values', 'sg_roughness_values'):
        assert (whole[name] - parts[name])[m].abs().max().item() < 2e-3, name


def test_runner_toggles_roughness_warmup_and_load_light(cuda_device, tmp_path):
    """What the reference trainer toggles on the material network during step 2 (idr_train.py:543-547,705-713: roughness
    warm-up via set_roughness_fake, --light_sg_path via load_light): same semantics through the accelerated forward."""
    import numpy as np
    from tests.util import load_golden
    dev = cuda_device
    net, _ = _build(dev)
    net.eval()
    inp, U, vecs = _inputs(dev, 24, 0, seed=3)
    with torch.no_grad():
        base = net.forward_with_uv(inp, uniforms=U)
        net.envmap_material_network.set_roughness_fake(True)
        fake = net.forward_with_uv(inp, uniforms=U)
        net.envmap_material_network.set_roughness_fake(False)
    hit = base['network_object_mask']
    assert int(hit.sum()) > 50
    assert torch.equal(fake['network_object_mask'], hit)
    assert (fake['sg_roughness_values'][hit] == 0.5).all()                       # sg_envmap_material.py:407-408
    assert not torch.allclose(fake['sg_specular_rgb_values'][hit], base['sg_specular_rgb_values'][hit])
    assert torch.equal(fake['sg_diffuse_albedo_values'], base['sg_diffuse_albedo_values'])
    # load_light: a [M, 7] .npy replaces lgtSGs (here the reference's shipped sunrise environment from the golden file)
    sunrise = load_golden("sg_render_cfg1.npz")["lgt_sunrise"]
    path = str(tmp_path / "sg_128.npy")
    np.save(path, sunrise)
    net.envmap_material_network.load_light(path)
    assert net.envmap_material_network.lgtSGs.shape == (128, 7) and net.envmap_material_network.lgtSGs.is_cuda
    assert torch.equal(net.envmap_material_network.get_light().cpu(), torch.from_numpy(sunrise))
    with torch.no_grad():
        lit = net.forward_with_uv(inp, uniforms=U)
    assert torch.isfinite(lit['sg_rgb_values']).all()
    assert not torch.allclose(lit['sg_rgb_values'][hit], base['sg_rgb_values'][hit])
    # the environment seen by miss rays is exactly the loaded light
    from nefii_b200.utils import rend_util
    from oracle import sg
    dirs, _ = rend_util.get_camera_params(inp['uv'], inp['pose'], inp['intrinsics'])
    bg = sg.background_sg(torch.from_numpy(sunrise).to(dev), dirs.reshape(-1, 3)[~hit])
    assert torch.allclose(lit['sg_rgb_values'][~hit], bg, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("training,rays", [(False, 0), (True, 4)])
def test_prefetched_trace_is_bit_identical(cuda_device, training, rays):
    """IDRNetwork.prefetch_trace: the primary trace of a later forward, run ahead on a side stream (bounded grid) while the
    caller's stream does other work, is the one forward computes itself -- every output bit-identical; a forward on other
    inputs in between does not pick it up."""
    dev = cuda_device
    net, _ = _build(dev)
    net.train(training)
    inp, U, _ = _inputs(dev, 32, rays, seed=3)
    inp2, U2, _ = _inputs(dev, 24, rays, seed=4)
    with torch.no_grad():
        torch.manual_seed(11)
        ref = net.forward_with_uv(inp, uniforms=U)
        torch.manual_seed(11)
        assert net.prefetch_trace(inp)
        other = net.forward_with_uv(inp2, uniforms=U2)          # the caller's stream is busy with another batch meanwhile
        assert len(net._prefetch_state["pending"]) == 1
        got = net.forward_with_uv(inp, uniforms=U)
        assert len(net._prefetch_state["pending"]) == 0
    assert other['points'].shape[0] == inp2['uv'].shape[1]
    assert int(ref['network_object_mask'].sum()) > 100
    for k in KEYS + ['network_object_mask', 'object_mask', 'secondary_points', 'secondary_mask', 'secondary_dir']:
        if ref[k] is None:
            assert got[k] is None
            continue
        assert torch.equal(got[k], ref[k]), k
    # a trainable geometry: the trace would depend on the parameter update before it -- refused
    net.train(True)
    net.unfreeze_geometry()
    assert net.prefetch_trace(inp) is False
