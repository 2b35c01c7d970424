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
    only = sys.argv[1:]                      # e.g. `python -m oracle.make_golden small_ops`: regenerate one fixture
    if only:
        for name in only:
            globals()["golden_" + name]()
        print("wrote", sorted(os.listdir(OUT)))
        return
    golden_sg_render()
    for extra in EXTRA:
        extra()
    print("wrote", sorted(os.listdir(OUT)))


def reference_idr_loss(inp, **conf):
    """The REAL model.loss.IDRLoss on oracle.loss.loss_inputs() -> (output dict, grads of idr/sg/normal/sdf)."""
    import contextlib
    import io
    from model.loss import IDRLoss
    with contextlib.redirect_stdout(io.StringIO()):
        crit = IDRLoss(**conf)
    leaves = {k: inp[k].clone().requires_grad_(True) for k in ("idr_rgb", "sg_rgb", "normal", "sdf_output")}
    outs = {'idr_rgb_values': leaves['idr_rgb'], 'sg_rgb_values': leaves['sg_rgb'], 'normal_values': leaves['normal'],
            'sdf_output': leaves['sdf_output'], 'network_object_mask': inp['net_mask'], 'object_mask': inp['obj_mask'],
            'grad_theta': None, 'sg_roughness_values': torch.zeros(inp['idr_rgb'].shape[0], 1),
            'sg_specular_rgb_values': torch.zeros_like(inp['sg_rgb'])}
    res = crit(outs, {'rgb': inp['rgb_gt'].unsqueeze(0)})
    res['loss'].backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return res, grads


LOSS_CONF = dict(idr_rgb_weight=1.0, sg_rgb_weight=1.0, eikonal_weight=0.1, mask_weight=100.0, alpha=50.0, normalsmooth_weight=1.0,
                 r_patch=1.0, loss_type='L1', env_loss_type='L2', background_rgb_weight=1.0)     # confs_sg/conf.conf:24-35


def golden_loss():
    """IDRLoss of the step-2 recipe (conf.conf loss{}) and an L2 / smooth-L1 variant, outputs and input gradients."""
    from oracle import loss as oloss
    out = {}
    for tag, seed, conf in (("conf", 0, LOSS_CONF), ("l2", 1, dict(LOSS_CONF, loss_type='L2', env_loss_type='L1')),
                            ("smooth", 2, dict(LOSS_CONF, loss_type='L1_smooth'))):
        inp = oloss.loss_inputs(n_pixels=512, seed=seed)
        res, grads = reference_idr_loss(inp, **conf)
        for k, v in inp.items():
            out["%s_in_%s" % (tag, k)] = v.numpy()
        for k in ('loss', 'idr_rgb_loss', 'sg_rgb_loss', 'mask_loss', 'normalsmooth_loss', 'background_rgb_loss'):
            out["%s_%s" % (tag, k)] = res[k].detach().numpy()
        for k, v in grads.items():
            out["%s_grad_%s" % (tag, k)] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "idr_loss.npz"), **out)


def golden_tracer():
    """The REAL RayTracing module on the analytic robot scene (eval + train) and on the seeded SDF MLP (eval)."""
    import contextlib
    import io
    from model.ray_tracing import RayTracing
    from . import mlp, ref_harness as rh, tracer
    cfg = tracer.TraceConfig()
    rt = RayTracing(**cfg.as_kwargs())
    uv, pose, K = rh.camera_batch(32, 0, seed=3, focal_scale=1.6)
    uv = uv + torch.rand(uv.shape, generator=torch.Generator().manual_seed(4)) - 0.5
    dirs, loc = tracer.camera_rays(uv, pose, K)
    obj = torch.ones(32 * 32, dtype=torch.bool)
    obj[::7] = False
    u = torch.rand(100, generator=torch.Generator().manual_seed(5))
    out = dict(dirs=dirs.numpy(), cam_loc=loc.numpy(), object_mask=obj.numpy(), uniforms=u.numpy(), prims=tracer.robot_scene().numpy())
    sdf = tracer.analytic_sdf(tracer.robot_scene())
    for training in (False, True):
        rt.train(training)
        with rh.injected_rng(None, [u.clone()]):
            p, m, t = rt(sdf=sdf, cam_loc=loc, object_mask=obj, ray_directions=dirs)
        tag = "train" if training else "eval"
        out["analytic_%s_points" % tag] = p.numpy()
        out["analytic_%s_mask" % tag] = m.numpy()
        out["analytic_%s_dists" % tag] = t.numpy()
    params = mlp.sdf_init(seed=1, bumps=0.03)
    with contextlib.redirect_stdout(io.StringIO()):
        from model.implicit_differentiable_renderer import ImplicitNetwork
        net = ImplicitNetwork(512, d_in=3, d_out=1, dims=[512] * 8, geometric_init=True, bias=0.6, skip_in=[4], weight_norm=True,
                              multires=6, use_last_as_f=True)
    net.load_state_dict(params.state_dict(""))
    rt.train(False)
    uv2, pose2, K2 = rh.camera_batch(24, 0, seed=6)
    dirs2, loc2 = tracer.camera_rays(uv2, pose2, K2)
    with torch.no_grad():
        p, m, t = rt(sdf=lambda x: net(x)[:, 0], cam_loc=loc2, object_mask=torch.ones(24 * 24, dtype=torch.bool), ray_directions=dirs2)
        x = torch.rand(256, 3, generator=torch.Generator().manual_seed(7)) * 1.6 - 0.8
        y = net(x)
    g = net.gradient(x.clone(), no_grad=True)[:, 0, :]
    out.update(mlp_dirs=dirs2.numpy(), mlp_cam_loc=loc2.numpy(), mlp_points=p.numpy(), mlp_mask=m.numpy(), mlp_dists=t.numpy(),
               mlp_x=x.numpy(), mlp_sdf=y[:, 0].numpy(), mlp_feat_first8=y[:, 1:9].numpy(), mlp_grad=g.detach().numpy())
    np.savez_compressed(os.path.join(OUT, "tracer_mlp.npz"), **out)


def golden_mis_and_pipeline():
    """Sampling functions of the REAL path_tracing_render.py and one full IDRNetwork.forward (eval + train)."""
    import unittest.mock as mock
    import model.path_tracing_render as ptr
    from . import ref_harness as rh
    n = 512
    normal, view, albedo = inputs.shading_inputs(n, seed=11)
    g = torch.Generator().manual_seed(12)
    rough = torch.rand(n, 1, generator=g) * 0.9 + 0.089
    lgt = inputs.synthetic_light_sgs(128, seed=13)
    u = torch.rand(n, 7, generator=g)
    cols = [u[:, 0:1], u[:, 1:2], u[:, 2:3], u[:, 3:4], u[:, 4:5].unsqueeze(-1), u[:, 5:6], u[:, 6:7]]
    it = iter(cols)
    L = lgt.reshape(1, 128, 7).expand(n, 128, 7)
    with mock.patch.object(torch, "rand", lambda shape, device=None: next(it).clone()):
        w0, p0 = ptr.cos_sampling(normal)
        w1, p1 = ptr.brdf_sampling(normal, rough, view)
        w2, p2 = ptr.mix_sg_sampling(normal, L)
    fns = [ptr.pdf_fn_cos, ptr.pdf_fn_brdf_gxx, ptr.pdf_fn_mix_sg]
    wi = [w0, w1, w2]
    pdf = [torch.clamp(p, min=1e-6) for p in (p0, p1, p2)]
    mat = torch.stack([torch.stack([pdf[i] if i == j else fns[j](wi[i], normal, view, rough, L) for j in range(3)]) for i in range(3)])
    out = dict(normal=normal.numpy(), view=view.numpy(), rough=rough.numpy(), lgt=lgt.numpy(), u=u.numpy(),
               wi=torch.stack(wi).numpy(), pdf=torch.stack(pdf).numpy(), pdf_matrix=mat.numpy())
    np.savez_compressed(os.path.join(OUT, "mis_sampling.npz"), **out)

    om = rh.small_model(seed=0)
    net = rh.build_reference_model(om)
    uv, pose, K = rh.camera_batch(12, 2, seed=1)
    S = uv.shape[1]
    obj = torch.ones(1, S, dtype=torch.bool)
    obj[0, ::5] = False
    g = torch.Generator().manual_seed(7)
    U = torch.rand(4096, 7, generator=g)
    vecs = [torch.rand(100, generator=g) for _ in range(2)]
    out = dict(uv=uv.numpy(), pose=pose.numpy(), intrinsics=K.numpy(), object_mask=obj.numpy(), U=U[:S * 2].numpy(),
               vec0=vecs[0].numpy(), vec1=vecs[1].numpy())
    for training in (False, True):
        net.train(training)
        with rh.injected_rng(lambda n_: U[:n_], [v.clone() for v in vecs]):
            ref = net({'uv': uv, 'pose': pose, 'intrinsics': K, 'object_mask': obj})
        tag = "train" if training else "eval"
        for k in ('points', 'idr_rgb_values', 'sg_rgb_values', 'normal_values', 'sdf_output', 'network_object_mask',
                  'sg_diffuse_rgb_values', 'sg_diffuse_albedo_values', 'sg_specular_rgb_values', 'sg_roughness_values',
                  'secondary_points', 'secondary_mask', 'secondary_dir'):
            out["%s_%s" % (tag, k)] = ref[k].detach().numpy()
    np.savez_compressed(os.path.join(OUT, "pipeline_small.npz"), **out)


def golden_small_ops():
    """SampleNetwork.forward + backward (model/sample_network.py:10-24) and get_camera_params in both pose forms
    (utils/rend_util.py:90-142) from the REAL reference."""
    from model.sample_network import SampleNetwork
    from utils import rend_util as ref_rend
    g = torch.Generator().manual_seed(11)
    n = 777
    s = (torch.randn(n, 1, generator=g) * 1e-3).requires_grad_(True)
    s0 = (s.detach() + torch.randn(n, 1, generator=g) * 1e-4).requires_grad_(True)
    grad = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    grad[:5] = 0.0                                   # exercises the |grad . v| < 1e-8 guard
    grad = grad.requires_grad_(True)
    dirs = dirs.requires_grad_(True)
    t0 = (torch.rand(n, 1, generator=g) * 3 + 0.5).requires_grad_(True)
    cam = torch.randn(n, 3, generator=g).requires_grad_(True)
    x = SampleNetwork()(s, s0, grad, t0, cam, dirs)
    g_out = torch.randn(n, 3, generator=g)
    grads = torch.autograd.grad(x, [s, s0, grad, t0, cam, dirs], g_out)
    out = dict(sn_s=s.detach().numpy(), sn_s0=s0.detach().numpy(), sn_grad=grad.detach().numpy(), sn_t0=t0.detach().numpy(),
               sn_cam=cam.detach().numpy(), sn_dirs=dirs.detach().numpy(), sn_x=x.detach().numpy(), sn_g_out=g_out.numpy())
    for name, t in zip(("s", "s0", "grad", "t0", "cam", "dirs"), grads):
        out["sn_g_" + name] = t.numpy()
    # camera rays: a rotated pose with skewed intrinsics, as a matrix and as quaternion + translation
    B, S = 2, 500
    uv = torch.rand(B, S, 2, generator=g) * 800
    q = torch.nn.functional.normalize(torch.randn(B, 4, generator=g), dim=-1)
    tr = torch.randn(B, 3, generator=g)
    pose7 = torch.cat([q, tr], -1)
    R = ref_rend.quat_to_rot(q)
    pose44 = torch.eye(4).repeat(B, 1, 1)
    pose44[:, :3, :3] = R
    pose44[:, :3, 3] = tr
    K = torch.eye(4).repeat(B, 1, 1)
    K[:, 0, 0] = torch.tensor([1900.0, 2100.0]); K[:, 1, 1] = torch.tensor([1950.0, 2050.0])
    K[:, 0, 1] = torch.tensor([0.7, -1.3]); K[:, 0, 2] = 400.5; K[:, 1, 2] = 399.25
    d44, c44 = ref_rend.get_camera_params(uv, pose44, K)
    d7, c7 = ref_rend.get_camera_params(uv, pose7, K)
    out.update(cr_uv=uv.numpy(), cr_pose44=pose44.numpy(), cr_pose7=pose7.numpy(), cr_K=K.numpy(), cr_dirs44=d44.numpy(),
               cr_cam44=c44.numpy(), cr_dirs7=d7.numpy(), cr_cam7=c7.numpy())
    np.savez_compressed(os.path.join(OUT, "small_ops.npz"), **out)


EXTRA = [golden_tracer, golden_mis_and_pipeline, golden_loss, golden_small_ops]

if __name__ == "__main__":
    main()
