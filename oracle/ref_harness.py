"""Helpers that drive the REAL reference with injected weights / random numbers (test infrastructure;
needs /root/reference, see ref_shim.py)."""
import contextlib
import io
import unittest.mock as mock

import torch

from . import inputs, mlp, pipeline, ref_shim, tracer


def load_oracle_weights(net, om):
    """Copy an OracleModel's weights into an IDRNetwork (the reference's or nefii_b200's: same state_dict keys)."""
    net.implicit_network.load_state_dict(om.sdf.state_dict(""))
    sd = {}
    for l, (w, b) in enumerate(zip(om.radiance.W, om.radiance.b)):
        sd['lin%d.weight_v' % l] = w.clone()
        sd['lin%d.weight_g' % l] = w.norm(dim=1, keepdim=True)
        sd['lin%d.bias' % l] = b.clone()
    net.rendering_network.load_state_dict(sd)
    msd = net.envmap_material_network.state_dict()
    for l, (w, b) in enumerate(zip(om.material.W, om.material.b)):
        msd['diffuse_albedo_layers.%d.weight' % (2 * l)] = w.clone()
        msd['diffuse_albedo_layers.%d.bias' % (2 * l)] = b.clone()
    msd['lgtSGs'] = om.lgtSGs.clone()
    net.envmap_material_network.load_state_dict(msd)
    net.freeze_geometry()
    return net


def build_reference_model(om, render_type="pt_render_indirect_mlp"):
    """IDRNetwork(conf.conf) of the REAL reference with the OracleModel's weights copied in, geometry frozen."""
    ref_shim.install()
    with contextlib.redirect_stdout(io.StringIO()):
        from model.implicit_differentiable_renderer import IDRNetwork
        net = IDRNetwork(ref_shim.model_conf(render_type=render_type, num_lgt_sgs=om.lgtSGs.shape[0],
                                             width=om.sdf.W[1].shape[0]))
    return load_oracle_weights(net, om)


@contextlib.contextmanager
def injected_rng(u7_fn, uniform_vectors, eikonal_points=None):
    """Patch torch.rand (7 draws per pt_render call, shapes [N,1] / [N,1,1]) and Tensor.uniform_ on
    torch.empty(n) (one [n_steps] vector per training-mode tracer call) and on torch.empty(n, 3) (the eikonal samples of a
    trainable geometry, implicit_differentiable_renderer.py:369)."""
    state = {"u7": None, "col": 0}
    vecs = list(uniform_vectors)
    real_empty = torch.empty

    def fake_rand(shape, device=None):
        n = shape[0]
        if u7_fn is None:
            raise RuntimeError('unexpected torch.rand call')
        if state["u7"] is None or state["col"] == 7:
            state["u7"], state["col"] = u7_fn(n), 0
        col = state["u7"][:, state["col"]].reshape(shape)
        state["col"] += 1
        return col.clone()

    class _Vec:
        def __init__(self, n):
            self.n = n

        def uniform_(self, a, b):
            return vecs.pop(0).clone()

    class _Eik:
        def uniform_(self, a, b):
            return eikonal_points.clone()

    def fake_empty(*size, **kw):
        if len(size) == 1 and isinstance(size[0], int) and not kw:
            return _Vec(size[0])
        if eikonal_points is not None and len(size) == 2 and size[1] == 3 and isinstance(size[0], int) and not kw:
            assert size[0] == eikonal_points.shape[0], (size, eikonal_points.shape)
            return _Eik()
        return real_empty(*size, **kw)

    with mock.patch.object(torch, "rand", fake_rand), mock.patch.object(torch, "empty", fake_empty):
        yield


def small_model(seed=0, n_sg=128, bumps=0.03):
    return pipeline.OracleModel(mlp.sdf_init(seed=seed + 1, bumps=bumps), mlp.radiance_init(seed=seed + 2),
                                mlp.material_init(seed=seed + 3), inputs.synthetic_light_sgs(n_sg, seed=seed + 4))


def camera_batch(n_side, rays_per_pixel, seed=0, focal_scale=2.4, cam=(0.0, 0.0, -3.0)):
    """One synthetic view: uv [1,S,R,2] (or [1,S,2] when rays_per_pixel == 0), pose, intrinsics."""
    g = torch.Generator().manual_seed(seed)
    K = torch.eye(4)
    K[0, 0] = K[1, 1] = focal_scale * n_side
    K[0, 2] = K[1, 2] = n_side / 2
    pose = torch.eye(4)
    pose[:3, 3] = torch.tensor(cam)
    ii, jj = torch.meshgrid(torch.arange(n_side).float(), torch.arange(n_side).float(), indexing="xy")
    uv = torch.stack([ii, jj], -1).reshape(1, -1, 2) + 0.5
    if rays_per_pixel > 0:
        jitter = torch.rand(rays_per_pixel, 2, generator=g) - 0.5          # shared by all pixels (scene_dataset.py:212-216)
        uv = uv.unsqueeze(2) + jitter.reshape(1, 1, rays_per_pixel, 2)
    return uv, pose[None], K[None]
