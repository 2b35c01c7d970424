"""Drop-in for the live part of the reference's code/model/path_tracing_render.py: the default renderer
``pt_render_indirect_mlp`` (:1255-1257 -> pt_render_diff_shadow_indirect_mlp with diff_geo=False, :1265-1487),
``get_visibility_and_indirect_light`` (:2109-2166) and ``sg_fn`` (:404-413).

Same signature and return dict.  One call = one sampling kernel, ONE batched secondary trace over the 3*N rays
(min-SDF sampling skipped: its lanes are the ones secondary_mask discards), one fused SDF forward+gradient and one
radiance-MLP pass over all secondary hits, one shading kernel; backward flows to lgtSGs, roughness, albedo,
specular reflectance and -- through the incoming radiance -- the radiance MLP's weights."""
import torch

from .. import _lib, integrator

TINY_NUMBER = 1e-6


def sg_fn(upsilon, xi, lamb, mu):
    return mu * torch.exp(lamb * (torch.sum(upsilon * xi, dim=-1, keepdim=True) - 1))


def draw_uniforms(n, device):
    """The 7 torch.rand calls of the reference, in its order and shapes (cos r1 r2, ggx r1 r2, mixture r0 r1 r2:
    path_tracing_render.py:138-139,73-74,201,219-220), so the same generator state yields the same samples."""
    cols = [torch.rand((n, 1), device=device) for _ in range(4)]
    cols.append(torch.rand((n, 1, 1), device=device).reshape(n, 1))
    cols += [torch.rand((n, 1), device=device) for _ in range(2)]
    return torch.cat(cols, dim=-1)


def get_visibility_and_indirect_light(light_points, hit_mask, wi, model):
    """All three sample types at once.  light_points [3,N,3], hit_mask [3,N,1] bool, wi [3,N,3]
    -> incoming radiance [3,N,3] (zero where the secondary ray escaped); visibility is 1 - hit_mask."""
    m = hit_mask.reshape(-1)
    # host sync 2 of 2 per forward: the number of secondary hits sizes the radiance query
    idx = torch.nonzero(m).squeeze(1)
    indirect = torch.zeros(light_points.shape, device=light_points.device, dtype=torch.float32).reshape(-1, 3)
    if idx.shape[0] > 0:
        xs = light_points.reshape(-1, 3).index_select(0, idx)
        if model.implicit_network.differentiable_features and torch.is_grad_enabled():
            # trainable geometry, diff_geo = False (:2111-2160): the FEATURES at the secondary hits keep their graph to the SDF
            # parameters (the reference evaluates the network there with grad enabled); the normals are detached
            feats = model.implicit_network(xs)[:, 1:]
            grads = model.implicit_network.gradient(xs, no_grad=True)[:, 0, :]
        else:
            _, feats, grads = model.implicit_network.evaluate(xs, want_feat=True, want_grad=True)
        normals = grads / (torch.norm(grads, dim=-1, keepdim=True) + 1e-6)
        view = -wi.reshape(-1, 3).index_select(0, idx)
        view = view / (torch.norm(view, dim=-1, keepdim=True) + 1e-6)
        rgb = model.rendering_network(xs, normals, view, feats)
        indirect = indirect.index_copy(0, idx, rgb)
    return indirect.reshape(light_points.shape)


def pt_render_indirect_mlp(lgtSGs, specular_reflectance, roughness, diffuse_albedo, normal, viewdirs, points, model,
                           blending_weights=None, diffuse_rgb=None, uniforms=None):
    """
    :param lgtSGs: [M, 7]
    :param specular_reflectance: [..., 3] / [1, 3] when fix_specular
    :param roughness: [..., 1]
    :param diffuse_albedo: [..., 3]
    :param normal, viewdirs, points: [..., 3]
    :param model: the IDRNetwork (ray_tracer, implicit_network, rendering_network)
    :param uniforms: optional [N,7] uniforms (parity tests); default: drawn like the reference
    """
    # blending_weights / diffuse_rgb: accepted and ignored, exactly as the reference's pt_render_diff_shadow_indirect_mlp does
    # (path_tracing_render.py:1265-1487 never reads them)
    dots_shape = list(normal.shape[:-1])
    n = normal.reshape(-1, 3).shape[0]
    dev = normal.device
    nrm = _lib.f32c(normal).reshape(n, 3)
    view = _lib.f32c(viewdirs).reshape(n, 3)
    pts = _lib.f32c(points).reshape(n, 3)
    with torch.no_grad():
        u = draw_uniforms(n, dev) if uniforms is None else uniforms[:n].contiguous()
        wi, pdf, weight, _ = integrator.mis_sample(lgtSGs, roughness.detach().reshape(n, 1), nrm, view, u)
        origins = pts.unsqueeze(0).expand(3, n, 3).reshape(-1, 3)
        obj = torch.ones(3 * n, dtype=torch.bool, device=dev)
        l_pts, l_hit, _ = model.ray_tracer(sdf=model.implicit_network, cam_loc=origins, object_mask=obj,
                                           ray_directions=wi.reshape(-1, 1, 3), skip_min_sdf=True)
        l_pts = l_pts.reshape(3, n, 3)
        l_hit = l_hit.reshape(3, n, 1)
    indirect = get_visibility_and_indirect_light(l_pts, l_hit, wi, model)
    # a normal that carries a graph (trainable geometry) keeps it through the shading: d / d normal is part of the backward
    nrm_in = normal.reshape(n, 3) if (torch.is_grad_enabled() and normal.requires_grad) else nrm
    out = integrator.mis_shade(lgtSGs, specular_reflectance.reshape(-1, 3), roughness.reshape(n, 1),
                               diffuse_albedo.reshape(n, 3), nrm_in, view, wi, pdf, weight, l_hit, indirect)
    return {'sg_rgb': out['sg_rgb'].reshape(dots_shape + [3]),
            'sg_specular_rgb': out['sg_specular_rgb'].reshape(dots_shape + [3]),
            'sg_diffuse_rgb': out['sg_diffuse_rgb'].reshape(dots_shape + [3]),
            'sg_diffuse_albedo': diffuse_albedo,
            'secondary_points': l_pts.reshape([3] + dots_shape + [3]),
            'secondary_mask': l_hit.reshape([3] + dots_shape + [1]),
            'secondary_dir': wi.reshape([3] + dots_shape + [3])}
