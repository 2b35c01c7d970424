"""Oracle (test infrastructure): the per-ray-batch rendering path end to end, restated.

Restates (reference code/model)
  * get_visibility_and_indirect_light (diff_geo=False) -- path_tracing_render.py:2109-2166
  * pt_render_indirect_mlp                             -- path_tracing_render.py:1255-1487
  * IDRNetwork.get_rbg_value                           -- implicit_differentiable_renderer.py:529-599
  * IDRNetwork.forward_with_uv (geometry frozen / eval) -- :312-501, mean_pixel :695-719,
    get_background_rgb :646-663
on top of the restated pieces in oracle/{mlp,tracer,mis,sg}.py.  All random numbers are explicit
inputs: `u7` [N_hit,7] for the importance samplers and one [n_steps] vector per tracer call in
training mode (the reference draws them with torch.rand / Tensor.uniform_).

Parity status: PINNED -- tests/test_oracle_hotpath.py runs the real IDRNetwork (weights copied in, RNG
patched) against this file when /root/reference is present; tests/golden/pipeline_*.npz otherwise.
"""
import torch

from . import mis, mlp, sg, tracer


class OracleModel:
    """Plain-tensor stand-in for IDRNetwork with conf.conf's model block (frozen geometry)."""

    def __init__(self, sdf, radiance, material, lgtSGs, specular_albedo=0.5, trace_cfg=None, render_background=True):
        self.sdf = sdf
        self.radiance = radiance
        self.material = material
        self.lgtSGs = lgtSGs
        self.specular_albedo = specular_albedo
        self.trace_cfg = trace_cfg or tracer.TraceConfig()
        self.render_background = render_background

    def to(self, *a, **k):
        return OracleModel(self.sdf.to(*a, **k), self.radiance.to(*a, **k), self.material.to(*a, **k),
                           self.lgtSGs.to(*a, **k), self.specular_albedo, self.trace_cfg, self.render_background)

    def sdf_fn(self, x):
        with torch.no_grad():
            return mlp.sdf_forward(self.sdf, x)[:, 0]

    def specular_reflectance(self, like):
        s = torch.full((1, 3), self.specular_albedo, dtype=like.dtype, device=like.device)
        return mlp.specular_remap(s)


def unit(v):
    return v / (torch.norm(v, dim=-1, keepdim=True) + 1e-6)


def secondary_query(model, light_points, hit_mask, wi, trainable_geometry=False):
    """visibility [N,1] and incoming radiance [N,3] for one sample type (diff_geo=False).  trainable_geometry: the network
    is evaluated with grad enabled there (path_tracing_render.py:2111-2112), so the feature vectors of the secondary hits keep
    their graph to the SDF parameters; the normals are detached either way (:2149, no_grad=not diff_geo)."""
    with torch.set_grad_enabled(trainable_geometry and torch.is_grad_enabled()):
        out = mlp.sdf_forward(model.sdf, light_points)
    visibility = 1 - hit_mask.to(light_points.dtype)
    m = hit_mask.reshape(-1)
    xs = light_points[m]
    with torch.no_grad():
        nrm = unit(mlp.sdf_gradient(model.sdf, xs))
    view = unit(-wi[m])
    rgb_hit = mlp.radiance_forward(model.radiance, xs, nrm, view, out[m, 1:])
    rgb = torch.zeros_like(light_points)
    rgb[m] = rgb_hit
    return visibility, rgb


def pt_render_indirect_mlp(model, lgtSGs, specular_reflectance, roughness, diffuse_albedo, normal, view, points,
                           u7, training, trace_uniforms=None, trainable_geometry=False):
    with torch.no_grad():
        wi, pdf, pdf_matrix = mis.sample_directions(lgtSGs, roughness, normal, view, u7)
        n = points.shape[0]
        origins = points.reshape(1, n, 3).expand(3, n, 3).reshape(-1, 3)
        dirs = wi.reshape(-1, 1, 3)
        obj = torch.ones(3 * n, dtype=torch.bool, device=points.device)
        l_pts, l_hit, l_dist, st = tracer.ray_trace(model.sdf_fn, origins, obj, dirs, model.trace_cfg, training=training,
                                                     uniforms=trace_uniforms)
    l_pts = l_pts.reshape(3, n, 3)
    l_hit = l_hit.reshape(3, n, 1)
    vis, ind = [], []
    for i in range(3):
        v, r = secondary_query(model, l_pts[i], l_hit[i], wi[i], trainable_geometry)
        vis.append(v)
        ind.append(r)
    ret = mis.shade(lgtSGs, specular_reflectance, roughness, diffuse_albedo, normal, view, wi, pdf, pdf_matrix,
                    torch.stack(vis), torch.stack(ind))
    ret.update(sg_diffuse_albedo=diffuse_albedo, secondary_points=l_pts, secondary_mask=l_hit, secondary_dir=wi,
               trace_stats=st)
    return ret


def get_rgb_value(model, points, view_dirs, u7, training, trace_uniforms=None):
    with torch.no_grad():
        feats = mlp.sdf_forward(model.sdf, points)[:, 1:]
        normals = unit(mlp.sdf_gradient(model.sdf, points))
    view_dirs = unit(view_dirs)
    idr_rgb = mlp.radiance_forward(model.radiance, points, normals, view_dirs, feats)
    albedo, rough = mlp.material_forward(model.material, points, feats)
    spec = model.specular_reflectance(points)
    ret = pt_render_indirect_mlp(model, model.lgtSGs, spec, rough, albedo, normals, view_dirs, points, u7, training,
                                 trace_uniforms)
    ret.update(normals=normals, idr_rgb=idr_rgb, sg_roughness=rough, sg_specular_reflectance=spec)
    return ret


def mean_pixel(x, bs, r, vector=False):
    flat = x.dim() == 1
    if flat:
        x = x[..., None]
    x = x.reshape(bs, r, x.shape[-1])
    if vector:
        x = x[:, 0, :]
    elif x.dtype == torch.bool:
        x = x.all(1)
    else:
        x = x.mean(1)
    return x[..., 0] if flat else x


def forward_with_uv(model, uv, pose, intrinsics, object_mask, u7_fn, training, primary_uniforms=None, secondary_uniforms=None):
    """IDRNetwork.forward_with_uv with frozen geometry.  uv [B,S,R,2] or [B,S,2].
    u7_fn(n_hit) -> [n_hit,7] uniforms (called once the number of surface hits is known)."""
    multi = uv.dim() == 4
    object_mask = object_mask.reshape(-1)
    if multi:
        B, S, R, _ = uv.shape
        uv = uv.reshape(B, S * R, 2)
        object_mask = object_mask.reshape(B, S, 1).expand(B, S, R).reshape(-1)
    dirs, cam = tracer.camera_rays(uv, pose, intrinsics)
    B, P, _ = dirs.shape
    with torch.no_grad():
        pts, net_mask, dists, st = tracer.ray_trace(model.sdf_fn, cam, object_mask, dirs, model.trace_cfg, training=training,
                                                    uniforms=primary_uniforms)
    points = (cam.unsqueeze(1) + dists.reshape(B, P, 1) * dirs).reshape(-1, 3)
    with torch.no_grad():
        sdf_output = mlp.sdf_forward(model.sdf, points)[:, 0:1]
    dirs = dirs.reshape(-1, 3)
    surface = net_mask
    ones = torch.ones_like(points)
    out = dict(idr_rgb_values=ones.clone(), sg_rgb_values=ones.clone(), normal_values=ones.clone(),
               sg_diffuse_rgb_values=ones.clone(), sg_diffuse_albedo_values=ones.clone(),
               sg_specular_rgb_values=torch.zeros_like(points), sg_roughness_values=torch.zeros_like(points[:, :1]),
               sg_specular_reflection_values=torch.zeros_like(points))
    ret = {}
    if int(surface.sum()) > 0:
        ret = get_rgb_value(model, points[surface], -dirs[surface], u7_fn(int(surface.sum())), training, secondary_uniforms)
        out['idr_rgb_values'][surface] = ret['idr_rgb']
        out['sg_rgb_values'][surface] = ret['sg_rgb']
        out['normal_values'][surface] = ret['normals']
        out['sg_diffuse_rgb_values'][surface] = ret['sg_diffuse_rgb']
        out['sg_diffuse_albedo_values'][surface] = ret['sg_diffuse_albedo']
        out['sg_specular_rgb_values'][surface] = ret['sg_specular_rgb']
        out['sg_roughness_values'][surface] = ret['sg_roughness']
        out['sg_specular_reflection_values'][surface] = ret['sg_specular_reflectance']
    bg = ~surface
    if model.render_background and bool(bg.any()):
        out['sg_rgb_values'][bg] = sg.background_sg(model.lgtSGs, dirs[bg])
    out.update(points=points, sdf_output=sdf_output, network_object_mask=net_mask, object_mask=object_mask, grad_theta=None,
               secondary_points=ret.get('secondary_points'), secondary_mask=ret.get('secondary_mask'),
               secondary_dir=ret.get('secondary_dir'), trace_stats=st, secondary_trace_stats=ret.get('trace_stats'))
    if multi:
        for key in ('idr_rgb_values', 'sg_rgb_values', 'network_object_mask', 'object_mask', 'sg_diffuse_rgb_values',
                    'sg_diffuse_albedo_values', 'sg_specular_rgb_values', 'sdf_output', 'points', 'sg_roughness_values',
                    'sg_specular_reflection_values'):
            out[key] = mean_pixel(out[key], B * S, R)
        out['normal_values'] = mean_pixel(out['normal_values'], B * S, R, vector=True)
    return out


# ------------------------------------------------------------------------------------------------
# trainable geometry (training and not freeze_geometry): implicit_differentiable_renderer.py:354-389, 529-599
# ------------------------------------------------------------------------------------------------
def sdf_gradient_graph(params, x):
    """ImplicitNetwork.gradient(x, no_grad=False) (:110-123): autograd with create_graph, [n,3]."""
    if not x.requires_grad:
        x = x.detach().requires_grad_(True)
    with torch.enable_grad():
        y = mlp.sdf_forward(params, x)[:, :1]
        return torch.autograd.grad(y, x, torch.ones_like(y), create_graph=True, retain_graph=True)[0]


def sample_network(surface_output, surface_sdf_values, surface_points_grad, surface_dists, surface_cam_loc, surface_ray_dirs):
    """SampleNetwork.forward (model/sample_network.py:10-24)."""
    dot = (surface_points_grad * surface_ray_dirs.detach()).sum(-1, keepdim=True)
    dot = torch.where(dot.abs() < 1e-8, torch.full_like(dot, 1e-8), dot)
    t = surface_dists - (surface_output - surface_sdf_values) / dot
    return surface_cam_loc + t * surface_ray_dirs


def forward_with_uv_trainable(model, uv, pose, intrinsics, object_mask, u7_fn, eikonal_points, primary_uniforms=None,
                              secondary_uniforms=None):
    """IDRNetwork.forward_with_uv in training mode with a TRAINABLE geometry (model.sdf's tensors require grad):
    eikonal samples, d sdf/dx with create_graph, the differentiable intersection, and features / normals that carry a graph
    through the radiance network, the material network and the shading.  eikonal_points: the [B*P//2, 3] uniform draw (:369)."""
    multi = uv.dim() == 4
    object_mask = object_mask.reshape(-1)
    if multi:
        B, S, R, _ = uv.shape
        uv = uv.reshape(B, S * R, 2)
        object_mask = object_mask.reshape(B, S, 1).expand(B, S, R).reshape(-1)
    dirs, cam = tracer.camera_rays(uv, pose, intrinsics)
    B, P, _ = dirs.shape
    with torch.no_grad():
        pts, net_mask, dists, st = tracer.ray_trace(model.sdf_fn, cam, object_mask, dirs, model.trace_cfg, training=True,
                                                    uniforms=primary_uniforms)
    points = (cam.unsqueeze(1) + dists.reshape(B, P, 1) * dirs).reshape(-1, 3)
    sdf_output = mlp.sdf_forward(model.sdf, points)[:, 0:1]
    dirs = dirs.reshape(-1, 3)
    surface = net_mask & object_mask
    n = int(surface.sum())
    surface_points = points[surface]
    eik = torch.cat([eikonal_points.to(points), points.clone().detach()], 0)
    points_all = torch.cat([surface_points, eik], 0)
    g = sdf_gradient_graph(model.sdf, points_all)
    grad_theta = g[n:]
    ones = torch.ones_like(points)
    out = dict(idr_rgb_values=ones.clone(), sg_rgb_values=ones.clone(), normal_values=ones.clone(),
               sg_diffuse_rgb_values=ones.clone(), sg_diffuse_albedo_values=ones.clone(),
               sg_specular_rgb_values=torch.zeros_like(points), sg_roughness_values=torch.zeros_like(points[:, :1]),
               sg_specular_reflection_values=torch.zeros_like(points))
    ret = {}
    if n > 0:
        surface_output = sdf_output[surface]
        cam_all = cam.unsqueeze(1).expand(B, P, 3).reshape(-1, 3)
        x = sample_network(surface_output, surface_output.detach(), g[:n].clone().detach(), dists[surface].unsqueeze(-1),
                           cam_all[surface], dirs[surface])
        view_dirs = unit(-dirs[surface])
        feats = mlp.sdf_forward(model.sdf, x)[:, 1:]
        normals = unit(sdf_gradient_graph(model.sdf, x))
        idr_rgb = mlp.radiance_forward(model.radiance, x, normals, view_dirs, feats)
        albedo, rough = mlp.material_forward(model.material, x, feats)
        spec = model.specular_reflectance(x)
        ret = pt_render_indirect_mlp(model, model.lgtSGs, spec, rough, albedo, normals, view_dirs, x, u7_fn(n), True,
                                     secondary_uniforms, trainable_geometry=True)
        ret.update(normals=normals, idr_rgb=idr_rgb, sg_roughness=rough, sg_specular_reflectance=spec)
        for key, src in (('idr_rgb_values', 'idr_rgb'), ('sg_rgb_values', 'sg_rgb'), ('normal_values', 'normals'),
                         ('sg_diffuse_rgb_values', 'sg_diffuse_rgb'), ('sg_diffuse_albedo_values', 'sg_diffuse_albedo'),
                         ('sg_specular_rgb_values', 'sg_specular_rgb'), ('sg_roughness_values', 'sg_roughness')):
            out[key] = out[key].index_put((torch.nonzero(surface).squeeze(1),), ret[src])
        out['sg_specular_reflection_values'] = out['sg_specular_reflection_values'].index_put(
            (torch.nonzero(surface).squeeze(1),), ret['sg_specular_reflectance'].expand(n, 3))
    bg = ~surface
    if model.render_background and bool(bg.any()):
        out['sg_rgb_values'] = out['sg_rgb_values'].index_put((torch.nonzero(bg).squeeze(1),), sg.background_sg(model.lgtSGs, dirs[bg]))
    out.update(points=points, sdf_output=sdf_output, network_object_mask=net_mask, object_mask=object_mask, grad_theta=grad_theta,
               secondary_points=ret.get('secondary_points'), secondary_mask=ret.get('secondary_mask'),
               secondary_dir=ret.get('secondary_dir'), differentiable_surface_points=x if n > 0 else None)
    if multi:
        for key in ('idr_rgb_values', 'sg_rgb_values', 'network_object_mask', 'object_mask', 'sg_diffuse_rgb_values',
                    'sg_diffuse_albedo_values', 'sg_specular_rgb_values', 'sdf_output', 'points', 'sg_roughness_values',
                    'sg_specular_reflection_values'):
            out[key] = mean_pixel(out[key], B * S, R)
        out['normal_values'] = mean_pixel(out['normal_values'], B * S, R, vector=True)
    return out
