"""Importance sampling, MIS shading and background lookup of the near-field integrator as autograd-aware
Python entry points over the C ABI (nefii_mis_sample / nefii_mis_shade_{fwd,bwd} / nefii_background_sg_*)."""
import torch

from . import _lib


def mis_sample(lgtSGs, roughness, normal, view, u, want_matrix=False):
    """-> wi [3,N,3], pdf [3,N], weight [3,N], pdf_matrix [3,3,N] | None.  u [N,7]: uniforms in the reference's
    torch.rand order (path_tracing_render.py:138-139,73-74,201,219-220)."""
    n = normal.shape[0]
    dev = normal.device
    lgt, r, nn, vv, uu = (_lib.f32c(lgtSGs), _lib.f32c(roughness).reshape(-1), _lib.f32c(normal), _lib.f32c(view), _lib.f32c(u))
    wi = torch.empty(3, n, 3, device=dev)
    pdf = torch.empty(3, n, device=dev)
    weight = torch.empty(3, n, device=dev)
    mat = torch.empty(3, 3, n, device=dev) if want_matrix else None
    if n:
        _lib.check(_lib.raw().nefii_mis_sample(_lib.stream_ptr(dev), n, lgt.shape[0], lgt.data_ptr(), r.data_ptr(), nn.data_ptr(),
                                               vv.data_ptr(), uu.data_ptr(), wi.data_ptr(), pdf.data_ptr(), weight.data_ptr(),
                                               mat.data_ptr() if mat is not None else None))
    return wi, pdf, weight, mat


class _MisShade(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lgtSGs, specular, roughness, albedo, normal, view, wi, pdf, weight, hit, indirect):
        n = normal.shape[0]
        dev = normal.device
        per_point = specular.shape[0] != 1
        t = [_lib.f32c(x) for x in (lgtSGs, specular, roughness.reshape(-1), albedo, normal, view, wi, pdf, weight, indirect)]
        hit8 = hit.reshape(3, n).to(torch.uint8).contiguous()
        out = torch.empty(3, n, 3, device=dev)
        light = torch.empty(3, n, 3, device=dev)
        if n:
            _lib.check(_lib.raw().nefii_mis_shade_fwd(
                _lib.stream_ptr(dev), n, t[0].shape[0], t[0].data_ptr(), t[1].data_ptr(), 1 if per_point else 0, t[2].data_ptr(),
                t[3].data_ptr(), t[4].data_ptr(), t[5].data_ptr(), t[6].data_ptr(), t[7].data_ptr(), t[8].data_ptr(),
                hit8.data_ptr(), t[9].data_ptr(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), light.data_ptr()))
        ctx.save_for_backward(*t, hit8, light)
        ctx.per_point = per_point
        ctx.shapes = (roughness.shape, specular.shape)
        return out

    @staticmethod
    def backward(ctx, g_out):
        *t, hit8, light = ctx.saved_tensors
        lgt, spec, rough, albedo, normal, view, wi, pdf, weight, indirect = t
        n = normal.shape[0]
        dev = normal.device
        g_out = _lib.f32c(g_out)
        g_rough = torch.zeros(n, device=dev)
        g_alb = torch.zeros(n, 3, device=dev)
        g_sr = torch.zeros(n, 3, device=dev)
        g_ind = torch.zeros(3, n, 3, device=dev)
        acc = torch.zeros(lgt.shape[0], 7, device=dev)
        g_lgt = torch.zeros_like(lgt)
        g_nrm = torch.zeros(n, 3, device=dev) if ctx.needs_input_grad[4] else None
        if n:
            lib = _lib.raw()
            _lib.check(lib.nefii_mis_shade_bwd(
                _lib.stream_ptr(dev), n, lgt.shape[0], lgt.data_ptr(), spec.data_ptr(), 1 if ctx.per_point else 0,
                rough.data_ptr(), albedo.data_ptr(), normal.data_ptr(), view.data_ptr(), wi.data_ptr(), pdf.data_ptr(),
                weight.data_ptr(), hit8.data_ptr(), indirect.data_ptr(), light.data_ptr(), g_out[0].data_ptr(),
                g_out[1].data_ptr(), g_out[2].data_ptr(), g_rough.data_ptr(), g_alb.data_ptr(), g_sr.data_ptr(),
                g_ind.data_ptr(), acc.data_ptr(), g_nrm.data_ptr() if g_nrm is not None else None))
            _lib.check(lib.nefii_sg_param_grad(_lib.stream_ptr(dev), lgt.shape[0], lgt.data_ptr(), acc.data_ptr(), 1e-6,
                                               g_lgt.data_ptr(), 0))
        rough_shape, spec_shape = ctx.shapes
        g_spec = g_sr if ctx.per_point else g_sr.sum(0, keepdim=True)
        return (g_lgt, g_spec.reshape(spec_shape), g_rough.reshape(rough_shape), g_alb, g_nrm, None, None, None, None, None, g_ind)


def mis_shade(lgtSGs, specular, roughness, albedo, normal, view, wi, pdf, weight, hit, indirect):
    """Shading half of pt_render_indirect_mlp (path_tracing_render.py:1406-1476).  hit [3,N(,1)] bool (secondary_mask),
    indirect [3,N,3].  -> dict(sg_rgb, sg_specular_rgb, sg_diffuse_rgb) [N,3]; differentiable w.r.t. lgtSGs, specular,
    roughness, albedo and indirect."""
    out = _MisShade.apply(lgtSGs, specular, roughness, albedo, normal, view, wi, pdf, weight, hit, indirect)
    return {'sg_rgb': out[0], 'sg_specular_rgb': out[1], 'sg_diffuse_rgb': out[2]}


class _BackgroundSG(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lgtSGs, dirs):
        lgt, d = _lib.f32c(lgtSGs), _lib.f32c(dirs).reshape(-1, 3)
        out = torch.empty(d.shape[0], 3, device=d.device)
        if d.shape[0]:
            _lib.check(_lib.raw().nefii_background_sg_fwd(_lib.stream_ptr(d.device), d.shape[0], lgt.shape[0], lgt.data_ptr(),
                                                          d.data_ptr(), out.data_ptr()))
        ctx.save_for_backward(lgt, d)
        return out

    @staticmethod
    def backward(ctx, g):
        lgt, d = ctx.saved_tensors
        g = _lib.f32c(g)
        acc = torch.zeros(lgt.shape[0], 7, device=d.device)
        g_lgt = torch.zeros_like(lgt)
        if d.shape[0]:
            lib = _lib.raw()
            _lib.check(lib.nefii_background_sg_bwd(_lib.stream_ptr(d.device), d.shape[0], lgt.shape[0], lgt.data_ptr(), d.data_ptr(),
                                                   g.data_ptr(), acc.data_ptr()))
            _lib.check(lib.nefii_sg_param_grad(_lib.stream_ptr(d.device), lgt.shape[0], lgt.data_ptr(), acc.data_ptr(), 1e-8,
                                               g_lgt.data_ptr(), 0))
        return g_lgt, None


def background_sg(lgtSGs, light_dir):
    """IDRNetwork.get_background_rgb for light_type 'sg' (implicit_differentiable_renderer.py:646-663)."""
    lead = list(light_dir.shape[:-1])
    return _BackgroundSG.apply(lgtSGs, light_dir).reshape(lead + [3])
