"""Drop-in for the reference's code/model/sg_render.py (render_with_sg, :164-295).

Same signature and return dict; the math runs in one fused CUDA kernel (csrc/sg_render.cu) through
the C ABI call nefii_sg_render_fwd.
"""
import torch

from .. import _lib

TINY_NUMBER = 1e-6


class _RenderWithSG(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lgtSGs, specular_reflectance, roughness, diffuse_albedo, normal, viewdirs, blending_weights):
        lib = _lib.raw()
        M = lgtSGs.shape[0]
        K = specular_reflectance.shape[0]
        n = _lib.f32c(normal).reshape(-1, 3)
        v = _lib.f32c(viewdirs).reshape(-1, 3)
        a = _lib.f32c(diffuse_albedo).reshape(-1, 3)
        N = n.shape[0]
        lgt = _lib.f32c(lgtSGs)
        spec = _lib.f32c(specular_reflectance.expand(K, 3))
        rough = _lib.f32c(roughness)
        bw = _lib.f32c(blending_weights).reshape(N, K) if blending_weights is not None else None
        out = torch.empty(3, N, 3, device=n.device, dtype=torch.float32)
        _lib.check(lib.nefii_sg_render_fwd(
            _lib.stream_ptr(n.device), N, M, K, _lib.dptr(lgt), _lib.dptr(spec), _lib.dptr(rough), _lib.dptr(a),
            _lib.dptr(n), _lib.dptr(v), _lib.dptr(bw, allow_none=True),
            _lib.dptr(out[0]), _lib.dptr(out[1]), _lib.dptr(out[2])))
        ctx.save_for_backward(lgt, spec, rough, a, n, v, out, *([bw] if bw is not None else []))
        ctx.has_blend = bw is not None
        ctx.blend_shape = blending_weights.shape if bw is not None else None
        ctx.shapes = (specular_reflectance.shape, roughness.shape, diffuse_albedo.shape)
        ctx.normal_shape = normal.shape
        return out

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.needs_input_grad[5]:
            raise _lib.NefiiError("nefii_b200: render_with_sg has no gradient w.r.t. the view directions (camera training is "
                                  "outside the accelerated path)")
        lgt, spec, rough, a, n, v, out = ctx.saved_tensors[:7]
        bw = ctx.saved_tensors[7] if ctx.has_blend else None
        lib = _lib.raw()
        dev = n.device
        N, M, K = n.shape[0], lgt.shape[0], spec.shape[0]
        g = _lib.f32c(grad_out)
        acc = torch.zeros(M, 7, device=dev)
        g_rough = torch.zeros(K, device=dev)
        g_spec = torch.zeros(K, 3, device=dev)
        g_alb = torch.zeros(N, 3, device=dev)
        g_nrm = torch.zeros(N, 3, device=dev) if ctx.needs_input_grad[4] else None
        g_bw = torch.zeros(N, K, device=dev) if (bw is not None and ctx.needs_input_grad[6]) else None
        g_lgt = torch.zeros_like(lgt)
        if N:
            _lib.check(lib.nefii_sg_render_bwd(
                _lib.stream_ptr(dev), N, M, K, lgt.data_ptr(), spec.data_ptr(), rough.data_ptr(), a.data_ptr(), n.data_ptr(),
                v.data_ptr(), out[1].data_ptr(), out[2].data_ptr(), g[0].data_ptr(), g[1].data_ptr(), g[2].data_ptr(),
                acc.data_ptr(), g_rough.data_ptr(), g_spec.data_ptr(), g_alb.data_ptr(),
                g_nrm.data_ptr() if g_nrm is not None else None, bw.data_ptr() if bw is not None else None,
                g_bw.data_ptr() if g_bw is not None else None))
            _lib.check(lib.nefii_sg_param_grad(_lib.stream_ptr(dev), M, lgt.data_ptr(), acc.data_ptr(), 1e-6, g_lgt.data_ptr(), 0))
        spec_shape, rough_shape, alb_shape = ctx.shapes
        g_spec_in = g_spec if spec_shape[-1] == 3 else g_spec.sum(-1, keepdim=True)
        g_nrm_in = g_nrm.reshape(ctx.normal_shape) if g_nrm is not None else None
        g_bw_in = g_bw.reshape(ctx.blend_shape) if g_bw is not None else None
        return g_lgt, g_spec_in.reshape(spec_shape), g_rough.reshape(rough_shape), g_alb.reshape(alb_shape), g_nrm_in, None, g_bw_in


def render_with_sg(lgtSGs, specular_reflectance, roughness, diffuse_albedo, normal, viewdirs,
                   blending_weights=None, diffuse_rgb=None):
    """
    :param lgtSGs: [M, 7]
    :param specular_reflectance: [K, 3]
    :param roughness: [K, 1]; values must be positive
    :param diffuse_albedo: [..., 3]; values must lie in [0,1]
    :param normal: [..., 3]; ----> camera; must have unit norm
    :param viewdirs: [..., 3]; ----> camera; must have unit norm
    :param blending_weights: [..., K]; values must be positive, and sum to one along last dimension
    :return dict with sg_rgb, sg_specular_rgb, sg_diffuse_rgb, sg_diffuse_albedo ([..., 3] each)
    """
    K = specular_reflectance.shape[0]
    assert (K == roughness.shape[0])          # same assert as sg_render.py:177
    dots_shape = list(normal.shape[:-1])
    out = _RenderWithSG.apply(lgtSGs, specular_reflectance, roughness, diffuse_albedo, normal, viewdirs,
                              blending_weights)
    specular_rgb = out[1].reshape(dots_shape + [3])
    if diffuse_rgb is None:
        diffuse_rgb = out[2].reshape(dots_shape + [3])
        rgb = out[0].reshape(dots_shape + [3])
    else:
        rgb = specular_rgb + diffuse_rgb
    return {'sg_rgb': rgb,
            'sg_specular_rgb': specular_rgb,
            'sg_diffuse_rgb': diffuse_rgb,
            'sg_diffuse_albedo': diffuse_albedo}
