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
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):  # pragma: no cover - filled in by the backward kernel milestone
        raise NotImplementedError("render_with_sg backward is not built yet")


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
