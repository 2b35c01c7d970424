"""Oracle (test infrastructure): spherical-Gaussian shading of the reference, restated.

Restates, in torch tensor ops that run on CPU or CUDA and in float32 or float64:
  * render_with_sg      -- reference code/model/sg_render.py:164-295
  * _sg_product         -- reference code/model/sg_render.py:141-158  (lambda_trick)
  * _hemisphere_integral-- reference code/model/sg_render.py:112-138  (hemisphere_int)
  * sg_eval / background_sg -- reference code/model/path_tracing_render.py:404-413 and
                           code/model/implicit_differentiable_renderer.py:646-663

Parity status: PINNED.  tests/test_oracle_vs_reference.py runs this file against the real
reference modules imported from /root/reference (when present) and against the committed
fixtures in tests/golden/ that oracle/make_golden.py generated from the reference.

The per-term floating-point operation order of the reference is kept on purpose (the SG
integrals cancel two large terms) -- only the code structure differs: rays are
flattened to one axis and every tensor is laid out [N, M, K, C].
"""
import math

import torch

EPS = 1e-6
COS_MU = 32.7080      # clamped-cosine lobe amplitude, sg_render.py:243
COS_LAMBDA = 0.0315   # clamped-cosine lobe sharpness, sg_render.py:244
COS_ALPHA = 31.7003   # constant-term correction,      sg_render.py:245


def _dot(a, b):
    return torch.sum(a * b, dim=-1, keepdim=True)


def _unit(v):
    return v / (torch.norm(v, dim=-1, keepdim=True) + EPS)


def unpack_light(lgt):
    """lgtSGs [..., 7] -> (unit lobe axis, sharpness >= 0, amplitude >= 0).  sg_render.py:194-196."""
    return _unit(lgt[..., :3]), torch.abs(lgt[..., 3:4]), torch.abs(lgt[..., -3:])


def _hemisphere_integral(sharp, cos_beta):
    """Fitted integral of an SG over the hemisphere around an axis at angle beta.  sg_render.py:112-138."""
    sharp = sharp + EPS
    inv = 1. / sharp
    t = torch.sqrt(sharp) * (1.6988 + 10.8438 * inv) / (1. + 6.2201 * inv + 10.2415 * inv * inv)
    ea = torch.exp(-t)
    front = (cos_beta >= 0).to(sharp.dtype)
    eb = torch.exp(-t * torch.clamp(cos_beta, min=0.))
    s_front = (1. - ea * eb) / (1. - ea + eb - ea * eb)
    b = torch.exp(t * torch.clamp(cos_beta, max=0.))
    s_back = (b - ea) / ((1. - ea) * (b + 1.))
    s = front * s_front + (1. - front) * s_back
    lower = 2. * math.pi / sharp * (torch.exp(-sharp) - torch.exp(-2. * sharp))
    upper = 2. * math.pi / sharp * (1. - torch.exp(-sharp))
    return lower * (1. - s) + upper * s


def _sg_product(axis1, sharp1, amp1, axis2, sharp2, amp2):
    """Product of two SGs assuming sharp1 << sharp2.  sg_render.py:141-158."""
    ratio = sharp1 / sharp2
    cosang = _dot(axis1, axis2)
    scale = torch.sqrt(ratio * ratio + 1. + 2. * ratio * cosang)
    scale = torch.min(scale, ratio + 1.)
    sharp3 = sharp2 * scale
    w1 = ratio / scale
    w2 = 1. / scale
    shift = sharp2 * (scale - ratio - 1.)
    axis3 = w1 * axis1 + w2 * axis2
    amp3 = amp1 * amp2 * torch.exp(shift)
    return axis3, sharp3, amp3


def _cosine_lobe_integral(normal, axis, sharp, amp):
    """(SG x clamped cosine) hemisphere integral, sg_render.py:243-252 / :279-285."""
    axis_p, sharp_p, amp_p = _sg_product(normal, COS_LAMBDA, COS_MU, axis, sharp, amp)
    c1 = _dot(axis_p, normal)
    c2 = _dot(axis, normal)
    return amp_p * _hemisphere_integral(sharp_p, c1) - amp * COS_ALPHA * _hemisphere_integral(sharp, c2)


def render_with_sg(lgtSGs, specular_reflectance, roughness, diffuse_albedo, normal, viewdirs,
                   blending_weights=None, diffuse_rgb=None):
    """Same signature / return dict as the reference's render_with_sg (sg_render.py:164-295)."""
    M = lgtSGs.shape[0]
    K = specular_reflectance.shape[0]
    assert K == roughness.shape[0]
    lead = list(normal.shape[:-1])
    n = normal.reshape(-1, 1, 1, 3).expand(-1, M, K, 3)
    v = viewdirs.reshape(-1, 1, 1, 3).expand(-1, M, K, 3)
    N = n.shape[0]

    lgt = lgtSGs.reshape(1, M, 1, 7).expand(N, M, K, 7)
    l_axis, l_sharp, l_amp = unpack_light(lgt)

    # GGX NDF as an SG around the normal, sg_render.py:199-203
    inv_r4 = 1. / (roughness * roughness * roughness * roughness)            # [K,1]
    b_sharp = (2. * inv_r4).reshape(1, 1, K, 1).expand(N, M, K, 1)
    b_amp = (inv_r4 / math.pi).expand(K, 3).reshape(1, 1, K, 3).expand(N, M, K, 3)

    # spherical warp to the reflection direction, sg_render.py:206-213
    nv = torch.clamp(_dot(n, v), min=0.)
    w_axis = _unit(2 * nv * n - v)
    w_sharp = b_sharp / (4 * nv + EPS)

    # Fresnel + shadowing evaluated at the warped lobe centre, sg_render.py:216-236
    half = _unit(w_axis + v)
    vh = torch.clamp(_dot(v, half), min=0.)
    spec = specular_reflectance.reshape(1, 1, K, 3).expand(N, M, K, 3)
    fresnel = spec + (1. - spec) * torch.pow(2.0, -(5.55473 * vh + 6.8316) * vh)
    d1 = torch.clamp(_dot(w_axis, n), min=0.)
    d2 = torch.clamp(_dot(v, n), min=0.)
    k = (roughness + 1.) * (roughness + 1.) / 8.
    g1 = d1 / (d1 * (1 - k) + k + EPS)
    g2 = d2 / (d2 * (1 - k) + k + EPS)
    moi = fresnel * (g1 * g2) / (4 * d1 * d2 + EPS)
    w_amp = b_amp * moi

    # light SG x BRDF SG, then x cosine, sg_render.py:239-252
    p_axis, p_sharp, p_amp = _sg_product(l_axis, l_sharp, l_amp, w_axis, w_sharp, w_amp)
    specular = _cosine_lobe_integral(n, p_axis, p_sharp, p_amp)
    if blending_weights is None:
        specular = specular.sum(dim=-2).sum(dim=-2)
    else:
        bw = blending_weights.reshape(N, K)
        specular = (specular.sum(dim=-3) * bw.unsqueeze(-1)).sum(dim=-2)
    specular = torch.clamp(specular, min=0.)

    # diffuse term, sg_render.py:268-286
    if diffuse_rgb is None:
        albedo = (diffuse_albedo.reshape(-1, 3) / math.pi).reshape(N, 1, 1, 3).expand(N, M, 1, 3)
        d_axis = l_axis[:, :, :1]
        d_amp = l_amp[:, :, :1] * albedo
        d_sharp = l_sharp[:, :, :1]
        # NB: the reference keeps the K axis on `normal` here, so for K > 1 the diffuse term is
        # broadcast over (and then summed across) the K base materials -- reproduced as is.
        diffuse = _cosine_lobe_integral(n, d_axis, d_sharp, d_amp)
        diffuse = diffuse.sum(dim=-2).sum(dim=-2)
        diffuse_rgb = torch.clamp(diffuse, min=0.)
    else:
        diffuse_rgb = diffuse_rgb.reshape(N, 3)

    specular = specular.reshape(lead + [3])
    diffuse_rgb = diffuse_rgb.reshape(lead + [3])
    return {
        'sg_rgb': specular + diffuse_rgb,
        'sg_specular_rgb': specular,
        'sg_diffuse_rgb': diffuse_rgb,
        'sg_diffuse_albedo': diffuse_albedo,
    }


def sg_eval(direction, axis, sharp, amp):
    """amp * exp(sharp * (direction . axis - 1)).  path_tracing_render.py:404-413."""
    return amp * torch.exp(sharp * (_dot(direction, axis) - 1))


def background_sg(lgtSGs, light_dir):
    """Environment radiance seen along miss rays.  implicit_differentiable_renderer.py:646-663.

    Note the reference normalises lobes with +1e-8 here (not 1e-6)."""
    M = lgtSGs.shape[0]
    lead = list(light_dir.shape[:-1])
    d = light_dir.reshape(-1, 1, 3).expand(-1, M, 3)
    lgt = lgtSGs.reshape(1, M, 7).expand(d.shape[0], M, 7)
    axis = lgt[..., :3] / (torch.norm(lgt[..., :3], dim=-1, keepdim=True) + 1e-8)
    sharp = torch.abs(lgt[..., 3:4])
    amp = torch.abs(lgt[..., -3:])
    return sg_eval(d, axis, sharp, amp).sum(-2).reshape(lead + [3])
