"""Oracle (test infrastructure): the near-field indirect-illumination integrator, restated.

Restates (reference code/model/path_tracing_render.py)
  * rotate_to_normal :12-33          * cos_sampling :128-156 / pdf_fn_cos :159-165
  * brdf_sampling :61-103 / pdf_fn_brdf_gxx :106-125
  * mix_sg_sampling :168-242 / pdf_fn_mix_sg :245-271
  * power_heuristic_list :390-401    * sg_fn :404-413
  * pt_render_diff_shadow_indirect_mlp (diff_geo=False, speed_first=True) :1265-1487, split into
      sample_directions()  -- the three importance-sampled directions + the 3x3 pdf matrix (:1290-1325)
      shade()              -- light evaluation, GGX BRDF, MIS weight, visibility blend (:1406-1476)
  * get_visibility_and_indirect_light (diff_geo=False) :2109-2166 is orchestration and lives with the
    renderer glue (oracle/pipeline.py).

The reference draws its random numbers with torch.rand inside the sampling functions; here the seven
uniforms per ray are an explicit input `u` [N,7] in the reference's draw order
(cos r1, cos r2, ggx r1, ggx r2, mix r0, mix r1, mix r2) so that both sides of a parity test consume
identical numbers.

Parity status: PINNED -- tests/test_oracle_mis.py patches torch.rand in the real reference with the same
uniforms (when /root/reference is present) and compares, and checks tests/golden/mis_*.npz.
"""
import math

import torch

EPS = 1e-6


def _dot(a, b):
    return torch.sum(a * b, dim=-1, keepdim=True)


def to_frame(local, n):
    """Rotate `local` (z = normal) into the frame of unit vector n.  path_tracing_render.py:12-33."""
    x_axis = torch.zeros_like(n)
    x_axis[..., 0] = 1
    y_axis = torch.zeros_like(n)
    y_axis[..., 1] = 1
    up = torch.where((n[..., 0:1] > 0.9).expand(n.shape), y_axis, x_axis)
    t = torch.cross(up, n, dim=-1)
    t = t / (torch.norm(t, dim=-1, keepdim=True) + EPS)
    s = torch.cross(t, n, dim=-1)
    return local[..., :1] * t + local[..., 1:2] * s + local[..., 2:] * n


def _sph(theta, phi):
    return torch.cat([theta.sin() * phi.cos(), theta.sin() * phi.sin(), theta.cos()], dim=-1)


def unpack(lgt):
    axis = lgt[..., :3] / (torch.norm(lgt[..., :3], dim=-1, keepdim=True) + EPS)
    return axis, torch.abs(lgt[..., 3:4]), torch.abs(lgt[..., -3:])


# ---- pdfs ---------------------------------------------------------------------------------------
def pdf_cos(wi, normal):
    return torch.clamp(_dot(wi, normal), min=EPS) / math.pi


def pdf_ggx(wi, normal, view, rough):
    h = wi + view
    h = h / torch.norm(h, dim=-1, keepdim=True)
    bad = torch.isnan(h)
    h = torch.where(bad, normal, h)
    c = torch.clamp(_dot(h, normal), min=EPS)
    root = c ** 2 + (1 - c ** 2) / (rough ** 4)
    pdf_h = c / (math.pi * (rough ** 4) * root * root)
    hv = torch.clamp(_dot(h, view), min=EPS)
    return pdf_h / (4 * hv)


def _mixture_weights(normal, lgt):
    """alpha_k proportional to (sum of amplitudes) * max(n . axis_k, 1e-6).  lgt [N,M,7]."""
    axis, sharp, amp = unpack(lgt)
    energy = amp.sum(dim=-1, keepdim=True)
    M = lgt.shape[-2]
    cosn = torch.sum(normal.unsqueeze(-2).expand(normal.shape[:-1] + (M, 3)) * axis, dim=-1, keepdim=True)
    w = energy * torch.clamp(cosn, EPS)
    return axis, sharp, energy, w / w.sum(dim=-2, keepdim=True)


def pdf_mix(wi, normal, lgt):
    axis, sharp, _, alpha = _mixture_weights(normal, lgt)
    c = sharp / (2 * math.pi * (1 - torch.exp(-2.0 * sharp)))
    M = lgt.shape[-2]
    d = torch.sum(wi.unsqueeze(-2).expand(wi.shape[:-1] + (M, 3)) * axis, dim=-1, keepdim=True)
    return (alpha * c * torch.exp(sharp * (d - 1))).sum(dim=-2)


# ---- samplers -----------------------------------------------------------------------------------
def sample_cos(normal, r1, r2):
    theta = torch.arccos(torch.sqrt(1 - r1))
    phi = 2 * math.pi * r2
    wi = to_frame(_sph(theta, phi), normal)
    return wi, theta.cos() / math.pi


def sample_ggx(normal, rough, view, r1, r2):
    theta = torch.arctan(rough ** 2 * torch.sqrt(r1 / (1 - r1)))
    phi = 2 * math.pi * r2
    h = to_frame(_sph(theta, phi), normal)
    wi = 2 * _dot(view, h) * h - view
    return wi, pdf_ggx(wi, normal, view, rough)


def sample_mix(normal, lgt, r0, r1, r2):
    """lgt [N,M,7] (expanded view), r0/r1/r2 [N,1]."""
    axis, sharp, energy, alpha = _mixture_weights(normal, lgt)
    right = torch.cumsum(alpha, dim=-2)
    left = right - alpha
    right = right.clone()
    left = left.clone()
    right[..., -1, :] = 1.0
    left[..., 0, :] = 0.0
    rr = r0.unsqueeze(-1)
    pick = (rr >= left) & (rr < right)                                   # exactly one lobe per ray
    k = torch.max(pick.float(), dim=-2, keepdim=True)[1]                # [N,1,1]
    axis_k = torch.gather(axis, -2, k.expand(k.shape[:-2] + (1, 3))).squeeze(-2)
    sharp_k = torch.gather(sharp, -2, k).squeeze(-2)
    c_k = sharp_k / (2 * math.pi * (1 - torch.exp(-2 * sharp_k)))
    theta = torch.arccos(1.0 / sharp_k * torch.log(torch.clamp(1 - sharp_k * r1 / (2 * math.pi * c_k), EPS)) + 1)
    phi = 2 * math.pi * r2
    wi = to_frame(_sph(theta, phi), axis_k)
    return wi, pdf_mix(wi, normal, lgt)


def sample_directions(lgtSGs, roughness, normal, view, u):
    """-> wi [3,N,3], pdf [3,N,1] (clamped at 1e-6), pdf_matrix [3(sample),3(strategy),N,1]."""
    M = lgtSGs.shape[0]
    N = normal.shape[0]
    lgt = lgtSGs.reshape(1, M, 7).expand(N, M, 7)
    cols = [u[:, i:i + 1] for i in range(7)]
    s = [sample_cos(normal, cols[0], cols[1]),
         sample_ggx(normal, roughness, view, cols[2], cols[3]),
         sample_mix(normal, lgt, cols[4], cols[5], cols[6])]
    wi = [a for a, _ in s]
    pdf = [torch.clamp(b, min=EPS) for _, b in s]
    fns = [lambda w: pdf_cos(w, normal), lambda w: pdf_ggx(w, normal, view, roughness), lambda w: pdf_mix(w, normal, lgt)]
    mat = [[pdf[i] if i == j else fns[j](wi[i]) for j in range(3)] for i in range(3)]
    return torch.stack(wi), torch.stack(pdf), torch.stack([torch.stack(r) for r in mat])


# ---- shading ------------------------------------------------------------------------------------
def shade(lgtSGs, specular_reflectance, roughness, diffuse_albedo, normal, view, wi, pdf, pdf_matrix,
          visibility, indirect):
    """Sum over the three sample types of the MIS-weighted specular + diffuse estimates.
    wi [3,N,3], pdf [3,N,1], pdf_matrix [3,3,N,1], visibility [3,N,1], indirect [3,N,3]."""
    M = lgtSGs.shape[0]
    N = normal.shape[0]
    lgt = lgtSGs.reshape(1, M, 7).expand(N, M, 7)
    axis, sharp, amp = unpack(lgt)
    spec_total, diff_total, rgb_total = 0, 0, 0
    for i in range(3):
        w = wi[i]
        light = (amp * torch.exp(sharp * (torch.sum(w.unsqueeze(-2).expand(N, M, 3) * axis, dim=-1, keepdim=True) - 1))).sum(-2)
        half = w + view
        half = half / (torch.norm(half, dim=-1, keepdim=True) + EPS)
        nh = torch.clamp(_dot(normal, half), min=0)
        r2 = roughness ** 2
        root = nh ** 2 + (1 - nh ** 2) / (r2 ** 2)
        D = 1.0 / (math.pi * (r2 ** 2) * root * root)
        vh = torch.clamp(_dot(view, half), min=0.)
        F = specular_reflectance + (1. - specular_reflectance) * torch.pow(2.0, -(5.55473 * vh + 6.8316) * vh)
        d1 = torch.clamp(_dot(view, normal), min=0.)
        d2 = torch.clamp(_dot(w, normal), min=0.)
        k = (roughness + 1.) * (roughness + 1.) / 8.
        G = (d1 / (d1 * (1 - k) + k + EPS)) * (d2 / (d2 * (1 - k) + k + EPS))
        fs = F * D * G / (4 * d1 * d2 + EPS)
        total = 0
        for j in range(3):
            total = total + (1 * pdf_matrix[i][j]) ** 2
        weight = (1 * pdf_matrix[i][i]) ** 2 / torch.clamp(total, min=EPS)
        light_all = light * visibility[i] + (1 - visibility[i]) * indirect[i]
        cosn = torch.clamp(_dot(w, normal), min=0)
        spec = torch.clamp(weight * light_all * fs * cosn / pdf[i], min=0.)
        diff = torch.clamp(weight * light_all * (diffuse_albedo / math.pi) * cosn / pdf[i], min=0.)
        spec_total = spec_total + spec
        diff_total = diff_total + diff
        rgb_total = rgb_total + (spec + diff)
    return {'sg_rgb': rgb_total, 'sg_specular_rgb': spec_total, 'sg_diffuse_rgb': diff_total}
