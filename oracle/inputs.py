"""Deterministic synthetic inputs for the BASELINE.json configs (test + bench infrastructure).

Everything is generated from fixed seeds with torch's CPU generator so that the container (no GPU),
the GPU box, the oracle and the CUDA path all see identical bytes.
"""
import math

import torch


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def synthetic_light_sgs(num_sgs=128, seed=0):
    """Random environment in the reference's init style (sg_envmap_material.py:127-152):
    sharpness 20 + 100|N(0,1)|, grey-ish amplitudes normalised to ~2*pi energy, Fibonacci lobes."""
    g = _gen(seed)
    lgt = torch.randn(num_sgs, 7, generator=g)
    lgt[:, -2:] = lgt[:, -3:-2].expand(-1, 2)
    lgt[:, 3:4] = 20. + torch.abs(lgt[:, 3:4] * 100.)
    lam = torch.abs(lgt[:, 3:4])
    mu = torch.abs(lgt[:, 4:])
    energy = mu * 2.0 * math.pi / lam * (1.0 - torch.exp(-2.0 * lam))
    lgt[:, 4:] = torch.abs(lgt[:, 4:]) / torch.sum(energy, dim=0, keepdim=True) * 2. * math.pi
    i = torch.arange(num_sgs, dtype=torch.float64)
    y = 1 - (i / float(num_sgs - 1)) * 2
    radius = torch.sqrt(1 - y * y)
    theta = math.pi * (3. - math.sqrt(5.)) * i
    lgt[:, 0] = (torch.cos(theta) * radius).float()
    lgt[:, 1] = y.float()
    lgt[:, 2] = (torch.sin(theta) * radius).float()
    return lgt


def shading_inputs(n_rays=1024, seed=0):
    """cfg 1: unit normals, view = normalize(n + 0.5 randn) (a few grazing /
    back-facing rays on purpose), albedo ~ U(0,1)."""
    g = _gen(seed)
    normal = torch.nn.functional.normalize(torch.randn(n_rays, 3, generator=g), dim=-1)
    view = torch.nn.functional.normalize(normal + 0.5 * torch.randn(n_rays, 3, generator=g), dim=-1)
    albedo = torch.rand(n_rays, 3, generator=g)
    return normal, view, albedo


ROUGHNESS_SWEEP = (0.089, 0.3, 0.5, 1.0)
