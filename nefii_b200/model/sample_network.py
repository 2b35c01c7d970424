"""SampleNetwork (reference code/model/sample_network.py:10-24): the differentiable ray/surface intersection of
IDR eq. 3.  Its forward value equals cam + t0 * v; it only matters when gradients flow into the geometry, which is
not the case on the accelerated path (step 2 freezes the geometry, eval has no gradients).  Kept as plain torch
so that code constructing IDRNetwork finds the attribute."""
import torch
import torch.nn as nn


class SampleNetwork(nn.Module):
    def forward(self, surface_output, surface_sdf_values, surface_points_grad, surface_dists, surface_cam_loc, surface_ray_dirs):
        dirs0 = surface_ray_dirs.detach()
        dot = torch.bmm(surface_points_grad.view(-1, 1, 3), dirs0.view(-1, 3, 1)).squeeze(-1)
        dot = torch.where(dot.abs() < 1e-8, torch.full_like(dot, 1e-8), dot)
        t_theta = surface_dists - (surface_output - surface_sdf_values) / dot
        return surface_cam_loc + t_theta * surface_ray_dirs
