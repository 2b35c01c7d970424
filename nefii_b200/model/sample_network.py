"""Drop-in for the reference's code/model/sample_network.py (:10-24): the differentiable ray/surface intersection of IDR
eq. 3, `x(theta) = c + (t0 - (s(x; theta) - s0) / (grad s . v0)) v`.

Same class, same ``forward(surface_output, surface_sdf_values, surface_points_grad, surface_dists, surface_cam_loc,
surface_ray_dirs)``; one fused CUDA kernel forward and one backward through the C ABI (nefii_sample_network_fwd / _bwd)
instead of a bmm, a masked assignment and five elementwise kernels.  The forward VALUE equals ``c + t0 v``; the operator
exists for its backward, dx/ds = -v / (grad s . v0), which carries the image loss into a trainable geometry
(implicit_differentiable_renderer.py:357-389, path_tracing_render.py:2132-2144)."""
import torch
import torch.nn as nn

from .. import _lib


class _SampleNetworkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, surface_output, surface_sdf_values, surface_points_grad, surface_dists, surface_cam_loc, surface_ray_dirs):
        n = surface_ray_dirs.reshape(-1, 3).shape[0]
        dev = surface_ray_dirs.device
        s, s0, t0 = [_lib.f32c(x).reshape(-1) for x in (surface_output, surface_sdf_values, surface_dists)]
        g, c, v = [_lib.f32c(x).reshape(-1, 3) for x in (surface_points_grad, surface_cam_loc, surface_ray_dirs)]
        if not (s.shape[0] == s0.shape[0] == t0.shape[0] == g.shape[0] == c.shape[0] == n):
            raise _lib.NefiiError("SampleNetwork: inconsistent input sizes")
        out = torch.empty(n, 3, device=dev, dtype=torch.float32)
        if n:
            _lib.check(_lib.raw().nefii_sample_network_fwd(_lib.stream_ptr(dev), n, s.data_ptr(), s0.data_ptr(), g.data_ptr(),
                                                           t0.data_ptr(), c.data_ptr(), v.data_ptr(), out.data_ptr()))
        ctx.save_for_backward(s, s0, g, t0, v)
        ctx.shapes = [x.shape for x in (surface_output, surface_sdf_values, surface_points_grad, surface_dists, surface_cam_loc,
                                        surface_ray_dirs)]
        return out

    @staticmethod
    def backward(ctx, g_out):
        s, s0, g, t0, v = ctx.saved_tensors
        n = v.shape[0]
        dev = v.device
        g_out = _lib.f32c(g_out).reshape(n, 3)
        need = ctx.needs_input_grad
        outs = [torch.empty(n, device=dev) if need[0] else None, torch.empty(n, device=dev) if need[1] else None,
                torch.empty(n, 3, device=dev) if need[2] else None, torch.empty(n, device=dev) if need[3] else None,
                torch.empty(n, 3, device=dev) if need[4] else None, torch.empty(n, 3, device=dev) if need[5] else None]
        if n:
            p = [o.data_ptr() if o is not None else None for o in outs]
            _lib.check(_lib.raw().nefii_sample_network_bwd(_lib.stream_ptr(dev), n, s.data_ptr(), s0.data_ptr(), g.data_ptr(),
                                                           t0.data_ptr(), v.data_ptr(), g_out.data_ptr(),
                                                           p[0], p[1], p[3], p[4], p[5], p[2]))
        return tuple(o.reshape(shape) if o is not None else None for o, shape in zip(outs, ctx.shapes))


class SampleNetwork(nn.Module):
    '''
    Represent the intersection (sample) point as differentiable function of the implicit geometry and camera parameters.
    See equation 3 in the paper for more details.
    '''

    def forward(self, surface_output, surface_sdf_values, surface_points_grad, surface_dists, surface_cam_loc, surface_ray_dirs):
        if not surface_ray_dirs.is_cuda:
            raise _lib.NefiiError("nefii_b200: SampleNetwork expects CUDA tensors (no CPU path exists)")
        return _SampleNetworkFn.apply(surface_output, surface_sdf_values, surface_points_grad, surface_dists, surface_cam_loc,
                                      surface_ray_dirs)
