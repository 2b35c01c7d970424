"""Drop-in for the reference's code/model/ray_tracing.py (RayTracing, :6-337).

Same constructor arguments and the same ``forward(sdf, cam_loc, object_mask, ray_directions)``
-> ``(points, network_object_mask, dists)`` contract; the whole trace (sphere tracing, sampler,
bisection, min-SDF sampling) runs in CUDA through the C ABI call ``nefii_ray_trace``.

The reference passes the SDF as a Python callable.  Here the callable is only used to *find* the
SDF source: an object exposing ``nefii_sdf_source()`` (our ImplicitNetwork, AnalyticSDF), a bound
method of such an object, or a closure over one -- e.g. the reference's own
``lambda x: self.implicit_network(x)[:, 0]`` (implicit_differentiable_renderer.py:344).
"""
import ctypes

import torch
import torch.nn as nn

from .. import _lib

c_void_p = ctypes.c_void_p


class TraceConfig(ctypes.Structure):
    """Mirror of ``nefii_trace_config``."""
    _fields_ = [("object_bounding_sphere", ctypes.c_float), ("sdf_threshold", ctypes.c_float),
                ("line_search_step", ctypes.c_float), ("line_step_iters", ctypes.c_int32),
                ("sphere_tracing_iters", ctypes.c_int32), ("n_steps", ctypes.c_int32),
                ("n_rootfind_steps", ctypes.c_int32)]


TRACE_TRAINING = 1
TRACE_SKIP_MIN_SDF = 2


class AnalyticSDF:
    """Union of spheres / boxes evaluated inside the tracer (bit-exact control-flow tests).
    prims: [n, 8] rows = kind (0 sphere, 1 box), centre xyz, radius | half extents, pad."""

    def __init__(self, prims, device):
        self.prims = prims.to(device=device, dtype=torch.float32).contiguous()

    def nefii_sdf_source(self):
        return 1, self.prims.data_ptr(), self.prims.shape[0], self.prims

    def __call__(self, x):
        x = _lib.f32c(x).reshape(-1, 3)
        out = torch.empty(x.shape[0], device=x.device)
        if x.shape[0]:
            _lib.check(_lib.raw().nefii_analytic_sdf_eval(_lib.stream_ptr(x.device), self.prims.data_ptr(),
                                                          self.prims.shape[0], x.shape[0], x.data_ptr(), out.data_ptr()))
        return out


def resolve_sdf_source(sdf):
    """-> (kind, pointer, n_prims, keepalive) for anything that leads to an SDF source."""
    seen = set()

    def probe(obj, depth):
        if obj is None or id(obj) in seen or depth > 3:
            return None
        seen.add(id(obj))
        if hasattr(obj, "nefii_sdf_source"):
            return obj.nefii_sdf_source()
        inner = getattr(obj, "implicit_network", None)
        if inner is not None and hasattr(inner, "nefii_sdf_source"):
            return inner.nefii_sdf_source()
        owner = getattr(obj, "__self__", None)
        if owner is not None:
            got = probe(owner, depth + 1)
            if got:
                return got
        for cell in getattr(obj, "__closure__", None) or ():
            try:
                got = probe(cell.cell_contents, depth + 1)
            except ValueError:
                got = None
            if got:
                return got
        return None

    src = probe(sdf, 0)
    if src is None:
        raise TypeError("nefii_b200 RayTracing: `sdf` must lead to an SDF source (ImplicitNetwork, AnalyticSDF, "
                        "a bound method of one, or a closure over one); arbitrary Python callables cannot run "
                        "inside the CUDA tracer")
    return src


class _TraceBuffers:
    """Persistent device buffers of one trace shape: the C side keys its CUDA-graph cache on the raw pointers, so inputs are
    copied into (and outputs out of) tensors that stay where they are from call to call."""

    def __init__(self, dev, n_batch, n_pix, n_steps):
        n = n_batch * n_pix
        self.cam = torch.empty(n_batch, 3, device=dev, dtype=torch.float32)
        self.dirs = torch.empty(n_batch, n_pix, 3, device=dev, dtype=torch.float32)
        self.obj = torch.empty(n, device=dev, dtype=torch.uint8)
        self.uniforms = torch.empty(n_steps, device=dev, dtype=torch.float32)
        self.points = torch.empty(n, 3, device=dev, dtype=torch.float32)
        self.hit = torch.empty(n, device=dev, dtype=torch.uint8)
        self.dists = torch.empty(n, device=dev, dtype=torch.float32)
        self.missing_from = n_batch      # rows [missing_from:] hold the padding rays of RayTracing._fill_missing_rays (none yet)


class RayTracing(nn.Module):
    # rays with their own origin (the integrator's secondary rays: n_pix == 1) are padded to a multiple of this many rays
    # with rays that miss the bounding sphere, so that a slowly varying ray count replays the same CUDA graph
    RAY_BUCKET = 8192
    MAX_SHAPES = 16
    # a new secondary-ray bucket also captures the graphs of the buckets within this distance (hit counts of consecutive pixel batches
    # of a scene differ by +-20 %: 72 k .. 107 k secondary rays in the bench scene)
    PRECAPTURE_RADIUS = 3

    def __init__(
            self,
            object_bounding_sphere=1.0,
            sdf_threshold=5.0e-5,
            line_search_step=0.5,
            line_step_iters=1,
            sphere_tracing_iters=10,
            n_steps=100,
            n_rootfind_steps=8,
    ):
        super().__init__()
        self.object_bounding_sphere = object_bounding_sphere
        self.sdf_threshold = sdf_threshold
        self.sphere_tracing_iters = sphere_tracing_iters
        self.line_step_iters = line_step_iters
        self.line_search_step = line_search_step
        self.n_steps = n_steps
        self.n_rootfind_steps = n_rootfind_steps
        # nefii_b200 extension (default keeps the reference's behaviour): minimal_sdf_points can be skipped
        # by callers that discard its lanes (the secondary-ray trace of the integrator).
        self.skip_min_sdf = False
        self.last_stats = None
        self.collect_stats = False
        self._ws = None
        self._linspace = {}
        self._shape_bufs = {}

    def __getstate__(self):       # device buffers whose raw pointers key the C side's graph cache are not copied / pickled
        state = dict(self.__dict__)
        state.update(_ws=None, _linspace={}, _shape_bufs={}, last_stats=None)
        return state

    def _config(self):
        return TraceConfig(self.object_bounding_sphere, self.sdf_threshold, self.line_search_step, self.line_step_iters,
                           self.sphere_tracing_iters, self.n_steps, self.n_rootfind_steps)

    def _shape_buffers(self, dev, n_batch, n_pix):
        key = (dev, n_batch, n_pix)
        b = self._shape_bufs.pop(key, None)
        if b is None:
            while len(self._shape_bufs) >= self.MAX_SHAPES:
                self._shape_bufs.pop(next(iter(self._shape_bufs)))          # oldest shape
            b = _TraceBuffers(dev, n_batch, n_pix, self.n_steps)
        self._shape_bufs[key] = b                                      # most recently used last
        return b

    def _fill_missing_rays(self, bufs, batch_size, n_rays, dev):
        """rays [batch_size:] of a bucket start far outside the bounding sphere and point away from it.  Written with fills (a
        `torch.tensor([...], device=dev)` is a synchronous host-to-device copy: it drained the stream in the middle of every forward),
        and only where the previous call left real rays: rows [bufs.missing_from:] already hold missing rays."""
        far = 4.0 * float(self.object_bounding_sphere) + 1.0
        end = min(bufs.missing_from, bufs.cam.shape[0])
        if batch_size < end:
            cam, dirs = bufs.cam[batch_size:end], bufs.dirs[batch_size:end]
            cam.zero_()
            cam[:, 2].fill_(far)
            dirs.zero_()
            dirs[:, :, 1].fill_(1.0)
            bufs.obj[n_rays:end * bufs.dirs.shape[1]].zero_()
        bufs.missing_from = batch_size

    def _launch(self, bufs, kind, ptr, n_prims, cap_batch, num_pixels, have_mask, flags, have_uniforms, dev, stats):
        key = (self.n_steps, dev)
        if key not in self._linspace:
            self._linspace[key] = torch.linspace(0, 1, steps=self.n_steps).to(dev)   # CPU linspace, as the reference
        lin = self._linspace[key]
        cfg = self._config()
        with torch.cuda.device(dev):
            _lib.check(_lib.raw().nefii_ray_trace(
                _lib.stream_ptr(dev), ctypes.byref(cfg), kind, c_void_p(ptr), n_prims, cap_batch, num_pixels,
                bufs.cam.data_ptr(), bufs.dirs.data_ptr(), bufs.obj.data_ptr() if have_mask else None, flags,
                lin.data_ptr(), bufs.uniforms.data_ptr() if have_uniforms else None,
                self._ws.data_ptr(), self._ws.numel(), bufs.points.data_ptr(), bufs.hit.data_ptr(), bufs.dists.data_ptr(), stats))

    def forward(self, sdf, cam_loc, object_mask, ray_directions, uniforms=None, skip_min_sdf=None):
        kind, ptr, n_prims, keep = resolve_sdf_source(sdf)
        dev = ray_directions.device
        if not ray_directions.is_cuda:
            raise _lib.NefiiError("nefii_b200: expected CUDA tensors (no CPU path exists)")
        batch_size, num_pixels, _ = ray_directions.shape
        n_rays = batch_size * num_pixels
        if n_rays == 0:
            return (torch.empty(0, 3, device=dev), torch.empty(0, dtype=torch.bool, device=dev), torch.empty(0, device=dev))
        # secondary-style rays: pad the batch to a bucket with rays that miss the bounding sphere
        cap_batch = batch_size
        bucketed = num_pixels == 1 and batch_size > 1
        if bucketed:
            cap_batch = (batch_size + self.RAY_BUCKET - 1) // self.RAY_BUCKET * self.RAY_BUCKET
        new_shape = (dev, cap_batch, num_pixels) not in self._shape_bufs
        bufs = self._shape_buffers(dev, cap_batch, num_pixels)
        bufs.cam[:batch_size].copy_(cam_loc.detach().reshape(batch_size, 3))
        bufs.dirs[:batch_size].copy_(ray_directions.detach())
        if bucketed:
            self._fill_missing_rays(bufs, batch_size, n_rays, dev)      # (also keeps bufs.missing_from current when nothing is padded)
        have_mask = object_mask is not None
        if have_mask:
            bufs.obj[:n_rays].copy_(object_mask.reshape(-1))
        flags = 0
        skip = self.skip_min_sdf if skip_min_sdf is None else skip_min_sdf
        have_uniforms = False
        if self.training:
            flags |= TRACE_TRAINING
            if skip:
                flags |= TRACE_SKIP_MIN_SDF
            elif uniforms is None:
                # same draw as the reference (CPU generator, ray_tracing.py:316)
                uniforms = torch.empty(self.n_steps).uniform_(0.0, 1.0)
        if uniforms is not None:
            bufs.uniforms.copy_(uniforms.detach().reshape(-1), non_blocking=True)
            have_uniforms = True
        # the workspace is shared by every shape and its pointer is part of the graph key: size it for the next bucket too, so that a
        # slowly growing ray count does not invalidate the graphs that exist
        n_cap = cap_batch * num_pixels
        grow = (cap_batch + (self.PRECAPTURE_RADIUS + 1) * self.RAY_BUCKET) * num_pixels if bucketed else n_cap
        lib = _lib.raw()
        need = int(lib.nefii_trace_workspace_bytes(kind, c_void_p(ptr), grow, self.n_steps))
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        stats = (ctypes.c_int64 * 8)() if self.collect_stats else None
        self._launch(bufs, kind, ptr, n_prims, cap_batch, num_pixels, have_mask, flags, have_uniforms, dev, stats)
        if bucketed and new_shape and _lib.raw().nefii_trace_graph_mode() == 1:
            # A ray count that hovers around a bucket boundary alternates between two shapes: capture the neighbours' graphs now (one
            # trace of rays that all miss the sphere: no SDF evaluation runs), not in the middle of somebody's timed loop.
            near = [cap_batch + d * self.RAY_BUCKET for k in range(1, self.PRECAPTURE_RADIUS + 1) for d in (k, -k)]
            for nb in near:
                if nb <= 0 or (dev, nb, num_pixels) in self._shape_bufs or len(self._shape_bufs) >= self.MAX_SHAPES - 1:
                    continue
                wb = self._shape_buffers(dev, nb, num_pixels)
                self._fill_missing_rays(wb, 0, 0, dev)
                if have_uniforms:
                    wb.uniforms.copy_(bufs.uniforms)
                self._launch(wb, kind, ptr, n_prims, nb, num_pixels, have_mask, flags, have_uniforms, dev, None)
            self._shape_bufs[(dev, cap_batch, num_pixels)] = self._shape_bufs.pop((dev, cap_batch, num_pixels))    # most recent again
        if stats is not None:
            self.last_stats = dict(n_sampler=stats[0], n_rootfind=stats[1], n_min_sdf=stats[2], n_evals=stats[3],
                                   n_rounds=stats[4])
        return bufs.points[:n_rays].clone(), bufs.hit[:n_rays].bool(), bufs.dists[:n_rays].clone()
