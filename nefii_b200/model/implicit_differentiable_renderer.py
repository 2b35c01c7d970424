"""Drop-in for the reference's code/model/implicit_differentiable_renderer.py.

``IDRNetwork(conf)`` / ``.forward(input, with_point=False)`` keep the reference's signature, output-dict keys
(:460-477), attributes (``implicit_network``, ``rendering_network``, ``envmap_material_network``, ``ray_tracer``,
``freeze_*``) and ``state_dict`` keys (``implicit_network.lin{l}.{weight_g,weight_v,bias}``, ...), so
``train.model_class = nefii_b200.model.implicit_differentiable_renderer.IDRNetwork`` is the only change a
conf needs (utils/general.py:10-16 get_class is the plug-in point).

Scope: the per-ray-batch rendering path -- step 2 training with FROZEN geometry (the fused fast path),
rendering, evaluation, and the same forward with a TRAINABLE geometry (`training and not freeze_geometry`, reference
:354-389: eikonal samples, d sdf/dx with a graph, SampleNetwork; a composition of twice-differentiable tcgen05 products).
Everything device-side runs through the C ABI; there is no PyTorch fallback.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, integrator, mlp, ops
from ..utils import rend_util
from .embedder import get_embedder
from .path_tracing_render import pt_render_indirect_mlp
from .ray_tracing import RayTracing
from .sample_network import SampleNetwork
from .sg_envmap_material import EnvmapMaterialNetwork
from .sg_render import render_with_sg


def _effective_weight(lin):
    """g * v / ||v|| for weight-normed layers (same op torch's weight_norm hook runs), else the plain weight."""
    if hasattr(lin, "weight_g"):
        return torch._weight_norm(lin.weight_v, lin.weight_g, 0)
    return lin.weight


class ImplicitNetwork(nn.Module):
    def __init__(
            self,
            feature_vector_size,
            d_in,
            d_out,
            dims,
            geometric_init=True,
            bias=1.0,
            skip_in=(),
            weight_norm=True,
            multires=0,
            use_last_as_f=False
    ):
        super().__init__()
        if d_in != 3 or d_out != 1 or len(skip_in) > 1:
            raise NotImplementedError("nefii_b200: ImplicitNetwork is built for d_in 3, d_out 1 and at most one skip connection "
                                      "(conf.conf, conf_neus.conf)")
        if use_last_as_f and feature_vector_size != dims[-1]:
            raise ValueError("use_last_as_f needs feature_vector_size == dims[-1]")       # the reference asserts the same
        if len(set(dims)) != 1:
            raise NotImplementedError("nefii_b200: hidden layers must share one width")
        self.feature_vector_size = feature_vector_size
        dims = [d_in] + list(dims) + [d_out if use_last_as_f else d_out + feature_vector_size]
        self.embed_fn = None
        self.multires = multires
        if multires > 0:
            self.embed_fn, input_ch = get_embedder(multires)
            dims[0] = input_ch
        self.num_layers = len(dims)
        self.skip_in = skip_in
        self.use_last_as_f = use_last_as_f
        for l in range(0, self.num_layers - 1):
            out_dim = dims[l + 1] - dims[0] if l + 1 in self.skip_in else dims[l + 1]
            lin = nn.Linear(dims[l], out_dim)
            if geometric_init:
                if l == self.num_layers - 2:
                    torch.nn.init.normal_(lin.weight, mean=np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                    torch.nn.init.constant_(lin.bias, -bias)
                elif multires > 0 and l == 0:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.constant_(lin.weight[:, 3:], 0.0)
                    torch.nn.init.normal_(lin.weight[:, :3], 0.0, np.sqrt(2) / np.sqrt(out_dim))
                elif multires > 0 and l in self.skip_in:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
                    torch.nn.init.constant_(lin.weight[:, -(dims[0] - 3):], 0.0)
                else:
                    torch.nn.init.constant_(lin.bias, 0.0)
                    torch.nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, "lin" + str(l), lin)
        self.softplus = nn.Softplus(beta=100)
        self._width = dims[1]
        self._n_hidden = self.num_layers - 2
        self._sdf_mlp = None
        self._packed_versions = None
        # set by IDRNetwork while it renders with a TRAINABLE geometry: forward() then keeps a graph from the feature columns to
        # the parameters (the fused trainable path returns them detached, which is all the step-1 recipe needs)
        self.differentiable_features = False

    # ---- packed-weight management ---------------------------------------------------------------
    def _layers(self):
        return [getattr(self, "lin" + str(l)) for l in range(self.num_layers - 1)]

    def _sync(self, device):
        params = list(self.parameters())
        versions = tuple((p.data_ptr(), p._version) for p in params) + (str(device),)
        if self._sdf_mlp is None or self._sdf_mlp.device != torch.device(device):
            skip = self.skip_in[0] if len(self.skip_in) else 0
            self._sdf_mlp = ops.SdfMlp(n_freqs=self.multires, width=self._width, n_hidden=self._n_hidden,
                                       skip_layer=skip, device=device,
                                       d_feat=0 if self.use_last_as_f else self.feature_vector_size)
            self._packed_versions = None
        if versions != self._packed_versions:
            with torch.no_grad():
                ws = [_effective_weight(l) for l in self._layers()]
                bs = [l.bias for l in self._layers()]
            self._sdf_mlp.set_weights(ws, bs)
            self._packed_versions = versions
        return self._sdf_mlp

    # copy.deepcopy / pickling of the module (EMA copies, torch.save(model)): the native handle and its packed weights belong to
    # THIS object (a copied raw pointer would be freed twice); the copy rebuilds them on first use
    def __getstate__(self):
        state = dict(self.__dict__)
        state["_sdf_mlp"] = None
        state["_packed_versions"] = None
        return state

    def nefii_sdf_source(self):
        dev = next(self.parameters()).device
        net = self._sync(dev)
        return 0, net.handle.value, 0, net

    def _check_frozen(self):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise _lib.NefiiError("nefii_b200: the SDF network runs inference-only on the accelerated path; call "
                                  "model.freeze_geometry() (step 2 / rendering) or wrap the call in torch.no_grad()")

    # accuracy tier of the stand-alone evaluations (sdf_output, features, normals): K blocks per TMEM partial of the layer GEMM,
    # 0 = library default (as accurate as 1 since the truncation compensation, csrc/mlp_gemm.cu)
    EVAL_FLUSH = 0

    def evaluate(self, x, want_feat=False, want_grad=False):
        """Fused pass: (sdf [N], feature [N,W] | None, d sdf/dx [N,3] | None)."""
        self._check_frozen()
        net = self._sync(x.device)
        return net.eval(x.detach(), want_feat=want_feat, want_grad=want_grad, k_flush=self.EVAL_FLUSH)

    def invalidate(self):
        """Drop the packed copy of the weights.  The cache is keyed on the parameters' (data_ptr, _version): in-place edits
        through `.data` (checkpoint surgery, converters) do not bump the version -- call this after such edits."""
        self._packed_versions = None

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate()
        return out

    def _trainable(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def _forward_trainable(self, input):
        """Step-1 geometry fitting (training/geometry_train.py:366-376: `geometry_model(points)[:, 0:1]` against sampled SDF
        values): the same layers on the trainable tcgen05 stack (nefii_b200/mlp.py: forward keeps the activation planes,
        backward = weight-gradient and data-gradient GEMMs), skip concat included; gradients reach every lin*.weight_g /
        weight_v / bias through torch's weight-norm fold.  The feature columns are returned detached (nothing in step 1
        reads them); d sdf/dx with create_graph (the eikonal term) is not built -- gradient() raises for it."""
        from ..mlp import dense_mlp
        if not self.use_last_as_f:
            raise _lib.NefiiError("nefii_b200: the trainable SDF stack (step-1 geometry fitting) is built for the use_last_as_f "
                                  "layout of conf.conf")
        layers = self._layers()
        ws = [_effective_weight(l) for l in layers]
        bs = [l.bias for l in layers]
        skip = self.skip_in[0] if len(self.skip_in) else 0
        y, feat = dense_mlp([(input.detach(), self.multires if self.multires > 0 else -1)], ws, bs, ops.ACT_SOFTPLUS100, skip=skip,
                            skip_scale=float(1.0 / np.sqrt(2)), want_hidden=True)
        return torch.cat([y, feat], dim=-1)

    def _forward_autograd(self, input):
        """The same network as a composition of twice-differentiable pieces: every Linear is `mlp.gemm_nt` (tcgen05 layer GEMM
        whose backward is built from itself), the positional encoding, bias, Softplus(100) and the skip concat are torch ops.
        Slower than the fused paths (activations round-trip in fp32, no epilogue fusion); used where a graph THROUGH the input
        is needed: gradient(create_graph=True) and inputs that carry a gradient (reference :85-123)."""
        from ..mlp import gemm_nt
        pe = self.embed_fn(input) if self.embed_fn is not None else input
        x = pe
        layers = self._layers()
        feature_vector = None
        for l, lin in enumerate(layers):
            if self.use_last_as_f and l == self.num_layers - 2:
                feature_vector = x
            if l in self.skip_in:
                x = torch.cat([x, pe], 1) / np.sqrt(2)
            x = gemm_nt(x, _effective_weight(lin)) + lin.bias
            if l < self.num_layers - 2:
                x = self.softplus(x)
        if self.use_last_as_f:
            x = torch.cat([x, feature_vector], dim=-1)
        return x

    # ---- reference API -----------------------------------------------------------------------------
    def forward(self, input, compute_grad=False):
        if torch.is_grad_enabled() and (input.requires_grad or (self.differentiable_features and self._trainable())):
            return self._forward_autograd(input)
        if self._trainable():
            return self._forward_trainable(input)
        sdf, feat, _ = self.evaluate(input, want_feat=True)
        return torch.cat([sdf.unsqueeze(-1), feat], dim=-1)

    def gradient(self, x, no_grad=False):
        """d sdf / d x, [N,1,3] (reference :110-123).  no_grad=True: the fused inference chain (closed-form reverse sweep on the
        layer GEMM).  no_grad=False: `create_graph=True` -- the result carries a graph to the parameters (eikonal term) and to
        x, through the twice-differentiable composition above."""
        if no_grad:
            if self._trainable():
                with torch.no_grad():
                    _, _, g = self._sync(x.device).eval(x.detach(), want_grad=True, k_flush=self.EVAL_FLUSH)
                return g.unsqueeze(1)
            _, _, g = self.evaluate(x, want_grad=True)
            return g.unsqueeze(1)
        x.requires_grad_(True)
        with torch.enable_grad():
            y = self._forward_autograd(x)[:, :1]
            d_output = torch.ones_like(y, requires_grad=False)
            gradients = torch.autograd.grad(outputs=y, inputs=x, grad_outputs=d_output, create_graph=True, retain_graph=True,
                                            only_inputs=True)[0]
        return gradients.unsqueeze(1)


class RenderingNetwork(nn.Module):
    def __init__(
            self,
            feature_vector_size,
            mode,
            d_in,
            d_out,
            dims,
            weight_norm=True,
            weight_init=False,
            multires_view=0,
            multires_xyz=0,
            normalize_output=True,
            clip_output=False,
            clip_method="relu",
    ):
        super().__init__()
        if mode != 'idr' or d_out > 4:
            raise NotImplementedError("nefii_b200: RenderingNetwork supports mode 'idr'")
        self.normalize_output = normalize_output
        self.clip_output = clip_output
        self.clip_method = clip_method
        self.feature_vector_size = feature_vector_size
        self.mode = mode
        dims = [d_in + feature_vector_size] + list(dims) + [d_out]
        self.multires_view, self.multires_xyz = multires_view, multires_xyz
        self.embedview_fn = None
        if multires_view > 0:
            self.embedview_fn, input_ch = get_embedder(multires_view)
            dims[0] += (input_ch - 3)
        self.embedxyz_fn = None
        if multires_xyz > 0:
            self.embedxyz_fn, input_ch = get_embedder(multires_xyz)
            dims[0] += (input_ch - 3)
        self.num_layers = len(dims)
        for l in range(0, self.num_layers - 1):
            lin = nn.Linear(dims[l], dims[l + 1])
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, "lin" + str(l), lin)
        if weight_init:
            for l in range(0, self.num_layers - 2):
                lin = getattr(self, "lin" + str(l))
                nn.init.kaiming_uniform_(lin.weight, mode='fan_in', nonlinearity='relu')
                nn.init.constant_(lin.bias, 0.0)
            lin = getattr(self, "lin" + str(self.num_layers - 2))
            nn.init.constant_(lin.bias, 0.0)
            if self.normalize_output:
                nn.init.xavier_uniform_(lin.weight, gain=nn.init.calculate_gain('tanh'))
            elif self.clip_method == "relu":
                nn.init.kaiming_uniform_(lin.weight, mode='fan_in', nonlinearity='relu')
        self.relu = nn.ReLU()
        self.tanh = nn.Tanh()

    def forward(self, points, normals, view_dirs, feature_vectors=None):
        segments = [(points, self.multires_xyz if self.multires_xyz > 0 else -1),
                    (view_dirs, self.multires_view if self.multires_view > 0 else -1),
                    (normals, -1)]
        if feature_vectors is not None:
            segments.append((feature_vectors, -1))
        layers = [getattr(self, "lin" + str(l)) for l in range(self.num_layers - 1)]
        x = mlp.dense_mlp(segments, [_effective_weight(l) for l in layers], [l.bias for l in layers], ops.ACT_RELU)
        if self.normalize_output:
            return (self.tanh(x) + 1.) / 2.
        elif not self.clip_output:
            return x
        elif self.clip_method == "relu":
            return self.relu(x)
        elif self.clip_method == "abs":
            return torch.abs(x)
        elif self.clip_method == "relu_init":
            return self.relu(x) + 0.5
        elif self.clip_method == "pow2":
            return x ** 2
        raise NotImplementedError(self.clip_method)


class IDRNetwork(nn.Module):
    def __init__(self, conf):
        super().__init__()
        self.feature_vector_size = conf.get_int('feature_vector_size')
        self.correct_normal = conf.get_bool('correct_normal', default=False)
        self.implicit_network = ImplicitNetwork(self.feature_vector_size, **conf.get_config('implicit_network'))
        self.rendering_network = RenderingNetwork(self.feature_vector_size, **conf.get_config('rendering_network'))
        self.envmap_material_network = EnvmapMaterialNetwork(correct_normal=self.correct_normal,
                                                             feature_vector_size=self.feature_vector_size,
                                                             **conf.get_config('envmap_material_network'))
        self.ray_tracer = RayTracing(**conf.get_config('ray_tracer'))
        self.sample_network = SampleNetwork()
        self.object_bounding_sphere = conf.get_float('ray_tracer.object_bounding_sphere')
        self.render_type = conf.get_string('render_type', default='sg')
        self.rgb_render = self.get_rgb_render(self.render_type)
        self.fast_multi_ray = conf.get_bool('fast_multi_ray', default=False)
        self.render_background = conf.get_bool('render_background', default=False)
        if self.fast_multi_ray:
            raise NotImplementedError("nefii_b200: fast_multi_ray is not used by the shipped confs")
        self.state_freeze_geo = False
        self.state_freeze_idr = False
        self.state_freeze_env_mat = False

    # ---- freezing API (idr_train.py:621-638) -------------------------------------------------------
    def freeze_geometry(self):
        for param in self.implicit_network.parameters():
            param.requires_grad = False
        self.state_freeze_geo = True

    def unfreeze_geometry(self):
        for param in self.implicit_network.parameters():
            param.requires_grad = True
        self.state_freeze_geo = False

    def freeze_idr(self):
        self.freeze_geometry()
        for param in self.rendering_network.parameters():
            param.requires_grad = False
        self.state_freeze_idr = True

    def unfreeze_idr(self):
        self.unfreeze_geometry()
        for param in self.rendering_network.parameters():
            param.requires_grad = True
        self.state_freeze_idr = False

    def freeze_decompose_render(self):
        for param in self.envmap_material_network.parameters():
            param.requires_grad = False
        self.state_freeze_env_mat = True

    def unfreeze_decompose_render(self):
        for param in self.envmap_material_network.parameters():
            param.requires_grad = True
        self.state_freeze_env_mat = False

    def train(self, mode: bool = True):
        nn.Module.train(self, mode)
        if self.state_freeze_idr:
            self.rendering_network.eval()
        if self.state_freeze_geo:
            self.implicit_network.eval()
        if self.state_freeze_env_mat:
            self.envmap_material_network.eval()
        return self

    def forward(self, input, with_point=False):
        if not with_point:
            return self.forward_with_uv(input)
        return self.forward_with_point(input)

    def __getstate__(self):       # the side stream / pending traces of prefetch_trace stay with this object
        state = dict(self.__dict__)
        state.pop("_prefetch_state", None)
        return state

    # ---- implicit_differentiable_renderer.py:312-501 ---------------------------------------------------
    # ---- the primary trace of a forward (reference :335-352), factored out so that it can run ahead of time ---------------
    def _primary_trace(self, uv, pose, intrinsics, object_mask, trace_uniforms, tracer, want_sdf):
        """uv [B, n, 2], object_mask [B * n] -> ray_dirs [B, n, 3], cam_loc, points [B * n, 3], hit mask, dists, sdf_output | None"""
        ray_dirs, cam_loc = rend_util.get_camera_params(uv, pose, intrinsics)
        batch_size, num_pixels, _ = ray_dirs.shape
        sdf_output = None
        with torch.no_grad():
            points, network_object_mask, dists = tracer(sdf=self.implicit_network, cam_loc=cam_loc, object_mask=object_mask,
                                                        ray_directions=ray_dirs, uniforms=trace_uniforms)
            points = (cam_loc.unsqueeze(1) + dists.reshape(batch_size, num_pixels, 1) * ray_dirs).reshape(-1, 3)
            if want_sdf:
                sdf_all, _, _ = self.implicit_network.evaluate(points)
                sdf_output = sdf_all.unsqueeze(-1)
        return ray_dirs, cam_loc, points, network_object_mask, dists, sdf_output

    # SMs the prefetched trace's bulk evaluations may occupy (the rest stay free for the latency-bound chain of the current
    # forward / backward on the caller's stream); NEFII_PREFETCH_SMS overrides
    PREFETCH_SMS = int(os.environ.get("NEFII_PREFETCH_SMS", "120"))

    @staticmethod
    def _prefetch_key(input):
        return tuple((input[k].data_ptr(), input[k]._version, tuple(input[k].shape)) for k in ("uv", "pose", "intrinsics", "object_mask"))

    def prefetch_trace(self, input):
        """nefii_b200 extension: start the primary trace of a LATER ``forward(input)`` now, on a side stream, next to whatever the
        caller's stream is doing (the shading, backward and optimizer step of the current batch).  With a frozen geometry (step 2,
        rendering) the trace of batch i + 1 does not depend on the parameter update of batch i, so the result is the one that
        forward would compute itself -- same kernels, same inputs -- and ``forward`` picks it up when it is called with the same
        input tensors.  Returns False (and does nothing) when the geometry is trainable.  The reference runs the two back to back
        (implicit_differentiable_renderer.py:343-349); at small per-GPU batches the trace is a chain of latency-bound rounds that
        leaves most of the GPU idle."""
        if self.training and not self.state_freeze_geo and torch.is_grad_enabled():
            return False
        uv, pose, intrinsics = input["uv"], input["pose"], input["intrinsics"]
        if pose.requires_grad or intrinsics.requires_grad or not uv.is_cuda:
            return False
        object_mask = input["object_mask"].reshape(-1)
        if len(uv.shape) == 4:
            B, S, R, D = uv.shape
            uv = uv.reshape(B, S * R, D)
            object_mask = object_mask.reshape(B, S, 1).expand(B, S, R).reshape(-1)
        dev = uv.device
        st = self.__dict__.setdefault("_prefetch_state", {})
        if st.get("device") != dev:
            rt = self.ray_tracer
            tracer = RayTracing(rt.object_bounding_sphere, rt.sdf_threshold, rt.line_search_step, rt.line_step_iters,
                                rt.sphere_tracing_iters, rt.n_steps, rt.n_rootfind_steps)     # own buffers, workspace and graphs
            st.clear()
            st.update(device=dev, stream=torch.cuda.Stream(device=dev), tracer=tracer)
        tracer, side = st["tracer"], st["stream"]
        tracer.train(self.ray_tracer.training)
        tracer.skip_min_sdf = self.ray_tracer.skip_min_sdf
        main = torch.cuda.current_stream(dev)
        side.wait_stream(main)                 # the inputs (and the weights) are ready where the caller's stream stands now
        lib = _lib.raw()
        _lib.check(lib.nefii_gemm_set_grid_cap(self.PREFETCH_SMS))
        try:
            with torch.cuda.stream(side), torch.no_grad():
                res = self._primary_trace(uv, pose, intrinsics, object_mask, None, tracer, True)
                done = torch.cuda.Event()
                done.record(side)
        finally:
            _lib.check(lib.nefii_gemm_set_grid_cap(0))
        # two slots: a pipelined loop starts the trace of batch i + 1 BEFORE it calls forward on batch i
        pending = st.setdefault("pending", {})
        pending[self._prefetch_key(input)] = (done, res, tracer.training, (uv, pose, intrinsics, object_mask))
        while len(pending) > 2:
            pending.pop(next(iter(pending)))
        return True

    def prefetch_join(self):
        """Make the current stream wait for a prefetched trace that nobody picked up (end of a timed loop)."""
        st = self.__dict__.get("_prefetch_state")
        for done, _res, _mode, _keep in (st or {}).get("pending", {}).values():
            torch.cuda.current_stream(st["device"]).wait_event(done)

    def _take_prefetched(self, input):
        st = self.__dict__.get("_prefetch_state")
        if not st or not st.get("pending"):
            return None
        got = st["pending"].pop(self._prefetch_key(input), None)
        if got is None:
            return None
        done, res, mode, _keep = got
        if mode != self.ray_tracer.training:
            return None
        main = torch.cuda.current_stream(st["device"])
        main.wait_event(done)
        for t in res:
            if t is not None:
                t.record_stream(main)          # allocated on the side stream, consumed (and freed) on this one
        return res

    def forward_with_uv(self, input, uniforms=None, trace_uniforms=None, eikonal_points=None):
        """uniforms / trace_uniforms / eikonal_points: optional injected random numbers (parity tests); by default they are
        drawn like the reference draws them."""
        unfrozen = self.training and not self.state_freeze_geo and torch.is_grad_enabled()
        intrinsics = input["intrinsics"]
        uv = input["uv"]
        pose = input["pose"]
        object_mask = input["object_mask"].reshape(-1)
        multi_ray_per_pix = len(uv.shape) == 4
        if multi_ray_per_pix:
            B, S, R, D = uv.shape
            uv = uv.reshape(B, S * R, D)
            object_mask = object_mask.reshape(B, S, 1).expand(B, S, R).reshape(-1)
        if pose.requires_grad or intrinsics.requires_grad:
            raise _lib.NefiiError("nefii_b200: camera parameters that require grad (--train_cameras) are outside the "
                                  "accelerated path: surface points are computed without a graph")
        pre = self._take_prefetched(input) if (trace_uniforms is None and not unfrozen) else None
        if pre is not None:
            ray_dirs, cam_loc, points, network_object_mask, dists, sdf_output = pre
        else:
            ray_dirs, cam_loc, points, network_object_mask, dists, sdf_output = self._primary_trace(
                uv, pose, intrinsics, object_mask, trace_uniforms, self.ray_tracer, not unfrozen)
        batch_size, num_pixels, _ = ray_dirs.shape
        if unfrozen:      # the mask loss reaches the geometry through sdf_output (:354): fused trainable stack
            sdf_output = self.implicit_network._forward_trainable(points)[:, 0:1]
        ray_dirs = ray_dirs.reshape(-1, 3)
        surface_mask = (network_object_mask & object_mask) if unfrozen else network_object_mask
        n_rays = points.shape[0]
        # host sync 1 of 2 per forward: the number of surface hits sizes everything downstream.  All gathers / scatters below
        # use the index list (index_select / index_copy): no further boolean-mask round trips.
        surface_idx = torch.nonzero(surface_mask).squeeze(1)
        n_hit = surface_idx.shape[0]
        differentiable_surface_points = points.index_select(0, surface_idx)
        grad_theta = None
        if unfrozen:
            # reference :357-389 -- eikonal samples, d sdf/dx with a graph (second-order path), differentiable intersection
            bound = self.object_bounding_sphere
            n_eik = batch_size * num_pixels // 2
            if eikonal_points is None:      # same draw as the reference (CPU generator, :369)
                eikonal_points = torch.empty(n_eik, 3).uniform_(-bound, bound)
            eik = torch.cat([eikonal_points.to(points.device, torch.float32), points.detach()], 0)
            points_all = torch.cat([differentiable_surface_points, eik], dim=0)
            g = self.implicit_network.gradient(points_all, False)
            grad_theta = g[n_hit:, 0, :]
            if n_hit > 0:
                surface_output = sdf_output.index_select(0, surface_idx)
                cam_all = cam_loc.unsqueeze(1).expand(batch_size, num_pixels, 3).reshape(-1, 3)
                differentiable_surface_points = self.sample_network(
                    surface_output, surface_output.detach(), g[:n_hit, 0, :].detach(), dists.index_select(0, surface_idx).unsqueeze(-1),
                    cam_all.index_select(0, surface_idx), ray_dirs.index_select(0, surface_idx))

        ones = torch.ones_like(points)
        idr_rgb_values, sg_rgb_values, normal_values = ones, ones, ones
        sg_diffuse_rgb_values, sg_diffuse_albedo_values = ones, ones
        sg_specular_rgb_values = torch.zeros_like(points)
        sg_roughness_values = torch.zeros_like(points[..., 0:1])
        sg_specular_reflection_values = sg_specular_rgb_values
        ret = {}
        if n_hit > 0:
            view_dirs = -ray_dirs.index_select(0, surface_idx)
            self.implicit_network.differentiable_features = unfrozen
            try:
                ret = self.get_rbg_value(differentiable_surface_points, view_dirs, uniforms=uniforms)
            finally:
                self.implicit_network.differentiable_features = False
            idr_rgb_values = ones.index_copy(0, surface_idx, ret['idr_rgb'])
            sg_rgb_values = ones.index_copy(0, surface_idx, ret['sg_rgb'])
            normal_values = ones.index_copy(0, surface_idx, ret['normals'])
            sg_diffuse_rgb_values = ones.index_copy(0, surface_idx, ret['sg_diffuse_rgb'])
            sg_diffuse_albedo_values = ones.index_copy(0, surface_idx, ret['sg_diffuse_albedo'])
            sg_specular_rgb_values = sg_specular_rgb_values.index_copy(0, surface_idx, ret['sg_specular_rgb'])
            sg_roughness_values = sg_roughness_values.index_copy(0, surface_idx, ret['sg_roughness'])
            spec = ret['sg_specular_reflectance']
            sg_specular_reflection_values = torch.zeros_like(points).index_copy(0, surface_idx, spec.expand(n_hit, 3))
        else:   # distinct tensors per output key, as the reference returns
            idr_rgb_values, sg_rgb_values, normal_values = ones.clone(), ones.clone(), ones.clone()
            sg_diffuse_rgb_values, sg_diffuse_albedo_values = ones.clone(), ones.clone()
            sg_specular_reflection_values = torch.zeros_like(points)

        if self.render_background and n_hit < n_rays:
            background_idx = torch.nonzero_static(~surface_mask, size=n_rays - n_hit).squeeze(1)     # size known: no sync
            background_rgb = self.get_background_rgb(ray_dirs.index_select(0, background_idx))
            sg_rgb_values = sg_rgb_values.index_copy(0, background_idx, background_rgb)

        output = {
            'points': points,
            'idr_rgb_values': idr_rgb_values,
            'sg_rgb_values': sg_rgb_values,
            'normal_values': normal_values,
            'sdf_output': sdf_output,
            'network_object_mask': network_object_mask,
            'object_mask': object_mask,
            'grad_theta': grad_theta,
            'sg_diffuse_rgb_values': sg_diffuse_rgb_values,
            'sg_diffuse_albedo_values': sg_diffuse_albedo_values,
            'sg_specular_rgb_values': sg_specular_rgb_values,
            'sg_roughness_values': sg_roughness_values,
            'sg_specular_reflection_values': sg_specular_reflection_values,
            'secondary_points': ret.get('secondary_points', None),
            'secondary_mask': ret.get('secondary_mask', None),
            'secondary_dir': ret.get('secondary_dir', None),
        }
        if multi_ray_per_pix:
            for key in ['idr_rgb_values', 'sg_rgb_values', 'network_object_mask', 'object_mask', 'sg_diffuse_rgb_values',
                        'sg_diffuse_albedo_values', 'sg_specular_rgb_values', 'sdf_output', 'points', 'sg_roughness_values',
                        'sg_specular_reflection_values']:
                output[key] = self.mean_pixel(output[key], B * S, R)
            output['normal_values'] = self.mean_pixel(output['normal_values'], B * S, R, vector=True)
        return output

    # ---- implicit_differentiable_renderer.py:503-527 ---------------------------------------------------
    def forward_with_point(self, input):
        points = input["points"]
        ray_dirs = input["ray_dirs"]
        N, R, _ = points.shape
        points = points.reshape(-1, 3)
        ray_dirs = ray_dirs.reshape(-1, 3)
        view_dirs = -ray_dirs
        state_freeze_geo = self.state_freeze_geo
        self.state_freeze_geo = True
        ret = self.get_rbg_value(points, view_dirs)
        self.state_freeze_geo = state_freeze_geo
        return {
            'idr_rgb_values': self.mean_pixel(ret['idr_rgb'], N, R),
            'sg_rgb_values': self.mean_pixel(ret['sg_rgb'], N, R),
        }

    # ---- implicit_differentiable_renderer.py:529-599 ---------------------------------------------------
    def get_rbg_value(self, points, view_dirs, multi_ray_data_shape=None, uniforms=None):
        if self.implicit_network.differentiable_features and torch.is_grad_enabled():
            # trainable geometry (:533-541): features and normals carry a graph to the SDF parameters and to the points
            feature_vectors = self.implicit_network(points)[:, 1:]
            g = self.implicit_network.gradient(points, False)[:, 0, :]
            normals = g / (torch.norm(g, dim=-1, keepdim=True) + 1e-6)
            view_dirs = view_dirs / (torch.norm(view_dirs, dim=-1, keepdim=True) + 1e-6)
        else:
            with torch.no_grad():
                _, feature_vectors, g = self.implicit_network.evaluate(points, want_feat=True, want_grad=True)
                normals = g / (torch.norm(g, dim=-1, keepdim=True) + 1e-6)
                view_dirs = view_dirs / (torch.norm(view_dirs, dim=-1, keepdim=True) + 1e-6)
        ret = {'normals': normals}
        idr_rgb = self.rendering_network(points, normals, view_dirs, feature_vectors)
        sg_envmap_material = self.envmap_material_network(points, feature_vectors, normals)
        ret['idr_rgb'] = idr_rgb
        if self.render_type in ("pt_render_indirect_mlp",):
            sg_ret = self.rgb_render(lgtSGs=sg_envmap_material['sg_lgtSGs'],
                                     specular_reflectance=sg_envmap_material['sg_specular_reflectance'],
                                     roughness=sg_envmap_material['sg_roughness'],
                                     diffuse_albedo=sg_envmap_material['sg_diffuse_albedo'],
                                     normal=normals, viewdirs=view_dirs,
                                     blending_weights=sg_envmap_material['sg_blending_weights'],
                                     points=points, model=self, uniforms=uniforms)
        else:
            sg_ret = self.rgb_render(lgtSGs=sg_envmap_material['sg_lgtSGs'],
                                     specular_reflectance=sg_envmap_material['sg_specular_reflectance'],
                                     roughness=sg_envmap_material['sg_roughness'],
                                     diffuse_albedo=sg_envmap_material['sg_diffuse_albedo'],
                                     normal=normals, viewdirs=view_dirs,
                                     blending_weights=sg_envmap_material['sg_blending_weights'])
        ret.update(sg_ret)
        ret.update({
            'sg_roughness': sg_envmap_material['sg_roughness'],
            'sg_specular_reflectance': sg_envmap_material['sg_specular_reflectance'],
            'sg_blending_weights': sg_envmap_material['sg_blending_weights']
        })
        return ret

    # ---- implicit_differentiable_renderer.py:646-663 ---------------------------------------------------
    def get_background_rgb(self, light_dir):
        return integrator.background_sg(self.envmap_material_network.get_lgtSGs(), light_dir)

    # ---- implicit_differentiable_renderer.py:695-719 ---------------------------------------------------
    def mean_pixel(self, x, bs, r, vector=False):
        assert x.shape[0] == bs * r
        no_dim = len(x.shape) == 1
        if no_dim:
            x = x[..., None]
        bsr, d = x.shape
        x = x.reshape(bs, r, d)
        if vector:
            x = x[:, 0, :]
        elif x.dtype == torch.float:
            x = x.mean(1)
        elif x.dtype == torch.bool:
            x = x.all(1)
        else:
            raise TypeError("mean_pixel: undefined type %s" % x.dtype)
        if no_dim:
            x = x[..., 0]
        return x

    # ---- implicit_differentiable_renderer.py:721-759 ---------------------------------------------------
    def get_rgb_render(self, render_type: str):
        if render_type == "sg":
            return render_with_sg
        if render_type == "pt_render_indirect_mlp":
            return pt_render_indirect_mlp
        raise NotImplementedError("nefii_b200: render_type '%s' is one of the reference's non-default integrator variants "
                                  "(SURVEY.md section 2: out of scope); supported: 'pt_render_indirect_mlp', 'sg'" % render_type)
