"""Import shim for the *real* reference (FuxiComputerVision/Nefii) -- TEST INFRASTRUCTURE ONLY.

The reference lives at /root/reference (read-only) and exists only in the build
container, never on the GPU box.  This module is used by
  * tests/ (``-m "not gpu"``) to validate the oracle restatement in oracle/nefii_oracle.py
  * oracle/make_golden.py to produce the committed fixtures under tests/golden/
It is never imported by the product package (nefii_b200/).

Three shims are needed:
  1. stub ``imageio`` / ``skimage`` modules (utils/rend_util.py imports them and calls
     imageio.plugins.freeimage.download() at import time);
  2. ``torch.Tensor.cuda`` -> identity when no GPU is present (the hot path hard-codes .cuda());
  3. a dict-backed stand-in for the pyhocon config object (``ConfStub``).
"""
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("NEFII_REFERENCE_ROOT", "/root/reference")
REF_CODE = os.path.join(REF_ROOT, "code")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_CODE, "model"))


class ConfStub(dict):
    """Minimal pyhocon.ConfigTree look-alike (get_int/get_float/get_bool/get_string/get_config)."""

    def _get(self, key, default=None):
        node = self
        for part in key.split("."):
            if not isinstance(node, dict) or part not in node:
                if default is None:
                    raise KeyError(key)
                return default
            node = node[part]
        return node

    def get_int(self, key, default=None):
        return int(self._get(key, default))

    def get_float(self, key, default=None):
        return float(self._get(key, default))

    def get_bool(self, key, default=None):
        v = self._get(key, default if default is not None else False)
        return bool(v)

    def get_string(self, key, default=None):
        return str(self._get(key, default))

    def get_config(self, key, default=None):
        v = self._get(key, default)
        return ConfStub(v)

    def get_list(self, key, default=None):
        return list(self._get(key, default))


_installed = False


def install():
    """Make ``import model.*`` / ``import utils.*`` resolve to the reference's packages."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REF_ROOT)
    # kornia: imported by model/loss.py for the SSIM term only (weight 0 in the step-2 recipe, never called here)
    for name in ("imageio", "imageio.plugins", "imageio.plugins.freeimage", "skimage", "kornia"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["imageio"].plugins = sys.modules["imageio.plugins"]
    sys.modules["imageio.plugins"].freeimage = sys.modules["imageio.plugins.freeimage"]
    sys.modules["imageio.plugins.freeimage"].download = lambda *a, **k: None
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REF_CODE not in sys.path:
        sys.path.insert(0, REF_CODE)
    _installed = True


def model_conf(render_type="pt_render_indirect_mlp", num_lgt_sgs=128, width=512):
    """The ``model{}`` block of code/confs_sg/conf.conf:36-95 as a ConfStub."""
    dims8 = [width] * 8
    return ConfStub(
        render_type=render_type,
        feature_vector_size=width,
        fast_multi_ray=False,
        render_background=True,
        implicit_network=dict(d_in=3, d_out=1, dims=dims8, geometric_init=True, bias=0.6,
                              skip_in=[4], weight_norm=True, multires=6, use_last_as_f=True),
        envmap_material_network=dict(multires=10, dims=dims8, white_specular=True, white_light=False,
                                     num_lgt_sgs=num_lgt_sgs, num_base_materials=1, upper_hemi=False,
                                     fix_specular_albedo=True, specular_albedo=[0.5, 0.5, 0.5],
                                     init_specular_reflectance=0.1, roughness_mlp=True, specular_mlp=True,
                                     dims_roughness=[width] * 4, dims_specular=[width] * 4, same_mlp=True),
        rendering_network=dict(mode="idr", d_in=9, d_out=3, dims=[width] * 4, weight_norm=True,
                               multires_view=4, multires_xyz=10, normalize_output=False,
                               clip_output=True, clip_method="pow2", weight_init=True),
        ray_tracer=dict(object_bounding_sphere=1.0, sdf_threshold=5.0e-5, line_search_step=0.5,
                        line_step_iters=3, sphere_tracing_iters=10, n_steps=100, n_rootfind_steps=32),
    )
