"""A dict-backed object with the pyhocon ConfigTree getters IDRNetwork uses (get_int / get_float / get_bool /
get_string / get_config / dotted keys), for callers that build the model without pyhocon (bench, smoke, tests).
The reference's runners pass a real pyhocon tree, which works the same way."""


class DictConf(dict):
    _MISSING = object()

    def _get(self, key, default=_MISSING):
        node = self
        for part in key.split("."):
            if not isinstance(node, dict) or part not in node:
                if default is DictConf._MISSING:
                    raise KeyError(key)
                return default
            node = node[part]
        return node

    def get_int(self, key, default=_MISSING):
        return int(self._get(key, default))

    def get_float(self, key, default=_MISSING):
        return float(self._get(key, default))

    def get_bool(self, key, default=_MISSING):
        return bool(self._get(key, default))

    def get_string(self, key, default=_MISSING):
        return str(self._get(key, default))

    def get_list(self, key, default=_MISSING):
        return list(self._get(key, default))

    def get_config(self, key, default=_MISSING):
        return DictConf(self._get(key, default))


def default_model_conf(render_type="pt_render_indirect_mlp", num_lgt_sgs=128, width=512):
    """The ``model{}`` block of the reference's code/confs_sg/conf.conf:36-95."""
    dims8 = [width] * 8
    return DictConf(
        render_type=render_type,
        feature_vector_size=width,
        fast_multi_ray=False,
        render_background=True,
        implicit_network=dict(d_in=3, d_out=1, dims=dims8, geometric_init=True, bias=0.6, skip_in=[4],
                              weight_norm=True, multires=6, use_last_as_f=True),
        envmap_material_network=dict(multires=10, dims=dims8, white_specular=True, white_light=False,
                                     num_lgt_sgs=num_lgt_sgs, num_base_materials=1, upper_hemi=False,
                                     fix_specular_albedo=True, specular_albedo=[0.5, 0.5, 0.5],
                                     init_specular_reflectance=0.1, roughness_mlp=True, specular_mlp=True,
                                     dims_roughness=[width] * 4, dims_specular=[width] * 4, same_mlp=True),
        rendering_network=dict(mode="idr", d_in=9, d_out=3, dims=[width] * 4, weight_norm=True, multires_view=4,
                               multires_xyz=10, normalize_output=False, clip_output=True, clip_method="pow2",
                               weight_init=True),
        ray_tracer=dict(object_bounding_sphere=1.0, sdf_threshold=5.0e-5, line_search_step=0.5, line_step_iters=3,
                        sphere_tracing_iters=10, n_steps=100, n_rootfind_steps=32),
    )
