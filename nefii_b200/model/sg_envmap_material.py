"""Drop-in for the reference's code/model/sg_envmap_material.py (EnvmapMaterialNetwork, :46-447).

Same constructor arguments, attributes, methods and ``state_dict`` keys (``lgtSGs``, ``specular_reflectance``,
``diffuse_albedo_layers.{2i}.{weight,bias}``), so reference checkpoints and ``--light_sg_path`` files load.
The albedo / roughness MLP runs on the tcgen05 layer GEMM (nefii_b200.mlp.dense_mlp) with weight gradients.
Supported configuration = what the shipped confs use on the hot path: SG light, one base material,
``same_mlp`` roughness head, fixed specular albedo; other combinations raise NotImplementedError.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import mlp, ops
from .embedder import get_embedder


def fibonacci_sphere(samples=1):
    points = []
    phi = np.pi * (3. - np.sqrt(5.))
    for i in range(samples):
        y = 1 - (i / float(samples - 1)) * 2
        radius = np.sqrt(1 - y * y)
        theta = phi * i
        points.append([np.cos(theta) * radius, y, np.sin(theta) * radius])
    return np.array(points)


def compute_energy(lgtSGs):
    lgtLambda = torch.abs(lgtSGs[:, 3:4])
    lgtMu = torch.abs(lgtSGs[:, 4:])
    return lgtMu * 2.0 * np.pi / lgtLambda * (1.0 - torch.exp(-2.0 * lgtLambda))


class EnvmapMaterialNetwork(nn.Module):
    def __init__(self, multires=0, dims=[256, 256, 256],
                 white_specular=False,
                 white_light=False,
                 num_lgt_sgs=32,
                 num_base_materials=2,
                 upper_hemi=False,
                 fix_specular_albedo=False,
                 specular_albedo=[-1., -1., -1.],
                 init_specular_reflectance=-1,
                 correct_normal=False,
                 roughness_mlp=False,
                 specular_mlp=False,
                 same_mlp=False,
                 dims_roughness=[256, 256, 256],
                 dims_specular=[256, 256, 256],
                 feature_vector_size=0,
                 use_normal=False,
                 light_type='sg'):
        super().__init__()
        if light_type != 'sg' or correct_normal or use_normal or white_light or num_base_materials != 1:
            raise NotImplementedError("nefii_b200: only the SG-light, single-material configuration of the shipped "
                                      "confs (conf.conf) is on the accelerated path")
        if not (roughness_mlp and same_mlp and fix_specular_albedo):
            raise NotImplementedError("nefii_b200: expected roughness_mlp + same_mlp + fix_specular_albedo (conf.conf)")
        self.correct_normal = correct_normal
        self.roughness_mlp = roughness_mlp
        self.specular_mlp = specular_mlp
        self.same_mlp = same_mlp
        self.feature_vector_size = feature_vector_size
        self.fix_specular_albedo = fix_specular_albedo
        self.fake_roughness = False
        self.fake_specular = False
        self.light_type = light_type
        self.use_normal = use_normal
        self.multires = multires

        input_dim = 3
        self.embed_fn = None
        if multires > 0:
            self.embed_fn, input_dim = get_embedder(multires)
        input_dim += feature_vector_size

        self.actv_fn = nn.ELU()
        layers = []
        dim = input_dim
        dim_o = 3 + 1
        for i in range(len(dims)):
            layers.append(nn.Linear(dim, dims[i]))
            layers.append(self.actv_fn)
            dim = dims[i]
        layers.append(nn.Linear(dim, dim_o))
        self.diffuse_albedo_layers = nn.Sequential(*layers)
        self.delta_normal_layers_layers = None

        self.numLgtSGs = num_lgt_sgs
        self.numBrdfSGs = num_base_materials
        self.white_light = white_light
        self.lgtSGs = nn.Parameter(torch.randn(self.numLgtSGs, 7), requires_grad=True)
        self.lgtSGs.data[:, -2:] = self.lgtSGs.data[:, -3:-2].expand((-1, 2))
        self.lgtSGs.data[:, 3:4] = 20. + torch.abs(self.lgtSGs.data[:, 3:4] * 100.)
        energy = compute_energy(self.lgtSGs.data)
        self.lgtSGs.data[:, 4:] = torch.abs(self.lgtSGs.data[:, 4:]) / torch.sum(energy, dim=0, keepdim=True) * 2. * np.pi
        lobes = fibonacci_sphere(self.numLgtSGs).astype(np.float32)
        self.lgtSGs.data[:, :3] = torch.from_numpy(lobes)
        self.upper_hemi = upper_hemi
        if self.upper_hemi:
            self.restrict_lobes_upper = lambda lgtSGs: torch.cat((lgtSGs[..., :1], torch.abs(lgtSGs[..., 1:2]), lgtSGs[..., 2:]), dim=-1)
            self.lgtSGs.data = self.restrict_lobes_upper(self.lgtSGs.data)

        self.white_specular = white_specular
        specular_albedo = np.array(specular_albedo).astype(np.float32)
        assert (np.all(np.logical_and(specular_albedo > 0., specular_albedo < 1.)))
        self.specular_reflectance = nn.Parameter(torch.from_numpy(specular_albedo).reshape((self.numBrdfSGs, 3)),
                                                 requires_grad=False)
        self.blending_weights_layers = []

    # ---- the small management API the runners use (idr_train.py:188-194,543-547,705-713; render.py:397,435-439) ----
    def freeze_light(self):
        self.lgtSGs.requires_grad = False

    def freeze_diffuse(self):
        for param in self.diffuse_albedo_layers.parameters():
            param.requires_grad = False

    def unfreeze_diffuse(self):
        for param in self.diffuse_albedo_layers.parameters():
            param.requires_grad = True

    def unfreeze_all(self):
        for param in self.parameters():
            param.requires_grad = True

    def freeze_all(self):
        for param in self.parameters():
            param.requires_grad = False

    def set_roughness_fake(self, state):
        self.fake_roughness = state

    def set_specular_fake(self, state):
        self.fake_specular = state

    def get_light(self):
        lgtSGs = self.lgtSGs.clone().detach()
        if self.upper_hemi:
            lgtSGs = self.restrict_lobes_upper(lgtSGs)
        return lgtSGs

    def load_light(self, path):
        assert (path.endswith('.npy'))
        device = self.lgtSGs.data.device
        self.lgtSGs = nn.Parameter(torch.from_numpy(np.load(path)).to(device), requires_grad=True)
        self.numLgtSGs = self.lgtSGs.data.shape[0]
        if self.lgtSGs.data.shape[1] == 7 or self.light_type != 'sg':
            self.white_light = False

    def get_base_materials(self):
        return torch.zeros(1, 1), self.specular_reflectance

    def get_lgtSGs(self):
        lgtSGs = self.lgtSGs
        if self.upper_hemi:
            lgtSGs = self.restrict_lobes_upper(lgtSGs)
        return lgtSGs

    @staticmethod
    def specular_remap(specular_reflectacne):
        return 0.16 * specular_reflectacne ** 2

    @staticmethod
    def specular_inv_remap(specular_reflectacne):
        return (specular_reflectacne / 0.16) ** 0.5

    def forward(self, points, feature_vector=None, normal=None):
        linears = [m for m in self.diffuse_albedo_layers if isinstance(m, nn.Linear)]
        segments = [(points, self.multires if self.multires > 0 else -1)]
        if feature_vector is not None:
            segments.append((feature_vector, -1))
        brdf = mlp.dense_mlp(segments, [l.weight for l in linears], [l.bias for l in linears], ops.ACT_ELU)
        diffuse_albedo = torch.sigmoid(brdf[..., :3])
        roughness = torch.sigmoid(brdf[..., 3:4])
        TINNY_ROUGHNESS = 0.089
        roughness = (1 - TINNY_ROUGHNESS) * roughness + TINNY_ROUGHNESS
        if self.fake_roughness:
            roughness = 0 * roughness + 0.5
        specular_reflectacne = self.specular_reflectance
        if self.fake_specular:
            specular_reflectacne = 0 * specular_reflectacne + 0.5
        specular_reflectacne = self.specular_remap(specular_reflectacne)
        return dict([
            ('sg_lgtSGs', self.get_lgtSGs()),
            ('sg_specular_reflectance', specular_reflectacne),
            ('sg_roughness', roughness),
            ('sg_diffuse_albedo', diffuse_albedo),
            ('sg_blending_weights', None)
        ])
