"""Oracle (test infrastructure): the three MLPs of the hot path, restated on plain weight tensors.

Restates
  * positional encoding           -- reference code/model/embedder.py:5-50
  * ImplicitNetwork.forward       -- code/model/implicit_differentiable_renderer.py:85-108
  * ImplicitNetwork.gradient      -- :110-123 (autograd there; closed-form reverse sweep here, which the
                                     tests check against autograd through `sdf_forward`)
  * RenderingNetwork.forward      -- :196-241 (mode 'idr', clip_method 'pow2')
  * EnvmapMaterialNetwork.forward -- code/model/sg_envmap_material.py:357-425 (same_mlp, roughness_mlp,
                                     fix_specular_albedo)
and the parameter initialisers those classes use (so synthetic weights look like the reference's).

Parity status: PINNED by tests/test_oracle_hotpath.py, which loads these weights into the real reference
modules (when /root/reference is present) and by tests/golden/mlp_*.npz generated from them.

Weights are kept as *effective* matrices: weight_norm layers are folded (w = g * v / ||v||_row,
torch.nn.utils.weight_norm with dim=0), which is what the CUDA path packs as well.
"""
import math

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2)


def embed(x, n_freqs):
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]  (embedder.py:22-36)."""
    out = [x]
    for k in range(n_freqs):
        f = 2.0 ** k
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


def fold_weight_norm(weight_g, weight_v):
    """torch.nn.utils.weight_norm(dim=0): w = v * (g / ||v||) with the norm over each output row."""
    norm = weight_v.reshape(weight_v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (weight_v.dim() - 1)))
    return weight_v * (weight_g / norm)


# ------------------------------------------------------------------------------------------------
# SDF / feature network
# ------------------------------------------------------------------------------------------------
class SdfParams:
    """Effective weights of ImplicitNetwork: W[l] [out_l, in_l], b[l] [out_l], l = 0..n_hidden."""

    def __init__(self, weights, biases, n_freqs=6, skip_layer=4, last_as_f=True):
        self.W = list(weights)
        self.b = list(biases)
        self.n_freqs = n_freqs
        self.skip_layer = skip_layer
        # True: feature vector = input of the last layer (conf.conf); False: the last Linear has 1 + F outputs and the
        # feature vector is its rows 1.. (conf_neus.conf; implicit_differentiable_renderer.py:39-42,105-106)
        self.last_as_f = last_as_f

    @property
    def n_layers(self):
        return len(self.W)

    def to(self, *a, **k):
        return SdfParams([w.to(*a, **k) for w in self.W], [b.to(*a, **k) for b in self.b], self.n_freqs, self.skip_layer,
                         self.last_as_f)

    def state_dict(self, prefix=""):
        """weight_g / weight_v / bias entries as the reference's checkpoints name them."""
        sd = {}
        for l, (w, b) in enumerate(zip(self.W, self.b)):
            sd["%slin%d.weight_v" % (prefix, l)] = w.clone()
            sd["%slin%d.weight_g" % (prefix, l)] = w.norm(dim=1, keepdim=True)
            sd["%slin%d.bias" % (prefix, l)] = b.clone()
        return sd


def sdf_init(seed=0, width=512, n_hidden=8, n_freqs=6, skip_layer=4, bias=0.6, bumps=0.0, d_feat=0):
    """Geometric initialisation as in ImplicitNetwork.__init__ (implicit_differentiable_renderer.py:60-74).
    `bumps` > 0 additionally gives the PE columns of layer 0 small random weights, which turns the
    sphere of radius `bias` into a bumpy blob (used by the synthetic 'robot-scale' scene)."""
    g = torch.Generator().manual_seed(seed)
    d_pe = 3 + 6 * n_freqs
    dims = [d_pe] + [width] * n_hidden + [1 + d_feat]      # d_feat > 0: use_last_as_f = False layout
    W, B = [], []
    n_lin = len(dims) - 1
    for l in range(n_lin):
        out_dim = dims[l + 1] - d_pe if (l + 1) == skip_layer else dims[l + 1]
        in_dim = dims[l]
        w = torch.empty(out_dim, in_dim)
        b = torch.zeros(out_dim)
        if l == n_lin - 1:
            w.normal_(math.sqrt(math.pi) / math.sqrt(in_dim), 0.0001, generator=g)
            b.fill_(-bias)
        elif l == 0:
            w.zero_()
            w[:, :3].normal_(0.0, math.sqrt(2) / math.sqrt(out_dim), generator=g)
            if bumps > 0:
                w[:, 3:].normal_(0.0, bumps * math.sqrt(2) / math.sqrt(out_dim), generator=g)
        elif l == skip_layer:
            w.normal_(0.0, math.sqrt(2) / math.sqrt(out_dim), generator=g)
            w[:, -(d_pe - 3):] = 0.0
        else:
            w.normal_(0.0, math.sqrt(2) / math.sqrt(out_dim), generator=g)
        W.append(w)
        B.append(b)
    return SdfParams(W, B, n_freqs, skip_layer, last_as_f=(d_feat == 0))


def sdf_forward(p, x, return_hidden=False):
    """-> [N, 1 + width]: column 0 the SDF, the rest the feature vector (input of the last layer)."""
    pe = embed(x, p.n_freqs)
    h = pe
    hidden = []
    n_lin = p.n_layers
    for l in range(n_lin):
        if l == n_lin - 1:
            feat = h
        if l == p.skip_layer:
            h = torch.cat([h, pe], 1) / SQRT2
        hidden.append(h)
        h = F.linear(h, p.W[l], p.b[l])
        if l < n_lin - 1:
            h = F.softplus(h, beta=100)
    out = torch.cat([h, feat], dim=-1) if p.last_as_f else h
    return (out, hidden) if return_hidden else out


def sdf_gradient(p, x):
    """d sdf / d x in closed form (reverse sweep through Softplus(beta=100), the skip concat and the PE).
    Equals ImplicitNetwork.gradient(x, no_grad=True)[:, 0, :]."""
    pe = embed(x, p.n_freqs)
    n_lin = p.n_layers
    h = pe
    pre = []
    for l in range(n_lin - 1):
        if l == p.skip_layer:
            h = torch.cat([h, pe], 1) / SQRT2
        z = F.linear(h, p.W[l], p.b[l])
        pre.append(z)
        h = F.softplus(z, beta=100)
    g = p.W[n_lin - 1][0:1].expand(x.shape[0], -1)
    g_pe = torch.zeros_like(pe)
    for l in range(n_lin - 2, -1, -1):
        g = g * torch.sigmoid(100 * pre[l])
        g = g @ p.W[l]
        if l == p.skip_layer:
            g = g / SQRT2
            n_h = g.shape[1] - pe.shape[1]
            g_pe = g_pe + g[:, n_h:]
            g = g[:, :n_h]
    g_pe = g_pe + g
    dx = g_pe[:, :3].clone()
    for k in range(p.n_freqs):
        f = 2.0 ** k
        gs = g_pe[:, 3 + 6 * k: 6 + 6 * k]
        gc = g_pe[:, 6 + 6 * k: 9 + 6 * k]
        dx = dx + f * (torch.cos(x * f) * gs - torch.sin(x * f) * gc)
    return dx


# ------------------------------------------------------------------------------------------------
# generic dense stack (radiance + material nets)
# ------------------------------------------------------------------------------------------------
class DenseParams:
    def __init__(self, weights, biases):
        self.W = list(weights)
        self.b = list(biases)

    def to(self, *a, **k):
        return DenseParams([w.to(*a, **k) for w in self.W], [b.to(*a, **k) for b in self.b])

    def requires_grad_(self, flag=True):
        for t in self.W + self.b:
            t.requires_grad_(flag)
        return self

    def tensors(self):
        return self.W + self.b


def _kaiming_uniform(out_dim, in_dim, g, a=0.0):
    gain = math.sqrt(2.0 / (1 + a * a))
    bound = gain * math.sqrt(3.0 / in_dim)
    return (torch.rand(out_dim, in_dim, generator=g) * 2 - 1) * bound


def radiance_init(seed=0, width=512, n_hidden=4, feature=512, xyz_freqs=10, view_freqs=4):
    """RenderingNetwork with weight_init=True (implicit_differentiable_renderer.py:179-191)."""
    g = torch.Generator().manual_seed(seed)
    d_in = (3 + 6 * xyz_freqs) + (3 + 6 * view_freqs) + 3 + feature
    dims = [d_in] + [width] * n_hidden + [3]
    W, B = [], []
    for l in range(len(dims) - 1):
        W.append(_kaiming_uniform(dims[l + 1], dims[l], g))
        B.append(torch.zeros(dims[l + 1]))
    return DenseParams(W, B)


def radiance_forward(p, points, normals, view_dirs, features, xyz_freqs=10, view_freqs=4):
    """mode 'idr': cat[PE(x), PE(view), normal, feature] -> ReLU stack -> x**2 (clip_method pow2)."""
    h = torch.cat([embed(points, xyz_freqs), embed(view_dirs, view_freqs), normals, features], dim=-1)
    n_lin = len(p.W)
    for l in range(n_lin):
        h = F.linear(h, p.W[l], p.b[l])
        if l < n_lin - 1:
            h = torch.relu(h)
    return h ** 2


def material_init(seed=0, width=512, n_hidden=8, feature=512, xyz_freqs=10, d_out=4):
    """EnvmapMaterialNetwork.diffuse_albedo_layers: nn.Linear default init (kaiming_uniform a=sqrt(5))."""
    g = torch.Generator().manual_seed(seed)
    d_in = (3 + 6 * xyz_freqs) + feature
    dims = [d_in] + [width] * n_hidden + [d_out]
    W, B = [], []
    for l in range(len(dims) - 1):
        W.append(_kaiming_uniform(dims[l + 1], dims[l], g, a=math.sqrt(5)))
        bound = 1 / math.sqrt(dims[l])
        B.append((torch.rand(dims[l + 1], generator=g) * 2 - 1) * bound)
    return DenseParams(W, B)


TINY_ROUGHNESS = 0.089


def material_forward(p, points, features, xyz_freqs=10):
    """-> (diffuse_albedo [N,3], roughness [N,1]) ; sg_envmap_material.py:357-405 with same_mlp."""
    h = torch.cat([embed(points, xyz_freqs), features], dim=-1)
    n_lin = len(p.W)
    for l in range(n_lin):
        h = F.linear(h, p.W[l], p.b[l])
        if l < n_lin - 1:
            h = F.elu(h)
    albedo = torch.sigmoid(h[..., :3])
    rough = torch.sigmoid(h[..., 3:4])
    rough = (1 - TINY_ROUGHNESS) * rough + TINY_ROUGHNESS
    return albedo, rough


def specular_remap(s):
    """sg_envmap_material.py:440-443."""
    return 0.16 * s ** 2
