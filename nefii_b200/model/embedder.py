"""Positional encoding (reference code/model/embedder.py).  The CUDA path encodes inside its own kernels
(csrc/sdf_mlp.cu encode_kernel, csrc/dense_mlp.cu assemble_input_kernel); this module only reports sizes and
offers a torch embed for callers that want the encoded tensor itself."""
import torch


class Embedder:
    def __init__(self, multires, input_dims=3):
        self.multires = multires
        self.input_dims = input_dims
        self.out_dim = input_dims + 2 * input_dims * multires

    def embed(self, inputs):
        out = [inputs]
        for k in range(self.multires):
            f = 2.0 ** k
            out.append(torch.sin(inputs * f))
            out.append(torch.cos(inputs * f))
        return torch.cat(out, -1)


def get_embedder(multires):
    e = Embedder(multires)

    def embed(x, eo=e):
        return eo.embed(x)
    embed.multires = multires
    return embed, e.out_dim
