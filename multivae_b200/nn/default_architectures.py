"""Default MLP architectures (reference: models/nn/default_architectures.py:21-258).

Parameter names: `layers.<i>.0.{weight,bias}`, `embedding`, `log_var`, `style_embedding`,
`style_log_var`.  On CUDA with bf16 compute the whole stack runs as ONE autograd Function over the native tcgen05 GEMM
(nn/linear_native.py: bias + ReLU / Sigmoid fused in the epilogue, native weight / data gradients); the fp32 path (and CPU
construction / state-dict handling) uses the library layers of nn/functional.py."""
import numpy as np
import torch
import torch.nn as nn

from ..configs import BaseAEConfig
from ..containers import ModelOutput
from . import functional as NF
from .base_architectures import BaseDecoder, BaseEncoder


def _native(x):
    """Native tensor-core path: CUDA tensor and bf16 compute (autocast active, set by the model's compute_dtype) or forced."""
    if not x.is_cuda or NF.backend() == "torch":
        return False
    return NF.backend() == "native" or torch.is_autocast_enabled()


def _hidden(n_in, n_out):
    return nn.Sequential(nn.Linear(n_in, n_out), nn.ReLU())


class Encoder_VAE_MLP(BaseEncoder):
    def __init__(self, args, n_hidden=1):
        super().__init__()
        self.input_dim = args.input_dim
        self.latent_dim = args.latent_dim
        self.layers = nn.ModuleList([_hidden(int(np.prod(args.input_dim)), 512)] + [_hidden(512, 512) for _ in range(n_hidden)])
        self.depth = len(self.layers)
        self.embedding = nn.Linear(512, self.latent_dim)
        self.log_var = nn.Linear(512, self.latent_dim)

    def forward(self, x, output_layer_levels=None):
        h = x.reshape(-1, int(np.prod(self.input_dim)))
        if _native(h):
            from .linear_native import mlp_chain
            out = mlp_chain(h, [l[0] for l in self.layers] + [[self.embedding, self.log_var]], ["relu"] * self.depth + ["none"], out_fp32=True)
            mu, lv = torch.split(out, [self.latent_dim, self.latent_dim], dim=-1)
            return ModelOutput(embedding=mu, log_covariance=lv)
        for layer in self.layers:
            h = NF.linear(h, layer[0].weight, layer[0].bias, act="relu")
        mu, lv = NF.linear_heads(h, [self.embedding, self.log_var])
        return ModelOutput(embedding=mu, log_covariance=lv)


class Encoder_VAE_MLP_Style(BaseEncoder):
    def __init__(self, args):
        super().__init__()
        self.input_dim = args.input_dim
        self.latent_dim = args.latent_dim
        self.style_dim = args.style_dim
        self.layers = nn.ModuleList([_hidden(int(np.prod(args.input_dim)), 512)])
        self.depth = 1
        self.embedding = nn.Linear(512, self.latent_dim)
        self.log_var = nn.Linear(512, self.latent_dim)
        self.style_embedding = nn.Linear(512, self.style_dim)
        self.style_log_var = nn.Linear(512, self.style_dim)

    def forward(self, x, output_layer_levels=None):
        h = x.reshape(-1, int(np.prod(self.input_dim)))
        if _native(h):
            from .linear_native import mlp_chain
            heads = [self.embedding, self.log_var, self.style_embedding, self.style_log_var]
            out = mlp_chain(h, [self.layers[0][0], heads], ["relu", "none"], out_fp32=True)
            mu, lv, smu, slv = torch.split(out, [hd.weight.shape[0] for hd in heads], dim=-1)
            return ModelOutput(embedding=mu, log_covariance=lv, style_embedding=smu, style_log_covariance=slv)
        h = NF.linear(h, self.layers[0][0].weight, self.layers[0][0].bias, act="relu")
        mu, lv, smu, slv = NF.linear_heads(h, [self.embedding, self.log_var, self.style_embedding, self.style_log_var])
        return ModelOutput(embedding=mu, log_covariance=lv, style_embedding=smu, style_log_covariance=slv)


class Decoder_AE_MLP(BaseDecoder):
    def __init__(self, args):
        super().__init__()
        self.input_dim = tuple(args.input_dim)
        self.layers = nn.ModuleList([
            _hidden(args.latent_dim, 512),
            nn.Sequential(nn.Linear(512, int(np.prod(args.input_dim))), nn.Sigmoid()),
        ])
        self.depth = 2

    def forward(self, z, **kwargs):
        if _native(z):
            from .linear_native import mlp_chain
            h = mlp_chain(z.reshape(-1, z.shape[-1]), [self.layers[0][0], self.layers[1][0]], ["relu", "sigmoid"])
            return ModelOutput(reconstruction=h.reshape(*z.shape[:-1], *self.input_dim))
        h = NF.linear(z.reshape(-1, z.shape[-1]), self.layers[0][0].weight, self.layers[0][0].bias, act="relu")
        h = NF.linear(h, self.layers[1][0].weight, self.layers[1][0].bias, act="sigmoid", out_dtype=kwargs.get("out_dtype"))
        return ModelOutput(reconstruction=h.reshape(*z.shape[:-1], *self.input_dim))


def BaseDictEncoders(input_dims, latent_dim):
    return nn.ModuleDict({m: Encoder_VAE_MLP(BaseAEConfig(input_dim=input_dims[m], latent_dim=latent_dim)) for m in input_dims})


def BaseDictEncoders_MultiLatents(input_dims, latent_dim, modality_dims):
    return nn.ModuleDict({m: Encoder_VAE_MLP_Style(BaseAEConfig(input_dim=input_dims[m], latent_dim=latent_dim, style_dim=modality_dims[m])) for m in input_dims})


def BaseDictDecoders(input_dims, latent_dim):
    return nn.ModuleDict({m: Decoder_AE_MLP(BaseAEConfig(input_dim=input_dims[m], latent_dim=latent_dim)) for m in input_dims})


def BaseDictDecodersMultiLatents(input_dims, latent_dim, modality_dims):
    return nn.ModuleDict({m: Decoder_AE_MLP(BaseAEConfig(input_dim=input_dims[m], latent_dim=latent_dim + modality_dims[m])) for m in input_dims})
