"""64 x 64 image ResNets of the CUB / CelebA-sized configurations (reference: models/nn/cub.py:144-293; the transformer text
encoder of that file is out of scope).

Parameter names follow the reference exactly (`conv_img`, `resnet.{i}.conv_0`, `conv_1`, `conv_s`, `fc_mu`, `fc_logvar`, `fc`),
so reference checkpoints load unchanged.  Differences from the PolyMNIST ResNets (nn/mmnist.py): pre-activation blocks
`x_s + 0.1 * conv_1(actvn(conv_0(actvn(x))))`, the fully connected heads read `actvn(out)`, the image head has no final
activation, and the stacks work at 64 -> 32 -> 16 pixels.  On CUDA with bf16 compute the stacks run as tcgen05 implicit-GEMM
kernels on the shared-halo layout (multivae_b200/nn/cub_native.py)."""
import math

import torch.nn as nn
import torch.nn.functional as F

from ..containers import ModelOutput
from . import functional as NF
from .base_architectures import BaseDecoder, BaseEncoder


def actvn(x):
    return F.leaky_relu(x, 2e-1)


class CUB_ResnetBlock(nn.Module):
    """Pre-activation residual block (cub.py:249-293)."""

    def __init__(self, fin, fout, fhidden=None, is_bias=True):
        super().__init__()
        self.is_bias = is_bias
        self.learned_shortcut = fin != fout
        self.fin, self.fout = fin, fout
        self.fhidden = min(fin, fout) if fhidden is None else fhidden
        self.conv_0 = nn.Conv2d(self.fin, self.fhidden, 3, stride=1, padding=1)
        self.conv_1 = nn.Conv2d(self.fhidden, self.fout, 3, stride=1, padding=1, bias=is_bias)
        if self.learned_shortcut:
            self.conv_s = nn.Conv2d(self.fin, self.fout, 1, stride=1, padding=0, bias=False)

    def forward(self, x):
        x_s = NF.conv2d(x, self.conv_s.weight) if self.learned_shortcut else x
        dx = NF.conv2d(actvn(x), self.conv_0.weight, self.conv_0.bias, padding=1)
        dx = NF.conv2d(actvn(dx), self.conv_1.weight, self.conv_1.bias, padding=1)
        return x_s + 0.1 * dx


class CUB_Resnet_Encoder(BaseEncoder):
    """cub.py:144-193: conv_img 3 -> nf, ResnetBlock(nf, nf), then per level AvgPool2d(3, 2, 1) + ResnetBlock(nf0, nf1), two
    Linear heads on actvn(features)."""

    def __init__(self, latent_dim, s0=16, nfilter=64, nfilter_max=1024):
        super().__init__()
        self.latent_dim = latent_dim
        size = 64
        self.s0, self.nf, self.nf_max = s0, nfilter, nfilter_max
        nf, nf_max = nfilter, nfilter_max
        nlayers = int(math.log2(size / s0))
        self.nf0 = min(nf_max, nf * 2 ** nlayers)
        blocks = [CUB_ResnetBlock(nf, nf)]
        for i in range(nlayers):
            nf0 = min(nf * 2 ** i, nf_max)
            nf1 = min(nf * 2 ** (i + 1), nf_max)
            blocks += [nn.AvgPool2d(3, stride=2, padding=1), CUB_ResnetBlock(nf0, nf1)]
        self.conv_img = nn.Conv2d(3, 1 * nf, 3, padding=1)
        self.resnet = nn.Sequential(*blocks)
        self.fc_mu = nn.Linear(self.nf0 * s0 * s0, self.latent_dim)
        self.fc_logvar = nn.Linear(self.nf0 * s0 * s0, self.latent_dim)

    def forward(self, x):
        from . import cub_native as CN
        if CN.use_native(self, x):
            return CN.encoder_forward(self, x)
        out = NF.conv2d(x, self.conv_img.weight, self.conv_img.bias, padding=1)
        out = self.resnet(out)
        out = actvn(out.reshape(x.size(0), self.nf0 * self.s0 * self.s0))
        mu, lv = NF.linear_heads(out, [self.fc_mu, self.fc_logvar])
        return ModelOutput(embedding=mu, log_covariance=lv)


class CUB_Resnet_Decoder(BaseDecoder):
    """cub.py:196-246: fc -> (nf0, s0, s0), per level ResnetBlock(nf0, nf1) + Upsample(2), ResnetBlock(nf, nf), conv_img nf -> 3
    on actvn(features); no output activation (the reference builds a Sigmoid and does not apply it, cub.py:236,244)."""

    def __init__(self, latent_dim, s0=16, nfilter=64, nfilter_max=512, **kwargs):
        super().__init__()
        size = 64
        self.latent_dim = latent_dim
        self.s0, self.nf, self.nf_max = s0, nfilter, nfilter_max
        nf, nf_max = nfilter, nfilter_max
        nlayers = int(math.log2(size / s0))
        self.nf0 = min(nf_max, nf * 2 ** nlayers)
        self.fc = nn.Linear(self.latent_dim, self.nf0 * s0 * s0)
        blocks = []
        for i in range(nlayers):
            nf0 = min(nf * 2 ** (nlayers - i), nf_max)
            nf1 = min(nf * 2 ** (nlayers - i - 1), nf_max)
            blocks += [CUB_ResnetBlock(nf0, nf1), nn.Upsample(scale_factor=2)]
        blocks += [CUB_ResnetBlock(nf, nf)]
        self.resnet = nn.Sequential(*blocks)
        self.conv_img = nn.Conv2d(nf, 3, 3, padding=1)
        self.sigmoid = nn.Sigmoid()

    def forward(self, z, out_dtype=None):
        from . import cub_native as CN
        if CN.use_native(self, z):
            return CN.decoder_forward(self, z)
        out = NF.linear(z.reshape(-1, z.size(-1)), self.fc.weight, self.fc.bias).view(-1, self.nf0, self.s0, self.s0)
        out = self.resnet(out)
        out = NF.conv2d(actvn(out), self.conv_img.weight, self.conv_img.bias, padding=1)
        return ModelOutput(reconstruction=out.view(*z.size()[:-1], *out.size()[1:]))
