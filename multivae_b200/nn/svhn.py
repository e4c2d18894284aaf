"""SVHN conv encoder/decoder (reference: models/nn/svhn.py:7-70).  Parameter names: `enc.{0,2,4}`,
`c1`, `c2`, `dec.{0,2,4,6}`."""
import torch.nn as nn

from ..containers import ModelOutput
from . import functional as NF
from .base_architectures import BaseDecoder, BaseEncoder
from .default_architectures import _native


class Encoder_VAE_SVHN(BaseEncoder):
    def __init__(self, args):
        super().__init__()
        self.input_dim, self.latent_dim = args.input_dim, args.latent_dim
        ch, f = args.input_dim[0], 32
        self.fBase = f
        self.enc = nn.Sequential(nn.Conv2d(ch, f, 4, 2, 1), nn.ReLU(True), nn.Conv2d(f, 2 * f, 4, 2, 1), nn.ReLU(True),
                                 nn.Conv2d(2 * f, 4 * f, 4, 2, 1), nn.ReLU(True))
        self.c1 = nn.Conv2d(4 * f, self.latent_dim, 4, 2, 0)
        self.c2 = nn.Conv2d(4 * f, self.latent_dim, 4, 2, 0)

    def forward(self, x):
        if _native(x):
            from .conv_native import conv_encoder
            mu, lv = conv_encoder(x, [self.enc[0], self.enc[2], self.enc[4]], [self.c1, self.c2])
            return ModelOutput(embedding=mu.squeeze(), log_covariance=lv.squeeze())
        h = x
        for i in (0, 2, 4):
            h = NF.conv2d(h, self.enc[i].weight, self.enc[i].bias, stride=2, padding=1, act="relu")
        mu = NF.conv2d(h, self.c1.weight, self.c1.bias, stride=2).squeeze()
        lv = NF.conv2d(h, self.c2.weight, self.c2.bias, stride=2).squeeze()
        return ModelOutput(embedding=mu, log_covariance=lv)


class Decoder_VAE_SVHN(BaseDecoder):
    def __init__(self, args):
        super().__init__()
        self.latent_dim = args.latent_dim
        f, ch = 32, args.input_dim[0]
        self.fBase, self.nb_channels = f, ch
        self.dec = nn.Sequential(nn.ConvTranspose2d(self.latent_dim, 4 * f, 4, 1, 0), nn.ReLU(True),
                                 nn.ConvTranspose2d(4 * f, 2 * f, 4, 2, 1), nn.ReLU(True),
                                 nn.ConvTranspose2d(2 * f, f, 4, 2, 1), nn.ReLU(True),
                                 nn.ConvTranspose2d(f, ch, 4, 2, 1), nn.Sigmoid())

    def forward(self, z):
        if _native(z):
            from .conv_native import convt_decoder
            h = convt_decoder(z.reshape(-1, z.shape[-1]), self.dec[0], "convt", [self.dec[2], self.dec[4], self.dec[6]], "sigmoid",
                              (4 * self.fBase, 4, 4))
            return ModelOutput(reconstruction=h.reshape(*z.shape[:-1], *h.shape[1:]))
        h = z.reshape(-1, z.shape[-1], 1, 1)
        h = NF.conv_transpose2d(h, self.dec[0].weight, self.dec[0].bias, stride=1, padding=0, act="relu")
        h = NF.conv_transpose2d(h, self.dec[2].weight, self.dec[2].bias, stride=2, padding=1, act="relu")
        h = NF.conv_transpose2d(h, self.dec[4].weight, self.dec[4].bias, stride=2, padding=1, act="relu")
        h = NF.conv_transpose2d(h, self.dec[6].weight, self.dec[6].bias, stride=2, padding=1, act="sigmoid")
        return ModelOutput(reconstruction=h.reshape(*z.shape[:-1], *h.shape[1:]))
