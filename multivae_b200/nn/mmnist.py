"""PolyMNIST architectures (reference: models/nn/mmnist.py:78-110,173-207,214-366).

Parameter names follow the reference exactly (`shared_encoder.{0,2,4}`, `class_mu`, `class_logvar`,
`decoder.{0,3,5,7}`, `conv_img_{u,w}`, `resnet_{u,w}.{0,2,4}.conv_layers.{0,2}`, `shortcut_layer`,
`fc_{mu,lv}_{u,w}`, `fc`, `resnet.{0,2,4}`, `conv_img.0`).  The ResNet pair is the north-star
architecture: on CUDA its forward/backward run as tcgen05 implicit-GEMM kernels on a bf16
shared-halo NHWC layout (multivae_b200/nn/resnet_native.py)."""
import torch.nn as nn
import torch.nn.functional as F

from ..containers import ModelOutput
from . import functional as NF
from .base_architectures import BaseDecoder, BaseEncoder
from .default_architectures import _native


class Unflatten(nn.Module):
    def __init__(self, ndims):
        super().__init__()
        self.ndims = ndims

    def forward(self, x):
        return x.view(x.size(0), *self.ndims)


class EncoderConvMMNIST_adapted(BaseEncoder):
    def __init__(self, model_config):
        super().__init__()
        self.latent_dim = model_config.latent_dim
        self.style_dim = 0
        self.shared_encoder = nn.Sequential(nn.Conv2d(3, 32, 3, 2, 1), nn.ReLU(), nn.Conv2d(32, 64, 3, 2, 1), nn.ReLU(),
                                            nn.Conv2d(64, 128, 3, 2, 1), nn.ReLU())
        self.class_mu = nn.Conv2d(128, self.latent_dim, 4, 2, 0)
        self.class_logvar = nn.Conv2d(128, self.latent_dim, 4, 2, 0)

    def forward(self, x):
        if _native(x):
            from .conv_native import conv_encoder
            e = self.shared_encoder
            mu, lv = conv_encoder(x, [e[0], e[2], e[4]], [self.class_mu, self.class_logvar])
            return ModelOutput(embedding=mu.squeeze(), log_covariance=lv.squeeze())
        h = x
        for i in (0, 2, 4):
            h = NF.conv2d(h, self.shared_encoder[i].weight, self.shared_encoder[i].bias, stride=2, padding=1, act="relu")
        mu = NF.conv2d(h, self.class_mu.weight, self.class_mu.bias, stride=2).squeeze()
        lv = NF.conv2d(h, self.class_logvar.weight, self.class_logvar.bias, stride=2).squeeze()
        return ModelOutput(embedding=mu, log_covariance=lv)


class DecoderConvMMNIST(BaseDecoder):
    def __init__(self, model_config):
        super().__init__()
        self.latent_dim = model_config.latent_dim
        self.decoder = nn.Sequential(
            nn.Linear(self.latent_dim, 2048), nn.ReLU(), Unflatten((128, 4, 4)),
            nn.ConvTranspose2d(128, 64, 3, 2, 1), nn.ReLU(),
            nn.ConvTranspose2d(64, 32, 3, 2, 1, output_padding=1), nn.ReLU(),
            nn.ConvTranspose2d(32, 3, 3, 2, 1, output_padding=1))

    def forward(self, z):
        d = self.decoder
        if _native(z):
            from .conv_native import convt_decoder
            h = convt_decoder(z.reshape(-1, z.size(-1)), d[0], "linear", [d[3], d[5], d[7]], "none", (128, 4, 4))
            return ModelOutput(reconstruction=h.view(*z.size()[:-1], *h.size()[1:]))
        h = NF.linear(z.reshape(-1, z.size(-1)), d[0].weight, d[0].bias, act="relu").view(-1, 128, 4, 4)
        h = NF.conv_transpose2d(h, d[3].weight, d[3].bias, stride=2, padding=1, act="relu")
        h = NF.conv_transpose2d(h, d[5].weight, d[5].bias, stride=2, padding=1, output_padding=1, act="relu")
        h = NF.conv_transpose2d(h, d[7].weight, d[7].bias, stride=2, padding=1, output_padding=1)
        return ModelOutput(reconstruction=h.view(*z.size()[:-1], *h.size()[1:]))


class ResnetBlock(nn.Module):
    """x_s + 0.1 * lrelu(conv3x3(lrelu(conv3x3(x))));  x_s = x or a bias-free 1x1 conv when the widths differ."""

    def __init__(self, nb_channels_in, nb_channels_out, nb_channels_hidden=None, bias=True):
        super().__init__()
        self.learn_shortcut = nb_channels_in != nb_channels_out
        hid = min(nb_channels_in, nb_channels_out) if nb_channels_hidden is None else nb_channels_hidden
        self.conv_layers = nn.Sequential(nn.Conv2d(nb_channels_in, hid, 3, 1, 1), nn.LeakyReLU(0.2),
                                         nn.Conv2d(hid, nb_channels_out, 3, 1, 1, bias=bias), nn.LeakyReLU(0.2))
        if self.learn_shortcut:
            self.shortcut_layer = nn.Conv2d(nb_channels_in, nb_channels_out, 1, 1, 0, bias=False)

    def forward(self, x):
        c = self.conv_layers
        xs = NF.conv2d(x, self.shortcut_layer.weight) if self.learn_shortcut else x
        h = NF.conv2d(x, c[0].weight, c[0].bias, padding=1, act="lrelu")
        dx = NF.conv2d(h, c[2].weight, c[2].bias, padding=1, act="lrelu")
        return xs + 0.1 * dx


def _resnet_stack(nf, widths, between):
    blocks = [ResnetBlock(nf, nf)] if between == "pool" else []
    for a, b in widths:
        if between == "pool":
            blocks += [nn.AvgPool2d(3, stride=2, padding=1), ResnetBlock(a, b)]
        else:
            blocks += [ResnetBlock(a, b), nn.Upsample(scale_factor=2)]
    if between != "pool":
        blocks += [ResnetBlock(nf, nf)]
    return nn.Sequential(*blocks)


class EncoderResnetMMNIST(BaseEncoder):
    def __init__(self, private_latent_dim, shared_latent_dim):
        super().__init__()
        self.latent_dim = shared_latent_dim
        self.s0, self.nf, self.nf_max = 7, 64, 1024
        self.nf0 = 256
        self.multiple_latent = private_latent_dim > 0
        widths = [(64, 128), (128, 256)]
        if self.multiple_latent:
            self.conv_img_w = nn.Conv2d(3, 64, 3, padding=1)
            self.resnet_w = _resnet_stack(64, widths, "pool")
            self.fc_mu_w = nn.Linear(256 * 49, private_latent_dim)
            self.fc_lv_w = nn.Linear(256 * 49, private_latent_dim)
        self.conv_img_u = nn.Conv2d(3, 64, 3, padding=1)
        self.resnet_u = _resnet_stack(64, widths, "pool")
        self.fc_mu_u = nn.Linear(256 * 49, shared_latent_dim)
        self.fc_lv_u = nn.Linear(256 * 49, shared_latent_dim)

    def _branch(self, x, tag):
        h = NF.conv2d(x, getattr(self, f"conv_img_{tag}").weight, getattr(self, f"conv_img_{tag}").bias, padding=1)
        h = getattr(self, f"resnet_{tag}")(h)
        h = h.reshape(h.size(0), -1)
        return NF.linear_heads(h, [getattr(self, f"fc_mu_{tag}"), getattr(self, f"fc_lv_{tag}")])

    def forward(self, x):
        from . import resnet_native as RN
        if RN.use_native(x):
            return RN.encoder_forward(self, x)
        mu, lv = self._branch(x, "u")
        out = ModelOutput(embedding=mu, log_covariance=lv)
        if self.multiple_latent:
            out["style_embedding"], out["style_log_covariance"] = self._branch(x, "w")
        return out


class DecoderResnetMMNIST(BaseDecoder):
    def __init__(self, latent_dim):
        super().__init__()
        self.s0, self.nf, self.nf_max, self.nf0 = 7, 64, 512, 256
        self.fc = nn.Linear(latent_dim, 256 * 49)
        self.resnet = _resnet_stack(64, [(256, 128), (128, 64)], "up")
        self.conv_img = nn.Sequential(nn.Conv2d(64, 3, 3, padding=1), nn.LeakyReLU(0.2))

    def forward(self, z, out_dtype=None):
        from . import resnet_native as RN
        if RN.use_native(z):
            return RN.decoder_forward(self, z, out_dtype=out_dtype)
        h = NF.linear(z.reshape(-1, z.size(-1)), self.fc.weight, self.fc.bias).view(-1, 256, 7, 7)
        h = self.resnet(h)
        h = NF.conv2d(h, self.conv_img[0].weight, self.conv_img[0].bias, padding=1, act="lrelu")
        return ModelOutput(reconstruction=h.view(*z.size()[:-1], *h.size()[1:]))
