"""CPU restatement of the reference's encoder/decoder networks as pure functions of a state_dict
(oracle; test infra only).  Keys/shapes are the reference's own (models/nn/*.py); `p` is a dict
name->tensor and `pre` the key prefix (e.g. "decoders.m0.").  Paths relative to
/root/reference/src/multivae/models/nn/."""
import math

import torch
import torch.nn.functional as F


def _lin(p, k, x):
    return F.linear(x, p[k + ".weight"], p.get(k + ".bias"))


# ---- default MLPs                                        default_architectures.py:21-73,75-140,225-258
def encoder_vae_mlp(p, pre, x, n_hidden=1):
    h = x.reshape(x.shape[0], -1)
    for i in range(1 + n_hidden):
        h = F.relu(_lin(p, f"{pre}layers.{i}.0", h))
    return _lin(p, pre + "embedding", h), _lin(p, pre + "log_var", h)


def encoder_vae_mlp_style(p, pre, x):
    h = F.relu(_lin(p, pre + "layers.0.0", x.reshape(x.shape[0], -1)))
    return (_lin(p, pre + "embedding", h), _lin(p, pre + "log_var", h),
            _lin(p, pre + "style_embedding", h), _lin(p, pre + "style_log_var", h))


def decoder_ae_mlp(p, pre, z, input_dim):
    h = F.relu(_lin(p, pre + "layers.0.0", z))
    h = torch.sigmoid(_lin(p, pre + "layers.1.0", h))
    return h.reshape(*z.shape[:-1], *input_dim)


# ---- SVHN conv                                           svhn.py:7-70
def encoder_vae_svhn(p, pre, x):
    h = x
    for i in (0, 2, 4):
        h = F.relu(F.conv2d(h, p[f"{pre}enc.{i}.weight"], p[f"{pre}enc.{i}.bias"], stride=2, padding=1))
    mu = F.conv2d(h, p[pre + "c1.weight"], p[pre + "c1.bias"], stride=2).squeeze()
    lv = F.conv2d(h, p[pre + "c2.weight"], p[pre + "c2.bias"], stride=2).squeeze()
    return mu, lv


def decoder_vae_svhn(p, pre, z):
    h = z.reshape(-1, z.shape[-1], 1, 1)
    h = F.relu(F.conv_transpose2d(h, p[pre + "dec.0.weight"], p[pre + "dec.0.bias"], stride=1, padding=0))
    h = F.relu(F.conv_transpose2d(h, p[pre + "dec.2.weight"], p[pre + "dec.2.bias"], stride=2, padding=1))
    h = F.relu(F.conv_transpose2d(h, p[pre + "dec.4.weight"], p[pre + "dec.4.bias"], stride=2, padding=1))
    h = torch.sigmoid(F.conv_transpose2d(h, p[pre + "dec.6.weight"], p[pre + "dec.6.bias"], stride=2, padding=1))
    return h.reshape(*z.shape[:-1], *h.shape[1:])


# ---- PolyMNIST conv                                      mmnist.py:78-110,173-207
def encoder_conv_mmnist_adapted(p, pre, x):
    h = x
    for i in (0, 2, 4):
        h = F.relu(F.conv2d(h, p[f"{pre}shared_encoder.{i}.weight"], p[f"{pre}shared_encoder.{i}.bias"], stride=2, padding=1))
    mu = F.conv2d(h, p[pre + "class_mu.weight"], p[pre + "class_mu.bias"], stride=2).squeeze()
    lv = F.conv2d(h, p[pre + "class_logvar.weight"], p[pre + "class_logvar.bias"], stride=2).squeeze()
    return mu, lv


def decoder_conv_mmnist(p, pre, z):
    h = F.relu(_lin(p, pre + "decoder.0", z.reshape(-1, z.shape[-1]))).view(-1, 128, 4, 4)
    h = F.relu(F.conv_transpose2d(h, p[pre + "decoder.3.weight"], p[pre + "decoder.3.bias"], stride=2, padding=1))
    h = F.relu(F.conv_transpose2d(h, p[pre + "decoder.5.weight"], p[pre + "decoder.5.bias"], stride=2, padding=1, output_padding=1))
    h = F.conv_transpose2d(h, p[pre + "decoder.7.weight"], p[pre + "decoder.7.bias"], stride=2, padding=1, output_padding=1)
    return h.view(*z.shape[:-1], *h.shape[1:])


# ---- PolyMNIST ResNet                                    mmnist.py:214-366
def resnet_block(p, pre, x):
    h = F.leaky_relu(F.conv2d(x, p[pre + "conv_layers.0.weight"], p[pre + "conv_layers.0.bias"], padding=1), 0.2)
    dx = F.leaky_relu(F.conv2d(h, p[pre + "conv_layers.2.weight"], p.get(pre + "conv_layers.2.bias"), padding=1), 0.2)
    xs = F.conv2d(x, p[pre + "shortcut_layer.weight"]) if (pre + "shortcut_layer.weight") in p else x
    return xs + 0.1 * dx


def _resnet_enc_branch(p, pre, x, tag):
    h = F.conv2d(x, p[f"{pre}conv_img_{tag}.weight"], p[f"{pre}conv_img_{tag}.bias"], padding=1)
    h = resnet_block(p, f"{pre}resnet_{tag}.0.", h)
    h = F.avg_pool2d(h, 3, stride=2, padding=1)
    h = resnet_block(p, f"{pre}resnet_{tag}.2.", h)
    h = F.avg_pool2d(h, 3, stride=2, padding=1)
    h = resnet_block(p, f"{pre}resnet_{tag}.4.", h)
    h = h.view(h.size(0), -1)
    return _lin(p, f"{pre}fc_mu_{tag}", h), _lin(p, f"{pre}fc_lv_{tag}", h)


def encoder_resnet_mmnist(p, pre, x):
    mu, lv = _resnet_enc_branch(p, pre, x, "u")
    if (pre + "conv_img_w.weight") in p:
        mu_w, lv_w = _resnet_enc_branch(p, pre, x, "w")
        return mu, lv, mu_w, lv_w
    return mu, lv


def decoder_resnet_mmnist(p, pre, z):
    h = _lin(p, pre + "fc", z).view(-1, 256, 7, 7)
    h = resnet_block(p, pre + "resnet.0.", h)
    h = F.interpolate(h, scale_factor=2)
    h = resnet_block(p, pre + "resnet.2.", h)
    h = F.interpolate(h, scale_factor=2)
    h = resnet_block(p, pre + "resnet.4.", h)
    h = F.leaky_relu(F.conv2d(h, p[pre + "conv_img.0.weight"], p[pre + "conv_img.0.bias"], padding=1), 0.2)
    return h.view(*z.shape[:-1], *h.shape[1:])


# ---- CUB 64 x 64 ResNets                                cub.py:144-293
def cub_resnet_block(p, pre, x):
    """Pre-activation block, cub.py:249-293: x_s + 0.1 * conv_1(actvn(conv_0(actvn(x))))."""
    xs = F.conv2d(x, p[pre + "conv_s.weight"]) if (pre + "conv_s.weight") in p else x
    dx = F.conv2d(F.leaky_relu(x, 0.2), p[pre + "conv_0.weight"], p[pre + "conv_0.bias"], padding=1)
    dx = F.conv2d(F.leaky_relu(dx, 0.2), p[pre + "conv_1.weight"], p.get(pre + "conv_1.bias"), padding=1)
    return xs + 0.1 * dx


def _cub_block_ids(p, pre):
    return sorted({int(k[len(pre + "resnet."):].split(".")[0]) for k in p if k.startswith(pre + "resnet.") and ".conv_0.weight" in k})


def cub_resnet_encoder(p, pre, x):
    """cub.py:185-193: conv_img, ResnetBlock, (AvgPool2d(3, 2, 1), ResnetBlock)*, heads on actvn(flattened features)."""
    h = F.conv2d(x, p[pre + "conv_img.weight"], p[pre + "conv_img.bias"], padding=1)
    for j, i in enumerate(_cub_block_ids(p, pre)):
        if j > 0:
            h = F.avg_pool2d(h, 3, stride=2, padding=1)
        h = cub_resnet_block(p, f"{pre}resnet.{i}.", h)
    h = F.leaky_relu(h.reshape(x.shape[0], -1), 0.2)
    return _lin(p, pre + "fc_mu", h), _lin(p, pre + "fc_logvar", h)


def cub_resnet_decoder(p, pre, z, s0=16):
    """cub.py:238-246: fc, (ResnetBlock, Upsample(2))*, ResnetBlock, conv_img on actvn(features); no output activation."""
    ids = _cub_block_ids(p, pre)
    nf0 = p[f"{pre}resnet.{ids[0]}.conv_0.weight"].shape[1]
    h = _lin(p, pre + "fc", z).view(z.shape[0], nf0, s0, s0)
    for j, i in enumerate(ids):
        h = cub_resnet_block(p, f"{pre}resnet.{i}.", h)
        if j + 1 < len(ids):
            h = F.interpolate(h, scale_factor=2)
    return F.conv2d(F.leaky_relu(h, 0.2), p[pre + "conv_img.weight"], p[pre + "conv_img.bias"], padding=1)


# ---- deterministic synthetic weights (shared by the golden generator and the tests) -------------
def synth_state_dict(shapes, seed=0, dtype=torch.float32):
    """name->shape  ->  name->tensor, PyTorch-default-like fan-in scaling, independent of module
    construction order: each tensor is drawn from its own generator seeded by (seed, index of the
    sorted key)."""
    out = {}
    for i, k in enumerate(sorted(shapes)):
        shp = tuple(shapes[k])
        g = torch.Generator().manual_seed(seed * 100003 + i)
        fan_in = 1
        if len(shp) >= 2:
            fan_in = int(torch.tensor(shp[1:]).prod())
        elif k.endswith(".bias"):
            w = shapes.get(k[:-5] + ".weight")
            fan_in = int(torch.tensor(tuple(w)[1:]).prod()) if w is not None else shp[0]
        bound = 1.0 / math.sqrt(max(fan_in, 1))
        out[k] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return out
