"""Encoder / decoder modules of multivae_b200.nn against the golden outputs of the REAL reference modules
(tests/golden/nets_*.pt, written by oracle/make_golden.py): outputs, input gradient, every parameter gradient.
Shared by the CPU tests (library-layer path) and the GPU tests (native sm_100a kernels)."""
import os

import torch

import multivae_b200 as mb
from multivae_b200 import nn as N
from oracle.port.nets import synth_state_dict

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _c(i, l, s=0):
    return mb.BaseAEConfig(input_dim=i, latent_dim=l, style_dim=s)


NETS = {
    "enc_resnet_mmnist": lambda: N.EncoderResnetMMNIST(32, 32), "dec_resnet_mmnist": lambda: N.DecoderResnetMMNIST(64),
    "enc_conv_mmnist": lambda: N.EncoderConvMMNIST_adapted(_c((3, 28, 28), 64)), "dec_conv_mmnist": lambda: N.DecoderConvMMNIST(_c((3, 28, 28), 64)),
    "enc_svhn": lambda: N.Encoder_VAE_SVHN(_c((3, 32, 32), 20)), "dec_svhn": lambda: N.Decoder_VAE_SVHN(_c((3, 32, 32), 20)),
    "enc_mlp": lambda: N.Encoder_VAE_MLP(_c((1, 28, 28), 20)), "enc_mlp_style": lambda: N.Encoder_VAE_MLP_Style(_c((3, 8, 8), 8, 4)),
    "dec_mlp": lambda: N.Decoder_AE_MLP(_c((1, 28, 28), 20)),
    "enc_cub_resnet": lambda: N.CUB_Resnet_Encoder(32), "dec_cub_resnet": lambda: N.CUB_Resnet_Decoder(32),
}


def golden_input(rec):
    g = torch.Generator().manual_seed(77)
    x = torch.rand(rec["in_shape"], generator=g) if rec["kind"] == "x" else torch.randn(rec["in_shape"], generator=g)
    return x


def check_net(name, device, rtol=1e-4, atol=1e-5, grad_tol=1e-3, autocast=False, verbose=False, l2_tol=None, collect=False):
    """Returns {"out": worst output error / output max, "grad": worst parameter-gradient error / gradient max}."""
    rec = torch.load(os.path.join(GOLD, f"nets_{name}.pt"), weights_only=False)
    net = NETS[name]()
    net.load_state_dict(synth_state_dict(rec["state_shapes"], seed=rec["sd_seed"]))
    net = net.to(device).train()
    x = golden_input(rec).to(device).requires_grad_(True)
    import contextlib
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
    with ctx:
        out = net(x)
    errs = {"out": 0.0, "grad": 0.0}
    for k, v in rec["outputs"].items():
        got = out[k].detach().float().cpu()
        assert got.shape == v.shape, (name, k, got.shape, v.shape)
        errs["out"] = max(errs["out"], float((got - v).abs().max()) / max(float(v.abs().max()), 1e-6))
        assert collect or torch.allclose(got, v, rtol=rtol, atol=atol * max(1.0, float(v.abs().max()))), (name, k, float((got - v).abs().max()))
    sum((out[k].float() * rec["cot"][k].to(device)).sum() for k in rec["outputs"]).backward()
    if x.grad is not None:   # the native ResNet encoders do not produce a gradient for the data (nothing trains on it)
        gi = x.grad.float().cpu()
        scale = max(float(rec["grad_in"].abs().max()), 1e-6)
        errs["grad_in"] = float((gi - rec["grad_in"]).abs().max()) / scale
        errs["grad_in_l2"] = float((gi - rec["grad_in"]).norm() / rec["grad_in"].norm().clamp(min=1e-12))
        assert collect or errs["grad_in"] <= grad_tol, (name, "grad_in", errs["grad_in"])
        if l2_tol is not None and not collect:
            assert errs["grad_in_l2"] <= l2_tol, (name, "grad_in_l2", errs["grad_in_l2"])
    for k, p in net.named_parameters():
        g = rec["grads"][k]
        pg = p.grad.detach().float().cpu()
        ref = g["full"] if g["full"] is not None else None
        if ref is not None:
            e = float((pg - ref).abs().max()) / max(float(ref.abs().max()), 1e-6)
        else:
            e = float((pg.flatten()[:8] - g["head"]).abs().max()) / max(float(g["head"].abs().max()), 1e-6)
            # the whole tensor through its sum
            assert collect or abs(float(pg.double().sum()) - g["sum"]) <= grad_tol * max(g["abssum"], 1e-3) + 1e-5, (name, k, "sum")
        errs["grad"] = max(errs["grad"], e)
        assert collect or (e <= grad_tol * 10 if ref is None else e <= grad_tol), (name, k, e)
    if verbose:
        print(name, errs)
    return errs
