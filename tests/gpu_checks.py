"""Shared GPU check helpers: replay golden cases through the product path (CUDA kernels via the C-ABI)
and compare with the reference's golden values and with the oracle port run on CPU.
Imports oracle/ as the checker only (tests + smoke)."""
import os

import torch

import multivae_b200 as mb
from oracle.cases import CASES, make_data
from oracle.port.nets import synth_state_dict
from oracle.replay import run_port

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = {"mmvaeplus": (mb.MMVAEPlus, mb.MMVAEPlusConfig), "mmvae": (mb.MMVAE, mb.MMVAEConfig),
          "mvtcae": (mb.MVTCAE, mb.MVTCAEConfig), "mvae": (mb.MVAE, mb.MVAEConfig), "mopoe": (mb.MoPoE, mb.MoPoEConfig),
          "cmvae": (mb.CMVAE, mb.CMVAEConfig), "crmvae": (mb.CRMVAE, mb.CRMVAEConfig)}


def load_golden(name):
    return torch.load(os.path.join(GOLD, f"elbo_{name}.pt"), weights_only=False)


def product_archs(spec):
    """multivae_b200.nn encoders / decoders for a spec's `arch` table (None, None = the model's own defaults)."""
    if "arch" not in spec:
        return None, None
    from multivae_b200 import nn as NN
    L, Lw = spec["cfg"]["latent_dim"], spec["cfg"].get("modalities_specific_dim")
    enc, dec = {}, {}
    for m, a in spec["arch"].items():
        c = mb.BaseAEConfig(input_dim=tuple(spec["dims"][m]), latent_dim=L)
        if a == "mlp":
            enc[m], dec[m] = NN.Encoder_VAE_MLP(c), NN.Decoder_AE_MLP(c)
        elif a == "svhn":
            enc[m], dec[m] = NN.Encoder_VAE_SVHN(c), NN.Decoder_VAE_SVHN(c)
        elif a == "conv_mmnist":
            enc[m], dec[m] = NN.EncoderConvMMNIST_adapted(c), NN.DecoderConvMMNIST(c)
        elif a == "resnet_mmnist":
            enc[m], dec[m] = NN.EncoderResnetMMNIST(Lw or 0, L), NN.DecoderResnetMMNIST(L + (Lw or 0))
        else:
            raise ValueError(a)
    return enc, dec


def build_model(spec, rec, device):
    import copy
    cls, cfgcls = MODELS[spec["model"]]
    cfg = cfgcls(n_modalities=len(spec["dims"]), input_dims=dict(spec["dims"]), **copy.deepcopy(spec["cfg"]))
    enc, dec = product_archs(spec)
    model = cls(cfg, enc, dec)
    sd = synth_state_dict(rec["state_shapes"], seed=rec["sd_seed"])
    missing = set(model.state_dict().keys()) ^ set(sd.keys())
    assert not missing, f"state_dict keys differ from the reference: {sorted(missing)[:5]}"
    model.load_state_dict(sd)
    return model.to(device).train()


def run_product(name, device="cuda", compute_dtype=None):
    """Returns (ModelOutput, model, golden record) after loss.backward().  compute_dtype = torch.bfloat16 runs the encoder /
    decoder contractions on the native tensor-core kernels (bf16 operands, fp32 accumulate)."""
    import numpy as np
    spec, rec = CASES[name], load_golden(name)
    model = build_model(spec, rec, device)
    if compute_dtype is not None:
        model.compute_dtype = compute_dtype
    # the fp32 comparison runs the library layers in true fp32 (cuDNN / cuBLAS default to TF32 convolutions)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    data, masks = make_data(spec)
    data = {k: v.to(device) for k, v in data.items()}
    q = [e.to(device) for e in rec["noise"]]

    def noise_source(shape, kind, dev):
        e = q.pop(0)
        assert tuple(e.shape) == tuple(shape), (e.shape, shape)
        return e

    model.noise_source = noise_source
    if spec["model"] == "mopoe" and masks is not None:
        model.choice_source = lambda probs: rec["choice"]
    if masks is not None:
        ds = mb.IncompleteDataset(data=data, masks={k: v.to(device) for k, v in masks.items()})
    else:
        ds = mb.MultimodalBaseDataset(data=data)
    if "np_seed" in spec:
        np.random.seed(spec["np_seed"])
    out = model(ds, **spec.get("fwd", {}))
    out.loss.backward()
    assert not q, "noise consumption order differs from the reference"
    return out, model, rec


def rel(a, b):
    a, b = float(a), float(b)
    return abs(a - b) / max(abs(b), 1e-12)


# Per-case tolerance overrides: none — every case, cfg5 (D = 12288, DReG) included, is held to 1e-4.
CASE_TOL = {}


def check_case(name, rtol_loss=1e-4, verbose=False):
    rtol_loss = CASE_TOL.get(name, {}).get("loss", rtol_loss)
    rtol_grad = CASE_TOL.get(name, {}).get("grad", 2e-3)
    out, model, rec = run_product(name)
    errs = {"loss": rel(out.loss.detach().cpu(), rec["loss"]), "loss_sum": rel(out.loss_sum.detach().cpu(), rec["loss_sum"])}
    assert errs["loss"] <= rtol_loss, (name, "loss", float(out.loss), float(rec["loss"]))
    assert errs["loss_sum"] <= rtol_loss, (name, "loss_sum")
    for k, v in rec["metrics"].items():
        got = out.metrics[k]
        got = got.detach().cpu() if torch.is_tensor(got) else got
        assert rel(got, v) <= 1e-4 or abs(float(got) - float(v)) < 1e-5, (name, "metric", k, float(got), float(v))
    if "lws" in rec:
        mods = [m for m in rec["lws"]]
        lw = model._last["lw"].cpu()
        for i, m in enumerate(mods):
            assert torch.allclose(lw[i], rec["lws"][m], rtol=1e-4, atol=1e-3), (name, "lw", m)
    # oracle port on CPU on the same inputs (the checker), besides the reference's golden numbers
    ploss, _, _, pp = run_port(CASES[name], rec)
    assert rel(out.loss.detach().cpu(), ploss.detach()) <= rtol_loss
    scale = max(1.0, abs(float(rec["loss"])))
    worst = 0.0
    big = "arch" in CASES[name] or CASES[name]["B"] >= 32
    for k, p in model.named_parameters():
        g = rec["grads"][k]
        if g is None:
            assert p.grad is None or float(p.grad.abs().sum()) == 0.0, (name, k, "expected no gradient")
            continue
        assert p.grad is not None, (name, k)
        pg = p.grad.detach().cpu()
        ref_full = pp[k].grad
        denom = max(float(ref_full.abs().max()), 1e-6 * scale)
        err = float((pg - ref_full).abs().max()) / denom
        l2 = float((pg - ref_full).double().norm()) / max(float(ref_full.double().norm()), 1e-6 * scale)
        worst = max(worst, err)
        # element-wise bound; at the configurations' full batch sizes a handful of elements sit on a discontinuity of the
        # gradient (sign(x - recon) of the Laplace likelihood, ReLU' at a pre-activation within rounding of 0) and flip
        # between two fp32 implementations, and cuDNN's fp32 convolution algorithms differ from the CPU's at the 1e-3 level on
        # single elements: for those cases outliers up to 5e-2 of the tensor maximum are accepted when the tensor as a whole
        # (relative L2) agrees within 5e-3 (one flipped ReLU unit of one sample moves a whole row of the first decoder weight's gradient by
        # ~6 % of that row; the loss itself agrees to 1e-7).  The small cases keep the element-wise 2e-3 bound.
        assert err <= rtol_grad or (big and err <= 5e-2 and l2 <= 5e-3), (name, k, err, l2)
        assert l2 <= (5e-3 if big else rtol_grad), (name, k, "l2", l2)
        assert abs(float(pg.double().sum()) - g["sum"]) <= rtol_grad * max(g["abssum"], 1e-3) + 1e-5 * scale, (name, k)
    errs["grad_max_rel"] = worst
    if verbose:
        print(name, errs)
    return errs


def check_case_bf16(name, verbose=False):
    """The same golden case with the encoder / decoder contractions on the native tensor-core path (bf16 operands, fp32
    accumulate, fp32 master weights, fp32 ELBO): returns the loss error and the worst parameter-gradient error against the
    fp32 reference — the distance is the operand precision BASELINE.json asks for, so it is reported and bounded, not 1e-4."""
    out, model, rec = run_product(name, compute_dtype=torch.bfloat16)
    _, _, _, pp = run_port(CASES[name], rec)
    errs = {"loss": rel(out.loss.detach().cpu(), rec["loss"]), "grad_max_rel": 0.0, "grad_rel_l2": 0.0}
    num = den = 0.0
    for k, p in model.named_parameters():
        if rec["grads"][k] is None:
            continue
        ref = pp[k].grad
        pg = p.grad.detach().float().cpu()
        e = float((pg - ref).abs().max()) / max(float(ref.abs().max()), 1e-12)
        if e > errs["grad_max_rel"]:
            errs["grad_max_rel"], errs["grad_worst"] = e, k
        num += float((pg - ref).double().pow(2).sum())
        den += float(ref.double().pow(2).sum())
    errs["grad_rel_l2"] = (num / max(den, 1e-30)) ** 0.5
    if verbose:
        print(name, "bf16 tensor path:", errs)
    return errs


def smoke_check():
    """One small invocation of the hot path on cuda:0, checked against the oracle."""
    assert torch.cuda.is_available(), "smoke() needs a GPU"
    torch.cuda.set_device(0)
    # 1) the north-star model on the tcgen05 path FIRST (so its kernels are inside the driver's launch window): the real
    #    reference's golden (ResNet encoders / decoders, K = 10, non-initial weights), bf16 operands vs the fp32 reference
    e16 = check_case_bf16("ns_mmvaeplus_resnet", verbose=True)
    assert e16["loss"] <= NS_BF16_LOSS_TOL and e16["grad_rel_l2"] <= NS_BF16_GRAD_L2_TOL, e16
    # 2) the fused ELBO kernels on a small MMVAE+ case: loss / lw / every gradient against the reference golden and the port
    e = check_case("mmvaeplus_dreg", verbose=True)
    print("smoke ok:", e, "north-star bf16 tensor path", e16)


# bounds of the bf16 tensor-core path on the north-star golden = 3x what was measured on B200
NS_BF16_LOSS_TOL = 5e-6      # measured 8.2e-7
NS_BF16_GRAD_L2_TOL = 3e-2   # measured 9.3e-3 (relative L2 over all parameter gradients; worst single tensor 0.11 of its max)
