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
          "mvtcae": (mb.MVTCAE, mb.MVTCAEConfig), "mvae": (mb.MVAE, mb.MVAEConfig), "mopoe": (mb.MoPoE, mb.MoPoEConfig)}


def load_golden(name):
    return torch.load(os.path.join(GOLD, f"elbo_{name}.pt"), weights_only=False)


def build_model(spec, rec, device):
    import copy
    cls, cfgcls = MODELS[spec["model"]]
    cfg = cfgcls(n_modalities=len(spec["dims"]), input_dims=dict(spec["dims"]), **copy.deepcopy(spec["cfg"]))
    model = cls(cfg)
    sd = synth_state_dict(rec["state_shapes"], seed=rec["sd_seed"])
    missing = set(model.state_dict().keys()) ^ set(sd.keys())
    assert not missing, f"state_dict keys differ from the reference: {sorted(missing)[:5]}"
    model.load_state_dict(sd)
    return model.to(device).train()


def run_product(name, device="cuda"):
    """Returns (ModelOutput, model) after loss.backward()."""
    import numpy as np
    spec, rec = CASES[name], load_golden(name)
    model = build_model(spec, rec, device)
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


# Per-case tolerances where fp32 itself cannot do better.  cfg5 (D = 12288, DReG): the importance weights are nearly one-hot
# over log-weights of magnitude 1.2e4, so fp32 rounding of the log-weights (1e-6 relative = 1e-2 absolute) moves the loss by
# ~1e-4 relative: the reference's own fp32 value differs from an fp64 evaluation of the same algorithm by 7.4e-5
# (tests/test_oracle_port.py::test_cfg5_fp32_conditioning), so 1e-4 agreement between two fp32 implementations is not defined.
CASE_TOL = {"cfg5_mmvaeplus_celeba": dict(loss=3e-4)}   # measured on B200: loss 1.24e-4, gradients 9e-4 of the tensor max


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
        worst = max(worst, err)
        assert err <= rtol_grad, (name, k, err)
        assert abs(float(pg.double().sum()) - g["sum"]) <= rtol_grad * max(g["abssum"], 1e-3) + 1e-5 * scale, (name, k)
    errs["grad_max_rel"] = worst
    if verbose:
        print(name, errs)
    return errs


def smoke_check():
    """One small invocation of the hot path on cuda:0, checked against the oracle."""
    assert torch.cuda.is_available(), "smoke() needs a GPU"
    torch.cuda.set_device(0)
    e = check_case("mmvaeplus_dreg", verbose=True)
    # the north-star model (ResNet encoders/decoders): library-network fp32 path and the tcgen05 decoder path (bf16
    # operands) against the fp32 CPU oracle, plus one backward through the native decoders
    import bench
    dev = torch.device("cuda", 0)
    r32 = bench.elbo_rel_err(dev)
    r16 = bench.elbo_rel_err(dev, torch.bfloat16)
    assert r32 <= 1e-4, r32
    assert r16 <= 2e-2, r16
    model = bench.north_star_model(dev)
    model.compute_dtype = torch.bfloat16
    out = model(mb.MultimodalBaseDataset(data={k: v.to(dev) for k, v in bench.synthetic_batch(2).items()}), K=2)
    out.loss.backward()
    g = model.decoders["m0"].resnet[4].conv_layers[0].weight.grad
    assert g is not None and bool(torch.isfinite(g).all()) and float(g.abs().sum()) > 0
    print("smoke ok:", e, "north-star ELBO rel err fp32 path", r32, "bf16 tensor path", r16)
