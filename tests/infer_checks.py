"""Inference path (encode / predict / compute_joint_nll[_paper] / compute_cond_nll) of the product against golden vectors of the
REAL reference (tests/golden/infer_*.pt, oracle/make_golden_infer.py).  The reference consumes its sampling noise one datapoint /
one sample at a time; the product draws whole (samples, batch, latent) tensors, so the recorded draws are re-assembled into the
product's request order here (same numbers, different batching)."""
import copy
import os

import numpy as np
import torch

import multivae_b200 as mb
from oracle.cases import CASES, make_data
from oracle.make_golden_infer import COND_K, INFER_CASES, NLL_BK, NLL_K  # noqa: F401
from oracle.port.nets import synth_state_dict
from tests.gpu_checks import MODELS

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _model(name, rec, device):
    spec = CASES[name]
    cls, cfgcls = MODELS[spec["model"]]
    model = cls(cfgcls(n_modalities=len(spec["dims"]), input_dims=dict(spec["dims"]), **copy.deepcopy(spec["cfg"])))
    model.load_state_dict(synth_state_dict(rec["state_shapes"], seed=rec["sd_seed"]))
    return model.to(device).eval()


def _stack_calls(log, per_call):
    """n sequential reference calls of `per_call` draws each -> per_call tensors stacked over the calls."""
    n = len(log) // per_call
    return [torch.stack([log[t * per_call + j] for t in range(n)]) for j in range(per_call)]


def _mmvaeplus_nll_plan(log, B, per_i, k, chunk):
    """Reference: per datapoint i, `per_i` draws of shape (k, 1, .).  Product: per chunk of samples, the same per_i requests
    with shape (n, B, .)."""
    out = []
    for k0 in range(0, k, chunk):
        n = min(chunk, k - k0)
        for j in range(per_i):
            out.append(torch.cat([log[i * per_i + j][k0:k0 + n] for i in range(B)], dim=1))
    return out


def _feed(model, draws, device):
    q = [d.to(device) for d in draws]

    def src(shape, kind, dev):
        e = q.pop(0)
        assert tuple(e.shape) == tuple(shape), (tuple(e.shape), tuple(shape))
        return e

    model.noise_source = src
    return q


def _close(got, ref, rtol, what):
    if isinstance(ref, dict):
        for k in ref:
            _close(got[k], ref[k], rtol, f"{what}.{k}")
        return
    if torch.is_tensor(ref):
        g = got.detach().float().cpu()
        assert g.shape == ref.shape, (what, g.shape, ref.shape)
        ref = ref.detach().float()
        fin = torch.isfinite(ref)
        if not bool(fin.all()):      # infinities (pruned cluster logits, unvisited cluster counts) must sit at the same places
            assert torch.equal(g[~fin], ref[~fin]), (what, "non-finite entries differ")
            g, ref = g[fin], ref[fin]
        err = float((g - ref).abs().max()) / max(float(ref.abs().max()), 1e-6)
        assert err <= rtol, (what, err)
    elif isinstance(ref, (bool, type(None))):
        assert got == ref, (what, got, ref)
    elif isinstance(ref, list):
        assert list(got) == list(ref), (what, got, ref)


def check_infer_case(name, device="cuda", rtol=2e-4):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    rec = torch.load(os.path.join(GOLD, f"infer_{name}.pt"), weights_only=False)
    spec = CASES[name]
    model = _model(name, rec, device)
    data, _ = make_data(spec)
    mods = list(spec["dims"])
    B = spec["B"]
    fresh = lambda: mb.MultimodalBaseDataset(data={k: v.to(device) for k, v in data.items()})  # noqa: E731
    fam = spec["model"]
    private = fam in ("mmvaeplus", "cmvae") or spec["cfg"].get("modalities_specific_dim") is not None
    done = []

    def run(key, fn, draws=None):
        c = rec["calls"][key]
        q = _feed(model, c["noise"] if draws is None else draws, device)
        np.random.seed(c["np_seed"])
        with torch.no_grad():
            out = fn()
        assert not q, (key, "noise left over", len(q))
        _close(dict(out) if hasattr(out, "keys") else out, c["out"], rtol, f"{name}.{key}")
        done.append(key)

    run("encode_mean", lambda: model.encode(fresh(), cond_mod="all", N=1, return_mean=True))
    run("encode_all_n3", lambda: model.encode(fresh(), cond_mod="all", N=3))
    run("encode_sub_n2_flat", lambda: model.encode(fresh(), cond_mod=[mods[0]], N=2, flatten=True))
    run("predict", lambda: model.predict(fresh(), cond_mod=[mods[0]], gen_mod="all", N=2, flatten=False))
    run("predict_all_to_one", lambda: model.predict(fresh(), cond_mod="all", gen_mod=mods[1]))
    c = rec["calls"]["joint_nll"]
    if fam in ("mmvaeplus", "cmvae"):
        # MMVAE+: the reference's popitem() quirk drops the last modality (mmvaePlus_model.py:497); CMVAE keeps all (cmvae_model.py:752)
        n_present = len(mods) - 1 if fam == "mmvaeplus" else len(mods)
        plan = _mmvaeplus_nll_plan(c["noise"], B, n_present * (2 + n_present - 1), NLL_K // len(mods), NLL_BK)
        run("joint_nll", lambda: model.compute_joint_nll(fresh(), K=NLL_K, batch_size_K=NLL_BK), plan)
    else:
        run("joint_nll", lambda: model.compute_joint_nll(fresh(), K=NLL_K, batch_size_K=NLL_BK))
    if "predict_clusters" in rec["calls"]:
        run("predict_clusters", lambda: model.predict_clusters(fresh()))
    if "predict_clusters_lliks" in rec["calls"]:
        run("predict_clusters_lliks", lambda: model.predict_clusters(fresh(), compute_lliks=True))
    if "prune_clusters" in rec["calls"]:
        def prune():
            m2 = copy.deepcopy(model)
            m2.noise_source = model.noise_source
            hv = m2.prune_clusters(mb.MultimodalBaseDataset(data=data), batch_size=4)
            return dict(h_values=torch.tensor([float(h) for h in hv]), n_clusters=torch.tensor(int(m2.n_clusters)), pc_params=m2._pc_params.detach().clone())
        run("prune_clusters", prune)
    if "joint_nll_paper" in rec["calls"]:
        run("joint_nll_paper", lambda: model.compute_joint_nll_paper(fresh(), K=30, batch_size_K=10))
    if "joint_nll_subset" in rec["calls"]:
        run("joint_nll_subset", lambda: model._compute_joint_nll_from_subset_encoding([mods[0], mods[2]], fresh(), K=40, batch_size_K=20))
    c = rec["calls"]["cond_nll"]
    pred = [mods[1], mods[2]] if len(mods) > 2 else [mods[1]]
    run("cond_nll", lambda: model.compute_cond_nll(fresh(), [mods[0]], pred, k_iwae=COND_K),
        _stack_calls(c["noise"], 1 + len(mods) if private else 1))
    assert set(done) == set(rec["calls"]), set(rec["calls"]) - set(done)
    return done
