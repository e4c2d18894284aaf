"""Golden vectors of the INFERENCE path (encode / predict / compute_joint_nll[_paper] / compute_cond_nll) from the REAL reference
(/root/reference/src through oracle/shim) with recorded sampling noise and fixed numpy seeds.  Build-container only.

    python oracle/make_golden_infer.py     # writes tests/golden/infer_<case>.pt
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402
from oracle.cases import CASES, make_data  # noqa: E402
from oracle.port.nets import synth_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
INFER_CASES = ["mvtcae", "mvae", "mopoe", "mopoe_private", "mmvae_dreg", "mmvae_iwae", "mmvaeplus_dreg", "mmvaeplus_normal",
               "cmvae_dreg", "crmvae"]
NLL_K, NLL_BK, COND_K = 100, 20, 6


def _cpu(o):
    if torch.is_tensor(o):
        return o.detach().clone()
    if isinstance(o, dict):
        return {k: _cpu(v) for k, v in o.items()}
    return o


def run_case(name):
    ref_harness.import_reference()
    from multivae.data.datasets.base import MultimodalBaseDataset
    from multivae.models import (CMVAE, CRMVAE, MMVAE, MVAE, MVTCAE, CMVAEConfig, CRMVAEConfig, MMVAEConfig, MMVAEPlus, MMVAEPlusConfig,
                                 MoPoE, MoPoEConfig, MVAEConfig, MVTCAEConfig)
    spec = CASES[name]
    cls = {"mmvaeplus": (MMVAEPlus, MMVAEPlusConfig), "mmvae": (MMVAE, MMVAEConfig), "mvtcae": (MVTCAE, MVTCAEConfig),
           "mvae": (MVAE, MVAEConfig), "mopoe": (MoPoE, MoPoEConfig), "cmvae": (CMVAE, CMVAEConfig),
           "crmvae": (CRMVAE, CRMVAEConfig)}[spec["model"]]
    cfg = cls[1](n_modalities=len(spec["dims"]), input_dims=dict(spec["dims"]), **copy.deepcopy(spec["cfg"]))
    model = cls[0](cfg)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synth_state_dict(shapes, seed=1))
    model.eval()
    data, _ = make_data(spec)
    mods = list(spec["dims"])
    fresh = lambda: MultimodalBaseDataset(data={k: v.clone() for k, v in data.items()})  # noqa: E731
    rec = dict(case=name, state_shapes=shapes, sd_seed=1, calls={})

    def record(key, fn, np_seed=5):
        q = ref_harness.NoiseQueue(record=True, generator=torch.Generator().manual_seed(3000 + len(rec["calls"])))
        np.random.seed(np_seed)
        with ref_harness.injected_noise(q), torch.no_grad():
            out = fn()
        rec["calls"][key] = dict(out=_cpu(dict(out) if hasattr(out, "keys") else out), noise=[e.clone() for e in q.log], np_seed=np_seed)
        print(f"  {name}.{key}: {len(q.log)} draws")

    record("encode_mean", lambda: model.encode(fresh(), cond_mod="all", N=1, return_mean=True))
    record("encode_all_n3", lambda: model.encode(fresh(), cond_mod="all", N=3))
    record("encode_sub_n2_flat", lambda: model.encode(fresh(), cond_mod=[mods[0]], N=2, flatten=True))
    record("predict", lambda: model.predict(fresh(), cond_mod=[mods[0]], gen_mod="all", N=2, flatten=False))
    record("predict_all_to_one", lambda: model.predict(fresh(), cond_mod="all", gen_mod=mods[1]))
    record("joint_nll", lambda: model.compute_joint_nll(fresh(), K=NLL_K, batch_size_K=NLL_BK))
    if hasattr(model, "compute_joint_nll_paper"):
        record("joint_nll_paper", lambda: model.compute_joint_nll_paper(fresh(), K=30, batch_size_K=10))
    if spec["model"] == "cmvae":
        record("predict_clusters", lambda: model.predict_clusters(fresh()))
        record("predict_clusters_lliks", lambda: model.predict_clusters(fresh(), compute_lliks=True))
    if spec["model"] == "mopoe":
        record("joint_nll_subset", lambda: model._compute_joint_nll_from_subset_encoding([mods[0], mods[2]], fresh(), K=40, batch_size_K=20))
    if spec["model"] == "cmvae":
        def prune():
            m2 = copy.deepcopy(model)
            hv = m2.prune_clusters(fresh(), batch_size=4)
            return dict(h_values=torch.tensor([float(h) for h in hv]), n_clusters=torch.tensor(int(m2.n_clusters)), pc_params=m2._pc_params.detach().clone())
        record("prune_clusters", prune)
    record("cond_nll", lambda: model.compute_cond_nll(fresh(), [mods[0]], [mods[1], mods[2]] if len(mods) > 2 else [mods[1]], k_iwae=COND_K))
    torch.save(rec, os.path.join(OUT, f"infer_{name}.pt"))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for n in sys.argv[1:] or INFER_CASES:
        print(n)
        run_case(n)
