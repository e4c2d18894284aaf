"""Generate tests/golden/*.pt by running the REAL reference (/root/reference/src through oracle/shim)
on seeded synthetic inputs with recorded sampling noise.  Build-container only.

    python oracle/make_golden.py            # writes tests/golden/elbo_<case>.pt and nets_<net>.pt
"""
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


def summarize_grads(model):
    g = {}
    for k, p in model.named_parameters():
        if p.grad is None:
            g[k] = None
            continue
        t = p.grad.detach()
        g[k] = dict(sum=float(t.double().sum()), abssum=float(t.double().abs().sum()),
                    head=t.flatten()[:8].clone(), full=t.clone() if t.numel() <= 600 else None)
    return g


def reference_archs(spec):
    """Encoders / decoders of the REAL reference for a spec's `arch` table (None, None = the model's own defaults)."""
    if "arch" not in spec:
        return None, None
    from multivae.models.base.base_config import BaseAEConfig
    from multivae.models.nn import default_architectures as da
    from multivae.models.nn import cub, mmnist, svhn
    L = spec["cfg"]["latent_dim"]
    Lw = spec["cfg"].get("modalities_specific_dim")
    enc, dec = {}, {}
    for m, a in spec["arch"].items():
        c = BaseAEConfig(input_dim=tuple(spec["dims"][m]), latent_dim=L)
        if a == "mlp":
            assert Lw is None
            enc[m], dec[m] = da.Encoder_VAE_MLP(c), da.Decoder_AE_MLP(c)
        elif a == "svhn":
            enc[m], dec[m] = svhn.Encoder_VAE_SVHN(c), svhn.Decoder_VAE_SVHN(c)
        elif a == "conv_mmnist":
            enc[m], dec[m] = mmnist.EncoderConvMMNIST_adapted(c), mmnist.DecoderConvMMNIST(c)
        elif a == "resnet_mmnist":
            enc[m], dec[m] = mmnist.EncoderResnetMMNIST(Lw or 0, L), mmnist.DecoderResnetMMNIST(L + (Lw or 0))
        else:
            raise ValueError(a)
    return enc, dec


def run_case(name, spec):
    ref_harness.import_reference()
    from multivae.data.datasets.base import IncompleteDataset, MultimodalBaseDataset
    from multivae.models import (CMVAE, CRMVAE, MMVAE, MVAE, MVTCAE, CMVAEConfig, CRMVAEConfig, MMVAEConfig, MMVAEPlus,
                                 MMVAEPlusConfig, MoPoE, MoPoEConfig, MVAEConfig, MVTCAEConfig)

    cls = {"mmvaeplus": (MMVAEPlus, MMVAEPlusConfig), "mmvae": (MMVAE, MMVAEConfig), "mvtcae": (MVTCAE, MVTCAEConfig),
           "mvae": (MVAE, MVAEConfig), "mopoe": (MoPoE, MoPoEConfig), "cmvae": (CMVAE, CMVAEConfig),
           "crmvae": (CRMVAE, CRMVAEConfig)}[spec["model"]]
    import copy
    cfg = cls[1](n_modalities=len(spec["dims"]), input_dims=dict(spec["dims"]), **copy.deepcopy(spec["cfg"]))
    enc, dec = reference_archs(spec)
    model = cls[0](cfg, enc, dec)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = synth_state_dict(shapes, seed=1)
    # non-trivial prior parameters so that their gradients are exercised
    model.load_state_dict(sd)
    model.train()
    data, masks = make_data(spec)
    ds = IncompleteDataset(data={k: v.clone() for k, v in data.items()}, masks={k: v.clone() for k, v in masks.items()}) \
        if masks is not None else MultimodalBaseDataset(data={k: v.clone() for k, v in data.items()})
    if "np_seed" in spec:
        np.random.seed(spec["np_seed"])
    torch.manual_seed(1234)  # only consumed by OneHotCategorical in masked MoPoE
    q = ref_harness.NoiseQueue(record=True, generator=torch.Generator().manual_seed(2000))
    extra = {}
    with ref_harness.injected_noise(q):
        if spec["model"] == "mopoe" and masks is not None:
            # capture the random mixture choice (mopoe_model.py:417-433)
            orig = model.random_mixture_component_selection

            def patched(mus, logvars, avail):
                import torch.distributions as dist
                choice = dist.OneHotCategorical(probs=avail.permute(1, 0)).sample()
                extra["choice"] = choice.argmax(-1).to(torch.int32)
                mus_ = mus.permute(1, 0, 2)[choice.bool()]
                lv_ = logvars.permute(1, 0, 2)[choice.bool()]
                return mus_, lv_

            model.random_mixture_component_selection = patched
        out = model(ds, **spec.get("fwd", {}))
        out.loss.backward()
    rec = dict(case=name, spec=spec, loss=out.loss.detach().clone(), loss_sum=out.loss_sum.detach().clone(),
               metrics={k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in out.metrics.items()},
               noise=[e.clone() for e in q.log], grads=summarize_grads(model), state_shapes=shapes, sd_seed=1, **extra)
    if spec["model"] == "mvae":
        rec["subsets"] = [list(s) for s in model.subsets]
    if spec["model"] == "mopoe":
        rec["subset_keys"] = list(model.subsets.keys())
    # per-term tensors for the MoE models: replay with the same noise
    if spec["model"] in ("mmvaeplus", "mmvae", "cmvae"):
        model.zero_grad()
        q2 = ref_harness.NoiseQueue(draws=[e.clone() for e in q.log])
        with ref_harness.injected_noise(q2), torch.no_grad():
            if spec["model"] == "mmvaeplus":
                emb, post, recs = model._compute_posteriors_and_embeddings(ds, detach=cfg.loss == "dreg_looser")
                lws, _ = model._compute_k_lws(post, emb, recs, ds)
            elif spec["model"] == "cmvae":
                post, emb, recs = model._compute_posteriors_and_embeddings(ds, detach=cfg.loss == "dreg_looser")
                lws, _, _ = model._compute_k_lws(post, emb, recs, ds)
            else:
                o = model(ds, detailed_output=True, compute_loss=False)
                qz = o["qz_xs_detach"] if cfg.loss == "dreg_looser" else o["qz_xs"]
                lws, _ = model.compute_k_lws(qz, o["zss"], o["recon"], ds)
        rec["lws"] = {k: v.clone() for k, v in lws.items()}
    torch.save(rec, os.path.join(OUT, f"elbo_{name}.pt"))
    print(f"{name}: loss={float(out.loss.detach()):.6f} noise_draws={len(q.log)}")


def run_nets():
    ref_harness.import_reference()
    from multivae.models.base.base_config import BaseAEConfig
    from multivae.models.nn import default_architectures as da
    from multivae.models.nn import cub, mmnist, svhn

    def cfgd(input_dim, latent_dim, style_dim=None):
        c = BaseAEConfig(input_dim=input_dim, latent_dim=latent_dim)
        if style_dim is not None:
            c.style_dim = style_dim
        return c

    nets = {
        "enc_resnet_mmnist": (lambda: mmnist.EncoderResnetMMNIST(32, 32), (2, 3, 28, 28), "x"),
        "dec_resnet_mmnist": (lambda: mmnist.DecoderResnetMMNIST(64), (3, 64), "z"),
        "enc_conv_mmnist": (lambda: mmnist.EncoderConvMMNIST_adapted(cfgd((3, 28, 28), 64)), (2, 3, 28, 28), "x"),
        "dec_conv_mmnist": (lambda: mmnist.DecoderConvMMNIST(cfgd((3, 28, 28), 64)), (3, 64), "z"),
        "enc_svhn": (lambda: svhn.Encoder_VAE_SVHN(cfgd((3, 32, 32), 20)), (2, 3, 32, 32), "x"),
        "dec_svhn": (lambda: svhn.Decoder_VAE_SVHN(cfgd((3, 32, 32), 20)), (3, 20), "z"),
        "enc_mlp": (lambda: da.Encoder_VAE_MLP(cfgd((1, 28, 28), 20)), (2, 1, 28, 28), "x"),
        "enc_mlp_style": (lambda: da.Encoder_VAE_MLP_Style(cfgd((3, 8, 8), 8, 4)), (2, 3, 8, 8), "x"),
        "dec_mlp": (lambda: da.Decoder_AE_MLP(cfgd((1, 28, 28), 20)), (3, 20), "z"),
        "enc_cub_resnet": (lambda: cub.CUB_Resnet_Encoder(latent_dim=32), (2, 3, 64, 64), "x"),
        "dec_cub_resnet": (lambda: cub.CUB_Resnet_Decoder(latent_dim=32), (2, 32), "z"),
    }
    only = [a[5:] for a in sys.argv[1:] if a.startswith("nets:")]
    if only:
        nets = {k: v for k, v in nets.items() if k in only}
    for name, (ctor, shp, kind) in nets.items():
        net = ctor()
        shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        net.load_state_dict(synth_state_dict(shapes, seed=2))
        g = torch.Generator().manual_seed(77)
        x = torch.rand(shp, generator=g) if kind == "x" else torch.randn(shp, generator=g)
        x.requires_grad_(True)
        out = net(x)
        outs = {k: v for k, v in out.items()}
        # scalar probe for gradients: sum(out * fixed random cotangent)
        cot = {k: torch.randn(v.shape, generator=g) for k, v in outs.items()}
        sum((outs[k] * cot[k]).sum() for k in outs).backward()
        rec = dict(net=name, in_shape=shp, kind=kind, state_shapes=shapes, sd_seed=2,
                   outputs={k: v.detach().clone() for k, v in outs.items()}, cot=cot,
                   grad_in=x.grad.clone(), grads=summarize_grads(net))
        torch.save(rec, os.path.join(OUT, f"nets_{name}.pt"))
        print(f"net {name}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in outs.items()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or list(CASES) + ["nets"]
    for n in which:
        if n == "nets" or n.startswith("nets:"):
            if n == "nets" or n == [a for a in which if a.startswith("nets:")][0]:
                run_nets()
        else:
            run_case(n, CASES[n])
