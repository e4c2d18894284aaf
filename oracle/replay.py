"""Replay a golden case (oracle/cases.py) through the oracle PORT.  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import torch

from .cases import make_data
from .port import elbo as E
from .port import nets as N
from .port.nets import synth_state_dict


def rescale_factors(spec):
    """models/base/base_ae_model.py:127-152."""
    dims = spec["dims"]
    if spec["cfg"].get("uses_likelihood_rescaling"):
        mx = max(int(np.prod(d)) for d in dims.values())
        return {m: mx / int(np.prod(d)) for m, d in dims.items()}
    return {m: 1 for m in dims}


def dist_tables(spec):
    dd = spec["cfg"].get("decoders_dist") or {m: "normal" for m in spec["dims"]}
    dp = spec["cfg"].get("decoder_dist_params") or {}
    return dd, {m: dp.get(m, {}).get("scale", 1.0) for m in spec["dims"]}


def default_nets(spec, p, dtype=torch.float32):
    """Closures over a parameter dict `p` for the reference's default MLP architectures."""
    mods = list(spec["dims"])
    if "arch" in spec:
        enc, dec = {}, {}
        for m in mods:
            a, e, d = spec["arch"][m], f"encoders.{m}.", f"decoders.{m}."
            if a == "mlp":
                enc[m] = lambda x, e=e: N.encoder_vae_mlp(p, e, x)
                dec[m] = lambda z, d=d, m=m: N.decoder_ae_mlp(p, d, z, spec["dims"][m])
            elif a == "svhn":
                enc[m] = lambda x, e=e: N.encoder_vae_svhn(p, e, x)
                dec[m] = lambda z, d=d: N.decoder_vae_svhn(p, d, z)
            elif a == "conv_mmnist":
                enc[m] = lambda x, e=e: N.encoder_conv_mmnist_adapted(p, e, x)
                dec[m] = lambda z, d=d: N.decoder_conv_mmnist(p, d, z)
            elif a == "resnet_mmnist":
                enc[m] = lambda x, e=e: N.encoder_resnet_mmnist(p, e, x)
                dec[m] = lambda z, d=d: N.decoder_resnet_mmnist(p, d, z)
            else:
                raise ValueError(a)
        return enc, dec
    if spec["model"] in ("mmvaeplus", "cmvae") or spec["cfg"].get("modalities_specific_dim") is not None:
        enc = {m: (lambda x, m=m: N.encoder_vae_mlp_style(p, f"encoders.{m}.", x)) for m in mods}
    else:
        enc = {m: (lambda x, m=m: N.encoder_vae_mlp(p, f"encoders.{m}.", x)) for m in mods}
    dec = {m: (lambda z, m=m: N.decoder_ae_mlp(p, f"decoders.{m}.", z, spec["dims"][m])) for m in mods}
    return enc, dec


def split_noise(spec, noise, mods_active=None):
    """Map the reference's FIFO consumption order onto the port's named noise dicts."""
    mods = mods_active or list(spec["dims"])
    it = iter(noise)
    model = spec["model"]
    if model in ("mmvaeplus", "cmvae"):
        out = {"u": {}, "w": {}, "prior": {}}
        for c in mods:
            out["u"][c] = next(it)
            out["w"][c] = next(it)
            out["prior"][c] = {}
            for r in mods:
                if r != c:
                    out["prior"][c][r] = next(it)
        return out
    if model == "mmvae":
        return {"z": {c: next(it) for c in mods}}
    if model in ("mvae", "crmvae"):
        return {"z": list(noise)}
    if model == "mopoe" and spec["cfg"].get("modalities_specific_dim") is not None:
        return {"z": next(it), "style": {m: next(it) for m in spec["dims"]}}   # shared, then one style draw per modality
    return {"z": next(it)}


def run_port(spec, rec, dtype=torch.float32, want_grads=True, details=None):
    """Returns (loss, loss_sum, metrics, params-with-grads)."""
    sd = synth_state_dict(rec["state_shapes"], seed=rec["sd_seed"])
    p = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    for k in p:  # frozen prior / cluster-scale parameters (requires_grad=False in the reference model)
        if ("prior" in k or "logvar_clusters" in k) and rec["grads"].get(k, 1) is None:
            p[k].requires_grad_(False)
    loss, loss_sum, metrics = _forward(spec, rec, p, dtype, details)
    if want_grads:
        loss.backward()
    return loss, loss_sum, metrics, p


def run_port_with_params(spec, rec, p, dtype=torch.float32):
    """One forward pass of the port on the caller's parameter dict (bench.py's CPU baseline keeps the parameters and an optimizer
    across steps); rec needs "noise" only.  Returns the loss."""
    return _forward(spec, rec, p, dtype, None)[0]


def _forward(spec, rec, p, dtype, details):
    data, masks = make_data(spec)
    data = {k: v.to(dtype) for k, v in data.items()}
    noise_l = [e.to(dtype) for e in rec["noise"]]
    enc, dec = default_nets(spec, p)
    dd, ds = dist_tables(spec)
    rs = rescale_factors(spec)
    cfg = spec["cfg"]
    model = spec["model"]
    mods = list(spec["dims"])
    active = [m for m in mods if masks is None or bool(masks[m].any())]
    noise = split_noise(spec, noise_l, active)
    metrics = {}
    if model == "mmvaeplus":
        pm = {m: p[f"mean_priors.{m}"] for m in mods + ["shared"]}
        pl = {m: p[f"logvars_priors.{m}"] for m in mods + ["shared"]}
        loss = E.mmvae_plus_forward(enc, dec, data, noise, K=cfg["K"], latent_dim=cfg["latent_dim"],
                                    style_dim=cfg["modalities_specific_dim"], beta=cfg["beta"],
                                    kind=cfg["prior_and_posterior_dist"], loss=cfg["loss"], dec_dist=dd, dec_scale=ds,
                                    rescale=rs, prior_mean=pm, prior_logvar=pl, masks=masks, details=details)
        loss_sum = loss
    elif model == "cmvae":
        n_c = cfg.get("number_of_clusters", 10)
        cm = dict(w_mean=p["w_mean_prior"], w_logvar=p["w_logvar_prior"], pc_logits=p["_pc_params"],
                  means=[p[f"mean_clusters.{i}"] for i in range(n_c)], logvars=[p[f"logvar_clusters.{i}"] for i in range(n_c)])
        loss = E.mmvae_plus_forward(enc, dec, data, noise, K=cfg["K"], latent_dim=cfg["latent_dim"],
                                    style_dim=cfg["modalities_specific_dim"], beta=cfg.get("beta", 1.0),
                                    kind=cfg["prior_and_posterior_dist"], loss=cfg["loss"], dec_dist=dd, dec_scale=ds,
                                    rescale=rs, prior_mean={m: p[f"r_mean_priors.{m}"] for m in mods},
                                    prior_logvar={m: p[f"r_logvars_priors.{m}"] for m in mods}, masks=masks, details=details, cmvae=cm)
        loss_sum = loss
    elif model == "crmvae":
        loss, loss_sum, metrics = E.crmvae_forward(enc, dec, data, noise, beta=cfg.get("beta", 2.5), dec_dist=dd, dec_scale=ds,
                                                   rescale=rs, masks=masks)
    elif model == "mmvae":
        loss = E.mmvae_forward(enc, dec, data, noise, K=cfg["K"], beta=cfg.get("beta", 1.0),
                               kind=cfg["prior_and_posterior_dist"], loss=cfg["loss"], dec_dist=dd, dec_scale=ds,
                               rescale=rs, prior_mean=p["prior_mean"], prior_logvar=p["prior_log_var"], masks=masks,
                               details=details)
        loss_sum = loss
    elif model == "mvtcae":
        loss, loss_sum, metrics = E.mvtcae_forward(enc, dec, data, noise, alpha=cfg["alpha"], beta=cfg["beta"],
                                                   dec_dist=dd, dec_scale=ds, rescale=rs, masks=masks)
    elif model == "mvae":
        fwd = spec.get("fwd", {})
        beta = E.mvae_beta(fwd.get("epoch", 1), fwd.get("batch_ratio", 0), cfg["warmup"], cfg["beta"])
        subsets = [list(mods)] + [[m] for m in mods]
        k = cfg.get("k", 0) if len(mods) > 2 else 0
        if k > 0:
            np.random.seed(spec["np_seed"])
            idx = np.random.choice(np.arange(len(rec["subsets"])), size=k, replace=False)
            subsets += [rec["subsets"][i] for i in idx]
        loss, loss_sum, metrics = E.mvae_forward(enc, dec, data, noise, subsets=subsets, beta=beta, dec_dist=dd,
                                                 dec_scale=ds, rescale=rs, masks=masks)
    elif model == "mopoe":
        loss, loss_sum, metrics = E.mopoe_forward(enc, dec, data, noise, latent_dim=cfg["latent_dim"], beta=cfg["beta"],
                                                  dec_dist=dd, dec_scale=ds, rescale=rs, masks=masks,
                                                  choice=rec.get("choice"), style=cfg.get("modalities_specific_dim") is not None,
                                                  beta_style=cfg.get("beta_style", 1.0))
    else:
        raise ValueError(model)
    return loss, loss_sum, metrics
