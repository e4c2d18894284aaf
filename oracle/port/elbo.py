"""CPU restatement of the posterior-aggregation + ELBO path of MultiVae (oracle; test infra only).

Plain PyTorch on CPU (fp32 by default, fp64 when the inputs are fp64).  Sampling noise is an explicit
argument (standard Laplace / Normal draws `e`, so that z = loc + scale * e), which is exactly what
torch.distributions' rsample computes (torch/distributions/laplace.py: loc - scale*sign(u)*log1p(-|u|);
normal.py: loc + eps*scale).  All paths below are relative to /root/reference/src/multivae/.
"""
import math
from itertools import chain, combinations

import torch
import torch.nn.functional as F

LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------------
# per-element decoder log-probabilities        models/base/base_utils.py:62-87
# --------------------------------------------------------------------------------------------
def recon_log_prob(dist, recon, target, scale=1.0):
    if dist == "normal":  # dist.Normal(recon, scale).log_prob(target)
        var = scale * scale
        return -((target - recon) ** 2) / (2 * var) - math.log(scale) - 0.5 * LOG_2PI
    if dist == "laplace":  # dist.Laplace(recon, scale).log_prob(target)
        return -math.log(2 * scale) - torch.abs(target - recon) / scale
    if dist == "bernoulli":  # dist.Bernoulli(logits=recon).log_prob(target)
        return -F.binary_cross_entropy_with_logits(recon, target.expand_as(recon), reduction="none")
    if dist == "categorical":  # base_utils.py:28-39
        return target * F.log_softmax(recon + 1e-6, dim=-1)
    raise ValueError(dist)


def laplace_log_prob(x, loc, scale):
    return -torch.log(2 * scale) - torch.abs(x - loc) / scale


def normal_log_prob(x, loc, scale):
    return -((x - loc) ** 2) / (2 * scale**2) - torch.log(scale) - 0.5 * LOG_2PI


def log_var_to_std(log_var, kind):
    """models/mmvaePlus/mmvaePlus_model.py:113-123, models/mmvae/mmvae_model.py:66-74."""
    if kind == "laplace_with_softmax":
        return F.softmax(log_var, dim=-1) * log_var.size(-1) + 1e-6
    if kind == "normal_with_softplus":
        return F.softplus(log_var) + 1e-6
    return torch.exp(0.5 * log_var)


def latent_log_prob(kind, x, loc, scale):
    return laplace_log_prob(x, loc, scale) if kind == "laplace_with_softmax" else normal_log_prob(x, loc, scale)


# --------------------------------------------------------------------------------------------
# PoE                                          models/base/base_utils.py:122-147
# --------------------------------------------------------------------------------------------
def poe(mus, logvars, eps=1e-8):
    var = torch.exp(logvars) + eps
    T = 1.0 / var
    pd_mu = torch.sum(mus * T, dim=0) / torch.sum(T, dim=0)
    pd_var = 1.0 / torch.sum(T, dim=0)
    return pd_mu, torch.log(pd_var)


def stable_poe(mus, logvars):
    if len(mus) == 1:
        return mus[0], logvars[0]
    ln_inv = -logvars
    ln_var = -torch.logsumexp(ln_inv, dim=0)
    mu = (torch.exp(ln_inv) * mus).sum(dim=0) * torch.exp(ln_var)
    return mu, ln_var


def kl_std_normal(mu, logvar):
    """-0.5 * (1 + lv - mu^2 - e^lv), elementwise (mvae_model.py:105, mvtcae_model.py:52, mopoe_model.py:126)."""
    return -0.5 * (1 - logvar.exp() - mu.pow(2) + logvar)


# --------------------------------------------------------------------------------------------
# MoPoE subset table + deterministic mixture selection   models/mopoe/mopoe_model.py:76-106,435-465
# --------------------------------------------------------------------------------------------
def mopoe_subsets(mod_names):
    """Insertion-ordered dict key->sorted list, the "" (empty) subset first, like the reference."""
    xs = list(mod_names)
    out = {}
    for names in chain.from_iterable(combinations(xs, n) for n in range(len(xs) + 1)):
        out["_".join(sorted(names))] = sorted(names)
    return out


def mopoe_subset_bitmasks(mod_names):
    """uint32 bitmask per non-empty subset (bit i = modality i in encoder order), reference order."""
    xs = list(mod_names)
    masks = []
    for key, mods in mopoe_subsets(xs).items():
        if key == "":
            continue
        masks.append(sum(1 << xs.index(m) for m in mods))
    return masks


def mopoe_sample_to_subset(num_samples, num_subsets):
    """Per-sample subset index; reproduces deterministic_mixture_component_selection exactly,
    including float32(1/S) and torch.floor on a float32 product."""
    w = (1 / float(num_subsets)) * torch.ones(num_subsets)  # float32, as in mopoe_model.py:337
    idx_start, idx_end = [], []
    for k in range(num_subsets):
        i_start = 0 if k == 0 else int(idx_end[k - 1])
        if k == num_subsets - 1:
            i_end = num_samples
        else:
            i_end = i_start + int(torch.floor(num_samples * w[k]))
        idx_start.append(i_start)
        idx_end.append(i_end)
    idx_end[-1] = num_samples
    out = torch.empty(num_samples, dtype=torch.int32)
    for k in range(num_subsets):
        out[idx_start[k] : idx_end[k]] = k
    return out


# --------------------------------------------------------------------------------------------
# MMVAE+ / MMVAE                               models/mmvaePlus/mmvaePlus_model.py:125-363
#                                              models/mmvae/mmvae_model.py:95-292
# --------------------------------------------------------------------------------------------
def _moe_terms(kind, z_u, post_u, masks, mods):
    """lq(u|X) = logsumexp_m log q_m(u) - log n_mods   (mmvaePlus_model.py:255-271)."""
    lqs = []
    for m in mods:
        q = latent_log_prob(kind, z_u, post_u[m][0], post_u[m][1]).sum(-1)  # (K,B)
        if masks is not None:
            q = q.masked_fill(torch.stack([masks[m] == False] * len(z_u)), -torch.inf)  # noqa: E712
        lqs.append(q)
    return torch.stack(lqs)


def mmvae_plus_forward(
    enc, dec, data, noise, *, K, latent_dim, style_dim, beta, kind, loss, dec_dist, dec_scale,
    rescale, prior_mean, prior_logvar, masks=None, details=None, cmvae=None,
):
    """One MMVAE+ forward.  enc[m](x)->(mu,lv,mu_w,lv_w); dec[m](z)->recon.
    prior_mean/prior_logvar: dict with per-modality (1,Lw) entries and "shared" (1,L+Lw).
    noise: dict  noise["u"][c], noise["w"][c] : (K,B,.) ; noise["prior"][c][r] : (K,B,Lw) for r != c.
    Returns the scalar loss (sum over batch), following mmvaePlus_model.py:200-363 line by line."""
    mods = [m for m in data if masks is None or bool(torch.any(masks[m]))]  # drop_unused_modalities
    detach = loss == "dreg_looser"
    emb, post, recons = {}, {}, {}
    for c in mods:
        mu, lv, mu_w, lv_w = enc[c](data[c])
        sig, sig_w = log_var_to_std(lv, kind), log_var_to_std(lv_w, kind)
        u = mu + sig * noise["u"][c]
        w = mu_w + sig_w * noise["w"][c]
        if detach:
            post[c] = {"u": (mu.detach(), sig.detach()), "w": (mu_w.detach(), sig_w.detach())}
        else:
            post[c] = {"u": (mu, sig), "w": (mu_w, sig_w)}
        recons[c] = {}
        for r in mods:
            if r == c:
                z = torch.cat([u, w], dim=-1)
            else:
                sig_p = log_var_to_std(prior_logvar[r], kind)
                wt = prior_mean[r] + sig_p * noise["prior"][c][r]
                z = torch.cat([u, wt], dim=-1)
            rec = dec[r](z.reshape(-1, z.shape[-1]))
            recons[c][r] = rec.reshape(*z.shape[:-1], *rec.shape[1:])
        emb[c] = {"u": u, "w": w}

    if masks is not None:
        n_mods = torch.sum(torch.stack(tuple(masks[m] for m in mods)).int(), dim=0)
    else:
        n_mods = torch.tensor([len(mods)])
    if cmvae is None:
        pz_mean, pz_std = prior_mean["shared"], log_var_to_std(prior_logvar["shared"], kind)
    lws = {}
    for c in mods:
        u, w = emb[c]["u"], emb[c]["w"]
        z = torch.cat([u, w], dim=-1)
        lpz = latent_log_prob(kind, z, pz_mean, pz_std).sum(-1) if cmvae is None else None
        lqu = torch.logsumexp(_moe_terms(kind, u, {m: post[m]["u"] for m in mods}, masks, mods), dim=0) - torch.log(n_mods)
        lqw = latent_log_prob(kind, w, *post[c]["w"]).sum(-1)
        lpx = 0
        for r in mods:
            rec = recons[c][r]
            Kk, B = rec.shape[0], rec.shape[1]
            t = recon_log_prob(dec_dist[r], rec, data[r], dec_scale[r]).view(Kk, B, -1).mul(rescale[r]).sum(-1)
            if masks is not None:
                t = t * masks[r].float()
            lpx = lpx + t
        if cmvae is None:
            lw = lpx + beta * (lpz - lqu - lqw)
        else:
            # CMVAE (models/cmvae/cmvae_model.py:263-345): fixed prior p(w), mixture-of-clusters prior over u with the explicit
            # expectation over q(c | u)
            lpw = latent_log_prob(kind, w, cmvae["w_mean"], log_var_to_std(cmvae["w_logvar"], kind)).sum(-1)
            n_c = len(cmvae["means"])
            lpc = torch.log(F.softmax(cmvae["pc_logits"], dim=-1))
            lpzc = torch.stack([latent_log_prob(kind, u, cmvae["means"][i], log_var_to_std(cmvae["logvars"][i], kind))
                                for i in range(n_c)], dim=0).sum(-1)
            qzc = torch.softmax(lpc.view(n_c, 1, 1) + lpzc, dim=0) + 1e-20
            lw = 0
            for ci, q_c in enumerate(qzc):
                lw = lw + q_c * (lpx + beta * (lpc[ci] + lpzc[ci] + lpw - lqu - lqw - q_c.log()))
        if masks is not None:
            lw = lw * masks[c].float()
        lws[c] = lw
        if details is not None:
            details[c] = {"lpz": lpz, "lqu_x": lqu, "lqw_x": lqw, "lpx_z": lpx, "lw": lw}

    if loss == "dreg_looser":
        wk = {}
        with torch.no_grad():
            for m, lw in lws.items():
                wk[m] = (lw - torch.logsumexp(lw, 0, keepdim=True)).exp()
        tot = torch.stack([lws[m] * wk[m] for m in lws], dim=0).sum(1)
        tot = tot.sum(0) / n_mods
        for m in emb:
            if emb[m]["w"].requires_grad:
                emb[m]["w"].register_hook(lambda g, w_=wk[m]: w_.unsqueeze(-1) * g)
            if emb[m]["u"].requires_grad:
                emb[m]["u"].register_hook(lambda g, w_=wk[m]: w_.unsqueeze(-1) * g)
        return -tot.sum()
    if loss == "iwae_looser":
        t = torch.stack(list(lws.values()), dim=0)
        t = torch.logsumexp(t, dim=1) - math.log(t.size(1))
        t = t.sum(0) / n_mods
        return -t.sum()
    raise NotImplementedError(loss)


def mmvae_forward(
    enc, dec, data, noise, *, K, beta, kind, loss, dec_dist, dec_scale, rescale, prior_mean,
    prior_logvar, masks=None, details=None,
):
    """MMVAE (single latent).  enc[m](x)->(mu,lv).  noise["z"][c]: (K,B,L).  mmvae_model.py:95-292
    (beta is accepted but unused, like the reference: lw = lpx_z + lpz - lqz_x, :228)."""
    mods = [m for m in data if masks is None or bool(torch.any(masks[m]))]
    emb, post, post_det, recons = {}, {}, {}, {}
    for c in mods:
        mu, lv = enc[c](data[c])
        sig = log_var_to_std(lv, kind)
        z = mu + sig * noise["z"][c]
        post[c], post_det[c] = (mu, sig), (mu.detach(), sig.detach())
        recons[c] = {}
        for r in mods:
            rec = dec[r](z.reshape(-1, z.shape[-1]))
            recons[c][r] = rec.reshape(*z.shape[:-1], *rec.shape[1:])
        emb[c] = z
    qs = post_det if loss == "dreg_looser" else post
    if masks is not None:
        n_mods = torch.sum(torch.stack(tuple(masks[m] for m in mods)).int(), dim=0)
    else:
        n_mods = torch.tensor([len(mods)])
    pz_std = log_var_to_std(prior_logvar, kind) if kind == "laplace_with_softmax" else torch.exp(0.5 * prior_logvar)
    lws = {}
    for c in mods:
        z = emb[c]
        lpz = latent_log_prob(kind, z, prior_mean, pz_std).sum(-1)
        lqz = torch.logsumexp(_moe_terms(kind, z, qs, masks, mods), dim=0) - torch.log(n_mods)
        lpx = 0
        for r in mods:
            rec = recons[c][r]
            Kk, B = rec.shape[0], rec.shape[1]
            t = recon_log_prob(dec_dist[r], rec, data[r], dec_scale[r]).view(Kk, B, -1).mul(rescale[r]).sum(-1)
            if masks is not None:
                t = t * masks[r].float()
            lpx = lpx + t
        lw = lpx + lpz - lqz
        if masks is not None:
            lw = lw * masks[c].float()
        lws[c] = lw
        if details is not None:
            details[c] = {"lpz": lpz, "lqz_x": lqz, "lpx_z": lpx, "lw": lw}
    if loss == "dreg_looser":
        wk = {}
        with torch.no_grad():
            for m, lw in lws.items():
                wk[m] = (lw - torch.logsumexp(lw, 0, keepdim=True)).exp()
        tot = torch.stack([lws[m] * wk[m] for m in emb], dim=0).sum(1)
        for m in emb:
            if emb[m].requires_grad:
                emb[m].register_hook(lambda g, w_=wk[m]: w_.unsqueeze(-1) * g)
        tot = tot.sum(0) / n_mods
        return -tot.sum()
    if loss == "iwae_looser":
        t = torch.stack(list(lws.values()), dim=0)
        t = torch.logsumexp(t, dim=1) - math.log(t.size(1))
        t = t.sum(0) / n_mods
        return -t.sum()
    raise NotImplementedError(loss)


# --------------------------------------------------------------------------------------------
# MVTCAE                                       models/mvtcae/mvtcae_model.py:42-169
# --------------------------------------------------------------------------------------------
def mvtcae_forward(enc, dec, data, noise, *, alpha, beta, dec_dist, dec_scale, rescale, masks=None):
    """noise["z"]: (B,L).  Returns (loss, loss_sum, metrics)."""
    mods = list(data.keys())
    outs = {}
    for m in mods:
        mu, lv = enc[m](data[m])
        if masks is not None:
            lv = lv.masked_fill(~masks[m].bool().unsqueeze(-1), torch.inf)
        outs[m] = (mu, lv)
    mus = torch.stack([outs[m][0] for m in mods])
    lvs = torch.stack([outs[m][1] for m in mods])
    jmu, jlv = poe(mus, lvs)
    z = jmu + torch.exp(0.5 * jlv) * noise["z"]
    ndata = len(z)
    res = {}
    joint_kld = kl_std_normal(jmu, jlv).sum()
    res["joint_divergence"] = joint_kld
    loss_rec = 0
    for m in mods:
        rec = dec[m](z)
        t = (-recon_log_prob(dec_dist[m], rec, data[m], dec_scale[m]) * rescale[m]).reshape(rec.size(0), -1).sum(-1)
        if masks is not None:
            t = masks[m].float() * t
        res[m] = t.sum()
        loss_rec = loss_rec + t.sum()
    kld_losses = 0.0
    for m in mods:
        mu, lv = outs[m]
        k = -0.5 * (1 - jlv.exp() / lv.exp() - (jmu - mu).pow(2) / lv.exp() + jlv - lv).reshape(mu.size(0), -1).sum(-1)
        if masks is not None:
            k = k.masked_fill((1 - masks[m].int()).bool(), 0)
        res["kld_" + m] = k.sum()
        kld_losses = kld_losses + k.sum()
    M = len(mods)
    rec_w, cvib_w, vib_w = (M - alpha) / M, alpha / M, 1 - alpha
    total = rec_w * loss_rec + beta * (cvib_w * kld_losses + vib_w * joint_kld)
    return total / ndata, total, res


# --------------------------------------------------------------------------------------------
# CRMVAE                                       models/crmvae/crmvae_model.py:44-180
# --------------------------------------------------------------------------------------------
def kl_gauss(mean, log_var, prior_mean, prior_log_var):
    """models/base/base_utils.py:90-119."""
    kl = 0.5 * (prior_log_var - log_var + torch.exp(log_var - prior_log_var) + ((mean - prior_mean) ** 2) / torch.exp(prior_log_var) - 1)
    return kl.sum(dim=-1)


def crmvae_forward(enc, dec, data, noise, *, beta, dec_dist, dec_scale, rescale, masks=None):
    """noise["z"]: list of (B,L): the joint draw, then one per modality.  Returns (loss, loss_sum, metrics)."""
    mods = list(data.keys())
    outs, masked = {}, {}
    for m in mods:
        mu, lv = enc[m](data[m])
        outs[m] = (mu, lv)
        lvm = lv.clone()
        if masks is not None:
            lvm = lvm.masked_fill((1 - masks[m].int()).bool().unsqueeze(-1), torch.inf)
        masked[m] = (mu, lvm)
    jmu, jlv = poe(torch.stack([masked[m][0] for m in mods]), torch.stack([masked[m][1] for m in mods]))
    it = iter(noise["z"])
    z = {"joint": jmu + torch.exp(0.5 * jlv) * next(it)}
    res = {}
    joint_kld = kl_gauss(jmu, jlv, torch.zeros_like(jmu), torch.zeros_like(jlv))
    res["joint_divergence"] = joint_kld.mean()
    divergence = joint_kld
    for m in mods:
        mu, lv = outs[m]
        z[m] = mu + torch.exp(0.5 * lv) * next(it)
        k = kl_gauss(jmu, jlv, mu, lv)
        if masks is not None:
            k = k * masks[m].float()
        divergence = divergence + k
        res[f"kl_{m}"] = k.mean()
    loss_rec = 0
    for g in mods:
        for src in ("joint", g):
            rec = dec[g](z[src])
            t = (-recon_log_prob(dec_dist[g], rec, data[g], dec_scale[g]) * rescale[g]).reshape(rec.size(0), -1).sum(-1)
            if masks is not None:
                t = masks[g].float() * t
            loss_rec = loss_rec + t
            res[f"recon_{g}_from_{src}"] = t.mean()
    M = len(mods)
    total = loss_rec / (2 * (M + 1)) + beta * divergence / (M + 1)
    return total.sum(), total.sum(), res


# --------------------------------------------------------------------------------------------
# MVAE                                         models/mvae/mvae_model.py:48-204
# --------------------------------------------------------------------------------------------
def mvae_beta(epoch, batch_ratio, warmup, beta):
    return 1 * beta if epoch >= warmup else (epoch - 1 + batch_ratio) / warmup * beta


def mvae_forward(enc, dec, data, noise, *, subsets, beta, dec_dist, dec_scale, rescale, masks=None):
    """subsets: list of lists (joint first, then unimodal, then sampled); noise["z"][i]: (B_i,L) per subset.
    With masks the batch of each subset is filtered like _filter_inputs_with_masks (mvae_model.py:115-135)."""
    total, metrics, len_batch = 0, {}, 0.0
    enc_order = list(enc.keys())
    for i, s in enumerate(subsets):
        d, mk = data, masks
        if masks is not None:
            filt = torch.zeros_like(masks[s[0]], dtype=torch.bool)
            for m in s:
                filt = torch.logical_or(filt, masks[m])
            if not bool(torch.any(filt)):
                total = total + torch.tensor(0.0, requires_grad=True)
                len_batch = 0.0
                continue
            d = {m: data[m][filt] for m in s}
            mk = {m: masks[m][filt] for m in s}
        mus, lvs = [], []
        for m in enc_order:
            if m in s:
                mu, lv = enc[m](d[m])
                if mk is not None:
                    lv = lv.masked_fill((1 - mk[m].int()).bool().flatten().unsqueeze(-1), torch.inf)
                mus.append(mu)
                lvs.append(lv)
        mus.append(torch.zeros_like(mus[0]))
        lvs.append(torch.zeros_like(lvs[0]))
        smu, slv = stable_poe(torch.stack(mus), torch.stack(lvs))
        z = smu + torch.exp(0.5 * slv) * noise["z"][i]
        elbo = 0
        for m in dec:
            if m in s:
                rec = dec[m](z)
                t = -(recon_log_prob(dec_dist[m], rec, d[m], dec_scale[m]) * rescale[m]).reshape(rec.size(0), -1).sum(-1)
                if mk is not None:
                    t = mk[m].float() * t
                elbo = elbo + t.sum()
        kld = -0.5 * torch.sum(1 + slv - smu.pow(2) - slv.exp())
        elbo = elbo + kld * beta
        # reference quirk (mvae_model.py:104-106): `recon = elbo_sub; elbo_sub += KLD * beta` is an
        # in-place add on the aliased tensor, so the logged "recon" metric equals the full subset ELBO.
        recon = elbo
        n = len(smu)
        key = "_".join(sorted(s))
        metrics[key] = elbo / n
        metrics["beta"] = beta
        metrics["kld" + key] = kld / n
        metrics["recon" + key] = recon / n
        total = total + elbo / n
        len_batch = n
    return total, total * len_batch, metrics


# --------------------------------------------------------------------------------------------
# MoPoE                                        models/mopoe/mopoe_model.py:108-350,417-465
# --------------------------------------------------------------------------------------------
def mopoe_forward(enc, dec, data, noise, *, latent_dim, beta, dec_dist, dec_scale, rescale,
                  masks=None, choice=None, style=False, beta_style=1.0):
    """Complete data: deterministic selection.  Incomplete data: `choice` (B,) int subset index per
    sample must be given (the reference draws it with OneHotCategorical; parity needs it injected).
    noise["z"]: (B,L); noise["style"][m]: (B,Lw) when style=True."""
    mods = list(enc.keys())
    outs = {m: enc[m](data[m]) for m in mods}
    bitmasks = mopoe_subset_bitmasks(mods)
    B = len(data[mods[0]])
    mus, lvs, avail = [], [], []
    for bm in bitmasks:
        sel = [m for i, m in enumerate(mods) if bm >> i & 1]
        sel = sorted(sel)  # the reference iterates sorted(mod_names) (mopoe_model.py:93)
        m_s = torch.stack([outs[m][0] for m in sel])
        l_s = torch.stack([outs[m][1] for m in sel])
        if masks is not None:
            f = torch.ones(B, dtype=torch.bool)
            for m in sel:
                f = torch.logical_and(f, masks[m])
            avail.append(f)
        if len(sel) == len(mods):
            m_s = torch.cat((m_s, torch.zeros(1, B, latent_dim, dtype=m_s.dtype)), dim=0)
            l_s = torch.cat((l_s, torch.zeros(1, B, latent_dim, dtype=m_s.dtype)), dim=0)
        a, b = poe(m_s, l_s)
        mus.append(a)
        lvs.append(b)
    mus, lvs = torch.stack(mus), torch.stack(lvs)
    S = mus.shape[0]
    if masks is not None:
        av = torch.stack(avail, dim=0).float()
        av = av / torch.sum(av, dim=0)
        idx = choice.long()
        weights = av
    else:
        idx = mopoe_sample_to_subset(B, S).long()
        weights = (1 / float(S)) * torch.ones(S, B)
    ar = torch.arange(B)
    jmu, jlv = mus[idx, ar], lvs[idx, ar]
    z = jmu + torch.exp(0.5 * jlv) * noise["z"]
    klds = kl_std_normal(mus, lvs).sum(-1)  # (S,B)
    kld = (weights.to(klds.dtype) * klds).sum(dim=0).mean()
    res = {"joint_divergence": kld}
    loss = 0
    for m in mods:
        if style:
            smu, slv = outs[m][2], outs[m][3]
            sz = smu + torch.exp(0.5 * slv) * noise["style"][m]
            full = torch.cat([z, sz], dim=-1)
        else:
            full = z
        rec = dec[m](full)
        t = (-recon_log_prob(dec_dist[m], rec, data[m], dec_scale[m]) * rescale[m]).view(rec.size(0), -1).sum(-1)
        if masks is not None:
            res["recon_" + m] = (t * masks[m].float()).mean()
        else:
            res["recon_" + m] = t.mean()
        loss = loss + res["recon_" + m]
        if style:
            sk = kl_std_normal(smu, slv).view(smu.size(0), -1).sum(-1)
            if masks is not None:
                sk = sk * masks[m].float()
            kld = kld + sk.mean() * beta_style
    # the reference adds the style KLs IN PLACE to the tensor it stored under "joint_divergence" (mopoe_model.py:166,222)
    res["joint_divergence"] = kld
    loss = loss + beta * kld
    return loss, loss * B, res
