"""Host side of the fused ELBO path: torch.autograd.Function wrappers whose forward/backward call
the CUDA kernels through the C-ABI (multivae_b200/_cabi.py).  No torch fallback.

MoE family (MMVAE / MMVAE+): replaces _compute_k_lws + _dreg_looser / _iwae_looser
(/root/reference/src/multivae/models/mmvaePlus/mmvaePlus_model.py:230-363,
 models/mmvae/mmvae_model.py:160-292).
"""
import math

import torch

from . import _cabi as C


def _f32c(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def _same_family(recons, xs, recon_meta):
    """All reconstructed modalities share D, element type and distribution family (and there are 2..8 of them)."""
    if not 2 <= len(recons) <= 8:
        return False
    d0, t0, f0 = xs[0][0].numel(), recons[0].dtype, recon_meta[0][0]
    if f0 == C.DIST["categorical"]:
        return False
    return all(x[0].numel() == d0 and r.dtype == t0 and m[0] == f0 for r, x, m in zip(recons, xs, recon_meta))


def lpx_fwd(lib, r, x, lpx, Cn, K, B, dist, scale, rescale, mask_r, accumulate):
    """One reconstructed modality: lpx[c,k,b] (+)= rescale * mask * sum_d log p(x[b,d] | recon[c,k,b,d])."""
    D = x[0].numel()
    if dist == C.DIST["categorical"]:
        V = r.shape[-1]
        C.check(lib.mv_moe_lpx_cat_fwd(C.ptr(r), C.dtype_code(r), C.ptr(x), C.ptr(lpx), Cn, K, B, D // V, V, rescale, C.ptr(mask_r),
                                       1 if accumulate else 0, C.stream()), "mv_moe_lpx_cat_fwd")
    else:
        C.check(lib.mv_moe_lpx_fwd(C.ptr(r), C.dtype_code(r), C.ptr(x), C.ptr(lpx), Cn, K, B, D, dist, scale, rescale, C.ptr(mask_r),
                                   1 if accumulate else 0, C.stream(), tag=f"|b={r.numel() * r.element_size() + x.numel() * 4 + Cn * K * B * 4}"),
                "mv_moe_lpx_fwd")


def lpx_bwd(lib, r, x, coef, g_loss, g, Cn, K, B, dist, scale, rescale, mask_r):
    D = x[0].numel()
    if dist == C.DIST["categorical"]:
        V = r.shape[-1]
        C.check(lib.mv_moe_lpx_cat_bwd(C.ptr(r), C.dtype_code(r), C.ptr(x), C.ptr(coef), C.ptr(g_loss), C.ptr(g), Cn, K, B, D // V, V,
                                       rescale, C.ptr(mask_r), C.stream()), "mv_moe_lpx_cat_bwd")
    else:
        C.check(lib.mv_moe_lpx_bwd(C.ptr(r), C.dtype_code(r), C.ptr(x), C.ptr(coef), C.ptr(g_loss), C.ptr(g), Cn, K, B, D, dist, scale,
                                   rescale, C.ptr(mask_r), C.stream(),
                                   tag=f"|b={2 * r.numel() * r.element_size() + x.numel() * 4 + Cn * K * B * 4}"), "mv_moe_lpx_bwd")


class MoEElboFn(torch.autograd.Function):
    """loss = -sum_b sum_c [ sum_k wk*lw  |  logsumexp_k lw - log K ] / n_mods(b)

    Tensor inputs (autograd): u (C,K,B,L), w (C,K,B,Lw) or None, mu_u/sig_u (C,B,L), mu_w/sig_w (C,B,Lw) or
    None, pz_std (L+Lw,), then one reconstruction tensor per recon modality r: (C,K,B,*dims_r), fp32 or bf16.
    `meta` carries the non-differentiable pieces: targets x_r (B,*dims) fp32, pz_mean, masks (C,B) uint8 or
    None, per-r (dist, scale, rescale, mask-row index), latent kind, loss kind, beta, detach flag.
    Also stores meta["wk"], meta["lw"] for the caller (DReG hooks, metrics, parity tests).
    """

    @staticmethod
    def forward(ctx, meta, u, w, mu_u, sig_u, mu_w, sig_w, pz_std, *recons):
        lib = C.lib()
        Cn, K, B, L = u.shape
        Lw = 0 if w is None else w.shape[-1]
        dev = u.device
        u, mu_u, sig_u, pz_std = _f32c(u), _f32c(mu_u), _f32c(sig_u), _f32c(pz_std)
        if Lw:
            w, mu_w, sig_w = _f32c(w), _f32c(mu_w), _f32c(sig_w)
        masks = meta.get("masks")
        lpx = torch.empty(Cn, K, B, device=dev, dtype=torch.float32)
        extra = None
        if meta.get("has_extra"):
            recons, extra = recons[:-1], recons[-1]
        recons = [r.contiguous() for r in recons]
        xs = meta["x"]
        for r, x in zip(recons, xs):
            assert r.shape[:3] == (Cn, K, B) and r[0, 0, 0].numel() == x[0].numel(), (r.shape, x.shape)
        ctx.batched = _same_family(recons, xs, meta["recon"])
        if ctx.batched:
            # one launch over all reconstructed modalities (same D / dtype / distribution family)
            mrows = [None if masks is None else masks[m[3]] for m in meta["recon"]]
            C.check(lib.mv_moe_lpx_fwd_multi(len(recons), C.ptr_array(recons), C.dtype_code(recons[0]), C.ptr_array(xs), C.ptr(lpx),
                                             Cn, K, B, xs[0][0].numel(), meta["recon"][0][0],
                                             C.float_array([m[1] for m in meta["recon"]]),
                                             C.float_array([m[2] for m in meta["recon"]]), C.ptr_array(mrows), C.stream()),
                    "mv_moe_lpx_fwd_multi")
        else:
            for i, (r, x) in enumerate(zip(recons, xs)):
                dist, scale, rescale, mrow = meta["recon"][i]
                mask_r = None if masks is None else masks[mrow]
                lpx_fwd(lib, r, x, lpx, Cn, K, B, dist, scale, rescale, mask_r, i > 0)
        if meta.get("has_extra"):   # an additive per-(c, k, b) log-weight term computed by the caller (CMVAE's cluster-mixture prior)
            lpx = lpx + _f32c(extra)
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        lw, wk, coef, loss_b = f(Cn, K, B), f(Cn, K, B), f(Cn, K, B), f(B)
        g_u, g_mu_u, g_sig_u = f(Cn, K, B, L), f(Cn, B, L), f(Cn, B, L)
        g_w = f(Cn, K, B, Lw) if Lw else None
        g_mu_w = f(Cn, B, Lw) if Lw else None
        g_sig_w = f(Cn, B, Lw) if Lw else None
        g_pz = f(B, L + Lw)
        C.check(lib.mv_moe_lw_fwd(C.ptr(u), C.ptr(w), C.ptr(mu_u), C.ptr(sig_u), C.ptr(mu_w), C.ptr(sig_w),
                                  C.ptr(meta["pz_mean"]), C.ptr(pz_std), C.ptr(lpx), C.ptr(masks), C.ptr(lw),
                                  C.ptr(wk), C.ptr(coef), C.ptr(loss_b), C.ptr(g_u), C.ptr(g_w), C.ptr(g_mu_u),
                                  C.ptr(g_sig_u), C.ptr(g_mu_w), C.ptr(g_sig_w), C.ptr(g_pz), Cn, K, B, L, Lw,
                                  meta["latent_kind"], meta["loss_kind"], float(meta["beta"]),
                                  1 if meta["detach"] else 0, 1 if meta.get("skip_u_prior") else 0, C.stream()), "mv_moe_lw_fwd")
        meta["wk"], meta["lw"], meta["lpx"] = wk, lw, lpx
        ctx.meta = meta
        ctx.dims = (Cn, K, B, L, Lw)
        ctx.save_for_backward(coef, g_u, g_w, g_mu_u, g_sig_u, g_mu_w, g_sig_w, g_pz, *recons)
        return loss_b.sum()

    @staticmethod
    def backward(ctx, g_loss):
        lib = C.lib()
        meta = ctx.meta
        Cn, K, B, L, Lw = ctx.dims
        coef, g_u, g_w, g_mu_u, g_sig_u, g_mu_w, g_sig_w, g_pz, *recons = ctx.saved_tensors
        g_loss = g_loss.reshape(1).float().contiguous()
        masks = meta.get("masks")
        g_recons = []
        need = ctx.needs_input_grad
        if ctx.batched and all(need[8 + i] for i in range(len(recons))):
            g_recons = [torch.empty_like(r) for r in recons]
            mrows = [None if masks is None else masks[m[3]] for m in meta["recon"]]
            C.check(lib.mv_moe_lpx_bwd_multi(len(recons), C.ptr_array(recons), C.dtype_code(recons[0]), C.ptr_array(meta["x"]),
                                             C.ptr(coef), C.ptr(g_loss), C.ptr_array(g_recons), Cn, K, B, meta["x"][0][0].numel(),
                                             meta["recon"][0][0], C.float_array([m[1] for m in meta["recon"]]),
                                             C.float_array([m[2] for m in meta["recon"]]), C.ptr_array(mrows), C.stream()),
                    "mv_moe_lpx_bwd_multi")
            recons = []
        for i, r in enumerate(recons):
            if not need[8 + i]:
                g_recons.append(None)
                continue
            dist, scale, rescale, mrow = meta["recon"][i]
            x = meta["x"][i]
            g = torch.empty_like(r)
            mask_r = None if masks is None else masks[mrow]
            lpx_bwd(lib, r, x, coef, g_loss, g, Cn, K, B, dist, scale, rescale, mask_r)
            g_recons.append(g)
        s = g_loss
        det = meta["detach"]
        g_extra = ((coef * s),) if meta.get("has_extra") else ()
        return (None, g_u * s if need[1] else None, (g_w * s) if (Lw and need[2]) else None,
                None if (det or not need[3]) else g_mu_u * s, None if (det or not need[4]) else g_sig_u * s,
                None if (det or not Lw or not need[5]) else g_mu_w * s,
                None if (det or not Lw or not need[6]) else g_sig_w * s,
                (g_pz.sum(0) * s) if need[7] else None, *g_recons, *g_extra)


class MoESampleFn(torch.autograd.Function):
    """sig = std(log-variance), K reparameterised samples of every unimodal posterior and the decoder inputs of all (cond, recon)
    modality pairs in one launch (mv_moe_sample_fwd / mv_moe_sample_bwd).

    forward(meta, mu_u, lv_u, mu_w, lv_w, prior_mean, prior_std, e_u, e_w, e_x) ->  (sig_u, sig_w, U, W, Z)
      mu_u, lv_u (C,B,L); mu_w, lv_w (C,B,Lw) or None; prior_mean, prior_std (C,Lw) of the private codes' priors;
      e_u (C,K,B,L), e_w (C,K,B,Lw), e_x (C,C-1,K,B,Lw) standard draws;  Z (C_recon, C_cond, K, B, L+Lw).
    meta is the dict MoEElboFn fills: the backward reads meta["wk"] (DReG: detach) to weight the samples' gradient."""

    @staticmethod
    def forward(ctx, meta, mu_u, lv_u, mu_w, lv_w, pm, sp, e_u, e_w, e_x):
        lib = C.lib()
        Cn, K, B, L = e_u.shape
        Lw = 0 if mu_w is None else mu_w.shape[-1]
        dev = mu_u.device
        mu_u, lv_u, e_u = _f32c(mu_u), _f32c(lv_u), _f32c(e_u)
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        sig_u, U = f(Cn, B, L), f(Cn, K, B, L)
        if Lw:
            mu_w, lv_w, pm, sp, e_w = _f32c(mu_w), _f32c(lv_w), _f32c(pm), _f32c(sp), _f32c(e_w)
            e_x = _f32c(e_x) if e_x is not None else None
            sig_w, W, Z = f(Cn, B, Lw), f(Cn, K, B, Lw), f(Cn, Cn, K, B, L + Lw)
        else:
            sig_w = W = Z = None
        kind = meta["std_kind"]
        C.check(lib.mv_moe_sample_fwd(C.ptr(mu_u), C.ptr(lv_u), C.ptr(mu_w), C.ptr(lv_w), C.ptr(pm) if Lw else None, C.ptr(sp) if Lw else None,
                                      C.ptr(e_u), C.ptr(e_w) if Lw else None, C.ptr(e_x) if Lw else None, kind, C.ptr(sig_u), C.ptr(sig_w),
                                      C.ptr(U), C.ptr(W), C.ptr(Z), Cn, K, B, L, Lw, C.stream()), "mv_moe_sample_fwd")
        ctx.meta, ctx.dims = meta, (Cn, K, B, L, Lw)
        ctx.save_for_backward(lv_u, sig_u, e_u, *((lv_w, sig_w, e_w) if Lw else ()), *((e_x,) if (Lw and e_x is not None) else ()))
        ctx.has_ex = bool(Lw and e_x is not None)
        empty = torch.empty(0, device=dev)
        return sig_u, (sig_w if Lw else empty), U, (W if Lw else empty), (Z if Lw else empty)

    @staticmethod
    def backward(ctx, g_sig_u, g_sig_w, g_U, g_W, g_Z):
        lib = C.lib()
        Cn, K, B, L, Lw = ctx.dims
        meta = ctx.meta
        sv = ctx.saved_tensors
        lv_u, sig_u, e_u = sv[:3]
        lv_w = sig_w = e_w = e_x = None
        if Lw:
            lv_w, sig_w, e_w = sv[3:6]
            e_x = sv[6] if ctx.has_ex else None
        c = lambda t: None if t is None else _f32c(t)  # noqa: E731
        g_sig_u, g_U = c(g_sig_u), c(g_U)
        if Lw:
            g_sig_w, g_W, g_Z = c(g_sig_w), c(g_W), c(g_Z)
        else:
            g_sig_w = g_W = g_Z = None
        wk = meta["wk"] if meta.get("detach") else None   # DReG: weight the samples' gradient once more (no tensor hooks needed)
        dev = lv_u.device
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        g_mu_u, g_lv_u = f(Cn, B, L), f(Cn, B, L)
        g_mu_w = g_lv_w = g_sp = None
        if Lw:
            g_mu_w, g_lv_w, g_sp = f(Cn, B, Lw), f(Cn, B, Lw), f(Cn, Lw)
        C.check(lib.mv_moe_sample_bwd(C.ptr(lv_u), C.ptr(lv_w), C.ptr(sig_u), C.ptr(sig_w), C.ptr(e_u), C.ptr(e_w), C.ptr(e_x), C.ptr(g_U),
                                      C.ptr(g_W), C.ptr(g_Z), C.ptr(g_sig_u), C.ptr(g_sig_w), C.ptr(wk), meta["std_kind"], C.ptr(g_mu_u),
                                      C.ptr(g_lv_u), C.ptr(g_mu_w), C.ptr(g_lv_w), C.ptr(g_sp), Cn, K, B, L, Lw, C.stream()),
                "mv_moe_sample_bwd")
        need = ctx.needs_input_grad
        return (None, g_mu_u if need[1] else None, g_lv_u if need[2] else None, g_mu_w if (Lw and need[3]) else None,
                g_lv_w if (Lw and need[4]) else None, None, g_sp if (Lw and need[6]) else None, None, None, None)


def log_var_to_std(log_var, kind):
    """mmvaePlus_model.py:113-123 / mmvae_model.py:66-74 (tiny (B,L) tensors; stays in torch)."""
    if kind == "laplace_with_softmax":
        return torch.softmax(log_var, dim=-1) * log_var.size(-1) + 1e-6
    if kind == "normal_with_softplus":
        return torch.nn.functional.softplus(log_var) + 1e-6
    return torch.exp(0.5 * log_var)


def standard_noise(shape, kind, device, generator=None):
    """Standard Laplace(0,1) / Normal(0,1) draws e such that z = loc + scale * e, i.e. what
    torch.distributions.{Laplace,Normal}.rsample draw internally."""
    if kind == "laplace_with_softmax":
        eps = torch.finfo(torch.float32).eps
        u = torch.empty(shape, device=device, dtype=torch.float32).uniform_(eps - 1, 1, generator=generator)
        return -u.sign() * torch.log1p(-u.abs().clamp(min=torch.finfo(torch.float32).tiny))
    return torch.empty(shape, device=device, dtype=torch.float32).normal_(generator=generator)


LOG_2PI = math.log(2 * math.pi)


class PoEFn(torch.autograd.Function):
    """Fused subset-PoE aggregation (mv_poe_fwd / mv_poe_bwd).

    forward(meta, mu, lv) with mu, lv (M,B,L) -> (z (B,L), kl_b (B,), kldm_b (M,B) or empty).
    meta: masks (M,B) u8|None, subsets (S,) int32 bitmasks, sel (B,) int32|None, w (S,B)|None, w_uniform,
    noise (B,L), prior_mode, stable, eps, want_kldm."""

    @staticmethod
    def forward(ctx, meta, mu, lv):
        lib = C.lib()
        M, B, L = mu.shape
        mu, lv = _f32c(mu), _f32c(lv)
        dev = mu.device
        z = torch.empty(B, L, device=dev, dtype=torch.float32)
        kl_b = torch.empty(B, device=dev, dtype=torch.float32)
        kldm = torch.empty(M, B, device=dev, dtype=torch.float32) if meta.get("want_kldm") else None
        S = meta["subsets"].numel()
        C.check(lib.mv_poe_fwd(C.ptr(mu), C.ptr(lv), C.ptr(meta.get("masks")), C.ptr(meta["subsets"]), S,
                               C.ptr(meta.get("sel")), C.ptr(meta.get("w")), float(meta.get("w_uniform", 1.0)),
                               C.ptr(meta["noise"]), meta["prior_mode"], 1 if meta["stable"] else 0,
                               float(meta.get("eps", 1e-8)), C.ptr(z), None, None, C.ptr(kl_b), C.ptr(kldm), M, B, L,
                               C.stream()), "mv_poe_fwd")
        ctx.meta = meta
        ctx.save_for_backward(mu, lv)
        ctx.has_kldm = kldm is not None
        if kldm is None:
            kldm = torch.empty(0, device=dev)
        return z, kl_b, kldm

    @staticmethod
    def backward(ctx, g_z, g_kl, g_kldm):
        lib = C.lib()
        meta = ctx.meta
        mu, lv = ctx.saved_tensors
        M, B, L = mu.shape
        g_mu, g_lv = torch.empty_like(mu), torch.empty_like(lv)
        S = meta["subsets"].numel()
        gz = None if g_z is None else _f32c(g_z)
        gk = None if g_kl is None else _f32c(g_kl)
        gm = _f32c(g_kldm) if (ctx.has_kldm and g_kldm is not None) else None
        C.check(lib.mv_poe_bwd(C.ptr(mu), C.ptr(lv), C.ptr(meta.get("masks")), C.ptr(meta["subsets"]), S,
                               C.ptr(meta.get("sel")), C.ptr(meta.get("w")), float(meta.get("w_uniform", 1.0)),
                               C.ptr(meta["noise"]), meta["prior_mode"], 1 if meta["stable"] else 0,
                               float(meta.get("eps", 1e-8)), C.ptr(gz), C.ptr(gk), C.ptr(gm), C.ptr(g_mu), C.ptr(g_lv),
                               M, B, L, C.stream()), "mv_poe_bwd")
        return None, g_mu, g_lv


class ReconNLLFn(torch.autograd.Function):
    """nll_b[b] = -rescale * mask[b] * sum_d log p(x[b,d] | recon[b,d])  for one modality (rows = samples).
    Same streaming kernels as the MoE family with C = K = 1 (mv_moe_lpx_fwd / mv_moe_lpx_bwd).
    Replaces `(-recon_log_probs[m](recon, x) * rescale).reshape(B,-1).sum(-1)` (+ mask multiply) of
    mvtcae_model.py:62-70, mvae_model.py:91-102, mopoe_model.py:186-201."""

    @staticmethod
    def forward(ctx, recon, x, mask, dist, scale, rescale):
        lib = C.lib()
        recon = recon.contiguous()
        B = recon.shape[0]
        D = recon[0].numel()
        lp = torch.empty(B, device=recon.device, dtype=torch.float32)
        lpx_fwd(lib, recon, x, lp, 1, 1, B, dist, scale, rescale, mask, False)
        ctx.save_for_backward(recon, x, mask if mask is not None else torch.empty(0, device=recon.device))
        ctx.args = (dist, scale, rescale, mask is not None)
        return -lp

    @staticmethod
    def backward(ctx, g):
        lib = C.lib()
        recon, x, mask = ctx.saved_tensors
        dist, scale, rescale, has_mask = ctx.args
        B, D = recon.shape[0], recon[0].numel()
        coef = (-g).float().contiguous()  # d(-lp)/d lp = -1, times upstream
        one = torch.ones(1, device=recon.device, dtype=torch.float32)
        gr = torch.empty_like(recon)
        lpx_bwd(lib, recon, x, coef, one, gr, 1, 1, B, dist, scale, rescale, mask if has_mask else None)
        return gr, None, None, None, None, None


class GaussKLFn(torch.autograd.Function):
    """kl_divergence(mean, log_var, prior_mean, prior_log_var) of models/base/base_utils.py:90-119: the general Gaussian KL summed
    over the last dimension (mv_gauss_kl_fwd / mv_gauss_kl_bwd); the prior is per sample or one broadcast row."""

    @staticmethod
    def forward(ctx, mean, log_var, prior_mean, prior_log_var):
        lib = C.lib()
        L = mean.shape[-1]
        mu, lv = _f32c(mean).reshape(-1, L), _f32c(log_var).reshape(-1, L)
        pm = _f32c(prior_mean.expand_as(prior_log_var) if prior_mean.numel() < prior_log_var.numel() else prior_mean).reshape(-1, L)
        pl = _f32c(prior_log_var.expand_as(prior_mean) if prior_log_var.numel() < prior_mean.numel() else prior_log_var).reshape(-1, L)
        rows = mu.shape[0]
        if pm.shape[0] not in (1, rows):
            pm, pl = pm.expand(rows, L).contiguous(), pl.expand(rows, L).contiguous()
        out = torch.empty(rows, device=mu.device, dtype=torch.float32)
        C.check(lib.mv_gauss_kl_fwd(C.ptr(mu), C.ptr(lv), C.ptr(pm), C.ptr(pl), C.ptr(out), rows, L, pm.shape[0], C.stream()),
                "mv_gauss_kl_fwd")
        ctx.save_for_backward(mu, lv, pm, pl)
        ctx.shapes = (mean.shape, log_var.shape, prior_mean.shape, prior_log_var.shape)
        return out.reshape(mean.shape[:-1])

    @staticmethod
    def backward(ctx, g):
        lib = C.lib()
        mu, lv, pm, pl = ctx.saved_tensors
        rows, L = mu.shape
        g = _f32c(g).reshape(-1)
        g_mu, g_lv, g_pm, g_pl = torch.empty_like(mu), torch.empty_like(lv), torch.empty_like(pm), torch.empty_like(pl)
        C.check(lib.mv_gauss_kl_bwd(C.ptr(mu), C.ptr(lv), C.ptr(pm), C.ptr(pl), C.ptr(g), C.ptr(g_mu), C.ptr(g_lv), C.ptr(g_pm),
                                    C.ptr(g_pl), rows, L, pm.shape[0], C.stream()), "mv_gauss_kl_bwd")
        s = ctx.shapes
        red = lambda t, shp: t.reshape(*([1] * (len(s[0]) - 1)), L).sum_to_size(shp) if t.shape[0] == 1 else t.reshape(s[0]).sum_to_size(shp)  # noqa: E731
        return g_mu.reshape(s[0]), g_lv.reshape(s[1]), red(g_pm, s[2]), red(g_pl, s[3])


def kl_divergence(mean, log_var, prior_mean, prior_log_var):
    """base_utils.py:90-119."""
    return GaussKLFn.apply(mean, log_var, prior_mean, prior_log_var)


def logmeanexp(lw):
    """lw (R, B) -> (B,): logsumexp over the R importance samples - log R (mv_logmeanexp)."""
    lw = _f32c(lw)
    R, B = lw.shape
    out = torch.empty(B, device=lw.device, dtype=torch.float32)
    C.check(C.lib().mv_logmeanexp(C.ptr(lw), R, B, C.ptr(out), C.stream()), "mv_logmeanexp")
    return out


def poe_joint(mu, lv, masks, bits, prior_mode, stable, eps=1e-8):
    """Parameters (joint_mu, joint_lv), each (B, L), of the product of the experts in ONE subset (bitmask tensor `bits` of one
    element): the forward kernel of the PoE family run for its joint_mu / joint_lv outputs only (inference paths: encode,
    compute_joint_nll).  mu, lv (M, B, L); masks (M, B) uint8 or None."""
    lib = C.lib()
    mu, lv = _f32c(mu), _f32c(lv)
    M, B, L = mu.shape
    jm = torch.empty(B, L, device=mu.device, dtype=torch.float32)
    jl = torch.empty_like(jm)
    kl = torch.empty(B, device=mu.device, dtype=torch.float32)
    C.check(lib.mv_poe_fwd(C.ptr(mu), C.ptr(lv), C.ptr(masks), C.ptr(bits), 1, None, None, 1.0, None, prior_mode,
                           1 if stable else 0, float(eps), None, C.ptr(jm), C.ptr(jl), C.ptr(kl), None, M, B, L, C.stream()),
            "mv_poe_fwd")
    return jm, jl


LOG_SQRT_2PI = 0.5 * math.log(2 * math.pi)


def normal_logpdf_sum(z, mu, log_var):
    """sum_l log N(z; mu, exp(log_var)) over the last dimension, torch.distributions.Normal.log_prob's formula."""
    var = torch.exp(log_var)
    return (-((z - mu) ** 2) / (2 * var) - 0.5 * log_var - LOG_SQRT_2PI).sum(-1)
