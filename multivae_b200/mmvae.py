"""MMVAE (mixture of experts, single latent) behind the reference's API
(reference: models/mmvae/mmvae_model.py:16-292).  Same fused kernels as MMVAE+ with Lw = 0 and beta
fixed to 1 in lw (the reference never uses config.beta in the loss, mmvae_model.py:228)."""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _cabi as C
from .base import BaseMultiVAE
from .containers import ModelOutput, drop_unused_modalities
from .elbo import MoEElboFn, MoESampleFn, log_var_to_std, standard_noise


class MMVAE(BaseMultiVAE):
    def __init__(self, model_config, encoders=None, decoders=None):
        super().__init__(model_config, encoders, decoders)
        if model_config.prior_and_posterior_dist not in ("laplace_with_softmax", "normal"):
            raise AttributeError(" The posterior_dist parameter must be  either 'laplace_with_softmax' or 'normal'. "
                                 f" {model_config.prior_and_posterior_dist} was provided.")
        self.prior_mean = nn.Parameter(torch.zeros(1, self.latent_dim), requires_grad=False)
        self.prior_log_var = nn.Parameter(torch.zeros(1, self.latent_dim), requires_grad=model_config.learn_prior)
        self.model_name = "MMVAE"
        self.noise_source = None

    def _noise(self, shape, device):
        kind = self.model_config.prior_and_posterior_dist
        if self.noise_source is not None:
            return self.noise_source(shape, kind, device)
        return standard_noise(shape, kind, device)

    def forward(self, inputs, **kwargs):
        inputs = drop_unused_modalities(inputs)
        self._unused_modalities = [m for m in self.encoders if m not in inputs.data]   # their parameters get NO gradient this step
        compute_loss = kwargs.pop("compute_loss", True)
        detailed = kwargs.pop("detailed_output", False)
        K = kwargs.pop("K", self.model_config.K)
        if self.model_config.loss not in C.LOSS:
            raise NotImplementedError()
        loss, meta, extra = self._elbo(inputs, K, self.model_config.loss, detailed)
        self._last = meta
        out = ModelOutput(loss=loss, loss_sum=loss, metrics={}) if compute_loss else ModelOutput()
        if detailed:
            out.update(extra)
        return out

    def _elbo(self, inputs, K, loss_name, detailed=False, rescale=None):
        """Encoders, K reparameterised samples per modality, one batched decoder call per reconstructed modality over the
        C*K*B rows, fused lpx / lw kernels.  Returns (loss, meta with lw / wk / lpx, detailed-output dict)."""
        kind = self.model_config.prior_and_posterior_dist
        detach = loss_name == "dreg_looser"
        mods = list(inputs.data.keys())
        dev = inputs.data[mods[0]].device
        B = len(inputs.data[mods[0]])
        enc_out = self._run_encoders(inputs, mods, dev)
        mu = torch.stack([enc_out[c].embedding.float() for c in mods])
        lv = torch.stack([enc_out[c].log_covariance.float() for c in mods])
        L_ = mu.shape[-1]
        if self.noise_source is not None:
            e = torch.stack([self._noise((K, B, L_), dev) for _ in mods])
        else:
            e = standard_noise((len(mods), K, B, L_), kind, dev)
        meta = dict(std_kind=C.STD_KIND[kind], detach=detach)
        # std(log-variance) + the K reparameterised samples of every posterior in one launch (mv_moe_sample_fwd)
        sig, _, Z, _, _ = MoESampleFn.apply(meta, mu, lv, None, None, None, None, e, None, None)
        zflat = Z.reshape(-1, Z.shape[-1])
        recs = self._run_decoders({r: zflat for r in mods}, dev)
        recons = [recs[r].reshape(len(mods), K, B, *recs[r].shape[1:]) for r in mods]
        pz_std = log_var_to_std(self.prior_log_var, kind).reshape(-1)
        rmeta = self._recon_meta(mods, mods)
        if rescale is not None:
            rmeta = [(d, sc, float(rescale), row) for d, sc, _, row in rmeta]
        meta.update(x=[self._target(inputs, r, rec) for r, rec in zip(mods, recons)],
                    pz_mean=self.prior_mean.detach().reshape(-1).float().contiguous(),
                    masks=self._stack_masks(inputs, mods), recon=rmeta,
                    latent_kind=C.LATENT[kind], loss_kind=C.LOSS[loss_name], beta=1.0)
        # DReG: MoESampleFn's backward multiplies the samples' gradient by meta["wk"] (filled in by the call below)
        loss = MoEElboFn.apply(meta, Z, None, mu, sig, None, None, pz_std, *recons)
        extra = {}
        if detailed:   # the reference's detailed_output fields (mmvae_model.py:151-156)
            extra = dict(qz_xs={c: self.post_dist(mu[i], sig[i]) for i, c in enumerate(mods)},
                         qz_xs_detach={c: self.post_dist(mu[i].detach(), sig[i].detach()) for i, c in enumerate(mods)},
                         zss={c: Z[i] for i, c in enumerate(mods)},
                         recon={c: {r: recons[j][i] for j, r in enumerate(mods)} for i, c in enumerate(mods)})
        return loss, meta, extra

    # ---- inference (mmvae_model.py:313-474) ---------------------------------------------------------------------------
    @property
    def post_dist(self):
        import torch.distributions as td
        return td.Laplace if self.model_config.prior_and_posterior_dist == "laplace_with_softmax" else td.Normal

    prior_dist = post_dist

    def log_var_to_std(self, log_var):
        return log_var_to_std(log_var, self.model_config.prior_and_posterior_dist)

    def _latent_logpdf_sum(self, z, mu, sigma):
        if self.model_config.prior_and_posterior_dist == "laplace_with_softmax":
            return (-torch.log(2 * sigma) - (z - mu).abs() / sigma).sum(-1)
        return (-((z - mu) ** 2) / (2 * sigma ** 2) - torch.log(sigma) - 0.5 * math.log(2 * math.pi)).sum(-1)

    def encode(self, inputs, cond_mod="all", N=1, return_mean=False, **kwargs):
        """Samples from ONE conditioning modality's posterior chosen at random (numpy, like the reference), or the mean of
        the posterior means (mmvae_model.py:313-363)."""
        cond_mod = super().encode(inputs, cond_mod, N, **kwargs).cond_mod
        if return_mean:
            with self._nn_ctx():
                emb = torch.stack([self.encoders[m](inputs.data[m]).embedding.float() for m in cond_mod]).mean(0)
            z = torch.stack([emb] * N) if N > 1 else emb
        else:
            mod = np.random.choice(cond_mod)
            with self._nn_ctx():
                o = self.encoders[mod](inputs.data[mod])
            mu, sigma = o.embedding.float(), self.log_var_to_std(o.log_covariance.float())
            shape = tuple(mu.shape) if N == 1 else (N,) + tuple(mu.shape)
            z = mu + sigma * self._noise(shape, mu.device)
        if kwargs.pop("flatten", False):
            z = z.reshape(-1, self.latent_dim)
        return ModelOutput(z=z, one_latent_space=True)

    def _encode_many(self, inputs, cond_mod, n):
        """n independent encode() calls (each with its own random conditioning modality, same numpy stream as n sequential
        reference calls) in one pass: (n, B, L)."""
        cond_mod = BaseMultiVAE.encode(self, inputs, cond_mod, 1).cond_mod
        picks = [np.random.choice(cond_mod) for _ in range(n)]
        post = {}
        for m in set(picks):
            with self._nn_ctx():
                o = self.encoders[m](inputs.data[m])
            post[m] = (o.embedding.float(), self.log_var_to_std(o.log_covariance.float()))
        mu = torch.stack([post[m][0] for m in picks])
        sg = torch.stack([post[m][1] for m in picks])
        return ModelOutput(z=mu + sg * self._noise(tuple(mu.shape), mu.device), one_latent_space=True)

    def generate_from_prior(self, n_samples, **kwargs):
        kind = self.model_config.prior_and_posterior_dist
        std = log_var_to_std(self.prior_log_var, kind)
        shape = (n_samples,) + tuple(std.shape) if n_samples > 1 else tuple(std.shape)
        z = self.prior_mean + std * self._noise(shape, std.device)
        return ModelOutput(z=z.squeeze(), one_latent_space=True)

    @torch.no_grad()
    def compute_joint_nll(self, inputs, K=1000, batch_size_K=100):
        """K samples from one (random) modality's posterior, proposal density = the mixture of all unimodal posteriors
        (mmvae_model.py:366-444); batched over datapoints."""
        from .elbo import logmeanexp
        self.eval()
        if hasattr(inputs, "masks"):
            raise AttributeError("The compute_joint_nll method is not yet implemented for incomplete datasets.")
        post = []
        for m in self.encoders:
            with self._nn_ctx():
                o = self.encoders[m](inputs.data[m])
            post.append((o.embedding.float(), self.log_var_to_std(o.log_covariance.float())))
        z_all = self.encode(inputs, N=K).z
        if K == 1:
            z_all = z_all.unsqueeze(0)
        pm, ps = self.prior_mean.float(), self.log_var_to_std(self.prior_log_var.float())
        lws = []
        for k0 in range(0, K, batch_size_K):
            z = z_all[k0:k0 + batch_size_K]
            lpx = self._iw_lpx(inputs, lambda m, z=z: z, z.shape[0])
            lq = torch.logsumexp(torch.stack([self._latent_logpdf_sum(z, mu, sg) for mu, sg in post]), dim=0) - math.log(self.n_modalities)
            lws.append(lpx + self._latent_logpdf_sum(z, pm, ps) - lq)
        return -logmeanexp(torch.cat(lws, dim=0)).sum()

    @torch.no_grad()
    def compute_joint_nll_paper(self, inputs, K=1000, batch_size_K=10):
        """The original paper's estimator: all mixture-of-experts samples, modality rescaling kept (mmvae_model.py:446-469;
        note that it log-sum-exps the batch SUMS of the chunks, like the reference)."""
        self.eval()
        lws, done = [], 0
        while done < K:
            n = min(batch_size_K, K - done)
            done += n
            _, meta, _ = self._elbo(inputs, n, "iwae_looser")
            lw = meta["lw"]                                                     # (C, n, B)
            per_c = torch.logsumexp(lw, dim=1) - math.log(n)                    # iwae(): log-mean-exp over k ...
            ll = torch.logsumexp(per_c, dim=0) - math.log(self.n_modalities)    # ... then over the modalities
            lws.append(ll.sum() + math.log(n * self.n_modalities))
        ll = torch.logsumexp(torch.stack(lws), dim=0) - math.log(done * self.n_modalities)
        return -ll
