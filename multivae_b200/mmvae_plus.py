"""MMVAE+ behind the reference's API (reference: models/mmvaePlus/mmvaePlus_model.py:28-363).

forward(inputs, **kwargs) -> ModelOutput(loss, loss_sum, metrics={}).  Differences in *how*, not what:
the M*M decoder invocations of the reference are batched into one call per decoder over the M*K*B rows
of all conditioning modalities, and _compute_k_lws + the IWAE/DReG looser are three fused CUDA kernels
(multivae_b200/csrc/elbo_moe.cu) instead of ~10^4 ATen calls."""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _cabi as C
from .base import BaseMultiVAE
from .containers import ModelOutput, drop_unused_modalities
from .elbo import MoEElboFn, MoESampleFn, log_var_to_std, standard_noise
from .nn.default_architectures import BaseDictDecodersMultiLatents, BaseDictEncoders_MultiLatents


class MoEPlusBase(BaseMultiVAE):
    """Machinery shared by the mixture-of-experts models with shared + private latent codes (MMVAE+, CMVAE): encoders, K
    reparameterised samples, the M*M decoder invocations batched into one call per decoder over the C*K*B rows, and the fused
    lpx / lw kernels.  Subclasses provide the priors (`_cross_prior`, `_shared_prior`) and, optionally, an additive log-weight
    term (`_extra_lw`)."""

    skip_u_prior = False

    def default_encoders(self, model_config):
        return BaseDictEncoders_MultiLatents(model_config.input_dims, model_config.latent_dim,
                                             {m: model_config.modalities_specific_dim for m in model_config.input_dims})

    def default_decoders(self, model_config):
        return BaseDictDecodersMultiLatents(model_config.input_dims, model_config.latent_dim,
                                            {m: model_config.modalities_specific_dim for m in model_config.input_dims})

    def _extra_lw(self, U, beta):
        return None

    def _noise(self, shape, device):
        kind = self.model_config.prior_and_posterior_dist
        if self.noise_source is not None:
            return self.noise_source(shape, kind, device)
        pool = getattr(self, "_noise_pool", None)
        if pool is not None:   # slice of the step's single batched draw (see _begin_noise_pool)
            n = 1
            for d in shape:
                n *= d
            flat, off = pool
            assert off + n <= flat.numel(), "noise pool exhausted"
            self._noise_pool = (flat, off + n)
            return flat[off:off + n].view(*shape)
        return standard_noise(shape, kind, device)

    def _begin_noise_pool(self, n_elements, device):
        """All standard draws of a forward pass in ONE batched draw (one RNG kernel + one transform chain instead of one per
        (conditioning, reconstructed) modality pair): the draws are i.i.d., so slicing a single draw is statistically the
        same as the reference's sequence of rsample calls (mmvaePlus_model.py:136-186).  Not used with an injected noise_source."""
        if self.noise_source is None:
            self._noise_pool = (standard_noise((n_elements,), self.model_config.prior_and_posterior_dist, device), 0)

    @property
    def post_dist(self):
        import torch.distributions as td
        return td.Laplace if self.model_config.prior_and_posterior_dist == "laplace_with_softmax" else td.Normal

    prior_dist = post_dist

    def _log_var_to_std(self, log_var):
        return log_var_to_std(log_var, self.model_config.prior_and_posterior_dist)

    def forward(self, inputs, **kwargs):
        if self.objective not in C.LOSS:
            raise NotImplementedError()
        inputs = drop_unused_modalities(inputs)
        self._unused_modalities = [m for m in self.encoders if m not in inputs.data]   # their parameters get NO gradient this step
        K = kwargs.pop("K", self.model_config.K)
        loss, meta = self._elbo(inputs, K, self.objective)
        self._last = meta
        return ModelOutput(loss=loss, loss_sum=loss, metrics=dict())

    def _elbo(self, inputs, K, loss_name, rescale=None, beta=None):
        """Encoders, reparameterised samples, one batched decoder call per reconstructed modality over the C*K*B rows, fused
        lpx / lw kernels.  Returns (loss, meta with lw / wk / lpx).  rescale / beta override the model's (likelihood estimation)."""
        kind = self.model_config.prior_and_posterior_dist
        detach = loss_name == "dreg_looser"
        mods = list(inputs.data.keys())
        dev = inputs.data[mods[0]].device
        B = len(inputs.data[mods[0]])

        # encoders, then ONE launch for std(log-variance), the K reparameterised samples of every posterior and the decoder
        # inputs of all (conditioning, reconstructed) pairs (mv_moe_sample_fwd).  The standard draws are consumed in the
        # reference's order (per conditioning modality: u, w, then one prior draw per *other* modality: mmvaePlus_model.py:136-186)
        L_, Lw_ = self.model_config.latent_dim, self.modalities_specific_dim
        Cn = len(mods)
        enc_out = self._run_encoders(inputs, mods, dev)
        mu_u = torch.stack([enc_out[c].embedding.float() for c in mods])
        lv_u = torch.stack([enc_out[c].log_covariance.float() for c in mods])
        mu_w = torch.stack([enc_out[c].style_embedding.float() for c in mods])
        lv_w = torch.stack([enc_out[c].style_log_covariance.float() for c in mods])
        if self.noise_source is not None:
            eu, ew, ex = [], [], []
            for c in mods:
                eu.append(self._noise((K, B, L_), dev))
                ew.append(self._noise((K, B, Lw_), dev))
                ex += [self._noise((K, B, Lw_), dev) for r in mods if r != c]
            e_u, e_w = torch.stack(eu), torch.stack(ew)
            e_x = torch.stack(ex).view(Cn, Cn - 1, K, B, Lw_) if Cn > 1 else None
        else:   # i.i.d. draws: one batched draw, three views of it (same distribution as the reference's sequence of rsample calls)
            n_u, n_w, n_x = Cn * K * B * L_, Cn * K * B * Lw_, Cn * (Cn - 1) * K * B * Lw_
            pool = standard_noise((n_u + n_w + n_x,), kind, dev)
            e_u, e_w = pool[:n_u].view(Cn, K, B, L_), pool[n_u:n_u + n_w].view(Cn, K, B, Lw_)
            e_x = pool[n_u + n_w:].view(Cn, Cn - 1, K, B, Lw_) if Cn > 1 else None
        priors = [self._cross_prior(r) for r in mods]
        pm = torch.cat([p_[0] for p_ in priors], dim=0)                               # (C, Lw)
        sp = log_var_to_std(torch.cat([p_[1] for p_ in priors], dim=0), kind)         # std per modality prior (softmax over its Lw dims)
        meta = dict(std_kind=C.STD_KIND[kind], detach=detach)
        sig_u, sig_w, U, W, Z = MoESampleFn.apply(meta, mu_u, lv_u, mu_w, lv_w, pm, sp, e_u, e_w, e_x)

        # one batched decoder call per reconstructed modality over all conditioning modalities
        z_by_mod = {r: Z[i].reshape(-1, L_ + Lw_) for i, r in enumerate(mods)}
        self._decoder_inputs = list(z_by_mod.values())   # the trainer hooks these: all have a gradient <=> every decoder's backward is done
        recs = self._run_decoders(z_by_mod, dev)
        recons = [recs[r].reshape(len(mods), K, B, *recs[r].shape[1:]) for r in mods]
        pz_mean, pz_std = self._shared_prior()
        pz_std = pz_std.reshape(-1)
        rmeta = self._recon_meta(mods, mods)
        if rescale is not None:
            rmeta = [(d, sc, float(rescale), row) for d, sc, _, row in rmeta]
        meta.update(x=[self._target(inputs, r, rec) for r, rec in zip(mods, recons)],
                    pz_mean=pz_mean.detach().reshape(-1).float().contiguous(),
                    masks=self._stack_masks(inputs, mods), recon=rmeta,
                    latent_kind=C.LATENT[kind], loss_kind=C.LOSS[loss_name], beta=self.beta if beta is None else beta,
                    skip_u_prior=self.skip_u_prior)
        extra = self._extra_lw(U, meta["beta"])
        meta["has_extra"] = extra is not None
        # DReG: the gradient reaching the samples is multiplied once more by wk (mmvaePlus_model.py:330-338): MoESampleFn's backward
        # reads meta["wk"], which the call below fills in
        loss = MoEElboFn.apply(meta, U, W, mu_u, sig_u, mu_w, sig_w, pz_std, *recons, *(() if extra is None else (extra,)))
        return loss, meta


class MMVAEPlus(MoEPlusBase):
    def __init__(self, model_config, encoders=None, decoders=None):
        if model_config.modalities_specific_dim is None:
            raise AttributeError("The modalities_specific_dim attribute must be provided in the model config.")
        super().__init__(model_config, encoders, decoders)
        if model_config.prior_and_posterior_dist not in ("laplace_with_softmax", "normal", "normal_with_softplus"):
            raise AttributeError(" The posterior_dist parameter must be either 'laplace_with_softmax','normal' or "
                                 f"'normal_with_softplus'.  {model_config.prior_and_posterior_dist} was provided.")
        self.mean_priors = nn.ParameterDict()
        self.logvars_priors = nn.ParameterDict()
        self.beta = model_config.beta
        self.modalities_specific_dim = model_config.modalities_specific_dim
        self.reconstruction_option = model_config.reconstruction_option
        self.multiple_latent_spaces = True
        self.style_dims = {m: self.modalities_specific_dim for m in self.encoders}
        Lw, L = model_config.modalities_specific_dim, model_config.latent_dim
        for mod in list(self.encoders.keys()):
            self.mean_priors[mod] = nn.Parameter(torch.zeros(1, Lw), requires_grad=False)
            self.logvars_priors[mod] = nn.Parameter(torch.zeros(1, Lw), requires_grad=model_config.learn_modality_prior)
        self.mean_priors["shared"] = nn.Parameter(torch.zeros(1, L + Lw), requires_grad=False)
        self.logvars_priors["shared"] = nn.Parameter(torch.zeros(1, L + Lw), requires_grad=model_config.learn_shared_prior)
        self.model_name = "MMVAEPlus"
        self.objective = model_config.loss
        self.noise_source = None  # test hook: callable(shape, kind, device) -> standard draws

    def _cross_prior(self, r):
        """Prior of modality r's private code used for cross-modal reconstructions (mmvaePlus_model.py:172-184)."""
        return self.mean_priors[r], self.logvars_priors[r]

    def _shared_prior(self):
        """Prior over cat[u, w] (mmvaePlus_model.py:97-108,249-250): (mean, std); the softmax of laplace_with_softmax runs over
        all L + Lw dimensions of the shared prior."""
        return self.mean_priors["shared"], log_var_to_std(self.logvars_priors["shared"], self.model_config.prior_and_posterior_dist)

    # ---- inference (mmvaePlus_model.py:365-533) ---------------------------------------------------------------------------
    def _style_prior(self, m, batch_size):
        """Prior parameters of modality m's private code for cross-modal generation (:424-437)."""
        if self.reconstruction_option == "single_prior":
            mu_m, lv_m = self.mean_priors[m], self.logvars_priors[m]
        else:   # joint_prior
            mu_m = self.mean_priors["shared"][:, self.latent_dim:]
            lv_m = self.logvars_priors["shared"][:, self.latent_dim:]
        return torch.cat([mu_m] * batch_size, dim=0), torch.cat([lv_m] * batch_size, dim=0)

    def encode(self, inputs, cond_mod="all", N=1, return_mean=False, **kwargs):
        """Shared code from ONE conditioning modality chosen at random (numpy, like the reference) or the mean of the posterior
        means; private codes from the posteriors of the conditioning modalities and from the priors for the others."""
        batch_size = len(list(inputs.data.values())[0])
        cond_mod = super().encode(inputs, cond_mod, N, **kwargs).cond_mod
        with self._nn_ctx():
            outs = {m: self.encoders[m](inputs.data[m]) for m in cond_mod}
        dev = outs[cond_mod[0]].embedding.device
        sample = (lambda mu, sg: mu + sg * self._noise(tuple(mu.shape) if N == 1 else (N,) + tuple(mu.shape), dev))
        if return_mean:
            emb = torch.stack([o.embedding.float() for o in outs.values()]).mean(0)
            z = torch.stack([emb] * N) if N > 1 else emb
        else:
            rm = np.random.choice(cond_mod)
            z = sample(outs[rm].embedding.float(), self._log_var_to_std(outs[rm].log_covariance.float()))
        flatten = kwargs.pop("flatten", False)
        if flatten:
            z = z.reshape(-1, self.latent_dim)
        style_z = {}
        for m in self.encoders:
            if m not in cond_mod:
                mu_m, lv_m = self._style_prior(m, batch_size)
            else:
                mu_m, lv_m = outs[m].style_embedding.float(), outs[m].style_log_covariance.float()
            if return_mean:
                style_z[m] = torch.stack([mu_m] * N) if N > 1 else mu_m
            else:
                style_z[m] = sample(mu_m, self._log_var_to_std(lv_m))
            if flatten:
                style_z[m] = style_z[m].reshape(-1, self.modalities_specific_dim)
        return ModelOutput(z=z, one_latent_space=False, modalities_z=style_z)

    def _encode_many(self, inputs, cond_mod, n):
        """n independent encode() calls (own random conditioning modality each, same numpy stream) in one pass."""
        batch_size = len(list(inputs.data.values())[0])
        cond_mod = BaseMultiVAE.encode(self, inputs, cond_mod, 1).cond_mod
        with self._nn_ctx():
            outs = {m: self.encoders[m](inputs.data[m]) for m in cond_mod}
        dev = outs[cond_mod[0]].embedding.device
        picks = [np.random.choice(cond_mod) for _ in range(n)]
        mu = torch.stack([outs[m].embedding.float() for m in picks])
        sg = torch.stack([self._log_var_to_std(outs[m].log_covariance.float()) for m in picks])
        z = mu + sg * self._noise(tuple(mu.shape), dev)
        style_z = {}
        for m in self.encoders:
            mu_m, lv_m = (self._style_prior(m, batch_size) if m not in cond_mod
                          else (outs[m].style_embedding.float(), outs[m].style_log_covariance.float()))
            style_z[m] = mu_m + self._log_var_to_std(lv_m) * self._noise((n,) + tuple(mu_m.shape), dev)
        return ModelOutput(z=z, one_latent_space=False, modalities_z=style_z)

    def generate_from_prior(self, n_samples, **kwargs):
        kind = self.model_config.prior_and_posterior_dist
        std = log_var_to_std(self.logvars_priors["shared"], kind)
        shape = (n_samples,) + tuple(std.shape) if n_samples > 1 else tuple(std.shape)
        z = self.mean_priors["shared"] + std * self._noise(shape, std.device)
        return ModelOutput(z=z.squeeze(), one_latent_space=True)

    @torch.no_grad()
    def compute_joint_nll(self, inputs, K=1000, batch_size_K=100, reference_quirk=True):
        """-sum_i ln p(x_i) from K // n_modalities importance samples per conditioning modality with rescale = beta = 1
        (mmvaePlus_model.py:478-533), batched over datapoints and chunked over the samples.

        reference_quirk: the reference measures the batch with `len(inputs.data.popitem()[1])` (:497), which REMOVES the last
        modality from the batch before the estimate is computed: the estimate then uses the remaining modalities only (while
        the mixture normaliser stays log n_modalities).  True reproduces those numbers (without mutating the caller's batch);
        False computes the estimator over all modalities."""
        from .containers import MultimodalBaseDataset
        from .elbo import logmeanexp
        self.eval()
        if hasattr(inputs, "masks"):
            raise AttributeError("The compute_joint_nll method is not yet implemented for incomplete datasets.")
        mods = list(inputs.data.keys())
        if reference_quirk:
            mods = mods[:-1]
        sub = MultimodalBaseDataset(data={m: inputs.data[m] for m in mods})
        k_iwae = K // self.n_modalities
        lws = []
        for k0 in range(0, k_iwae, batch_size_K):
            _, meta = self._elbo(sub, min(batch_size_K, k_iwae - k0), "iwae_looser", rescale=1.0, beta=1.0)
            lws.append(meta["lw"])                                   # (C, n, B)
        lw = torch.cat(lws, dim=1)
        # the kernel normalises the mixture-of-experts density by the number of modalities PRESENT; the reference by n_modalities
        lw = lw + (math.log(self.n_modalities) - math.log(len(mods)))
        return -logmeanexp(lw.reshape(-1, lw.shape[-1])).sum()
