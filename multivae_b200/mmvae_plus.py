"""MMVAE+ behind the reference's API (reference: models/mmvaePlus/mmvaePlus_model.py:28-363).

forward(inputs, **kwargs) -> ModelOutput(loss, loss_sum, metrics={}).  Differences in *how*, not what:
the M*M decoder invocations of the reference are batched into one call per decoder over the M*K*B rows
of all conditioning modalities, and _compute_k_lws + the IWAE/DReG looser are three fused CUDA kernels
(multivae_b200/csrc/elbo_moe.cu) instead of ~10^4 ATen calls."""
import torch
import torch.nn as nn

from . import _cabi as C
from .base import BaseMultiVAE
from .containers import ModelOutput, drop_unused_modalities
from .elbo import MoEElboFn, log_var_to_std, standard_noise
from .nn.default_architectures import BaseDictDecodersMultiLatents, BaseDictEncoders_MultiLatents


class MMVAEPlus(BaseMultiVAE):
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

    def default_encoders(self, model_config):
        return BaseDictEncoders_MultiLatents(model_config.input_dims, model_config.latent_dim,
                                             {m: model_config.modalities_specific_dim for m in model_config.input_dims})

    def default_decoders(self, model_config):
        return BaseDictDecodersMultiLatents(model_config.input_dims, model_config.latent_dim,
                                            {m: model_config.modalities_specific_dim for m in model_config.input_dims})

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

    def forward(self, inputs, **kwargs):
        if self.objective not in C.LOSS:
            raise NotImplementedError()
        inputs = drop_unused_modalities(inputs)
        kind = self.model_config.prior_and_posterior_dist
        K = kwargs.pop("K", self.model_config.K)
        detach = self.objective == "dreg_looser"
        mods = list(inputs.data.keys())
        dev = inputs.data[mods[0]].device
        B = len(inputs.data[mods[0]])

        # encoders + reparameterised samples, in the reference's noise-consumption order (per cond modality:
        # u, w, then one prior draw per *other* modality: mmvaePlus_model.py:136-186)
        mu_u, sig_u, mu_w, sig_w, u, w, w_cross = [], [], [], [], [], [], {}
        L_, Lw_ = self.model_config.latent_dim, self.modalities_specific_dim
        self._begin_noise_pool(len(mods) * K * B * (L_ + Lw_ + (len(mods) - 1) * Lw_), dev)
        for c in mods:
            with self._nn_ctx():
                o = self.encoders[c](inputs.data[c])
            su = log_var_to_std(o.log_covariance.float(), kind)
            sw = log_var_to_std(o.style_log_covariance.float(), kind)
            mu_u.append(o.embedding.float()); sig_u.append(su)
            mu_w.append(o.style_embedding.float()); sig_w.append(sw)
            u.append(mu_u[-1] + su * self._noise((K, B, su.shape[-1]), dev))
            w.append(mu_w[-1] + sw * self._noise((K, B, sw.shape[-1]), dev))
            for r in mods:
                if r != c:
                    sp = log_var_to_std(self.logvars_priors[r], kind)
                    w_cross[(c, r)] = self.mean_priors[r] + sp * self._noise((K, B, sp.shape[-1]), dev)
        self._noise_pool = None
        U, W = torch.stack(u), torch.stack(w)  # (C,K,B,L), (C,K,B,Lw)

        # one batched decoder call per reconstructed modality over all conditioning modalities
        recons = []
        for r in mods:
            wz = torch.stack([W[i] if c == r else w_cross[(c, r)] for i, c in enumerate(mods)])
            z = torch.cat([U, wz], dim=-1)
            with self._nn_ctx():
                rec = self.decoders[r](z.reshape(-1, z.shape[-1]))["reconstruction"]
            recons.append(rec.reshape(len(mods), K, B, *rec.shape[1:]))

        pz_std = log_var_to_std(self.logvars_priors["shared"], kind).reshape(-1)
        meta = dict(x=[inputs.data[r].float().contiguous() for r in mods],
                    pz_mean=self.mean_priors["shared"].detach().reshape(-1).float().contiguous(),
                    masks=self._stack_masks(inputs, mods), recon=self._recon_meta(mods, mods),
                    latent_kind=C.LATENT[kind], loss_kind=C.LOSS[self.objective], beta=self.beta, detach=detach)
        loss = MoEElboFn.apply(meta, U, W, torch.stack(mu_u), torch.stack(sig_u), torch.stack(mu_w),
                               torch.stack(sig_w), pz_std, *recons)
        if detach:
            # DReG: the gradient reaching the samples is multiplied once more by wk (mmvaePlus_model.py:330-338)
            wk = meta["wk"].unsqueeze(-1)
            if U.requires_grad:
                U.register_hook(lambda g: g * wk)
                W.register_hook(lambda g: g * wk)
        self._last = meta
        return ModelOutput(loss=loss, loss_sum=loss, metrics=dict())
