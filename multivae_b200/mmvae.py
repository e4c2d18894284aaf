"""MMVAE (mixture of experts, single latent) behind the reference's API
(reference: models/mmvae/mmvae_model.py:16-292).  Same fused kernels as MMVAE+ with Lw = 0 and beta
fixed to 1 in lw (the reference never uses config.beta in the loss, mmvae_model.py:228)."""
import torch
import torch.nn as nn

from . import _cabi as C
from .base import BaseMultiVAE
from .containers import ModelOutput, drop_unused_modalities
from .elbo import MoEElboFn, log_var_to_std, standard_noise


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
        kwargs.pop("compute_loss", True)
        kwargs.pop("detailed_output", False)
        K = kwargs.pop("K", self.model_config.K)
        if self.model_config.loss not in C.LOSS:
            raise NotImplementedError()
        kind = self.model_config.prior_and_posterior_dist
        detach = self.model_config.loss == "dreg_looser"
        mods = list(inputs.data.keys())
        dev = inputs.data[mods[0]].device
        B = len(inputs.data[mods[0]])
        mus, sigs, zs = [], [], []
        for c in mods:
            with self._nn_ctx():
                o = self.encoders[c](inputs.data[c])
            s = log_var_to_std(o.log_covariance.float(), kind)
            mus.append(o.embedding.float()); sigs.append(s)
            zs.append(mus[-1] + s * self._noise((K, B, s.shape[-1]), dev))
        Z = torch.stack(zs)
        recons = []
        for r in mods:
            with self._nn_ctx():
                rec = self.decoders[r](Z.reshape(-1, Z.shape[-1]))["reconstruction"]
            recons.append(rec.reshape(len(mods), K, B, *rec.shape[1:]))
        pz_std = log_var_to_std(self.prior_log_var, kind).reshape(-1)
        meta = dict(x=[inputs.data[r].float().contiguous() for r in mods],
                    pz_mean=self.prior_mean.detach().reshape(-1).float().contiguous(),
                    masks=self._stack_masks(inputs, mods), recon=self._recon_meta(mods, mods),
                    latent_kind=C.LATENT[kind], loss_kind=C.LOSS[self.model_config.loss], beta=1.0, detach=detach)
        loss = MoEElboFn.apply(meta, Z, None, torch.stack(mus), torch.stack(sigs), None, None, pz_std, *recons)
        if detach and Z.requires_grad:
            wk = meta["wk"].unsqueeze(-1)
            Z.register_hook(lambda g: g * wk)
        self._last = meta
        return ModelOutput(loss=loss, loss_sum=loss, metrics={})
