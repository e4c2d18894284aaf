"""MVTCAE behind the reference's API (reference: models/mvtcae/mvtcae_model.py:16-169)."""
import torch

from . import _cabi as C
from .base import BaseMultiVAE
from .containers import ModelOutput
from .elbo import PoEFn, ReconNLLFn


class MVTCAE(BaseMultiVAE):
    def __init__(self, model_config, encoders=None, decoders=None):
        super().__init__(model_config, encoders, decoders)
        self.alpha = model_config.alpha
        self.beta = model_config.beta
        self.model_name = "MVTCAE"
        self.noise_source = None

    def forward(self, inputs, **kwargs):
        mods = list(inputs.data.keys())
        dev = inputs.data[mods[0]].device
        with self._nn_ctx():
            outs = [self.encoders[m](inputs.data[m]) for m in mods]
        mu = torch.stack([o.embedding.float() for o in outs])
        lv = torch.stack([o.log_covariance.float() for o in outs])
        M, B, L = mu.shape
        masks = self._stack_masks(inputs, mods)
        noise = (self.noise_source((B, L), "normal", dev) if self.noise_source
                 else torch.randn(B, L, device=dev)).contiguous()
        full = torch.tensor([(1 << M) - 1], dtype=torch.int32, device=dev)
        meta = dict(masks=masks, subsets=full, sel=None, w=None, w_uniform=1.0, noise=noise,
                    prior_mode=0, stable=False, eps=1e-8, want_kldm=True)  # poe without prior expert (:166)
        z, kl_b, kldm = PoEFn.apply(meta, mu, lv)
        results = {}
        joint_kld = kl_b.sum()
        results["joint_divergence"] = joint_kld
        loss_rec = 0
        for i, m in enumerate(mods):
            # the reference iterates self.encoders.keys(); with complete inputs the orders coincide
            with self._nn_ctx():
                rec = self.decoders[m](z).reconstruction
            dist, scale = self.recon_dists[m]
            nll = ReconNLLFn.apply(rec, inputs.data[m].float().contiguous(), None if masks is None else masks[i],
                                   dist, scale, float(self.rescale_factors[m]))
            results[m] = nll.sum()
            loss_rec = loss_rec + results[m]
        kld_losses = 0.0
        for i, m in enumerate(mods):
            results["kld_" + m] = kldm[i].sum()
            kld_losses = kld_losses + results["kld_" + m]
        rec_weight = (self.n_modalities - self.alpha) / self.n_modalities
        cvib_weight = self.alpha / self.n_modalities
        vib_weight = 1 - self.alpha
        total = rec_weight * loss_rec + self.beta * (cvib_weight * kld_losses + vib_weight * joint_kld)
        return ModelOutput(loss=total / B, loss_sum=total, metrics=results)
