"""MVTCAE behind the reference's API (reference: models/mvtcae/mvtcae_model.py:16-169)."""
import torch

from . import _cabi as C
from .base import BaseMultiVAE
from .containers import ModelOutput
from .elbo import PoEFn, ReconNLLFn, logmeanexp, normal_logpdf_sum, poe_joint


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
        enc_out = self._run_encoders(inputs, mods, dev)
        outs = [enc_out[m] for m in mods]
        mu = torch.stack([o.embedding.float() for o in outs])
        lv = torch.stack([o.log_covariance.float() for o in outs])
        M, B, L = mu.shape
        masks = self._stack_masks(inputs, mods)
        noise = (self.noise_source((B, L), "normal", dev) if self.noise_source
                 else torch.randn(B, L, device=dev)).contiguous()
        full = self._const(("full", M, str(dev)), lambda: torch.tensor([(1 << M) - 1], dtype=torch.int32, device=dev))
        meta = dict(masks=masks, subsets=full, sel=None, w=None, w_uniform=1.0, noise=noise,
                    prior_mode=0, stable=False, eps=1e-8, want_kldm=True)  # poe without prior expert (:166)
        z, kl_b, kldm = PoEFn.apply(meta, mu, lv)
        results = {}
        joint_kld = kl_b.sum()
        results["joint_divergence"] = joint_kld
        loss_rec = 0
        recs = self._run_decoders({m: z for m in mods}, dev)
        for i, m in enumerate(mods):
            # the reference iterates self.encoders.keys(); with complete inputs the orders coincide
            rec = recs[m]
            dist, scale = self.recon_dists[m]
            nll = ReconNLLFn.apply(rec, self._target(inputs, m, rec), None if masks is None else masks[i],
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

    # ---- inference (mvtcae_model.py:134-289) --------------------------------------------------------------------------
    def _joint_posterior(self, inputs, mods):
        """(mu, lv) of the PoE (no prior expert, eps 1e-8) of the unimodal posteriors of `mods` (_inference, :134-169)."""
        dev = inputs.data[mods[0]].device
        with self._nn_ctx():
            outs = [self.encoders[m](inputs.data[m]) for m in mods]
        mu = torch.stack([o.embedding.float() for o in outs])
        lv = torch.stack([o.log_covariance.float() for o in outs])
        M = len(mods)
        full = self._const(("full", M, str(dev)), lambda: torch.tensor([(1 << M) - 1], dtype=torch.int32, device=dev))
        return poe_joint(mu, lv, self._stack_masks(inputs, mods), full, 0, False, 1e-8)

    def encode(self, inputs, cond_mod="all", N=1, return_mean=False, **kwargs):
        cond_mod = super().encode(inputs, cond_mod, N, **kwargs).cond_mod
        from .containers import MultimodalBaseDataset
        cond_inputs = MultimodalBaseDataset(data={k: inputs.data[k] for k in cond_mod})   # masks dropped, like the reference (:201-203)
        mu, lv = self._joint_posterior(cond_inputs, list(cond_mod))
        z = self._rsample_gaussian(mu, lv, N=N, return_mean=return_mean, flatten=kwargs.pop("flatten", False))
        return ModelOutput(z=z, one_latent_space=True)

    @torch.no_grad()
    def compute_joint_nll(self, inputs, K=1000, batch_size_K=100):
        """-sum_i ln p(x_i) with K importance samples from the joint posterior (mvtcae_model.py:214-289), batched: the
        decoders run on chunks of batch_size_K * B rows instead of one datapoint at a time."""
        self.eval()
        if hasattr(inputs, "masks"):
            raise AttributeError("The compute_joint_nll method is not yet implemented for incomplete datasets.")
        mods = list(inputs.data.keys())
        mu, lv = self._joint_posterior(inputs, mods)
        return _gaussian_iw_nll(self, inputs, mu, lv, K, batch_size_K)


def _gaussian_iw_nll(model, inputs, mu, lv, K, batch_size_K, lq_fn=None):
    """Shared by MVTCAE / MVAE / MoPoE: z_k ~ N(mu, e^lv), lw = sum_m ln p(x_m|z) + ln N(z;0,I) - ln q(z|X), ll = logmeanexp_k.
    lq_fn(z) overrides the proposal density (MoPoE: mixture over subsets)."""
    B, L = mu.shape
    z_all = mu + torch.exp(0.5 * lv) * model._draw((K, B, L), "normal", mu.device)   # one draw of K samples, like rsample([K])
    lws = []
    for k0 in range(0, K, batch_size_K):
        z = z_all[k0:k0 + batch_size_K]
        lpx = model._iw_lpx(inputs, lambda m, z=z: z, z.shape[0])
        lpz = normal_logpdf_sum(z, torch.zeros_like(mu), torch.zeros_like(lv))
        lq = normal_logpdf_sum(z, mu, lv) if lq_fn is None else lq_fn(z)
        lws.append(lpx + lpz - lq)
    return -logmeanexp(torch.cat(lws, dim=0)).sum()
