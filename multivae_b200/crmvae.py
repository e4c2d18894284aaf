"""CRMVAE behind the reference's API (reference: models/crmvae/crmvae_model.py:14-295): the MVTCAE aggregation (product of the
unimodal experts, KL(q(z|X) || N(0,I)) and KL(q(z|X) || q(z|x_m)) in ONE fused kernel, mv_poe_fwd / mv_poe_bwd) plus unimodal
reconstruction terms from samples of every q(z|x_m)."""
import torch

from .base import BaseMultiVAE
from .containers import ModelOutput, MultimodalBaseDataset
from .elbo import PoEFn, ReconNLLFn, poe_joint


class CRMVAE(BaseMultiVAE):
    def __init__(self, model_config, encoders=None, decoders=None):
        super().__init__(model_config, encoders, decoders)
        self.model_name = "CRMVAE"
        self.noise_source = None

    def _draw_n(self, shape, dev):
        return (self.noise_source(tuple(shape), "normal", dev) if self.noise_source else torch.randn(shape, device=dev)).contiguous()

    def forward(self, inputs, **kwargs):
        mods = list(inputs.data.keys())
        dev = inputs.data[mods[0]].device
        enc_out = self._run_encoders(inputs, mods, dev)
        outs = [enc_out[m] for m in mods]
        mu = torch.stack([o.embedding.float() for o in outs])
        lv = torch.stack([o.log_covariance.float() for o in outs])
        M, B, L = mu.shape
        masks = self._stack_masks(inputs, mods)
        full = self._const(("full", M, str(dev)), lambda: torch.tensor([(1 << M) - 1], dtype=torch.int32, device=dev))
        # joint posterior = PoE of the available experts (no prior expert, eps 1e-8), its sample, KL to N(0, I) and to every expert
        meta = dict(masks=masks, subsets=full, sel=None, w=None, w_uniform=1.0, noise=self._draw_n((B, L), dev), prior_mode=0,
                    stable=False, eps=1e-8, want_kldm=True)
        z_joint, joint_kld, kldm = PoEFn.apply(meta, mu, lv)
        results = {"joint_divergence": joint_kld.mean()}
        divergence = joint_kld
        z_samples = {"joint": z_joint}
        for i, m in enumerate(mods):
            # sample of the (unmasked) unimodal posterior q(z | x_m) (crmvae_model.py:68-71)
            z_samples[m] = mu[i] + torch.exp(0.5 * lv[i]) * self._draw_n((B, L), dev)
            divergence = divergence + kldm[i]            # masked samples contribute 0 (the kernel skips unavailable experts)
            results[f"kl_{m}"] = kldm[i].mean()
        loss_rec = 0
        for i, g in enumerate(self.decoders.keys()):
            for src in ("joint", g):
                with self._nn_ctx():
                    rec = self._logits(self.decoders[g](z_samples[src]).reconstruction)
                dist, scale = self.recon_dists[g]
                mrow = None if masks is None else masks[mods.index(g)]
                m_rec = ReconNLLFn.apply(rec, self._target(inputs, g, rec), mrow, dist, scale, float(self.rescale_factors[g]))
                loss_rec = loss_rec + m_rec
                results[f"recon_{g}_from_{src}"] = m_rec.mean()
        loss_rec = loss_rec / (2 * (self.n_modalities + 1))
        divergence = divergence / (self.n_modalities + 1)
        total = (loss_rec + self.model_config.beta * divergence).sum()
        return ModelOutput(loss=total, loss_sum=total, metrics=results)

    # ---- inference (crmvae_model.py:180-295) --------------------------------------------------------------------------
    def _joint_posterior(self, inputs, mods):
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
        flatten = kwargs.pop("flatten", False)
        cond_inputs = MultimodalBaseDataset(data={k: inputs.data[k] for k in cond_mod})
        mu, lv = self._joint_posterior(cond_inputs, list(cond_mod))
        return ModelOutput(z=self._rsample_gaussian(mu, lv, N=N, return_mean=return_mean, flatten=flatten), one_latent_space=True)

    @torch.no_grad()
    def compute_joint_nll(self, inputs, K=1000, batch_size_K=100):
        from .mvtcae import _gaussian_iw_nll
        self.eval()
        if hasattr(inputs, "masks"):
            raise AttributeError("The compute_joint_nll method is not yet implemented for incomplete datasets.")
        mu, lv = self._joint_posterior(inputs, list(inputs.data.keys()))
        return _gaussian_iw_nll(self, inputs, mu, lv, K, batch_size_K)
