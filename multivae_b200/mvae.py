"""MVAE (PoE with prior expert, joint + unimodal + k random subset ELBOs, KL warm-up) behind the
reference's API (reference: models/mvae/mvae_model.py:16-204).  The encoders run ONCE per step; the
reference re-runs them for every subset (compute_mu_log_var_subset, :53-80) with identical results."""
import numpy as np
import torch
from numpy.random import choice

from .base import BaseMultiVAE
from .containers import ModelOutput
from .elbo import PoEFn, ReconNLLFn, poe_joint
from .subsets import mvae_random_subsets, subset_bitmask


class MVAE(BaseMultiVAE):

    def __init__(self, model_config, encoders=None, decoders=None):
        super().__init__(model_config, encoders, decoders)
        self.subsampling = model_config.use_subsampling
        self.k = model_config.k
        if self.n_modalities <= 2:
            self.k = 0
        self.subsets = mvae_random_subsets(self.encoders.keys())
        self.warmup = model_config.warmup
        self.start_keep_best_epoch = model_config.warmup + 1
        self.beta = model_config.beta
        self.model_name = "MVAE"
        self.noise_source = None
        self._beta_dev = None   # device scalar holding the KL weight while the step runs as a CUDA graph

    @property
    def graph_safe(self):
        # the k random subsets are drawn on the host every step (numpy, like the reference): not capturable
        return not (self.subsampling and self.k > 0)

    def _beta_value(self, epoch, batch_ratio):
        """KL warm-up (mvae_model.py:153-156)."""
        return 1 * self.beta if epoch >= self.warmup else (epoch - 1 + batch_ratio) / self.warmup * self.beta

    def prepare_step(self, epoch=1, batch_ratio=0.0):
        """Graph mode: the KL weight lives in a device scalar refreshed here (outside capture / before each replay), so ONE
        captured step serves every (epoch, batch_ratio)."""
        dev = next(self.parameters()).device
        if self._beta_dev is None or self._beta_dev.device != dev:
            self._beta_dev = torch.zeros((), device=dev, dtype=torch.float32)
        self._beta_dev.fill_(float(self._beta_value(epoch, batch_ratio)))

    def forward(self, inputs, **kwargs):
        epoch = kwargs.pop("epoch", 1)
        batch_ratio = kwargs.pop("batch_ratio", 0)
        beta = self._beta_dev if kwargs.pop("beta_on_device", False) else self._beta_value(epoch, batch_ratio)
        order = list(self.encoders.keys())
        subsets = [list(order)]
        if self.subsampling:
            subsets.extend([[m] for m in order])
            if self.k > 0 and self.training:
                for i in choice(np.arange(len(self.subsets)), size=self.k, replace=False):
                    subsets.append(self.subsets[i])
        dev = inputs.data[order[0]].device
        outs = self._run_encoders(inputs, order, dev)
        mu = torch.stack([outs[m].embedding.float() for m in order])
        lv = torch.stack([outs[m].log_covariance.float() for m in order])
        M, B, L = mu.shape
        has_masks = hasattr(inputs, "masks")
        masks = self._stack_masks(inputs, order)
        # phase 1: the subset posteriors and their samples (one fused launch per subset)
        jobs = []   # (subset, n, keep, z, kl_b) or (subset, 0, ...) for a subset no sample has a modality of
        for s in subsets:
            # samples with at least one available modality of the subset (_filter_inputs_with_masks, :115-135)
            if has_masks:
                keep = torch.zeros(B, dtype=torch.bool, device=dev)
                for m in s:
                    keep = keep | inputs.masks[m].bool()
                n = int(keep.sum())
                if n == 0:
                    jobs.append((s, 0, None, None, None))
                    continue
                wrow = keep.float().reshape(1, B).contiguous()
            else:
                keep, n, wrow = None, B, None
            if self.noise_source:
                e = self.noise_source((n, L), "normal", dev)
                noise = torch.zeros(B, L, device=dev)
                if keep is None:
                    noise = e
                else:
                    noise[keep] = e
            else:
                noise = torch.randn(B, L, device=dev)
            bm = subset_bitmask(s, order)
            bits = self._const(("bits", bm, str(dev)), lambda: torch.tensor([bm], dtype=torch.int32, device=dev))
            meta = dict(masks=masks, subsets=bits, sel=None, w=wrow, w_uniform=1.0, noise=noise.contiguous(),
                        prior_mode=1, stable=True, eps=0.0, want_kldm=False)  # prior expert always + stable_poe (:75-79)
            z, kl_b, _ = PoEFn.apply(meta, mu, lv)
            jobs.append((s, n, keep, z, kl_b))
        # phase 2: every (subset, modality) decoder pass, independent of each other (own streams for the small decoders)
        recs = self._run_decoders({(j, m): job[3] for j, job in enumerate(jobs) if job[1] for m in self.decoders if m in job[0]}, dev,
                                  mod_of=lambda k: k[1])
        # phase 3: the subset ELBOs
        total, metrics, len_batch = 0, {}, 0.0
        for j, (s, n, keep, z, kl_b) in enumerate(jobs):
            if n == 0:
                total = total + torch.tensor(0.0, requires_grad=True, device=dev)
                len_batch = 0.0
                continue
            elbo = 0
            for m in self.decoders:
                if m in s:
                    rec = recs[(j, m)]
                    dist, scale = self.recon_dists[m]
                    mk = None
                    if has_masks:
                        mk = (inputs.masks[m].bool() & keep).to(torch.uint8).contiguous()
                    nll = ReconNLLFn.apply(rec, self._target(inputs, m, rec), mk, dist, scale, float(self.rescale_factors[m]))
                    elbo = elbo + nll.sum()
            kld = kl_b.sum()
            elbo = elbo + kld * beta
            key = "_".join(sorted(s))
            metrics[key] = elbo / n
            metrics["beta"] = beta
            metrics["kld" + key] = kld / n
            metrics["recon" + key] = elbo / n  # reference quirk: `recon` aliases the in-place updated elbo (:104-106)
            total = total + elbo / n
            len_batch = n
        return ModelOutput(loss=total, loss_sum=total * len_batch, metrics=metrics)

    # ---- inference (mvae_model.py:53-80, 206-317) -----------------------------------------------------------------------
    def compute_mu_log_var_subset(self, inputs, subset):
        """PoE (prior expert always, stable form) of the posteriors of `subset`; unavailable samples are excluded (:53-80)."""
        subset = list(subset)
        dev = inputs.data[subset[0]].device
        with self._nn_ctx():
            outs = [self.encoders[m](inputs.data[m]) for m in subset]
        mu = torch.stack([o.embedding.float() for o in outs])
        lv = torch.stack([o.log_covariance.float() for o in outs])
        M = len(subset)
        full = self._const(("full", M, str(dev)), lambda: torch.tensor([(1 << M) - 1], dtype=torch.int32, device=dev))
        return poe_joint(mu, lv, self._stack_masks(inputs, subset), full, 1, True, 0.0)

    def encode(self, inputs, cond_mod="all", N=1, return_mean=False, **kwargs):
        cond_mod = super().encode(inputs, cond_mod, N, **kwargs).cond_mod
        mu, lv = self.compute_mu_log_var_subset(inputs, cond_mod)
        z = self._rsample_gaussian(mu, lv, N=N, return_mean=return_mean, flatten=kwargs.pop("flatten", False))
        return ModelOutput(z=z, one_latent_space=True)

    @torch.no_grad()
    def compute_joint_nll(self, inputs, K=1000, batch_size_K=100):
        """mvae_model.py:241-317, batched over datapoints and samples."""
        from .mvtcae import _gaussian_iw_nll
        self.eval()
        if hasattr(inputs, "masks"):
            raise AttributeError("The compute_joint_nll method is not yet implemented for incomplete datasets.")
        mu, lv = self.compute_mu_log_var_subset(inputs, list(self.encoders.keys()))
        return _gaussian_iw_nll(self, inputs, mu, lv, K, batch_size_K)
