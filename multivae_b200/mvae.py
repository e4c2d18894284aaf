"""MVAE (PoE with prior expert, joint + unimodal + k random subset ELBOs, KL warm-up) behind the
reference's API (reference: models/mvae/mvae_model.py:16-204).  The encoders run ONCE per step; the
reference re-runs them for every subset (compute_mu_log_var_subset, :53-80) with identical results."""
import numpy as np
import torch
from numpy.random import choice

from .base import BaseMultiVAE
from .containers import ModelOutput
from .elbo import PoEFn, ReconNLLFn
from .subsets import mvae_random_subsets, subset_bitmask


class MVAE(BaseMultiVAE):
    step_depends_on_epoch = True  # KL warm-up reads (epoch, batch_ratio): a captured step is only valid for one value

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

    def forward(self, inputs, **kwargs):
        epoch = kwargs.pop("epoch", 1)
        batch_ratio = kwargs.pop("batch_ratio", 0)
        beta = 1 * self.beta if epoch >= self.warmup else (epoch - 1 + batch_ratio) / self.warmup * self.beta
        order = list(self.encoders.keys())
        subsets = [list(order)]
        if self.subsampling:
            subsets.extend([[m] for m in order])
            if self.k > 0 and self.training:
                for i in choice(np.arange(len(self.subsets)), size=self.k, replace=False):
                    subsets.append(self.subsets[i])
        dev = inputs.data[order[0]].device
        with self._nn_ctx():
            outs = {m: self.encoders[m](inputs.data[m]) for m in order}
        mu = torch.stack([outs[m].embedding.float() for m in order])
        lv = torch.stack([outs[m].log_covariance.float() for m in order])
        M, B, L = mu.shape
        has_masks = hasattr(inputs, "masks")
        masks = self._stack_masks(inputs, order)
        total, metrics, len_batch = 0, {}, 0.0
        for s in subsets:
            # samples with at least one available modality of the subset (_filter_inputs_with_masks, :115-135)
            if has_masks:
                keep = torch.zeros(B, dtype=torch.bool, device=dev)
                for m in s:
                    keep = keep | inputs.masks[m].bool()
                n = int(keep.sum())
                if n == 0:
                    total = total + torch.tensor(0.0, requires_grad=True, device=dev)
                    len_batch = 0.0
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
            bits = torch.tensor([subset_bitmask(s, order)], dtype=torch.int32, device=dev)
            meta = dict(masks=masks, subsets=bits, sel=None, w=wrow, w_uniform=1.0, noise=noise.contiguous(),
                        prior_mode=1, stable=True, eps=0.0, want_kldm=False)  # prior expert always + stable_poe (:75-79)
            z, kl_b, _ = PoEFn.apply(meta, mu, lv)
            elbo = 0
            for m in self.decoders:
                if m in s:
                    with self._nn_ctx():
                        rec = self.decoders[m](z).reconstruction
                    dist, scale = self.recon_dists[m]
                    mk = None
                    if has_masks:
                        mk = (inputs.masks[m].bool() & keep).to(torch.uint8).contiguous()
                    nll = ReconNLLFn.apply(rec, inputs.data[m].float().contiguous(), mk, dist, scale,
                                           float(self.rescale_factors[m]))
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
