"""MoPoE behind the reference's API (reference: models/mopoe/mopoe_model.py:21-465): 2^M-1 subset PoEs,
per-sample mixture component selection, weighted analytic KL — one fused kernel instead of the
reference's per-subset cat/poe loop."""
import torch

from .base import BaseMultiVAE
from .containers import ModelOutput
from .elbo import PoEFn, ReconNLLFn
from .subsets import all_subsets, deterministic_selection, subset_bitmask


class MoPoE(BaseMultiVAE):
    def __init__(self, model_config, encoders=None, decoders=None):
        super().__init__(model_config, encoders, decoders)
        self.multiple_latent_spaces = model_config.modalities_specific_dim is not None
        if self.multiple_latent_spaces:
            raise NotImplementedError("MoPoE with modality-specific latent spaces is not covered yet")
        self.model_name = "MoPoE"
        list_subsets = self.model_config.subsets
        if isinstance(list_subsets, dict):
            list_subsets = list(list_subsets.values())
        if list_subsets is None:
            self.subsets = all_subsets(self.encoders.keys())
        else:
            self.set_subsets(list_subsets)
        self.model_config.subsets = self.subsets
        self.noise_source = None
        self.choice_source = None  # test hook: callable(probs (B,S)) -> (B,) int subset index

    def set_subsets(self, subsets_list):
        subsets = {}
        for mod_names in subsets_list:
            for mod_name in sorted(mod_names):
                if (mod_name not in self.encoders.keys()) and (mod_name != ""):
                    raise AttributeError(f"The provided subsets list contains unknown modality name {mod_name}."
                                         " that is not the encoders dictionary or inputs_dim dictionary.")
            subsets["_".join(sorted(mod_names))] = sorted(mod_names)
        self.subsets = subsets

    def subset_table(self):
        order = list(self.encoders.keys())
        return [subset_bitmask(v, order) for k, v in self.subsets.items() if k != ""]

    def forward(self, inputs, **kwargs):
        order = list(self.encoders.keys())
        dev = inputs.data[order[0]].device
        with self._nn_ctx():
            outs = [self.encoders[m](inputs.data[m]) for m in order]
        mu = torch.stack([o.embedding.float() for o in outs])
        lv = torch.stack([o.log_covariance.float() for o in outs])
        M, B, L = mu.shape
        table = self.subset_table()
        S = len(table)
        bits = torch.tensor(table, dtype=torch.int32, device=dev)
        if hasattr(inputs, "masks"):
            mk = torch.stack([inputs.masks[m].bool() for m in order])  # (M,B)
            tb = torch.tensor([[(t >> i) & 1 for i in range(M)] for t in table], dtype=torch.bool, device=dev)  # (S,M)
            avail = (~tb.unsqueeze(-1) | mk.unsqueeze(0)).all(dim=1).float()  # (S,B): all members available
            w = (avail / avail.sum(0)).contiguous()
            probs = w.permute(1, 0)
            if self.choice_source is not None:
                sel = self.choice_source(probs)
            else:
                sel = torch.distributions.OneHotCategorical(probs=probs).sample().argmax(-1)
            sel = sel.to(torch.int32).to(dev).contiguous()
            w_uniform = 0.0
        else:
            sel = deterministic_selection(B, S).to(dev)
            w, w_uniform = None, 1 / float(S)
        noise = (self.noise_source((B, L), "normal", dev) if self.noise_source else torch.randn(B, L, device=dev))
        meta = dict(masks=None, subsets=bits, sel=sel, w=w, w_uniform=w_uniform, noise=noise.contiguous(),
                    prior_mode=2, stable=False, eps=1e-8, want_kldm=False)  # prior expert only for the full subset (:252-261)
        z, kl_b, _ = PoEFn.apply(meta, mu, lv)
        results = {"joint_divergence": kl_b.mean()}
        loss = 0
        for i, m in enumerate(order):
            with self._nn_ctx():
                rec = self.decoders[m](z).reconstruction
            dist, scale = self.recon_dists[m]
            mrow = inputs.masks[m].to(torch.uint8).contiguous() if hasattr(inputs, "masks") else None
            nll = ReconNLLFn.apply(rec, inputs.data[m].float().contiguous(), mrow, dist, scale,
                                   float(self.rescale_factors[m]))
            results["recon_" + m] = nll.mean()
            loss = loss + results["recon_" + m]
        loss = loss + self.model_config.beta * results["joint_divergence"]
        self._last = dict(sel=sel, subsets=bits)
        return ModelOutput(loss=loss, loss_sum=loss * B, metrics=results)
