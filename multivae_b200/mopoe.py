"""MoPoE behind the reference's API (reference: models/mopoe/mopoe_model.py:21-465): 2^M-1 subset PoEs,
per-sample mixture component selection, weighted analytic KL — one fused kernel instead of the
reference's per-subset cat/poe loop."""
import math

import torch

from .base import BaseMultiVAE
from .containers import ModelOutput
from .elbo import PoEFn, ReconNLLFn, logmeanexp, normal_logpdf_sum, poe_joint
from .nn.default_architectures import BaseDictDecodersMultiLatents, BaseDictEncoders_MultiLatents
from .subsets import all_subsets, deterministic_selection, subset_bitmask


class MoPoE(BaseMultiVAE):
    def __init__(self, model_config, encoders=None, decoders=None):
        super().__init__(model_config, encoders, decoders)
        self.multiple_latent_spaces = model_config.modalities_specific_dim is not None
        self.model_name = "MoPoE"
        if self.multiple_latent_spaces:   # mopoe_model.py:57-74
            self.style_dims = model_config.modalities_specific_dim
            if encoders is None:
                self.set_encoders(BaseDictEncoders_MultiLatents(model_config.input_dims, model_config.latent_dim,
                                                                model_config.modalities_specific_dim))
            if decoders is None:
                self.set_decoders(BaseDictDecodersMultiLatents(model_config.input_dims, model_config.latent_dim,
                                                               model_config.modalities_specific_dim))
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
        enc_out = self._run_encoders(inputs, order, dev)
        outs = [enc_out[m] for m in order]
        mu = torch.stack([o.embedding.float() for o in outs])
        lv = torch.stack([o.log_covariance.float() for o in outs])
        M, B, L = mu.shape
        table = self.subset_table()
        S = len(table)
        bits = self._const(("bits", str(dev)), lambda: torch.tensor(table, dtype=torch.int32, device=dev))
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
            sel = self._const(("sel", B, S, str(dev)), lambda: deterministic_selection(B, S).to(dev))
            w, w_uniform = None, 1 / float(S)
        noise = (self.noise_source((B, L), "normal", dev) if self.noise_source else torch.randn(B, L, device=dev))
        meta = dict(masks=None, subsets=bits, sel=sel, w=w, w_uniform=w_uniform, noise=noise.contiguous(),
                    prior_mode=2, stable=False, eps=1e-8, want_kldm=False)  # prior expert only for the full subset (:252-261)
        z, kl_b, _ = PoEFn.apply(meta, mu, lv)
        results = {}
        kld = kl_b.mean()
        loss = 0
        one = self._const(("one", str(dev)), lambda: torch.tensor([1], dtype=torch.int32, device=dev))
        z_by_mod, mrows = {}, {}
        for i, m in enumerate(order):
            full = z
            mrows[m] = inputs.masks[m].to(torch.uint8).contiguous() if hasattr(inputs, "masks") else None
            if self.multiple_latent_spaces:
                # modality-specific latent: sample + analytic KL(q(w_m|x_m) || N(0,I)) in one fused launch (a one-expert "PoE"
                # without prior expert and eps = 0 is the Gaussian itself) - mopoe_model.py:171-225
                if not hasattr(outs[i], "style_embedding"):
                    raise AttributeError(" model_config.modality_specific_dims is not None, "
                                         f"but encoder output for modality {m} doesn't have a style_embedding attribute. ")
                smu, slv = outs[i].style_embedding.float(), outs[i].style_log_covariance.float()
                sn = (self.noise_source(tuple(smu.shape), "normal", dev) if self.noise_source
                      else torch.randn(smu.shape, device=dev)).contiguous()
                smeta = dict(masks=None, subsets=one, sel=None, w=None, w_uniform=1.0, noise=sn, prior_mode=0, stable=False,
                             eps=0.0, want_kldm=False)
                zs, skl, _ = PoEFn.apply(smeta, smu.unsqueeze(0), slv.unsqueeze(0))
                full = torch.cat([z, zs], dim=-1)
                if mrows[m] is not None:
                    skl = skl * inputs.masks[m].float()
                kld = kld + skl.mean() * self.model_config.beta_style
            z_by_mod[m] = full
        recs = self._run_decoders(z_by_mod, dev)
        for m in order:
            rec = recs[m]
            dist, scale = self.recon_dists[m]
            nll = ReconNLLFn.apply(rec, self._target(inputs, m, rec), mrows[m], dist, scale, float(self.rescale_factors[m]))
            results["recon_" + m] = nll.mean()
            loss = loss + results["recon_" + m]
        # the reference accumulates the style KLs in place into the tensor it also reports as "joint_divergence" (:166,222)
        results = {"joint_divergence": kld, **results}
        loss = loss + self.model_config.beta * kld
        self._last = dict(sel=sel, subsets=bits)
        return ModelOutput(loss=loss, loss_sum=loss * B, metrics=results)

    # ---- inference (mopoe_model.py:274-416, 468-717) ------------------------------------------------------------------
    def _unimodal(self, inputs, mods):
        with self._nn_ctx():
            outs = {m: self.encoders[m](inputs.data[m]) for m in mods}
        return outs

    def _subset_posterior(self, outs, mods, dev):
        """PoE (prior expert only for the full set of modalities, eps 1e-8) of the unimodal posteriors of `mods` (_poe_fusion)."""
        mu = torch.stack([outs[m].embedding.float() for m in mods])
        lv = torch.stack([outs[m].log_covariance.float() for m in mods])
        n = len(mods)
        full = self._const(("full", n, str(dev)), lambda: torch.tensor([(1 << n) - 1], dtype=torch.int32, device=dev))
        # prior_mode 2 adds the prior expert when popcount(subset) == M of the launch: here only when all modalities take part
        return poe_joint(mu, lv, None, full, 2 if n == self.n_modalities else 0, False, 1e-8)

    def encode(self, inputs, cond_mod="all", N=1, return_mean=False, **kwargs):
        cond_mod = super().encode(inputs, cond_mod, N, **kwargs).cond_mod
        key = "_".join(sorted(cond_mod))
        if key not in self.subsets:
            raise KeyError(key)
        dev = inputs.data[cond_mod[0]].device
        need_all = return_mean and len(cond_mod) == self.n_modalities
        outs = self._unimodal(inputs, list(self.encoders.keys()) if need_all else list(self.subsets[key]))
        mu, lv = self._subset_posterior(outs, self.subsets[key], dev)
        if need_all:   # aggregate: mean over the subset posteriors' means (:388-392)
            mu = torch.stack([self._subset_posterior(outs, v, dev)[0] for k, v in self.subsets.items() if k != ""]).mean(0)
        flatten = kwargs.pop("flatten", False)
        z = self._rsample_gaussian(mu, lv, N, return_mean, flatten=flatten)
        if not self.multiple_latent_spaces:
            return ModelOutput(z=z, one_latent_space=True)
        modalities_z = {}
        for m in self.encoders:
            if m in cond_mod:
                smu, slv = outs[m].style_embedding.float(), outs[m].style_log_covariance.float()
            else:
                smu = torch.zeros(len(mu), self.style_dims[m], device=dev)
                slv = torch.zeros_like(smu)
            modalities_z[m] = self._rsample_gaussian(smu, slv, N, return_mean, flatten)
        return ModelOutput(z=z, one_latent_space=False, modalities_z=modalities_z)

    def _iw_nll(self, inputs, outs, mu, lv, K, batch_size_K, lq_shared):
        """ln p(X) with K samples from N(mu, e^lv) (+ the private posteriors), proposal density lq_shared(z) for the shared code."""
        mods = list(inputs.data.keys())
        B, L = mu.shape
        dev = mu.device
        z_all = mu + torch.exp(0.5 * lv) * self._draw((K, B, L), "normal", dev)
        priv = {}
        if self.multiple_latent_spaces:
            for m in mods:
                smu, slv = outs[m].style_embedding.float(), outs[m].style_log_covariance.float()
                priv[m] = (smu, slv, smu + torch.exp(0.5 * slv) * self._draw((K,) + tuple(smu.shape), "normal", dev))
        lws = []
        for k0 in range(0, K, batch_size_K):
            z = z_all[k0:k0 + batch_size_K]
            n = z.shape[0]
            z_of = (lambda m, z=z: z) if not priv else (lambda m, z=z, k0=k0: torch.cat([z, priv[m][2][k0:k0 + z.shape[0]]], dim=-1))
            lw = self._iw_lpx(inputs, z_of, n, mods) + normal_logpdf_sum(z, torch.zeros_like(mu), torch.zeros_like(lv)) - lq_shared(z)
            for m in priv:
                smu, slv, zp = priv[m]
                zp = zp[k0:k0 + n]
                lw = lw + normal_logpdf_sum(zp, torch.zeros_like(smu), torch.zeros_like(slv)) - normal_logpdf_sum(zp, smu, slv)
            lws.append(lw)
        return -logmeanexp(torch.cat(lws, dim=0)).sum()

    @torch.no_grad()
    def compute_joint_nll(self, inputs, K=1000, batch_size_K=100):
        """-sum_i ln p(x_i): samples from the per-sample selected subset posterior, proposal density = the uniform mixture of
        ALL subset posteriors (mopoe_model.py:468-595); batched over datapoints."""
        self.eval()
        if hasattr(inputs, "masks"):
            raise AttributeError("The compute_joint_nll method is not yet implemented for incomplete datasets.")
        order = list(self.encoders.keys())
        dev = inputs.data[order[0]].device
        outs = self._unimodal(inputs, order)
        post = [self._subset_posterior(outs, v, dev) for k, v in self.subsets.items() if k != ""]
        mus, lvs = torch.stack([p[0] for p in post]), torch.stack([p[1] for p in post])      # (S, B, L)
        S, B, L = mus.shape
        sel = self._const(("sel", B, S, str(dev)), lambda: deterministic_selection(B, S).to(dev)).long()
        ar = torch.arange(B, device=dev)
        mu, lv = mus[sel, ar], lvs[sel, ar]

        def lq(z):   # log-mean-exp over the S subset posteriors, one subset at a time (bounded temporaries)
            acc = torch.stack([normal_logpdf_sum(z, mus[s], lvs[s]) for s in range(S)])
            return torch.logsumexp(acc, dim=0) - math.log(S)

        return self._iw_nll(inputs, outs, mu, lv, K, batch_size_K, lq)

    @torch.no_grad()
    def _compute_joint_nll_from_subset_encoding(self, subset, inputs, K=1000, batch_size_K=100):
        """Same estimator with the PoE posterior of `subset` as proposal (mopoe_model.py:597-701)."""
        self.eval()
        if hasattr(inputs, "masks"):
            raise AttributeError("The compute_joint_nll method is not yet implemented for incomplete datasets.")
        order = list(self.encoders.keys())
        dev = inputs.data[order[0]].device
        outs = self._unimodal(inputs, order)
        mu, lv = self._subset_posterior(outs, self.subsets["_".join(sorted(subset))], dev)
        return self._iw_nll(inputs, outs, mu, lv, K, batch_size_K, lambda z: normal_logpdf_sum(z, mu, lv))

    def compute_joint_nll_paper(self, inputs, K=1000, batch_size_K=100):
        """The original paper's estimator: PoE of all modalities as proposal (mopoe_model.py:703-717)."""
        return self._compute_joint_nll_from_subset_encoding(list(self.encoders.keys()), inputs, K, batch_size_K)
