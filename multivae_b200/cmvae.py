"""CMVAE behind the reference's API (reference: models/cmvae/cmvae_model.py:30-560): MMVAE+ with a mixture-of-clusters prior
over the shared code.  The step reuses the whole MMVAE+ path (batched decoders, fused lpx / lw kernels, DReG hooks); what differs
is the prior term of the log-weights,

    lw = lpx + beta * ( sum_c q(c|u) [log pi_c + log p(u|c) - log q(c|u)] + log p(w) - log q(u|X) - log q(w|x) ),

whose cluster expectation (a [n_clusters, C, K, B] computation on the latent samples) is evaluated by the host on the sampled
codes and handed to the latent kernel as an additive term; the kernel leaves its own log p(u) out (`skip_u_prior`)."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .containers import ModelOutput
from .mmvae_plus import MoEPlusBase


class CMVAE(MoEPlusBase):
    skip_u_prior = True

    def __init__(self, model_config, encoders=None, decoders=None):
        if model_config.modalities_specific_dim is None:
            raise AttributeError("The modalities_specific_dim attribute must be provided in the model config.")
        super().__init__(model_config, encoders, decoders)
        self.model_name = "CMVAE"
        if model_config.prior_and_posterior_dist not in ("laplace_with_softmax", "normal", "normal_with_softplus"):
            raise AttributeError(" The posterior_dist parameter must be either 'laplace_with_softmax','normal' or "
                                 f"'normal_with_softplus'.  {model_config.prior_and_posterior_dist} was provided.")
        self.multiple_latent_spaces = True
        self.n_clusters = model_config.number_of_clusters
        self.beta = model_config.beta
        self.objective = model_config.loss
        self.modalities_specific_dim = model_config.modalities_specific_dim
        self.style_dims = {m: model_config.modalities_specific_dim for m in self.encoders}
        Lw, L = model_config.modalities_specific_dim, model_config.latent_dim
        # priors of the private codes used for cross-modal generation ("r" in the paper): fixed mean, learnable scale
        self.r_mean_priors = nn.ParameterDict()
        self.r_logvars_priors = nn.ParameterDict()
        for mod in list(self.encoders.keys()):
            self.r_mean_priors[mod] = nn.Parameter(torch.zeros(1, Lw), requires_grad=False)
            self.r_logvars_priors[mod] = nn.Parameter(torch.zeros(1, Lw), requires_grad=model_config.learn_modality_prior)
        # regularising prior p(w_m) of the private codes
        self.w_mean_prior = nn.Parameter(torch.zeros(1, Lw), requires_grad=False)
        self.w_logvar_prior = nn.Parameter(torch.zeros(1, Lw), requires_grad=False)
        # cluster weights, means (learnable) and scales (fixed, like the original code)
        self._pc_params = nn.Parameter(torch.zeros(self.n_clusters), requires_grad=True)
        self.mean_clusters = nn.ParameterList([nn.Parameter((2 * torch.rand(1, L)) - 1, requires_grad=True) for _ in range(self.n_clusters)])
        self.logvar_clusters = nn.ParameterList([nn.Parameter(torch.zeros(1, L), False) for _ in range(self.n_clusters)])
        self.noise_source = None

    @property
    def pc_params(self):
        return F.softmax(self._pc_params, dim=-1)

    # ---- the pieces MoEPlusBase._elbo asks for ------------------------------------------------------------------------
    def _cross_prior(self, r):
        return self.r_mean_priors[r], self.r_logvars_priors[r]

    def _shared_prior(self):
        """Prior over cat[u, w] as the latent kernel sees it, (mean, std): the u part is skipped (`skip_u_prior`), the w part is
        p(w) with its scale computed over the Lw private dimensions only (cmvae_model.py:278-280)."""
        L = self.model_config.latent_dim
        mean = torch.cat([self.w_mean_prior.new_zeros(1, L), self.w_mean_prior], dim=-1)
        std = torch.cat([self.w_logvar_prior.new_ones(1, L), self._log_var_to_std(self.w_logvar_prior)], dim=-1)
        return mean, std

    def _log_p_u_given_c(self, u):
        """[n_clusters, *u.shape[:-1]]: log p(u | c) summed over the latent dimensions (cmvae_model.py:309-316)."""
        kind = self.model_config.prior_and_posterior_dist
        mu = torch.stack([m for m in self.mean_clusters]).reshape(self.n_clusters, *([1] * (u.dim() - 1)), -1)
        sg = torch.stack([self._log_var_to_std(lv) for lv in self.logvar_clusters]).reshape(self.n_clusters, *([1] * (u.dim() - 1)), -1)
        if kind == "laplace_with_softmax":
            return (-torch.log(2 * sg) - (u.unsqueeze(0) - mu).abs() / sg).sum(-1)
        return (-((u.unsqueeze(0) - mu) ** 2) / (2 * sg ** 2) - torch.log(sg) - 0.5 * np.log(2 * np.pi)).sum(-1)

    def _extra_lw(self, U, beta):
        """beta * sum_c q(c|u) (log pi_c + log p(u|c) - log q(c|u)) for every sample u (cmvae_model.py:303-341); U (C, K, B, L)."""
        lpc = torch.log(self.pc_params).reshape(self.n_clusters, 1, 1, 1)
        lpzc = self._log_p_u_given_c(U)                       # (n_clusters, C, K, B)
        qzc = torch.softmax(lpc + lpzc, dim=0) + 1e-20
        return beta * (qzc * (lpc + lpzc - qzc.log())).sum(0)

    # ---- inference (cmvae_model.py:395-560) ---------------------------------------------------------------------------
    def _style_prior(self, m, batch_size):
        if self.model_config.reconstruction_option == "single_prior":
            mu_m, lv_m = self.r_mean_priors[m], self.r_logvars_priors[m]
        else:
            mu_m, lv_m = self.w_mean_prior, self.w_logvar_prior
        return torch.cat([mu_m] * batch_size, dim=0), torch.cat([lv_m] * batch_size, dim=0)

    def encode(self, inputs, cond_mod="all", N=1, return_mean=False, **kwargs):
        cond_mod = super().encode(inputs, cond_mod, N, **kwargs).cond_mod
        with self._nn_ctx():
            outs = {m: self.encoders[m](inputs.data[m]) for m in cond_mod}
        dev = outs[cond_mod[0]].embedding.device
        sample = (lambda mu, sg: mu + sg * self._noise(tuple(mu.shape) if N == 1 else (N,) + tuple(mu.shape), dev))
        rm = np.random.choice(cond_mod)
        mu, sigma = outs[rm].embedding.float(), self._log_var_to_std(outs[rm].log_covariance.float())
        if return_mean:
            z = torch.stack([mu] * N) if N > 1 else mu
        else:
            z = sample(mu, sigma)
        flatten = kwargs.pop("flatten", False)
        if flatten:
            z = z.reshape(-1, self.latent_dim)
        style_z = {}
        for m in self.encoders:
            if m not in cond_mod:
                mu_m, lv_m = self._style_prior(m, len(mu))
            else:
                mu_m, lv_m = outs[m].style_embedding.float(), outs[m].style_log_covariance.float()
            if return_mean:
                style_z[m] = torch.stack([mu_m] * N) if N > 1 else mu_m
            else:
                style_z[m] = sample(mu_m, self._log_var_to_std(lv_m))
            if flatten:
                style_z[m] = style_z[m].reshape(-1, self.model_config.modalities_specific_dim)
        return ModelOutput(z=z, one_latent_space=False, modalities_z=style_z)

    def generate_from_prior(self, n_samples, **kwargs):
        """Cluster assignment, then the shared code from its cluster and the private codes from their priors (:506-545)."""
        dev = self._pc_params.device
        clusters = torch.distributions.Categorical(logits=self._pc_params).sample([n_samples])
        means = torch.cat([self.mean_clusters[int(c)] for c in clusters], dim=0)
        lvs = torch.cat([self.logvar_clusters[int(c)] for c in clusters], dim=0)
        z = means + self._log_var_to_std(lvs) * self._noise(tuple(means.shape), dev)
        style_z = {}
        for m in self.encoders:
            mu_m, lv_m = self._style_prior(m, n_samples)
            style_z[m] = mu_m + self._log_var_to_std(lv_m) * self._noise(tuple(mu_m.shape), dev)
        return ModelOutput(z=z.detach(), one_latent_space=False, modalities_z={k: v.detach() for k, v in style_z.items()})

    def predict_clusters(self, inputs, **kwargs):
        """Cluster of every sample: argmax of the product over modalities of q(c | mean of q(u | x_m)) (cmvae_model.py:547-600)."""
        with torch.no_grad():
            lpc = torch.log(self.pc_params).reshape(self.n_clusters, 1)
            pc_zs, acc = {}, []
            for m in inputs.data:
                with self._nn_ctx():
                    mu = self.encoders[m](inputs.data[m]).embedding.float()
                pc = torch.softmax(lpc + self._log_p_u_given_c(mu), dim=0)
                pc_zs[m] = pc
                acc.append(pc)
            clusters = torch.stack(acc, dim=0).prod(0).argmax(0)
        return ModelOutput(clusters=clusters, pc_zs=pc_zs)
