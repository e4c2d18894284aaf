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

    def _log_p_u_given_c(self, u, n=None):
        """[n, *u.shape[:-1]]: log p(u | c) summed over the latent dimensions for the first n clusters (training: n_clusters,
        cmvae_model.py:299-304; cluster prediction: every cluster ever created, :572-581 — pruned ones carry zero weight)."""
        kind = self.model_config.prior_and_posterior_dist
        n = self.n_clusters if n is None else n
        mu = torch.stack([m for m in self.mean_clusters][:n]).reshape(n, *([1] * (u.dim() - 1)), -1)
        sg = torch.stack([self._log_var_to_std(lv) for lv in self.logvar_clusters][:n]).reshape(n, *([1] * (u.dim() - 1)), -1)
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
        """Cluster of every sample (cmvae_model.py:546-619): per modality, one SAMPLE u ~ q(u | x_m), p(c | u) ∝ p(u | c) p(c), the
        modality's vote = argmax_c; the result is the majority vote over the modalities (`torch.mode`).  `compute_lliks=True` also
        returns the mean over modalities of sum_c p(c|u) (log p(u|c) + log p(c) - log p(c|u)) / latent_dim (used by the pruning)."""
        with torch.no_grad():
            compute_norm_lliks = kwargs.pop("compute_lliks", False)
            votes, pc_zs, norm_lliks = [], {}, []
            lpc = torch.log(self.pc_params + 1e-20).view(-1, 1)
            for m in inputs.data:
                with self._nn_ctx():
                    o = self.encoders[m](inputs.data[m])
                mu, sigma = o.embedding.float(), self._log_var_to_std(o.log_covariance.float())
                z = mu + sigma * self._noise(tuple(mu.shape), mu.device)
                lpz_c = self._log_p_u_given_c(z, len(self.mean_clusters))   # (all clusters, batch)
                pc_z = torch.softmax(lpc + lpz_c, dim=0)
                votes.append(torch.argmax(pc_z, dim=0))
                pc_zs[m] = pc_z
                if compute_norm_lliks:
                    norm_lliks.append(((lpz_c + lpc - pc_z.log()) * pc_z).sum(0).squeeze(-1) / self.latent_dim)
            clusters = torch.mode(torch.stack(votes, dim=-1), dim=-1)[0]
            if compute_norm_lliks:
                return ModelOutput(clusters=clusters, pc_zs=pc_zs, norm_lliks=torch.stack(norm_lliks, dim=0).mean(0))
            return ModelOutput(clusters=clusters, pc_zs=pc_zs)

    def prune_clusters(self, train_data, batch_size=128):
        """The paper's post-hoc selection of the number of clusters (cmvae_model.py:621-711): repeatedly measure the penalised
        normalised entropy beta * H(p(c | u_m)) - normalised likelihood on `train_data`, remove the cluster with the least mass
        (its logit becomes -inf), and finally keep the cluster set with the lowest value.  Returns the list of entropy values
        indexed by the number of clusters."""
        from scipy.stats import entropy
        from torch.utils.data import DataLoader

        from .containers import MultimodalBaseDataset
        with torch.no_grad():
            dev = self._pc_params.device
            n = len(next(iter(train_data.data.values())))
            n_cluster_params = [None] * (self.n_clusters + 1)
            h_values = [torch.inf] * (self.n_clusters + 1)
            while self.n_clusters >= 2:
                mass = torch.zeros_like(self._pc_params)
                h_data = []
                for idx in DataLoader(range(n), batch_size=batch_size):
                    batch = MultimodalBaseDataset(data={k: v[idx].to(dev) for k, v in train_data.data.items()})
                    cp = self.predict_clusters(batch, compute_lliks=True)
                    for i in range(len(mass)):
                        mass[i] += (cp.clusters == i).int().sum()
                    h_pzc = []
                    for pc_z in cp.pc_zs.values():
                        p = pc_z.squeeze(1).cpu().numpy() if pc_z.dim() > 2 else pc_z.cpu().numpy()
                        h_pzc.append(torch.Tensor(entropy(p, axis=0) / np.log(np.count_nonzero(p, axis=0))).to(dev))
                    h_data.append(self.model_config.beta * torch.stack(h_pzc, dim=0).mean(0) - cp.norm_lliks)
                h = torch.cat(h_data, dim=-1).mean(-1)
                h_values[self.n_clusters] = h
                n_cluster_params[self.n_clusters] = self._pc_params.clone()
                assert torch.all(mass[torch.argwhere(self._pc_params == -torch.inf)] == 0)
                self.n_clusters = self.n_clusters - 1
                mass[self._pc_params.isinf()] = torch.inf
                self._pc_params[torch.argmin(mass)] = -torch.inf
                assert torch.sum(~self._pc_params.isinf()) == self.n_clusters
            self.n_clusters = int(torch.argmin(torch.Tensor(h_values)))
            self._pc_params = torch.nn.Parameter(n_cluster_params[self.n_clusters])
            return h_values

    def compute_joint_nll(self, inputs, K=1000, batch_size_K=100, **kwargs):
        """-sum_i ln p(x_i) from K // n_modalities importance samples per conditioning modality: the log-mean-exp of the training
        log-weights (cluster-mixture prior included) with rescale = beta = 1 (cmvae_model.py:733-791), batched over the datapoints
        and chunked over the samples instead of the reference's per-datapoint loop."""
        from .elbo import logmeanexp
        self.eval()
        if hasattr(inputs, "masks"):
            raise AttributeError("The compute_joint_nll method is not yet implemented for incomplete datasets.")
        k_iwae = K // self.n_modalities
        with torch.no_grad():
            lws = []
            for k0 in range(0, k_iwae, batch_size_K):
                _, meta = self._elbo(inputs, min(batch_size_K, k_iwae - k0), "iwae_looser", rescale=1.0, beta=1.0)
                lws.append(meta["lw"])                                   # (C, n, B)
            lw = torch.cat(lws, dim=1)
        return -logmeanexp(lw.reshape(-1, lw.shape[-1])).sum()
