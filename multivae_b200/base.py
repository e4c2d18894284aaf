"""BaseMultiVAE: constructor contract, sanity checks, rescale factors and decoder-distribution table of
the reference's base class (/root/reference/src/multivae/models/base/base_ae_model.py:24-180), without
its save/load/hub plumbing (out of scope: SURVEY section 2 rows 12, 15)."""
from copy import deepcopy

import numpy as np
import torch
import torch.nn as nn

from . import _cabi as C
from .containers import ModelOutput  # noqa: F401
from .nn.base_architectures import BaseDecoder, BaseEncoder
from .nn.default_architectures import BaseDictDecoders, BaseDictEncoders


class BaseMultiVAE(nn.Module):
    def __init__(self, model_config, encoders=None, decoders=None):
        super().__init__()
        self.model_name = "BaseMultiVAE"
        self.model_config = model_config
        self.n_modalities = model_config.n_modalities
        self.input_dims = model_config.input_dims
        self.latent_dim = model_config.latent_dim
        self.device = None
        self.multiple_latent_spaces = False
        # arithmetic type of the encoder/decoder contractions (bf16 operands, fp32 accumulate) — the
        # parameters themselves and the whole ELBO path stay fp32 like the reference
        self.compute_dtype = torch.float32
        self.use_likelihood_rescaling = model_config.uses_likelihood_rescaling
        if model_config.input_dims is not None and len(model_config.input_dims) != model_config.n_modalities:
            raise AttributeError(
                f"The provided number of input_dims {len(model_config.input_dims)} doesn't"
                f"match the number of modalities ({model_config.n_modalities} in model config ")
        if encoders is None:
            if self.input_dims is None:
                raise AttributeError("Please provide encoders or input dims for the modalities in the model_config.")
            encoders = self.default_encoders(model_config)
        else:
            self.model_config.custom_architectures.append("encoders")
        if decoders is None:
            if self.input_dims is None:
                raise AttributeError("Please provide decoders or input dims for the modalities in the model_config.")
            decoders = self.default_decoders(model_config)
        else:
            self.model_config.custom_architectures.append("decoders")
        self.sanity_check(encoders, decoders)
        self.set_decoders(decoders)
        self.set_encoders(encoders)
        self.modalities_name = list(self.decoders.keys())
        self.rescale_factors = self.set_rescale_factors()
        if model_config.decoders_dist is None:
            model_config.decoders_dist = {k: "normal" for k in self.encoders}
        if model_config.decoder_dist_params is None:
            model_config.decoder_dist_params = {}
        self.set_decoders_dist(model_config.decoders_dist, deepcopy(model_config.decoder_dist_params))

    # ---- decoder distributions: (name, scale) per modality consumed by the fused kernels ---------
    def set_decoders_dist(self, recon_dict, dist_params_dict):
        self.recon_dists = {}
        for k, name in recon_dict.items():
            if name not in ("normal", "laplace", "bernoulli", "categorical"):
                raise ValueError("The distribution type 'dist' is not supported")
            self.recon_dists[k] = (C.DIST[name], float(dist_params_dict.get(k, {}).get("scale", 1.0)))

    def set_rescale_factors(self):
        if self.use_likelihood_rescaling:
            if self.model_config.rescale_factors is not None:
                return self.model_config.rescale_factors
            if self.input_dims is None:
                raise AttributeError(
                    " inputs_dim is None but (use_likelihood_rescaling = True in model_config)"
                    " To compute default likelihood rescalings we need the input dimensions.")
            max_dim = max(*[np.prod(self.input_dims[k]) for k in self.input_dims])
            return {k: max_dim / np.prod(self.input_dims[k]) for k in self.input_dims}
        return {k: 1 for k in self.encoders}

    def sanity_check(self, encoders, decoders):
        if self.n_modalities != len(encoders.keys()):
            raise AttributeError(f"The provided number of encoders {len(encoders.keys())} doesn't"
                                 f"match the number of modalities ({self.n_modalities} in model config ")
        if self.n_modalities != len(decoders.keys()):
            raise AttributeError(f"The provided number of decoders {len(decoders.keys())} doesn't"
                                 f"match the number of modalities ({self.n_modalities} in model config ")
        if encoders.keys() != decoders.keys():
            raise AttributeError("The names of the modalities in the encoders dict doesn't match the names of the "
                                 "modalities in the decoders dict.")
        if self.input_dims is not None and self.input_dims.keys() != encoders.keys():
            raise KeyError(f"Warning! : The modalities names in model_config.input_dims : {list(self.input_dims.keys())}"
                           f" do not match the modalities names in encoders : {list(encoders.keys())}")

    def default_encoders(self, model_config):
        return BaseDictEncoders(self.input_dims, model_config.latent_dim)

    def default_decoders(self, model_config):
        return BaseDictDecoders(self.input_dims, model_config.latent_dim)

    def set_encoders(self, encoders):
        self.encoders = nn.ModuleDict()
        for m, enc in encoders.items():
            if not isinstance(enc, BaseEncoder):
                raise AttributeError(f"For modality {m}, encoder must inherit from BaseEncoder class. Refer to documentation.")
            self.encoders[m] = enc

    def set_decoders(self, decoders):
        self.decoders = nn.ModuleDict()
        for m, dec in decoders.items():
            if not isinstance(dec, BaseDecoder):
                raise AttributeError(f"For modality {m}, decoder must inherit from BaseDecoder class. Refer to documentation.")
            self.decoders[m] = dec

    def __getstate__(self):
        """copy.deepcopy / pickling (the reference's trainer deep-copies the best model, base_trainer.py): the per-device caches —
        side streams of the parallel encoders / decoders, device-resident constants — are rebuilt on demand, never copied."""
        state = self.__dict__.copy()
        for k in ("_enc_streams", "_dec_streams", "_const_cache"):
            state.pop(k, None)
        return state

    def _nn_ctx(self):
        """Context in which the encoders / decoders run (library layers: bf16 autocast when
        compute_dtype is bf16; the native tcgen05 layers read `compute_dtype` themselves)."""
        import contextlib
        if self.compute_dtype == torch.bfloat16 and torch.cuda.is_available():
            return torch.autocast("cuda", dtype=torch.bfloat16)
        return contextlib.nullcontext()

    def _const(self, key, build):
        """Device-resident constants of a forward pass (subset tables, sample->subset maps), built once per (key, device):
        a step must not create device tensors from host data — that is a synchronising pageable copy and illegal during
        CUDA-graph capture."""
        cache = self.__dict__.setdefault("_const_cache", {})
        if key not in cache:
            cache[key] = build()
        return cache[key]

    # Whether a captured CUDA graph of the training step stays valid from one step to the next (no host-side randomness,
    # no Python scalar that changes per step baked into the kernels' arguments).
    graph_safe = True

    def prepare_step(self, epoch=1, batch_ratio=0.0):
        """Called by the trainer OUTSIDE graph capture/replay: refresh device scalars the step reads (MVAE's KL weight)."""

    # The M encoders are independent and small (B images each, against C*K*B for a decoder): their kernels are bound by launch
    # latency and ramp, not by the GPU.  On CUDA each modality's encoder runs on its own stream (forked from and joined to the
    # current one, so it also works inside CUDA-graph capture); autograd replays every backward node on the stream of its
    # forward, so the encoders' backward passes overlap in the same way.
    parallel_encoders = True

    def _run_encoders(self, inputs, mods, dev):
        if not (self.parallel_encoders and dev.type == "cuda" and len(mods) > 1):
            out = {}
            for c in mods:
                with self._nn_ctx():
                    out[c] = self.encoders[c](inputs.data[c])
            return out
        streams = self.__dict__.setdefault("_enc_streams", {})
        cur = torch.cuda.current_stream()
        out = {}
        for c in mods:
            st = streams.get((c, dev.index))
            if st is None:
                st = streams[(c, dev.index)] = torch.cuda.Stream(device=dev)
            st.wait_stream(cur)
            with torch.cuda.stream(st), self._nn_ctx():
                out[c] = self.encoders[c](inputs.data[c])
        for c in mods:
            cur.wait_stream(streams[(c, dev.index)])
        return out

    def _run_decoders(self, z_by_mod, dev, mod_of=None):
        """reconstruction = decoders[m](z) for every (m, z): like the encoders, small decoders (the MLP / strided-convolution
        architectures on B or C*K*B rows) are bound by launch latency, so on CUDA each runs on its own stream.  The ResNet decoders
        of the north star saturate the GPU on their own and stay on the current stream."""
        from .nn.mmnist import DecoderResnetMMNIST
        mod_of = mod_of or (lambda k: k)   # keys are modality names unless the caller decodes a modality several times
        keys = list(z_by_mod)
        heavy = any(isinstance(self.decoders[mod_of(k)], DecoderResnetMMNIST) for k in keys)
        if heavy or not (self.parallel_encoders and dev.type == "cuda" and len(keys) > 1):
            out = {}
            for k in keys:
                with self._nn_ctx():
                    out[k] = self._logits(self.decoders[mod_of(k)](z_by_mod[k]).reconstruction)
            return out
        streams = self.__dict__.setdefault("_dec_streams", {})
        cur = torch.cuda.current_stream()
        out = {}
        for k in keys:
            st = streams.get((k, dev.index))
            if st is None:
                st = streams[(k, dev.index)] = torch.cuda.Stream(device=dev)
            st.wait_stream(cur)
            with torch.cuda.stream(st), self._nn_ctx():
                out[k] = self._logits(self.decoders[mod_of(k)](z_by_mod[k]).reconstruction)
        for k in keys:
            cur.wait_stream(streams[(k, dev.index)])
        return out

    def update(self):
        """Called by the trainer at the end of each epoch (base_trainer.py:738-741)."""

    # ---- inference API (base_ae_model.py:182-311, 374-442) ------------------------------------------
    def _draw(self, shape, kind, device):
        """Standard draws e (z = loc + scale * e); `noise_source` is the tests' injection hook."""
        from .elbo import standard_noise
        src = getattr(self, "noise_source", None)
        if src is not None:
            return src(tuple(shape), kind, device)
        return standard_noise(tuple(shape), kind, device)

    def _rsample_gaussian(self, mu, log_var, N=1, return_mean=False, flatten=False):
        """rsample_from_gaussian (base_utils.py:150-172)."""
        if return_mean:
            z = torch.stack([mu] * N) if N > 1 else mu
        else:
            shape = tuple(mu.shape) if N == 1 else (N,) + tuple(mu.shape)
            z = mu + torch.exp(0.5 * log_var) * self._draw(shape, "normal", mu.device)
        if N > 1 and flatten:
            if z.dim() == 2:
                z = z.unsqueeze(0)
            z = z.reshape(-1, *z.shape[2:])
        return z

    def encode(self, inputs, cond_mod="all", N=1, return_mean=False, **kwargs):
        """Argument checks shared by every model's encode (base_ae_model.py:182-224): returns the list of conditioning
        modalities; the model classes add the latent codes."""
        if isinstance(cond_mod, str):
            if cond_mod == "all":
                cond_mod = list(self.encoders.keys())
            elif cond_mod in self.encoders.keys():
                cond_mod = [cond_mod]
            else:
                raise AttributeError('If cond_mod is a string, it must either be "all" or a modality name'
                                     f" The provided string {cond_mod} is neither.")
        ignore_incomplete = kwargs.pop("ignore_incomplete", False)
        if hasattr(inputs, "masks") and not ignore_incomplete:
            avail = torch.tensor(True)
            for m in cond_mod:
                avail = torch.logical_and(avail.to(inputs.masks[m].device), inputs.masks[m])
            if not torch.all(avail):
                raise AttributeError("You tried to encode a incomplete dataset conditioning on",
                                     f"modalities {cond_mod}, but some samples are not available"
                                     "in all those modalities.")
        return ModelOutput(cond_mod=cond_mod, z=None, one_latent_space=None)

    def decode(self, embedding, modalities="all"):
        """Decode latent codes in the requested modalities (base_ae_model.py:226-267); forward-only through the decoders
        (native tensor-core stacks when compute_dtype is bf16: nothing is saved for a backward pass under no_grad)."""
        self.eval()
        with torch.no_grad():
            if modalities == "all":
                modalities = list(self.decoders.keys())
            elif isinstance(modalities, str):
                modalities = [modalities]
            try:
                outputs = ModelOutput()
                for m in modalities:
                    z = embedding.z if embedding.one_latent_space else torch.cat([embedding.z, embedding.modalities_z[m]], dim=-1)
                    with self._nn_ctx():
                        outputs[m] = self.decoders[m](z).reconstruction
                return outputs
            except Exception:
                raise ValueError("There was an error during decode. "
                                 " Check that the format for the embedding is correct:"
                                 "it must be a ModelOuput instance and "
                                 "embedding.z must be a Tensor of shape (batch_size, *latent_shape)"
                                 "If you used the encode function with N>1 to generate the embedding,"
                                 " you need to pass flatten=True to have the right format for decoding.")

    def predict(self, inputs, cond_mod="all", gen_mod="all", N=1, flatten=False, **kwargs):
        """Generate in `gen_mod` conditioning on `cond_mod` (base_ae_model.py:269-311)."""
        self.eval()
        ignore_incomplete = kwargs.pop("ignore_incomplete", False)
        with torch.no_grad():
            z = self.encode(inputs, cond_mod, N=N, flatten=True, ignore_incomplete=ignore_incomplete, **kwargs)
        output = self.decode(z, gen_mod)
        n_data = len(z.z) // N
        if not flatten and N > 1:
            for m in output.keys():
                output[m] = output[m].reshape(N, n_data, *output[m].shape[1:])
        return output

    def compute_joint_nll(self, inputs, K=1000, batch_size_K=100):
        raise NotImplementedError

    def generate_from_prior(self, n_samples, **kwargs):
        """Static N(0, I) prior (base_ae_model.py:379-394)."""
        shape = (n_samples, self.latent_dim) if n_samples > 1 else (self.latent_dim,)
        dev = next(self.parameters()).device
        return ModelOutput(z=self._draw(shape, "normal", dev), one_latent_space=True)

    def _iw_lpx(self, inputs, z_of, n_k, mods=None):
        """sum_m log p(x_m | z_k) for n_k latent samples per datapoint: one batched decoder pass per modality over the
        n_k * B rows + the streaming log-prob kernel (rescale 1, like the reference's likelihood estimators).
        z_of(m) -> (n_k, B, latent) decoder input of modality m.  Returns lpx (n_k, B) fp32."""
        from .elbo import lpx_fwd
        lib = C.lib()
        mods = list(inputs.data.keys()) if mods is None else mods
        B = len(inputs.data[mods[0]])
        dev = inputs.data[mods[0]].device
        lpx = torch.empty(1, n_k, B, device=dev, dtype=torch.float32)
        for i, m in enumerate(mods):
            z = z_of(m)
            with self._nn_ctx():
                rec = self._logits(self.decoders[m](z.reshape(n_k * B, z.shape[-1])).reconstruction)
            rec = rec.reshape(1, n_k, B, *rec.shape[1:]).contiguous()
            dist, scale = self.recon_dists[m]
            lpx_fwd(lib, rec, self._target(inputs, m, rec), lpx, 1, n_k, B, dist, scale, 1.0, None, i > 0)
        return lpx[0]

    def _encode_many(self, inputs, cond_mod, n):
        """n samples per datapoint, distributed like n independent encode() calls, as (n, B, latent) tensors.  Models whose
        encode draws a random conditioning modality per call (MMVAE, MMVAE+) override this."""
        emb = self.encode(inputs, cond_mod, N=n)
        if n == 1:
            emb["z"] = emb.z.unsqueeze(0)
            if not emb.one_latent_space:
                emb["modalities_z"] = {m: v.unsqueeze(0) for m, v in emb.modalities_z.items()}
        return emb

    @torch.no_grad()
    def compute_cond_nll(self, inputs, subset, pred_mods, k_iwae=1000, batch_size_K=100):
        """-ln p(x_pred | x_cond) estimated with k_iwae samples z ~ q(z | x_cond) (base_ae_model.py:396-442).  The reference
        runs k_iwae sequential encode + decode passes; here the encoders run once per chunk of `batch_size_K` samples
        (encode(N=chunk)) and the decoders on chunk * B rows."""
        from .elbo import logmeanexp
        self.eval()
        lps = {m: [] for m in pred_mods}
        done = 0
        while done < k_iwae:
            n = min(batch_size_K, k_iwae - done)
            done += n
            emb = self._encode_many(inputs, subset, n)
            zs = emb.z
            for m in pred_mods:
                if emb.one_latent_space:
                    z_of = lambda _m, zs=zs: zs  # noqa: E731
                else:
                    z_of = lambda _m, zs=zs, zm=emb.modalities_z[m]: torch.cat([zs, zm], dim=-1)  # noqa: E731
                lps[m].append(self._iw_lpx(inputs, z_of, n, mods=[m]))
        out = {}
        for m in pred_mods:
            ll = logmeanexp(torch.cat(lps[m], dim=0))
            out[m] = -torch.sum(ll) / len(ll)
        return out

    # ---- helpers shared by the model variants ------------------------------------------------------
    @staticmethod
    def _stack_masks(inputs, mods):
        if not hasattr(inputs, "masks"):
            return None
        return torch.stack([inputs.masks[m].to(torch.uint8) for m in mods]).contiguous()

    @staticmethod
    def _target(inputs, m, recon=None):
        """fp32 target tensor of modality m.  Categorical targets may come as dicts like the reference's text modalities
        (base_utils.py:41-59): {"one_hot": probs} or {"tokens": class ids} (expanded to one-hot over the decoder's classes)."""
        x = inputs.data[m]
        if isinstance(x, dict):
            if "one_hot" in x:
                x = x["one_hot"]
            elif "tokens" in x:
                x = torch.nn.functional.one_hot(x["tokens"], recon.shape[-1])
            else:
                raise NotImplementedError()
        return x.float().contiguous()

    @staticmethod
    def _logits(rec):
        """Decoders of categorical modalities may return {"one_hot": logits} (base_utils.py:45-49)."""
        if isinstance(rec, dict):
            if "one_hot" not in rec:
                raise NotImplementedError()
            return rec["one_hot"]
        return rec

    def _recon_meta(self, mods_recon, mods_rows):
        """(dist, scale, rescale, row of the stacked mask tensor) per reconstructed modality."""
        return [(self.recon_dists[m][0], self.recon_dists[m][1], float(self.rescale_factors[m]), mods_rows.index(m))
                for m in mods_recon]
