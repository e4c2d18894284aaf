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
            if name == "categorical":
                raise NotImplementedError("categorical decoders are not covered by the fused ELBO kernels yet")
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

    def _nn_ctx(self):
        """Context in which the encoders / decoders run (library layers: bf16 autocast when
        compute_dtype is bf16; the native tcgen05 layers read `compute_dtype` themselves)."""
        import contextlib
        if self.compute_dtype == torch.bfloat16 and torch.cuda.is_available():
            return torch.autocast("cuda", dtype=torch.bfloat16)
        return contextlib.nullcontext()

    def update(self):
        """Called by the trainer at the end of each epoch (base_trainer.py:738-741)."""

    # ---- helpers shared by the model variants ------------------------------------------------------
    @staticmethod
    def _stack_masks(inputs, mods):
        if not hasattr(inputs, "masks"):
            return None
        return torch.stack([inputs.masks[m].to(torch.uint8) for m in mods]).contiguous()

    def _recon_meta(self, mods_recon, mods_rows):
        """(dist, scale, rescale, row of the stacked mask tensor) per reconstructed modality."""
        return [(self.recon_dists[m][0], self.recon_dists[m][1], float(self.rescale_factors[m]), mods_rows.index(m))
                for m in mods_recon]
