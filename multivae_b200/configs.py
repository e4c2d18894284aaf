"""Model configurations: same field names, defaults and validation errors as the reference's
pydantic dataclasses (models/base/base_config.py:9-46, models/*/*_config.py)."""
import json
import os
from dataclasses import asdict, dataclass, field
from typing import Dict, List, Optional, Union


@dataclass
class BaseConfig:
    def __post_init__(self):
        self.name = self.__class__.__name__

    def to_dict(self):
        d = asdict(self)
        d["name"] = self.__class__.__name__
        return d

    def to_json_string(self):
        return json.dumps(self.to_dict(), default=str)

    def save_json(self, dir_path, filename):
        with open(os.path.join(dir_path, f"{filename}.json"), "w", encoding="utf-8") as fp:
            fp.write(self.to_json_string())

    @classmethod
    def from_dict(cls, d):
        d = dict(d)
        d.pop("name", None)
        return cls(**d)

    @classmethod
    def from_json_file(cls, path):
        with open(path) as f:
            return cls.from_dict(json.load(f))


@dataclass
class BaseAEConfig(BaseConfig):
    input_dim: Optional[tuple] = None
    latent_dim: int = 10
    style_dim: int = 0


_DISTS = ("normal", "bernoulli", "laplace", "categorical")


@dataclass
class BaseMultiVAEConfig(BaseConfig):
    n_modalities: int = None
    latent_dim: int = 10
    input_dims: Optional[dict] = None
    uses_likelihood_rescaling: bool = False
    rescale_factors: Optional[dict] = None
    decoders_dist: Optional[Dict[str, str]] = None
    decoder_dist_params: Optional[dict] = None
    custom_architectures: list = field(default_factory=list)

    def __post_init__(self):
        super().__post_init__()
        if self.n_modalities is None:
            raise TypeError("n_modalities is required")
        if self.input_dims is not None:
            self.input_dims = {k: tuple(self.input_dims[k]) for k in self.input_dims}
        if self.decoders_dist is not None:
            for k, v in self.decoders_dist.items():
                if v not in _DISTS:
                    raise ValueError(f"decoders_dist[{k}] must be one of {_DISTS}, got {v}")


def _choice(name, value, allowed):
    if value not in allowed:
        raise ValueError(f"{name} must be one of {allowed}, got {value!r}")


@dataclass
class MMVAEPlusConfig(BaseMultiVAEConfig):
    K: int = 10
    prior_and_posterior_dist: str = "laplace_with_softmax"
    learn_shared_prior: bool = False
    learn_modality_prior: bool = True
    beta: float = 1.0
    modalities_specific_dim: int = None
    reconstruction_option: str = "joint_prior"
    loss: str = "dreg_looser"

    def __post_init__(self):
        super().__post_init__()
        _choice("prior_and_posterior_dist", self.prior_and_posterior_dist,
                ("laplace_with_softmax", "normal_with_softplus", "normal"))
        _choice("loss", self.loss, ("iwae_looser", "dreg_looser"))
        _choice("reconstruction_option", self.reconstruction_option, ("single_prior", "joint_prior"))


@dataclass
class MMVAEConfig(BaseMultiVAEConfig):
    K: int = 10
    prior_and_posterior_dist: str = "laplace_with_softmax"
    learn_prior: bool = True
    beta: float = 1.0
    loss: str = "dreg_looser"

    def __post_init__(self):
        super().__post_init__()
        _choice("prior_and_posterior_dist", self.prior_and_posterior_dist, ("laplace_with_softmax", "normal"))
        _choice("loss", self.loss, ("iwae_looser", "dreg_looser"))


@dataclass
class MoPoEConfig(BaseMultiVAEConfig):
    subsets: Union[List[list], Dict[str, list], None] = None
    beta: float = 1.0
    beta_style: float = 1.0
    modalities_specific_dim: Optional[dict] = None


@dataclass
class MVAEConfig(BaseMultiVAEConfig):
    use_subsampling: bool = True
    k: int = 0
    warmup: int = 10
    beta: float = 1


@dataclass
class MVTCAEConfig(BaseMultiVAEConfig):
    alpha: float = 0.1
    beta: float = 2.5


@dataclass
class CRMVAEConfig(BaseMultiVAEConfig):
    beta: float = 2.5


@dataclass
class CMVAEConfig(BaseMultiVAEConfig):
    K: int = 10
    prior_and_posterior_dist: str = "laplace_with_softmax"
    learn_modality_prior: bool = True
    beta: float = 1.0
    modalities_specific_dim: int = None
    reconstruction_option: str = "joint_prior"
    loss: str = "dreg_looser"
    number_of_clusters: int = 10

    def __post_init__(self):
        super().__post_init__()
        _choice("prior_and_posterior_dist", self.prior_and_posterior_dist,
                ("laplace_with_softmax", "normal_with_softplus", "normal"))
        _choice("loss", self.loss, ("iwae_looser", "dreg_looser"))
        _choice("reconstruction_option", self.reconstruction_option, ("single_prior", "joint_prior"))
