from typing import Optional, Tuple, Union

from pydantic.dataclasses import dataclass

from pythae.config import BaseConfig


@dataclass
class BaseAEConfig(BaseConfig):
    input_dim: Union[Tuple[int, ...], None] = None
    latent_dim: int = 10
    uses_default_encoder: bool = True
    uses_default_decoder: bool = True


@dataclass
class EnvironmentConfig(BaseConfig):
    python_version: str = "3.8"
