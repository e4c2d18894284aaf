from pythae.models.nn.base_architectures import BaseDecoder, BaseEncoder  # noqa
from .base_config import BaseAEConfig  # noqa
