from .base_architectures import BaseDecoder, BaseEncoder  # noqa
