from .base_config import BaseAEConfig  # noqa
