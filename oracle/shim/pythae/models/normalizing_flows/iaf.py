from .base import BaseNF as IAF, BaseNFConfig as IAFConfig  # noqa
