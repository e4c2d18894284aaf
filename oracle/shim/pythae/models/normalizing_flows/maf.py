from .base import BaseNF as MAF, BaseNFConfig as MAFConfig  # noqa
