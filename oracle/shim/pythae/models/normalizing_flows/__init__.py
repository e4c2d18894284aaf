from .base import BaseNF, BaseNFConfig, NFModel  # noqa
from .maf import MAF, MAFConfig  # noqa
from .iaf import IAF, IAFConfig  # noqa
