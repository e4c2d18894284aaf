"""Import-time placeholders (JNF and the samplers only; not on the hot path)."""
import torch.nn as nn


class BaseNFConfig:
    def __init__(self, *a, **k):
        raise NotImplementedError("oracle shim placeholder")


class BaseNF(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("oracle shim placeholder")


class NFModel(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("oracle shim placeholder")
