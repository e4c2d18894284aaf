"""Native (tcgen05) execution of the PolyMNIST ResNet encoder/decoder — filled in by the conv kernels."""
from . import functional as NF


def use_native(x):
    return False
