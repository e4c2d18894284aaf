"""Native (tcgen05) execution of the PolyMNIST ResNet encoder/decoder — filled in by the conv kernels."""


def use_native(x):
    return False


def status(which):
    return "torch (cuDNN/cuBLAS library calls)"
