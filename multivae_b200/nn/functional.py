"""Layer primitives used by the architectures.

Two execution paths exist, selected per call by `backend()`:
  * "native": hand-written sm_100a kernels (tcgen05 GEMM / implicit-GEMM convolution with fused
    epilogues) reached through the C-ABI — the product path for the architectures it covers;
  * "torch": ATen (cuBLAS/cuDNN) ops: the fp32 path of the parity checks and CPU-side module handling.
    This is a *library* path on the GPU, never a CPU fallback of the fused ELBO kernels.
"""
import os

import torch
import torch.nn.functional as F

_BACKEND = os.environ.get("MULTIVAE_B200_NN", "auto")


def set_backend(name):
    global _BACKEND
    assert name in ("auto", "native", "torch")
    _BACKEND = name


def backend():
    return _BACKEND


_ACT = {"none": lambda t: t, "relu": F.relu, "sigmoid": torch.sigmoid, "lrelu": lambda t: F.leaky_relu(t, 0.2)}


def linear(x, weight, bias=None, act="none", out_dtype=None):
    y = _ACT[act](F.linear(x, weight, bias))
    return y if out_dtype is None else y.to(out_dtype)


def linear_heads(x, heads):
    """Several Linear heads sharing one input: one GEMM over the concatenated weights."""
    w = torch.cat([h.weight for h in heads], dim=0)
    b = torch.cat([h.bias for h in heads], dim=0)
    y = F.linear(x, w, b)
    return torch.split(y, [h.weight.shape[0] for h in heads], dim=-1)


def conv2d(x, weight, bias=None, stride=1, padding=0, act="none"):
    return _ACT[act](F.conv2d(x, weight, bias, stride=stride, padding=padding))


def conv_transpose2d(x, weight, bias=None, stride=1, padding=0, output_padding=0, act="none"):
    return _ACT[act](F.conv_transpose2d(x, weight, bias, stride=stride, padding=padding, output_padding=output_padding))


def backend_summary():
    """Which implementation each layer family runs on with bf16 compute (reported by bench.py)."""
    from . import resnet_native as RN
    return {"elbo": "native sm_100a kernels (C-ABI)", "resnet_decoder": RN.status("decoder"),
            "resnet_encoder": RN.status("encoder"),
            "cub_resnets": "native sm_100a: the same tap-GEMM / wgrad / halo kernels (two-box TMA windows at 64 pixels) + mv_lrelu_fwd",
            "mlp": "native sm_100a: tcgen05 GEMM fwd / dgrad / wgrad with fused bias + ReLU / Sigmoid (mv_gemm)",
            "strided_and_transposed_conv": "native sm_100a: im2col / col2im gathers around the tcgen05 GEMM (mv_im2col, mv_col2im, mv_gemm)",
            "fp32_path": "library layers (cuDNN / cuBLAS), used only by the fp32 parity checks"}
