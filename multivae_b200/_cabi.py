"""ctypes binding of the C-ABI shared library (include/multivae_b200.h).

This is the thin layer the Python host uses to reach the CUDA kernels: raw device pointers
(`tensor.data_ptr()`), sizes and the current CUDA stream handle; no torch types cross the boundary.
There is NO fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmultivae_b200.so")

MV_F32, MV_BF16 = 0, 1
DIST = {"normal": 0, "laplace": 1, "bernoulli": 2}
LATENT = {"laplace_with_softmax": 0, "normal": 1, "normal_with_softplus": 1}
LOSS = {"iwae_looser": 0, "dreg_looser": 1}
ACT = {"none": 0, "relu": 1, "lrelu": 2, "sigmoid": 3}

_lib = None

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

# symbol -> argtypes, exactly the prototypes in include/multivae_b200.h
_PROTOS = {
    "mv_version": [ctypes.POINTER(c_int)] * 3,
    "mv_moe_lpx_fwd": [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int, c_float, c_float,
                       c_void_p, c_int, c_void_p],
    "mv_moe_lpx_bwd": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int,
                       c_float, c_float, c_void_p, c_void_p],
    "mv_moe_lw_fwd": [c_void_p] * 21 + [c_int] * 7 + [c_float, c_int, c_void_p],
    "mv_poe_fwd": [c_void_p] * 4 + [c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_float] +
                  [c_void_p] * 5 + [c_int] * 3 + [c_void_p],
    "mv_poe_bwd": [c_void_p] * 4 + [c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_float] +
                  [c_void_p] * 5 + [c_int] * 3 + [c_void_p],
}


class NativeLibraryError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "multivae_b200 has no CPU or eager fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mv_last_error.restype = ctypes.c_char_p
        for name, args in _PROTOS.items():
            fn = getattr(_lib, name)
            fn.argtypes = args
            fn.restype = c_int
    return _lib


def exported_symbols():
    return ["mv_last_error"] + list(_PROTOS)


def check(status, what):
    if status != 0:
        msg = lib().mv_last_error().decode()
        if status == 1:
            raise ValueError(f"{what}: {msg}")
        if status == 3:
            raise NotImplementedError(f"{what}: {msg}")
        raise NativeLibraryError(f"{what}: {msg}")


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeLibraryError("multivae_b200 kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise ValueError("non-contiguous tensor passed to a native kernel")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def dtype_code(t):
    if t.dtype == torch.float32:
        return MV_F32
    if t.dtype == torch.bfloat16:
        return MV_BF16
    raise NotImplementedError(f"dtype {t.dtype} not supported by the native kernels")
