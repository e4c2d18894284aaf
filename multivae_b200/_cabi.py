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
DIST = {"normal": 0, "laplace": 1, "bernoulli": 2, "categorical": 3}
LATENT = {"laplace_with_softmax": 0, "normal": 1, "normal_with_softplus": 1}
STD_KIND = {"laplace_with_softmax": 0, "normal": 1, "normal_with_softplus": 2}   # _log_var_to_std variants
LOSS = {"iwae_looser": 0, "dreg_looser": 1}
ACT = {"none": 0, "relu": 1, "lrelu": 2, "sigmoid": 3}

_lib = None
_raw = None

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

class TapGemmArgs(ctypes.Structure):
    """mv_tapgemm_args of include/multivae_b200.h (field order and types must match)."""
    _fields_ = [
        ("A", c_void_p), ("a_rows", c_int64), ("a_ld", ctypes.c_int32), ("Cin", ctypes.c_int32),
        ("Wt", c_void_p), ("T", ctypes.c_int32), ("tap_off", ctypes.c_int32 * 9),
        ("N_total", ctypes.c_int32), ("BN", ctypes.c_int32), ("P", c_int64),
        ("bias", c_void_p), ("act", ctypes.c_int32), ("alpha", c_float),
        ("res", c_void_p), ("res_ld", ctypes.c_int32),
        ("dact1", c_void_p), ("dact1_ld", ctypes.c_int32), ("slope1", c_float),
        ("out", c_void_p), ("out_ld", ctypes.c_int32),
        ("out2", c_void_p), ("out2_ld", ctypes.c_int32), ("out2_pre", ctypes.c_int32), ("alpha2", c_float),
        ("dact2", c_void_p), ("dact2_ld", ctypes.c_int32), ("slope2", c_float),
        ("img_stride", ctypes.c_int32), ("Wp", ctypes.c_int32), ("W", ctypes.c_int32), ("H", ctypes.c_int32),
        ("n_img", ctypes.c_int32), ("out_mode", ctypes.c_int32), ("n_valid", ctypes.c_int32),
        ("out2_mask", c_void_p), ("dmask2", c_void_p), ("dmask1", c_void_p),
        ("res_mask", c_void_p), ("res_scale_pos", c_float), ("res_scale_neg", c_float),
        ("A2", c_void_p), ("a2_ld", ctypes.c_int32), ("W2", c_void_p),
    ]


class GemmArgs(ctypes.Structure):
    """mv_gemm_args of include/multivae_b200.h."""
    _fields_ = [("A", c_void_p), ("M", c_int64), ("a_ld", ctypes.c_int32), ("a_mn", ctypes.c_int32),
                ("B", c_void_p), ("N", ctypes.c_int32), ("b_ld", ctypes.c_int32), ("b_mn", ctypes.c_int32),
                ("K", ctypes.c_int32), ("bias", c_void_p), ("act", ctypes.c_int32), ("alpha", c_float),
                ("dact", c_void_p), ("dact_ld", ctypes.c_int32), ("dslope", c_float),
                ("out", c_void_p), ("out_ld", ctypes.c_int32), ("out_kind", ctypes.c_int32)]


class ConvGeom(ctypes.Structure):
    """mv_conv_geom of include/multivae_b200.h."""
    _fields_ = [(k, ctypes.c_int32) for k in ("n_img", "H", "W", "C", "nchw", "kh", "kw", "stride", "pad", "grid_h", "grid_w", "ld", "tc_order")]


class PackItem(ctypes.Structure):
    """mv_pack_item of include/multivae_b200.h."""
    _fields_ = [("src", c_void_p), ("dst_fwd", c_void_p), ("dst_dgrad", c_void_p), ("N", ctypes.c_int32), ("C", ctypes.c_int32),
                ("T", ctypes.c_int32), ("Npad", ctypes.c_int32), ("Cpad", ctypes.c_int32)]


class UnpackItem(ctypes.Structure):
    """mv_unpack_item of include/multivae_b200.h."""
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("N", ctypes.c_int32), ("C", ctypes.c_int32), ("T", ctypes.c_int32),
                ("Npad", ctypes.c_int32), ("Cpad", ctypes.c_int32), ("swapped", ctypes.c_int32)]


# symbol -> argtypes, exactly the prototypes in include/multivae_b200.h
_PROTOS = {
    "mv_version": [ctypes.POINTER(c_int)] * 3,
    "mv_moe_lpx_fwd": [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int, c_float, c_float,
                       c_void_p, c_int, c_void_p],
    "mv_moe_lpx_bwd": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int,
                       c_float, c_float, c_void_p, c_void_p],
    "mv_moe_lpx_fwd_multi": [c_int, ctypes.POINTER(c_void_p), c_int, ctypes.POINTER(c_void_p), c_void_p, c_int, c_int, c_int, c_int64,
                             c_int, ctypes.POINTER(c_float), ctypes.POINTER(c_float), ctypes.POINTER(c_void_p), c_void_p],
    "mv_moe_lpx_bwd_multi": [c_int, ctypes.POINTER(c_void_p), c_int, ctypes.POINTER(c_void_p), c_void_p, c_void_p,
                             ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int64, c_int, ctypes.POINTER(c_float),
                             ctypes.POINTER(c_float), ctypes.POINTER(c_void_p), c_void_p],
    "mv_moe_lpx_cat_fwd": [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_int, c_void_p],
    "mv_moe_lpx_cat_bwd": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                           c_void_p, c_void_p],
    "mv_logmeanexp": [c_void_p, c_int, c_int, c_void_p, c_void_p],
    "mv_gauss_kl_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p],
    "mv_gauss_kl_bwd": [c_void_p] * 9 + [c_int64, c_int, c_int, c_void_p],
    "mv_moe_lw_fwd": [c_void_p] * 21 + [c_int] * 7 + [c_float, c_int, c_int, c_void_p],
    "mv_moe_sample_fwd": [c_void_p] * 9 + [c_int] + [c_void_p] * 5 + [c_int] * 5 + [c_void_p],
    "mv_moe_sample_bwd": [c_void_p] * 13 + [c_int] + [c_void_p] * 5 + [c_int] * 5 + [c_void_p],
    "mv_poe_fwd": [c_void_p] * 4 + [c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_float] +
                  [c_void_p] * 5 + [c_int] * 3 + [c_void_p],
    "mv_tapgemm": [ctypes.POINTER(TapGemmArgs), c_void_p],
    "mv_gemm": [ctypes.POINTER(GemmArgs), c_void_p],
    "mv_im2col": [c_void_p, c_int, c_void_p, ctypes.POINTER(ConvGeom), c_void_p],
    "mv_col2im": [c_void_p, c_int, c_void_p, ctypes.POINTER(ConvGeom), c_void_p, c_int, c_void_p, c_float, c_void_p],
    "mv_chan_sum_nchw": [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    "mv_pack_tc": [ctypes.POINTER(PackItem), c_int, c_void_p],
    "mv_unpack_tc_add": [ctypes.POINTER(PackItem), c_int, c_void_p],
    "mv_colsum_any": [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p],
    "mv_act_bwd": [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p],
    "mv_upsample2x_fwd": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "mv_upsample2x_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p],
    "mv_head_grad_pack": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p],
    "mv_colsum": [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p],
    "mv_avgpool3s2_fwd": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "mv_avgpool3s2_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p],
    "mv_scale_dact": [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_float, c_void_p],
    "mv_lrelu_fwd": [c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p],
    "mv_gather_cast": [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_void_p],
    "mv_gather_f32": [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_float, c_void_p],
    "mv_pack_conv_weights": [ctypes.POINTER(PackItem), c_int, c_void_p],
    "mv_unpack_wgrad_add": [ctypes.POINTER(UnpackItem), c_int, c_void_p],
    "mv_wgrad_slice": [c_void_p, c_int64, c_int, c_int, c_void_p, c_int64, c_int, c_int, c_int, ctypes.POINTER(ctypes.c_int32),
                       c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p],
    "mv_wgrad": [c_void_p, c_int64, c_int, c_int, c_void_p, c_int64, c_int, c_int, c_int, ctypes.POINTER(ctypes.c_int32),
                 c_int64, c_void_p, c_void_p, c_void_p],
    "mv_wgrad_nct": [c_void_p, c_int64, c_int, c_int, c_void_p, c_int64, c_int, c_int, c_int, ctypes.POINTER(ctypes.c_int32),
                     c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p],
    "mv_poe_bwd": [c_void_p] * 4 + [c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_float] +
                  [c_void_p] * 5 + [c_int] * 3 + [c_void_p],
}


class NativeLibraryError(RuntimeError):
    pass


class KernelTimer:
    """Optional per-entry-point device timing: when active, every C-ABI call is bracketed by CUDA
    events on the launching (current) stream; `summary()` synchronises once and returns
    {symbol: (calls, total_ms)}.  Used by bench.py for the live roofline numbers."""

    def __init__(self):
        self.records = []

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for key, e0, e1 in self.records:
            c, t = out.get(key, (0, 0.0))
            out[key] = (c + 1, t + e0.elapsed_time(e1))
        return out


_timer = None


def set_timer(t):
    global _timer
    _timer = t


class _Lib:
    """Attribute proxy over the CDLL: same callables, plus the optional KernelTimer bracket."""

    def __init__(self, raw):
        self._raw = raw

    def __getattr__(self, name):
        fn = getattr(self._raw, name)
        if name not in _PROTOS or name == "mv_version":
            return fn

        def call(*args, tag=None):
            if _timer is None:
                return fn(*args)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*args)
            e1.record()
            _timer.records.append((name if tag is None else f"{name}:{tag}", e0, e1))
            return r

        self.__dict__[name] = call
        return call


def lib():
    """Load (once) and return the shared library; raise loudly if it has not been built."""
    global _lib, _raw
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "multivae_b200 has no CPU or eager fallback.")
        _raw = ctypes.CDLL(LIB_PATH)
        _raw.mv_last_error.restype = ctypes.c_char_p
        _raw.mv_launch_count.restype = ctypes.c_ulonglong
        _raw.mv_launch_count.argtypes = []
        for name, args in _PROTOS.items():
            fn = getattr(_raw, name)
            fn.argtypes = args
            fn.restype = c_int
        _lib = _Lib(_raw)
    return _lib


def launch_count():
    return int(lib().mv_launch_count())


def exported_symbols():
    return ["mv_last_error", "mv_launch_count"] + list(_PROTOS)


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


def ptr_array(tensors):
    """Host array of device pointers (None -> NULL) for the batched entry points."""
    return (c_void_p * len(tensors))(*[None if t is None else ptr(t) for t in tensors])


def float_array(values):
    return (c_float * len(values))(*[float(v) for v in values])


def stream():
    return torch.cuda.current_stream().cuda_stream


def dtype_code(t):
    if t.dtype == torch.float32:
        return MV_F32
    if t.dtype == torch.bfloat16:
        return MV_BF16
    raise NotImplementedError(f"dtype {t.dtype} not supported by the native kernels")
