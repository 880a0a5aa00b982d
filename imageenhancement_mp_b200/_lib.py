"""ctypes binding of ``libimgenh_b200.so`` (the C ABI declared in include/imgenh_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is
raised.  ``torch`` is only the tensor container - tensors are passed as raw device pointers
and the current CUDA stream as ``void*``.
"""
from __future__ import annotations

import collections
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libimgenh_b200.so")

IE_EPI_BF16_RASTER, IE_EPI_F32_NHWC, IE_EPI_F32_SOFTMAX = 0, 1, 2
IE_LAYOUT_X_DENSE, IE_LAYOUT_Y_DENSE = 1, 2


class ImgEnhError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    """struct ie_conv_desc (include/imgenh_b200.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "n_img", "h", "w", "hv", "wv", "kh", "kw", "cin", "x_pitch", "x_coff",
        "cout", "y_pitch", "y_coff", "relu", "epilogue", "dense")]


_P, _I, _LL, _F = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> argument ctypes (every function returns int unless noted)
SIZE_QUERIES = ("ie_maxpool2_stat_scratch_bytes", "ie_channel_mean_scratch_bytes")      # return a byte count, not a status
SIGNATURES = {
    "ie_version": [],
    "ie_sm_count": [],
    "ie_pack_conv_weights": [_P, _I, _I, _I, _I, _I, _P, _P],
    "ie_pack_input_im2col3x3": [_P, _I, _I, _I, _I, _I, _I, _P, _P],
    "ie_conv_first_layer_f32": [_P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P, _I, _I, _P],
    "ie_conv2d_nhwc_bf16": [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P, _LL, _P],
    "ie_conv_set_mode": [_I, _I],
    "ie_debug_conv2d_naive": [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P, _LL, _P],
    "ie_maxpool2_stat_scratch_bytes": [_I, _I, _I, _I, _I],
    "ie_maxpool2_nhwc_bf16": [_P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P, _P, _LL, _I, _I, _LL, _I, _P],
    "ie_upsample_bilinear_nhwc_bf16": [_P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P],
    "ie_channel_mean_scratch_bytes": [_I, _I, _I, _I, _I, _I],
    "ie_channel_mean_nhwc_bf16": [_P, _I, _I, _I, _I, _I, _I, _P, _P, _LL, _I, _I, _LL, _I, _P],
    "ie_broadcast_hw_bf16": [_P, _I, _I, _I, _I, _P, _I, _I, _I, _P],
    "ie_raster_to_nhwc_f32": [_P, _I, _I, _I, _I, _I, _I, _P, _I, _P],
    "ie_softmax_taps_f32": [_P, _I, _I, _I, _P, _P],
    "ie_cost_volume_f32": [_P, _I, _I, _I, _I, _P, _P, _P],
    "ie_kpn_apply_f32": [_P, _I, _P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ie_kpn_apply_tf32": [_P, _I, _P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ie_kpn_apply_tc": [_P, _I, _P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ie_convolve_filts_f32": [_P, _I, _P, _P, _I, _I, _I, _I, _I, _P],
    "ie_mean_hw_f32": [_P, _I, _I, _I, _I, _I, _P, _P],
    "ie_invert_preproc_f32": [_P, _I, _I, _I, _P, _I, _I, _I, _I, _P, _P],
    "ie_eval_metrics_f32": [_P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P, _P],
    "ie_eval_metrics_crops_f32": [_P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "ie_eval_metrics_tune": [_I, _I, _I],
    "ie_preprocess_tune": [_I],
    "ie_ssim_tune": [_I],
    "ie_metric_totals_ssim_f64": [_P, _P, _I, _I, _I, _I, _I, _P, _P],
    "ie_metric_totals_f64": [_P, _I, _I, _I, _I, _I, _P, _P],
    "ie_sqdiff_sum_f32": [_P, _P, _I, _LL, _P, _P],
    "ie_img_loss_sums_f32": [_P, _P, _I, _I, _I, _P, _P],
    "ie_ssim_f32": [_P, _P, _I, _I, _I, _P, _P],
    "ie_preprocess_u8_rng": [_P, _I, _I, _I, _I, _P, _I, _F, _P, _P, _P, C.c_ulonglong, _I, _I, _I, _I, _P, _P, _P],
    "ie_preprocess_u8": [_P, _I, _I, _I, _I, _P, _I, _F, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P],
}

_lib = None
LAUNCHES = collections.Counter()      # C-ABI calls made (each enqueues at least one kernel of ours; exactly one per call in bench.py's step)
TRACE = None                          # set to a list to collect (name, start_event, end_event) per call (profiling)


def load():
    """Load the shared library (once) and attach prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImgEnhError(
            f"{LIB_PATH} not found: build it with `python -m imageenhancement_mp_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = C.c_longlong if name in SIZE_QUERIES else C.c_int
    lib.ie_last_error.argtypes = []
    lib.ie_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    lib = load()
    LAUNCHES[name] += 1
    if TRACE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        TRACE.append((name, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.ie_last_error().decode("utf-8", "replace")
        raise ImgEnhError(f"{name} failed ({rc}): {msg}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ImgEnhError("imageenhancement_mp_b200 runs on CUDA tensors only (no CPU fallback)")
