"""ctypes binding of libtmglow_b200.so (C ABI declared in include/tmglow_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an
exception is raised.  Build it with ``python deep-turbulence_b200/build.py``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TMGLOW_B200_LIB") or os.path.join(_HERE, "lib", "libtmglow_b200.so")   # override: A/B-testing builds

TMG_MAX_LEVELS = 6
TMG_FLAG_BN_TRAIN = 1
TMG_FLAG_SHARED_X = 2
PREC_FP32, PREC_TF32X3, PREC_TF32, PREC_F16X3, PREC_F16 = 0, 1, 2, 3, 4
PRECISIONS = {"fp32": PREC_FP32, "tf32x3": PREC_TF32X3, "tf32": PREC_TF32, "f16x3": PREC_F16X3, "f16": PREC_F16}

OK, ERR_BAD_CONFIG, ERR_BAD_SHAPE, ERR_NULL, ERR_WORKSPACE, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOT_READY = \
    0, -1, -2, -3, -4, -5, -6, -7


class TmgConfig(C.Structure):
    _fields_ = [("in_features", C.c_int32), ("out_features", C.c_int32), ("n_levels", C.c_int32),
                ("enc_blocks", C.c_int32 * TMG_MAX_LEVELS), ("glow_blocks", C.c_int32 * TMG_MAX_LEVELS),
                ("cond_features", C.c_int32), ("cglow_upscale", C.c_int32), ("growth_rate", C.c_int32),
                ("init_features", C.c_int32), ("rec_features", C.c_int32)]


_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_I = C.c_int
_SZ = C.c_size_t
_U32 = C.c_uint32
_I64 = C.c_int64

# name -> (restype, argtypes); every symbol include/tmglow_b200.h declares
SIGNATURES = {
    "tmg_version": (_I, []),
    "tmg_last_error": (C.c_char_p, []),
    "tmg_device_count": (_I, []),
    "tmg_model_create": (_I, [C.POINTER(TmgConfig), C.POINTER(_P)]),
    "tmg_model_destroy": (None, [_P]),
    "tmg_model_set_precision": (_I, [_P, _I]),
    "tmg_model_get_precision": (_I, [_P]),
    "tmg_conv3x3_workspace_bytes": (_SZ, [_I, _I]),
    "tmg_conv3x3": (_I, [_I, _P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P, _P, _SZ, _P]),
    "tmg_conv3x3_backward_workspace_bytes": (_SZ, [_I, _I, _I, _I, _I]),
    "tmg_conv3x3_backward": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "tmg_conv3x3_wgrad_tc_workspace_bytes": (_SZ, [_I, _I, _I, _I, _I]),
    "tmg_conv3x3_wgrad_tc": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _SZ, _P]),
    "tmg_flow_step_backward_workspace_bytes": (_SZ, [_P, _I, _I, _I, _I]),
    "tmg_flow_step_backward": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "tmg_model_overflow": (_I, [_P, _I, _P]),
    "tmg_model_param_entries": (_I64, [_P]),
    "tmg_model_param_name": (C.c_char_p, [_P, _I64]),
    "tmg_model_param_offset": (_I64, [_P, _I64]),
    "tmg_model_param_numel": (_I64, [_P, _I64]),
    "tmg_model_param_shape": (_I, [_P, _I64, C.POINTER(_I64)]),
    "tmg_model_param_total": (_I64, [_P]),
    "tmg_model_refresh": (_I, [_P, _P, _P]),
    "tmg_model_get_conv1x1": (_I, [_P, _I, _I, _I, _P, _P]),
    "tmg_workspace_bytes": (_SZ, [_P, _I, _I, _I]),
    "tmg_reconstruct": (_I, [_P, _I, _I, _I, _P, _PP, _PP, _PP, _P, _P, _PP, _PP, _P, _SZ, _U32, _P]),
    "tmg_tape_bytes": (_SZ, [_P, _I, _I, _I]),
    "tmg_reconstruct_train": (_I, [_P, _I, _I, _I, _P, _PP, _PP, _PP, _P, _P, _PP, _PP, _P, _SZ, _P, _SZ, _U32, _P]),
    "tmg_reconstruct_backward_workspace_bytes": (_SZ, [_P, _I, _I, _I]),
    "tmg_reconstruct_backward": (_I, [_P, _I, _I, _I, _P, _PP, _PP, _PP, _P, _P, _P, _PP, _PP, _PP, _PP, _P, _P, _SZ, _U32, _P]),
    "tmg_bptt_tape_bytes": (_SZ, [_P, _I, _I, _I, _I]),
    "tmg_bptt_workspace_bytes": (_SZ, [_P, _I, _I, _I, _I]),
    "tmg_bptt_forward": (_I, [_P, _I, _I, _I, _I, _P, _PP, _PP, _PP, _P, _P, _PP, _PP, _P, _SZ, _P, _SZ, _U32, _P]),
    "tmg_bptt_backward": (_I, [_P, _I, _I, _I, _I, _P, _PP, _PP, _PP, _P, _P, _P, _PP, _PP, _PP, _PP, _P, _P, _SZ, _U32, _P]),
    "tmg_backward_finalize": (_I, [_P, _P, _P]),
    "tmg_adam_workspace_bytes": (_SZ, []),
    "tmg_adam_step": (_I, [_P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _SZ, _P]),
    "tmg_backward_graph_stats": (_I, [_P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "tmg_forward": (_I, [_P, _I, _I, _I, _P, _P, _PP, _PP, _P, _P, _PP, _PP, _PP, _P, _SZ, _U32, _P]),
    "tmg_encoder_forward": (_I, [_P, _I, _I, _I, _P, _PP, _P, _P, _SZ, _U32, _P]),
    "tmg_squeeze_forward": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "tmg_squeeze_reverse": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "tmg_flow_step": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "tmg_split_forward": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "tmg_split_reverse": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "tmg_tmglow_loss_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "tmg_tmglow_loss": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, C.c_double, C.c_double, C.c_double, _P, _P, _P, _P,
                             _P, _SZ, _P]),
    "tmg_nchw_to_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "tmg_nhwc_to_nchw": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "tmg_launch_count": (_I64, [_I]),
    "tmg_profile_enable": (_I, [_I]),
    "tmg_profile_classes": (_I, []),
    "tmg_profile_class_name": (C.c_char_p, [_I]),
    "tmg_profile_query": (_I, [_I, C.POINTER(C.c_double), C.POINTER(_I64), C.POINTER(C.c_double),
                               C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it was not built -- no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libtmglow_b200.so not found at %s: build it with `python deep-turbulence_b200/build.py` "
                "(there is no CPU/PyTorch fallback for the TM-Glow hot path)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)       # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class TmgError(RuntimeError):
    pass


def check(status):
    """Map a tmg_status to the exception the reference would raise for the same precondition."""
    if status == OK:
        return
    msg = load().tmg_last_error().decode("utf-8", "replace")
    if status in (ERR_BAD_SHAPE, ERR_BAD_CONFIG):
        raise AssertionError(msg)            # the reference uses `assert` for shapes (flowUtils.py:112,136)
    raise TmgError("libtmglow_b200 error %d: %s" % (status, msg))


def ptr_array(ptrs):
    """void*[] from a list of ints/None."""
    arr = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return C.cast(arr, _PP)
