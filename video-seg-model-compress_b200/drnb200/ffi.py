"""ctypes binding of libdrnb200.so (include/drnb200.h).

The library is the product: there is no Python/CPU fallback.  Loading fails loudly when the shared
object has not been built (run ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C video-seg-model-compress_b200/csrc``), and every call raises :class:`Drnb200Error` on a non-zero status.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdrnb200.so")

BF16, F16 = 0, 1
IMPL_AUTO, IMPL_DIRECT, IMPL_TCGEN05 = 0, 1, 2
HOST_PINNED, HOST_WC, HOST_HUGE = 0, 1, 2
KB_PROJ = 3 << 20        # DRNB200_KB_PROJ: tile-list entries of the residual projection (include/drnb200.h)


class Drnb200Error(RuntimeError):
    pass


class ConvDesc(C.Structure):
    """mirror of drnb200_conv_desc"""
    _fields_ = [(n, C.c_int32) for n in (
        "N", "H", "W", "Cin", "Cout", "ksize", "stride", "dilation", "relu", "has_residual",
        "act_dtype", "out_f32", "tile_o", "tile_ci", "impl", "x_cpitch", "res_cpitch", "res_coffset", "relu_n",
        "proj_cin", "acc_layout")]


# name -> (restype, argtypes); every symbol include/drnb200.h declares
_P = C.c_void_p
SIGNATURES = {
    "drnb200_version": (C.c_int, []),
    "drnb200_last_error": (C.c_char_p, []),
    "drnb200_compact_mask": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       _P, _P, _P, _P]),
    "drnb200_pack_weights": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       _P, _P, C.c_int, _P, _P]),
    "drnb200_conv_plan_create": (C.c_int, [C.POINTER(_P), C.POINTER(ConvDesc), _P, _P, _P, _P, _P]),
    "drnb200_conv_forward": (C.c_int, [_P, _P, _P, _P, _P]),
    "drnb200_conv_plan_impl": (C.c_int, [_P]),
    "drnb200_conv_plan_mode": (C.c_int, [_P]),
    "drnb200_conv_plan_tile_macs": (C.c_int64, [_P]),
    "drnb200_conv_plan_destroy": (None, [_P]),
    "drnb200_stem_forward": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       _P, _P]),
    "drnb200_stem_plan_create": (C.c_int, [C.POINTER(_P), _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, _P]),
    "drnb200_stem_plan_forward": (C.c_int, [_P, _P, _P, _P]),
    "drnb200_stem_plan_destroy": (None, [_P]),
    "drnb200_stem_plan_forward_u8": (C.c_int, [_P, _P, _P, C.c_int, _P, _P]),
    "drnb200_ingest_lut": (C.c_int, [_P, _P, C.c_int, _P]),
    "drnb200_head_plan_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, _P, _P, _P]),
    "drnb200_head_forward": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "drnb200_head_plan_destroy": (None, [_P]),
    "drnb200_head_plan_fused": (C.c_int, [_P]),
    "drnb200_head_plan_set_upsample": (C.c_int, [_P, C.c_int]),
    "drnb200_confusion": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int, _P, _P]),
    "drnb200_colorize": (C.c_int, [_P, C.c_int64, _P, C.c_int, _P, C.c_float, _P, _P]),
    "drnb200_labels_to_i64": (C.c_int, [_P, C.c_int64, _P, _P]),
    "drnb200_ms_accumulate": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int,
                                        _P, _P, _P, C.c_int, _P, _P, _P, C.c_int, C.c_int, _P]),
    "drnb200_ms_argmax": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "drnb200_resize_u8": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, _P, _P, C.c_int,
                                    _P, _P, _P, C.c_int, _P, _P]),
    "drnb200_host_alloc": (C.c_int, [C.POINTER(_P), C.c_uint64, C.c_int]),
    "drnb200_host_free": (C.c_int, [_P]),
}

_lib = None


def lib():
    """Load (once) and return the ctypes handle with typed signatures."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Drnb200Error(
                "libdrnb200.so is not built (%s missing). Build it with "
                "`make -C video-seg-model-compress_b200/csrc`; there is no CPU fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status, what=""):
    if status != 0:
        msg = lib().drnb200_last_error()
        raise Drnb200Error("%s failed with status %d: %s" % (
            what or "drnb200 call", status, msg.decode() if msg else "?"))


def ptr(t):
    """device pointer of a torch tensor (or None)"""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
