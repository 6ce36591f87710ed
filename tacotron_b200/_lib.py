"""ctypes binding of libtaco_b200.so (include/taco_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, an
exception is raised.  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtaco_b200.so")

ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3
IMPL_TC, IMPL_SIMT = 0, 1
EPI_NORMAL, EPI_HIGHWAY = 0, 1
DEC_INFER, DEC_TEACHER, DEC_SCHED = 0, 1, 2


class LinearDesc(C.Structure):
    _fields_ = [
        ("X", C.c_void_p), ("ldx", C.c_int64),
        ("B", C.c_int32), ("T", C.c_int32), ("C", C.c_int32),
        ("taps", C.c_int32), ("tap0", C.c_int32), ("N", C.c_int32),
        ("bank_K", C.c_int32), ("bank_cout", C.c_int32),
        ("W", C.c_void_p), ("Wp", C.c_void_p), ("ldwp", C.c_int64),
        ("Y", C.c_void_p), ("ldy", C.c_int64),
        ("bias", C.c_void_p), ("act", C.c_int32),
        ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("keep", C.c_void_p), ("keep_scale", C.c_float),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("epilogue", C.c_int32),
        ("hx", C.c_void_p), ("ldhx", C.c_int64),
        ("pool", C.c_int32), ("impl", C.c_int32),
    ]


class DecoderWeights(C.Structure):
    _fields_ = [
        ("pre_W1", C.c_void_p), ("pre_b1", C.c_void_p), ("pre_W2", C.c_void_p), ("pre_b2", C.c_void_p),
        ("in_W", C.c_void_p), ("in_b", C.c_void_p),
        ("gru_Wg", C.c_void_p * 3), ("gru_bg", C.c_void_p * 3), ("gru_Wc", C.c_void_p * 3), ("gru_bc", C.c_void_p * 3),
        ("out_W", C.c_void_p), ("out_b", C.c_void_p),
        ("att_Wq", C.c_void_p), ("att_v", C.c_void_p), ("att_Wa", C.c_void_p),
    ]


class DecoderArgs(C.Structure):
    _fields_ = [
        ("weights", C.POINTER(DecoderWeights)),
        ("packed", C.c_void_p), ("keys", C.c_void_p), ("values", C.c_void_p), ("text_length", C.c_void_p),
        ("mel", C.c_void_p), ("sample_mask", C.c_void_p), ("keep1", C.c_void_p), ("keep2", C.c_void_p),
        ("keep_scale", C.c_float), ("mode", C.c_int32),
        ("B", C.c_int32), ("Tx", C.c_int32), ("T", C.c_int32), ("r", C.c_int32),
        ("y", C.c_void_p), ("align", C.c_void_p), ("workspace", C.c_void_p), ("step_ns", C.c_void_p),
    ]


class TacoError(RuntimeError):
    pass


_lib = None

# every symbol include/taco_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "taco_last_error", "taco_version", "taco_device_info", "taco_linear_fwd", "taco_pack_weight",
    "taco_maxpool_fwd", "taco_gather_rows", "taco_mask_rows", "taco_bigru_fwd",
    "taco_decoder_packed_bytes", "taco_decoder_workspace_bytes", "taco_decoder_pack", "taco_decoder_fwd",
    "taco_l1_loss_fwd", "taco_l1_partial_count", "taco_launch_count",
]


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TacoError(
            f"{LIB_PATH} not found: the CUDA extension has not been built (run `python -m tacotron_b200.build` "
            "or __graft_entry__.build()). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.taco_last_error.restype = C.c_char_p
    L.taco_version.restype = C.c_int
    L.taco_device_info.argtypes = [C.POINTER(C.c_int)] * 3
    L.taco_linear_fwd.argtypes = [C.POINTER(LinearDesc), C.c_void_p]
    L.taco_pack_weight.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    L.taco_maxpool_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.taco_gather_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float,
                                   C.c_void_p, C.c_void_p]
    L.taco_mask_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.taco_bigru_fwd.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_void_p]
    L.taco_decoder_packed_bytes.argtypes = [C.c_int]
    L.taco_decoder_packed_bytes.restype = C.c_size_t
    L.taco_decoder_workspace_bytes.argtypes = [C.c_int] * 4
    L.taco_decoder_workspace_bytes.restype = C.c_size_t
    L.taco_decoder_pack.argtypes = [C.POINTER(DecoderWeights), C.c_int, C.c_void_p, C.c_void_p]
    L.taco_decoder_fwd.argtypes = [C.POINTER(DecoderArgs), C.c_void_p]
    L.taco_l1_loss_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.taco_l1_partial_count.restype = C.c_int
    L.taco_launch_count.restype = C.c_ulonglong
    for name in EXPORTS:                      # every declared symbol must resolve (AttributeError otherwise)
        getattr(L, name)
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = lib().taco_last_error().decode("utf-8", "replace")
        raise TacoError(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
