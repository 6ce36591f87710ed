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
IMPL_TC, IMPL_SIMT, IMPL_TC3 = 0, 1, 2
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
        ("h_save", C.c_void_p),
    ]


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_int64), ("B", C.c_void_p), ("ldb", C.c_int64), ("C", C.c_void_p), ("ldc", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("ta", C.c_int32), ("tb", C.c_int32), ("beta", C.c_float),
        ("shift", C.c_int32), ("period", C.c_int32), ("taps", C.c_int32), ("dshift", C.c_int32), ("kper", C.c_int32),
        ("b_tap_stride", C.c_int64), ("batch", C.c_int32),
        ("a_bstride", C.c_int64), ("b_bstride", C.c_int64), ("c_bstride", C.c_int64), ("bshift", C.c_int32),
    ]


class DecoderBwdArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("T", C.c_int32), ("Tx", C.c_int32), ("r", C.c_int32), ("keep_scale", C.c_float),
        ("W_a", C.c_void_p), ("W_q", C.c_void_p), ("W_out", C.c_void_p), ("W_in", C.c_void_p), ("W1", C.c_void_p),
        ("W2", C.c_void_p), ("v", C.c_void_p),
        ("Wg", C.c_void_p * 3), ("Wc", C.c_void_p * 3),
        ("dy_ext", C.c_void_p),
        ("RU", C.c_void_p * 3), ("C", C.c_void_p * 3), ("H", C.c_void_p * 3),
        ("align", C.c_void_p), ("values", C.c_void_p), ("keys", C.c_void_p),
        ("PQ", C.c_void_p), ("PN1", C.c_void_p), ("PN2", C.c_void_p), ("sel", C.c_void_p),
        ("DATT", C.c_void_p), ("DY", C.c_void_p), ("DPQ", C.c_void_p), ("DSCORE", C.c_void_p),
        ("DCTX", C.c_void_p), ("DG", C.c_void_p * 3), ("DC", C.c_void_p * 3), ("DZ", C.c_void_p), ("DPN2", C.c_void_p),
        ("DPN1", C.c_void_p), ("DX", C.c_void_p), ("workspace", C.c_void_p),
    ]


class TacoError(RuntimeError):
    pass


_lib = None

# every symbol include/taco_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "taco_last_error", "taco_version", "taco_device_info", "taco_linear_fwd", "taco_pack_weight", "taco_pack_weight_x3",
    "taco_maxpool_fwd", "taco_gather_rows", "taco_mask_rows", "taco_bigru_fwd",
    "taco_decoder_packed_bytes", "taco_decoder_workspace_bytes", "taco_decoder_pack", "taco_decoder_fwd",
    "taco_l1_loss_fwd", "taco_l1_partial_count", "taco_launch_count",
    # training path
    "taco_gemm", "taco_set_gemm_impl", "taco_conv_dw", "taco_colsum", "taco_bias_act", "taco_mul_shift", "taco_epi_bwd", "taco_epi_fwd_keep", "taco_bn_param_grad",
    "taco_maxpool_bwd", "taco_highway_fwd", "taco_highway_bwd", "taco_l1_bwd", "taco_l1_bwd_ld", "taco_scatter_add_rows", "taco_bigru_bwd",
    "taco_dec_inputs", "taco_decoder_bwd_workspace_bytes", "taco_decoder_bwd", "taco_attn_bwd_post", "taco_sumsq",
    "taco_adam_step",
    # spectrogram inversion (Griffin-Lim glue kernels)
    "taco_gl_init", "taco_gl_ola", "taco_gl_frame", "taco_gl_phase", "taco_rfft2048", "taco_irfft2048",
    # input data format
    "taco_normalize_f16",
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
    L.taco_conv_dw.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64] + [C.c_int32] * 6 + [C.c_void_p]
    L.taco_pack_weight_x3.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
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
    vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
    L.taco_gemm.argtypes = [C.POINTER(GemmDesc), vp]
    L.taco_colsum.argtypes = [vp, vp, i64, vp, i64, vp, i64, i32, i32, vp]
    L.taco_bias_act.argtypes = [vp, i64, i32, i32, vp, i32, vp]
    L.taco_mul_shift.argtypes = [vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, vp]
    L.taco_epi_bwd.argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, i32, i32, i32, vp, vp, f32, vp]
    L.taco_epi_fwd_keep.argtypes = [vp, i64, vp, i32, i32, f32, vp]
    L.taco_bn_param_grad.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp]
    L.taco_maxpool_bwd.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    L.taco_highway_fwd.argtypes = [vp, i64, vp, i64, vp, i64, i32, i32, vp]
    L.taco_highway_bwd.argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, i32, i32, vp]
    L.taco_l1_bwd.argtypes = [vp, vp, vp, i64, f32, vp]
    L.taco_l1_bwd_ld.argtypes = [vp, i64, vp, vp, i64, i64, f32, vp]
    L.taco_scatter_add_rows.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    L.taco_bigru_bwd.argtypes = [vp] * 8 + [i32, i32, vp]
    L.taco_dec_inputs.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    L.taco_decoder_bwd_workspace_bytes.restype = C.c_size_t
    L.taco_decoder_bwd.argtypes = [C.POINTER(DecoderBwdArgs), vp]
    L.taco_attn_bwd_post.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]
    L.taco_sumsq.argtypes = [vp, i64, vp, vp, vp]
    L.taco_adam_step.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, vp, vp]
    L.taco_gl_init.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]
    L.taco_gl_ola.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    L.taco_gl_frame.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    L.taco_gl_phase.argtypes = [vp, vp, vp, i64, vp]
    L.taco_rfft2048.argtypes = [vp, vp, i64, vp]
    L.taco_irfft2048.argtypes = [vp, vp, i64, vp]
    L.taco_normalize_f16.argtypes = [vp, vp, vp, vp, i64, i32, vp]
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
