"""Host logic that decides which backward products go to the tcgen05 kernels (kernels._dw_tc_ok / _gemm_tc_ok): pure
functions of shapes, strides and alignment -- checked on CPU tensors (no compute, no GPU)."""
import torch

from tacotron_b200 import kernels as K


def _dw(A, B, Cm, **kw):
    a = dict(tb=False, beta=1.0, period=0, taps=1, dshift=0, kper=0, batch=1, a_bstride=0, b_bstride=0, c_bstride=0, bshift=0)
    a.update(kw)
    return K._dw_tc_ok(Cm, A, B, a["tb"], a["beta"], a["period"], a["taps"], a["dshift"], a["kper"], a["batch"], a["a_bstride"],
                       a["b_bstride"], a["c_bstride"], a["bshift"])


def test_weight_gradient_route_conditions():
    X, dZ, gW = torch.zeros(4000, 256), torch.zeros(4000, 128), torch.zeros(256, 128)
    assert _dw(X, dZ, gW)
    assert _dw(X, dZ, gW, beta=0.0)
    assert _dw(X, dZ, gW, batch=3, bshift=1, c_bstride=256 * 128, period=1000)         # conv taps as batch entries
    assert _dw(torch.zeros(4000, 384)[:, 128:], torch.zeros(4000, 1024)[:, 256:384], gW)    # column slices keep 16-byte rows
    assert not _dw(X, dZ, gW, tb=True)
    assert not _dw(X, dZ, gW, beta=0.5)
    assert not _dw(X, dZ, gW, batch=3, bshift=1, a_bstride=7)                            # per-batch operands: the attention products
    assert not _dw(X, dZ, gW, period=999)                                                # rows not a whole number of utterances
    assert not _dw(torch.zeros(4000, 1025), dZ, torch.zeros(1025, 128))                 # 1025-float rows: not TMA-readable
    assert not _dw(torch.zeros(4000, 257)[:, 1:], dZ, gW)                                # misaligned base
    assert not _dw(torch.zeros(100, 256), torch.zeros(100, 128), gW)                     # too few rows to be worth a launch
    assert _dw(K.padded_rows(4000, 1025, X), dZ, torch.zeros(1025, 128))                 # ... the padded-pitch view is


def test_plain_product_route_conditions():
    A, Cm = torch.zeros(6400, 512), torch.zeros(6400, 256)
    ok = lambda Cm_, A_, **kw: K._gemm_tc_ok(Cm_, A_, None, kw.get("beta", 0.0), kw.get("period", 0), kw.get("taps", 1), kw.get("dshift", 0),
                                            kw.get("kper", 0), kw.get("batch", 1))
    assert ok(Cm, A) and ok(Cm, A, beta=1.0) and ok(Cm, A, period=200)
    assert not ok(Cm, A, taps=3, kper=512)              # K-segmented form stays on taco_gemm (conv_dx has its own route)
    assert not ok(Cm, A, batch=4)
    assert not ok(Cm, A, period=300)                     # 6400 rows are not a multiple of 300
    assert not ok(torch.zeros(6400, 1025), A)           # output rows must be 16-byte aligned for the vector epilogue
    assert not ok(Cm, torch.zeros(6400, 513)[:, 1:])    # misaligned input base
    assert not ok(torch.zeros(100, 256), torch.zeros(100, 512))


def test_padded_rows_view():
    v = K.padded_rows(10, 1025, torch.zeros(1))
    assert v.shape == (10, 1025) and v.stride(0) == 1028 and v.stride(1) == 1 and v.data_ptr() % 16 == 0
    assert K.padded_rows(10, 128, torch.zeros(1)).stride(0) == 128
