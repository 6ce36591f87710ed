"""GPU tests of the tcgen05 weight-gradient kernel (taco_conv_dw, kernels.DW_TC): dW[j,c,n] += sum_{b,t} x[b,t+tap0+j,c] dz[b,t,n]
(the tf.gradients of tf.layers.conv1d / tf.layers.dense w.r.t. the kernel: models/ops.py:54,80, tacotron.py:40,42,148,170).
Error-compensated 3xTF32 with fp32 accumulation in tensor memory: compared with a float64 evaluation of the same sums; the
bar is 2e-5 of max|ref| (measured ~3e-6; single-pass TF32 would be ~1e-3)."""
import pytest
import torch

from tests import mirror_kernels as MK

pytestmark = pytest.mark.gpu
TOL = 2e-5


@pytest.fixture()
def K():
    from tacotron_b200 import kernels as Kn
    prev, Kn.DW_TC = Kn.DW_TC, True
    prev_min, Kn.DW_TC_MIN_ROWS = Kn.DW_TC_MIN_ROWS, 1
    try:
        yield Kn
    finally:
        Kn.DW_TC = prev
        Kn.DW_TC_MIN_ROWS = prev_min


@pytest.mark.parametrize("taps,Cin,Cout,B,T", [(1, 256, 128, 4, 50), (3, 1024, 256, 2, 40), (3, 256, 80, 2, 70), (2, 80, 128, 3, 17),
                                               (7, 128, 128, 2, 33), (16, 128, 128, 2, 40), (5, 100, 132, 3, 95), (1, 128, 1028, 1, 333)])
def test_conv_dw_matches_float64(K, taps, Cin, Cout, B, T):
    g = torch.Generator().manual_seed(100 + taps)
    X = torch.randn(B * T, Cin, generator=g)
    dZ = torch.randn(B * T, Cout, generator=g)
    tap0 = -((taps - 1) // 2)
    for beta in (0.0, 1.0):
        gW = torch.randn(taps, Cin, Cout, generator=g)
        kw = dict(beta=beta, shift=tap0, bshift=1, batch=taps, c_bstride=Cin * Cout, period=T)
        ref = gW.double().clone()
        MK.gemm(ref.reshape(taps * Cin, Cout)[:Cin], X.double(), dZ.double(), ta=True, **kw)
        got = gW.cuda()
        K.gemm(got.reshape(taps * Cin, Cout)[:Cin], X.cuda(), dZ.cuda(), ta=True, **kw)
        torch.cuda.synchronize()
        err = (got.cpu().double() - ref).abs().max().item()
        assert err <= TOL * ref.abs().max().item(), (taps, beta, err, ref.abs().max().item())


def test_conv_dw_column_slices_and_row_shift(K):
    """operands as column slices of wider buffers (bank gradient, GRU [gates|cand] blocks) and the decoder-style row shift
    (time-major rows, shift = -B over the whole buffer, no period)"""
    g = torch.Generator().manual_seed(7)
    R, Bt = 600, 8
    wideX = torch.randn(R, 384, generator=g)
    wideZ = torch.randn(R, 1024, generator=g)
    X, dZ = wideX[:, 128:384], wideZ[:, 256:384]
    gW = torch.randn(256, 128, generator=g)
    ref = gW.double().clone()
    MK.gemm(ref, X.double(), dZ.double(), ta=True, beta=1.0, shift=-Bt)
    got = gW.cuda()
    Xg, Zg = wideX.cuda()[:, 128:384], wideZ.cuda()[:, 256:384]
    K.gemm(got, Xg, Zg, ta=True, beta=1.0, shift=-Bt)
    torch.cuda.synchronize()
    err = (got.cpu().double() - ref).abs().max().item()
    assert err <= TOL * ref.abs().max().item(), err


def test_conv_dw_full_size_post_projection(K):
    """the largest weight gradient of the model (post-net conv projection 3x1024 -> 256 over 32 x 1000 frames) against
    float64, and that the routed call is one launch"""
    g = torch.Generator().manual_seed(11)
    B, T, Cin, Cout, taps = 32, 1000, 1024, 256, 3
    X = (torch.randn(B * T, Cin, generator=g) * 0.5).cuda()
    dZ = torch.randn(B * T, Cout, generator=g).cuda()
    kw = dict(beta=1.0, shift=-1, bshift=1, batch=taps, c_bstride=Cin * Cout, period=T)
    got = torch.zeros(taps, Cin, Cout, device="cuda")
    n0 = K.L.lib().taco_launch_count()
    K.gemm(got.reshape(taps * Cin, Cout)[:Cin], X, dZ, ta=True, **kw)
    assert K.L.lib().taco_launch_count() - n0 == 1
    # float64 evaluation of the same sums on the device (library matmul as the yardstick)
    Xv, Zd = X.view(B, T, Cin).double(), dZ.double()
    ref = torch.zeros(taps, Cin, Cout, device="cuda", dtype=torch.float64)
    for j in range(taps):
        sh = -1 + j
        Xs = torch.zeros_like(Xv)
        if sh < 0:
            Xs[:, -sh:] = Xv[:, :T + sh]
        elif sh > 0:
            Xs[:, :T - sh] = Xv[:, sh:]
        else:
            Xs = Xv
        ref[j] = Xs.reshape(B * T, Cin).t() @ Zd
    torch.cuda.synchronize()
    err = (got.double() - ref).abs().max().item()
    assert err <= TOL * ref.abs().max().item(), (err, ref.abs().max().item())


@pytest.mark.parametrize("M,Kc,N,tb,beta,shift,period", [(6400, 512, 256, False, 1.0, -32, 0), (4096, 256, 128, True, 0.0, 0, 0),
                                                         (3000, 128, 768, False, 0.0, 1, 1000), (6400, 80, 256, False, 0.0, 0, 0),
                                                         (6400, 256, 400, True, 1.0, 0, 0), (3000, 256, 128, True, 1.0, -1, 1000)])
def test_plain_products_on_tensor_cores(M, Kc, N, tb, beta, shift, period):
    """kernels.GEMM_TC: y (+)= shift(x) . W and dx (+)= shift(dy) . W^T of the backward (activation recomputation and data
    gradients, with the row shift / period of the recurrent layers) through the tcgen05 forward kernel, 3xTF32"""
    from tacotron_b200 import kernels as Kn
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, Kc, generator=g)
    Bm = torch.randn(N, Kc, generator=g) if tb else torch.randn(Kc, N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    ref = C0.double().clone()
    MK.gemm(ref, A.double(), Bm.double(), tb=tb, beta=beta, shift=shift, period=period)
    got = C0.cuda()
    prev, Kn.GEMM_TC = Kn.GEMM_TC, True
    try:
        n0 = Kn.L.lib().taco_launch_count()
        Kn.gemm(got, A.cuda(), Bm.cuda(), tb=tb, beta=beta, shift=shift, period=period)
        assert Kn.L.lib().taco_launch_count() - n0 == 2          # weight pack + the tensor-core kernel (not taco_gemm)
    finally:
        Kn.GEMM_TC = prev
    torch.cuda.synchronize()
    err = (got.cpu().double() - ref).abs().max().item()
    assert err <= TOL * ref.abs().max().item(), (err, ref.abs().max().item())
