"""-m gpu: parity of taco_linear_fwd (tcgen05 TF32 path and exact-fp32 SIMT path) and of the small
bandwidth kernels against the CPU oracle primitives, through the C-ABI."""
import pytest
import torch

from oracle import tf12
from tests.util import assert_close

pytestmark = pytest.mark.gpu

TOL = {"fp32": 2e-5, "fp32x3": 2e-5, "tf32": 3e-3}   # max-abs error relative to max|ref|; TF32 has a 10-bit mantissa;
# fp32x3 = 3xTF32 on tcgen05 is held to the SAME bar as the exact-product FFMA kernel


def _rt(precision):
    from tacotron_b200.models import ops
    from tacotron_b200.params import ParamStore
    st = ParamStore([("dummy", (4,), "zeros")], "cuda")
    return ops.Runtime(st, precision), ops


def _pack(ops, rt, W, taps, Cin, N):
    if rt.precision == "fp32":
        return None
    return ops._pack(rt, None, W.contiguous(), taps, Cin, N)


@pytest.mark.parametrize("precision", ["fp32", "tf32", "fp32x3"])
@pytest.mark.parametrize("B,T,Cin,N", [(1, 300, 256, 1025), (2, 77, 80, 128), (3, 128, 128, 768), (1, 5, 32, 8),
                                       (24, 1000, 128, 768), (8, 1000, 256, 1025)])   # last two: several waves of tiles
                                                                                      # (16-byte-aligned rows / 4-byte-aligned rows = the shifted-store epilogue)
def test_dense(precision, B, T, Cin, N):
    rt, ops = _rt(precision)
    g = torch.Generator().manual_seed(B * 1000 + T)
    x = torch.randn(B, T, Cin, generator=g)
    W = torch.randn(Cin, N, generator=g) / Cin ** 0.5
    b = torch.randn(N, generator=g)
    ref = tf12.dense(x, W, b, torch.relu)
    Wd = W.cuda()
    y = ops.linear(rt, x.cuda(), Wd, _pack(ops, rt, Wd, 1, Cin, N), N, bias=b.cuda(), act=1)
    torch.cuda.synchronize()
    assert_close(y, ref, TOL[precision], f"dense {precision} {B}x{T}x{Cin}->{N}")


@pytest.mark.parametrize("precision", ["fp32", "tf32", "fp32x3"])
@pytest.mark.parametrize("B,T,Cin,N,taps", [(2, 128, 256, 128, 3), (3, 200, 1024, 256, 3), (2, 50, 80, 128, 4), (1, 1, 128, 128, 3),
                                          (2, 131, 128, 128, 7), (20, 1000, 256, 256, 3)])   # last: several waves of tiles
def test_conv_same_with_bn_residual(precision, B, T, Cin, N, taps):
    rt, ops = _rt(precision)
    g = torch.Generator().manual_seed(T)
    x = torch.randn(B, T, Cin, generator=g)
    W = torch.randn(taps, Cin, N, generator=g) / (taps * Cin) ** 0.5
    b = torch.randn(N, generator=g) * 0.1
    scale = 1 + 0.1 * torch.randn(N, generator=g)
    shift = 0.1 * torch.randn(N, generator=g)
    res = torch.randn(B, T, N, generator=g)
    ref = tf12.conv1d_same(x, W, b, torch.relu) * scale + shift + res
    Wd = W.cuda()
    y = ops.linear(rt, x.cuda(), Wd, _pack(ops, rt, Wd, taps, Cin, N), N, taps=taps, bias=b.cuda(), act=1,
                   scale=scale.cuda(), shift=shift.cuda(), residual=res.cuda())
    torch.cuda.synchronize()
    assert_close(y, ref, TOL[precision], f"conv{taps} {precision}")


@pytest.mark.parametrize("precision", ["fp32", "tf32", "fp32x3"])
@pytest.mark.parametrize("B,T,Cin,K", [(2, 128, 128, 16), (2, 300, 80, 8), (1, 127, 80, 8), (1, 255, 128, 4), (3, 20, 80, 8),
                                       (6, 1000, 80, 8)])           # last: several waves of tiles
def test_conv_bank_bn_pool(precision, B, T, Cin, K):
    """models/ops.py:54-71: K filters, concat, BN affine, max-pool(2,1,'same') -- one grouped call."""
    from tacotron_b200.models import ops as O2
    from tacotron_b200.params import ParamStore, cbhg_shapes
    st = ParamStore(cbhg_shapes("x/cbhg", Cin, K, (128, 128, Cin)), "cuda")
    g = torch.Generator().manual_seed(K * T)
    p = {}
    for n, (s, kind) in st.shapes.items():
        if "bank" in n:
            t = torch.randn(s, generator=g) * (0.5 if kind in ("gamma", "beta", "mean", "zeros") else 1.0)
            if kind == "conv":
                t = t / (s[0] * s[1]) ** 0.5
            if kind == "var":
                t = 1 + t.abs()
            p[n] = t
        else:
            p[n] = torch.zeros(s)
    st.load(p)
    O2.runtime(st, precision)
    x = torch.randn(B, T, Cin, generator=g)
    bank = torch.cat([tf12.conv1d_same(x, p[f"x/cbhg/bank/W{k}"], p[f"x/cbhg/bank/b{k}"], torch.relu) for k in range(1, K + 1)], -1)
    bank = tf12.batch_norm_inference(bank, p["x/cbhg/bank/bn_gamma"], p["x/cbhg/bank/bn_beta"], p["x/cbhg/bank/bn_mean"],
                                     p["x/cbhg/bank/bn_var"])
    ref = tf12.max_pool_2_1_same(bank)
    y = O2.conv1d_banks(x.cuda(), K=K, cout=128, scope=O2.Scope(st, "x/cbhg/bank"))
    torch.cuda.synchronize()
    assert_close(y, ref, TOL[precision], f"bank K={K} T={T} {precision}")


@pytest.mark.parametrize("precision", ["fp32", "tf32", "fp32x3"])
@pytest.mark.parametrize("M", [64, 1000, 40000])          # 40000 rows: several waves of tiles
def test_highway(precision, M):
    from tacotron_b200.models import ops as O2
    from tacotron_b200.params import ParamStore
    shapes = [("h/WT", (128, 128), "dense"), ("h/bT", (128,), "zeros"), ("h/WH", (128, 128), "dense"), ("h/bH", (128,), "zeros")]
    st = ParamStore(shapes, "cuda")
    g = torch.Generator().manual_seed(M)
    p = {n: torch.randn(s, generator=g) * (0.1 if len(s) == 2 else 0.3) for n, (s, _) in st.shapes.items()}
    st.load(p)
    O2.runtime(st, precision)
    x = torch.randn(1, M, 128, generator=g)
    T_ = torch.sigmoid(x @ p["h/WT"] + p["h/bT"]); H = torch.relu(x @ p["h/WH"] + p["h/bH"])
    ref = H * T_ + x * (1 - T_)
    y = O2.highway(x.cuda(), scope=O2.Scope(st, "h"))
    torch.cuda.synchronize()
    assert_close(y, ref, TOL[precision], f"highway {precision}")


@pytest.mark.parametrize("precision", ["fp32", "tf32", "fp32x3"])
@pytest.mark.parametrize("B,T", [(2, 40), (40, 1000)])      # (40, 1000): several waves of tiles
def test_dropout_keep_mask(precision, B, T):
    rt, ops = _rt(precision)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, T, 256, generator=g); W = torch.randn(256, 128, generator=g) / 16; b = torch.randn(128, generator=g)
    keep = (torch.rand(B, T, 128, generator=g) >= 0.5).to(torch.uint8)
    ref = tf12.dropout(tf12.dense(x, W, b, torch.relu), keep, 0.5)
    Wd = W.cuda()
    y = ops.linear(rt, x.cuda(), Wd, _pack(ops, rt, Wd, 1, 256, 128), 128, bias=b.cuda(), act=1, keep=keep.cuda(), keep_scale=2.0)
    torch.cuda.synchronize()
    assert_close(y, ref, TOL[precision], "dropout")
    assert torch.all(y.cpu()[keep == 0] == 0)


def test_tf32_vs_fp32_paths_agree_at_full_size():
    """Post-net projection at the C2 shape (32x1000x1024 -> 256): the two GPU paths against each other."""
    rt32, ops = _rt("fp32")
    rtt, _ = _rt("tf32")
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(32, 1000, 1024, device="cuda", generator=g)
    W = torch.randn(3, 1024, 256, device="cuda", generator=g) / 55.0
    a = ops.linear(rt32, x, W, None, 256, taps=3)
    b = ops.linear(rtt, x, W, _pack(ops, rtt, W, 3, 1024, 256), 256, taps=3)
    torch.cuda.synchronize()
    err = float((a - b).abs().max() / a.abs().max())
    assert err < 3e-3, err
    # linearity property of the tensor-core path: f(2x) == 2 f(x) exactly (power-of-two scaling)
    b2 = ops.linear(rtt, x * 2, W, _pack(ops, rtt, W, 3, 1024, 256), 256, taps=3, tag="lin2")
    torch.cuda.synchronize()
    assert torch.equal(b2, b * 2)
    # the 3xTF32 tensor-core path is fp32-grade at this size (K = 3072): 2e-5 of max|ref| against the exact-product kernel
    rt3, _ = _rt("fp32x3")
    c = ops.linear(rt3, x, W, _pack(ops, rt3, W, 3, 1024, 256), 256, taps=3, tag="lin3")
    torch.cuda.synchronize()
    err3 = float((a - c).abs().max() / a.abs().max())
    assert err3 < 5e-5, err3       # (both sides carry fp32 accumulation rounding over K = 3072; 2.6e-5 measured)
    # power-of-two scaling commutes exactly with the hi / lo split as well
    c2 = ops.linear(rt3, x * 2, W, _pack(ops, rt3, W, 3, 1024, 256), 256, taps=3, tag="lin3b")
    torch.cuda.synchronize()
    assert torch.equal(c2, c * 2)


def test_small_kernels():
    from tacotron_b200 import _lib as L
    lib = L.lib()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(3, 17, 64, generator=g)
    xd = x.cuda()
    y = torch.empty_like(x).cuda()
    L.check(lib.taco_maxpool_fwd(L.ptr(xd), L.ptr(y), 3, 17, 64, L.current_stream()))
    assert torch.equal(y.cpu(), tf12.max_pool_2_1_same(x))
    ln = torch.tensor([17, 0, 5], dtype=torch.int32)
    lnd = ln.cuda()
    L.check(lib.taco_mask_rows(L.ptr(xd), L.ptr(lnd), L.ptr(y), 3, 17, 64, L.current_stream()))
    ref = x * (torch.arange(17)[None, :, None] < ln[:, None, None])
    assert torch.equal(y.cpu(), ref)
    table = torch.randn(20, 32, generator=g); ids = torch.randint(0, 20, (50,), generator=g, dtype=torch.int32)
    out = torch.empty(50, 32).cuda()
    td, idd = table.cuda(), ids.cuda()           # keep the device tensors alive across the async launch
    L.check(lib.taco_gather_rows(L.ptr(td), L.ptr(idd), 50, 32, 20, None, 1.0, L.ptr(out), L.current_stream()))
    assert torch.equal(out.cpu(), table[ids.long()])
    a = torch.randn(100003, generator=g); b = torch.randn(100003, generator=g)
    part = torch.empty(lib.taco_l1_partial_count()).cuda(); res = torch.empty(1).cuda()
    ad, bd = a.cuda(), b.cuda()
    L.check(lib.taco_l1_loss_fwd(L.ptr(ad), L.ptr(bd), a.numel(), L.ptr(part), L.ptr(res), L.current_stream()))
    ref = (a.double() - b.double()).abs().sum()
    assert abs(float(res) - float(ref)) / float(ref) < 1e-6


def test_error_reporting_through_abi():
    from tacotron_b200 import _lib as L
    import ctypes as C
    d = L.LinearDesc()
    rc = L.lib().taco_linear_fwd(C.byref(d), None)
    assert rc != 0 and b"NULL" in L.lib().taco_last_error()
