"""GPU parity tests of the training path (backward + clip + Adam), through the C-ABI (tacotron_b200/kernels.py):
every training kernel against its torch-CPU mirror (tests/mirror_kernels.py -- the same functions the host logic is
pinned with on CPU), then the whole backward / optimizer step against torch.autograd over the oracle.

Tolerances (written per test): kernels are fp32 with re-ordered sums -> 1e-4 of the tensor's max magnitude; whole-model
gradients in 'fp32' precision mode -> 2e-3 per tensor.  In 'tf32' mode the FORWARD runs on TF32 tensor cores, so a
few units sit on the other side of a derivative discontinuity (sign() of the L1 loss, relu masks, max-pool argmax)
than in the fp32 oracle: per-tensor maxima then differ by several percent by construction, and the meaningful bar is
the agreement of the whole gradient vector (relative L2 <= 0.15, cosine >= 0.98).

First hardware run (profiles/r01_train_diag.txt): every fp32 check within 12 % of its allowance.
"""
import pytest
import torch

from tests import train_checks as TC

pytestmark = pytest.mark.gpu


def _assert(res, tol, floor=1e-6):
    bad = {k: v for k, v in res.items() if not k.startswith("_") and v[0] > tol * (v[1] + floor) + 1e-7}
    assert not bad, "mismatch:\n" + TC.fmt(bad)


def test_gemm_variants():
    _assert(TC.check_gemm(), 1e-4)


def test_elementwise_and_reductions():
    _assert(TC.check_elementwise(), 1e-4)


@pytest.mark.parametrize("B,T", [(3, 9), (2, 1), (32, 40)])
def test_bigru_bwd(B, T):
    _assert(TC.check_bigru_bwd(B, T), 1e-4)


@pytest.mark.parametrize("r,sched", [(2, True), (5, False), (5, True)])
def test_decoder_bwd_kernel(r, sched):
    _assert(TC.check_decoder_bwd(r, sched), 2e-4)


def test_decoder_bwd_kernel_full_batch():
    _assert(TC.check_decoder_bwd(5, True, B=32, Tx=32, T=6), 2e-4)


@pytest.mark.parametrize("r,sched,precision,tol", [(2, True, "fp32", 2e-4), (5, False, "fp32", 2e-4), (2, True, "tf32", 2e-2),
                                                   (2, True, "fp32x3", 2e-4)])
def test_train_forward_saves(r, sched, precision, tol):
    res = TC.check_train_forward(r, sched, precision)
    assert res["_missing"][0] == 0, [k for k in res if k.startswith("_missing_names")]
    _assert(res, tol)


@pytest.mark.parametrize("r,sched", [(2, True), (5, False)])
def test_model_backward_matches_autograd(r, sched):
    res = TC.check_model_bwd(r, sched, "fp32")
    _assert(res, 2e-3, floor=1e-3)
    assert res["_rel_l2"][0] <= 1e-4 and res["_cosine"][0] >= 0.9999


@pytest.mark.parametrize("r,sched", [(2, True), (5, False)])
def test_model_backward_fp32x3_matches_autograd(r, sched):
    """the default precision mode: 3xTF32 tcgen05 forward, data gradients on the same tcgen05 kernel (3xTF32), weight gradients on
    the 3xTF32 mma.sync GEMM -- held to the SAME bars as the exact-product fp32 mode"""
    res = TC.check_model_bwd(r, sched, "fp32x3")
    _assert(res, 2e-3, floor=1e-3)
    assert res["_rel_l2"][0] <= 1e-4 and res["_cosine"][0] >= 0.9999


def test_model_backward_tf32_forward():
    res = TC.check_model_bwd(5, True, "tf32")
    assert res["_rel_l2"][0] <= 0.15 and res["_cosine"][0] >= 0.98, (res["_rel_l2"], res["_cosine"])


@pytest.mark.parametrize("precision", ["fp32", "fp32x3"])
def test_train_step_matches_oracle(precision):
    res = TC.check_train_step(2, True, precision, steps=2)
    params = {k: v for k, v in res.items() if k.startswith("param_worst")}
    _assert({k: v for k, v in res.items() if k not in params}, 1e-4, floor=1e-3)      # losses and global gradient norms
    # parameters after 2 Adam steps of lr 1e-3 (each moves a parameter by ~1e-3): 5e-4 catches a missing or
    # wrong-signed update, and tolerates the sign noise of parameters whose true gradient is at rounding level
    # (measured on B200: 2.4e-5)
    for k, (e, r) in params.items():
        assert e <= 5e-4, (k, e, r)
