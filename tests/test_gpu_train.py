"""GPU parity tests of the training path (backward + clip + Adam), through the C-ABI (tacotron_b200/kernels.py):
every training kernel against its torch-CPU mirror (tests/mirror_kernels.py -- the same functions the host logic is
pinned with on CPU), then the whole backward / optimizer step against torch.autograd over the oracle.

Tolerances (written per test): kernels are fp32 with re-ordered sums -> 1e-4 of the tensor's max magnitude; whole-model
gradients in 'fp32' precision mode -> 2e-3; in 'tf32' mode (TF32 tensor-core forward, fp32 backward) -> 3e-2.

STATUS (round 1): this file was written after the round's GPU budget was spent; the markers below say so.  A test
that passes shows up as XPASS in the round-end run; the xfail markers are to be removed once a hardware run is green.
"""
import pytest
import torch

from tests import train_checks as TC

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="training path: first hardware run pending (round-1 GPU budget spent)")]


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


@pytest.mark.parametrize("r,sched,precision,tol", [(2, True, "fp32", 2e-4), (5, False, "fp32", 2e-4), (2, True, "tf32", 2e-2)])
def test_train_forward_saves(r, sched, precision, tol):
    res = TC.check_train_forward(r, sched, precision)
    assert res["_missing"][0] == 0, [k for k in res if k.startswith("_missing_names")]
    _assert(res, tol)


@pytest.mark.parametrize("r,sched,precision,tol", [(2, True, "fp32", 2e-3), (5, False, "fp32", 2e-3), (5, True, "tf32", 3e-2)])
def test_model_backward_matches_autograd(r, sched, precision, tol):
    _assert(TC.check_model_bwd(r, sched, precision), tol, floor=1e-3)


def test_train_step_matches_oracle():
    _assert(TC.check_train_step(2, True, "fp32", steps=2), 2e-3, floor=1e-3)
