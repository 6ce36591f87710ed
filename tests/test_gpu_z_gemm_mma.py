"""GPU tests of the opt-in tensor-core variant of taco_gemm (3xTF32 mma.sync, csrc/train.cu::gemm_mma_kernel): the same
checks as the default FFMA kernel (every addressing form against the torch-CPU mirror, then the whole backward against
autograd over the oracle), with the implementation switched by taco_set_gemm_impl(1).  Tolerance 2e-5 of max|ref| for
the kernel (3xTF32 keeps ~21 bits per product), 2e-3 per tensor for whole-model gradients (same bar as the default).

STATUS (round 1): written after the round's GPU budget was spent -- no hardware run yet, hence the non-strict xfail
markers (XPASS in the round-end run = it works; the variant becomes the default only after that).
"""
import pytest

from tests import train_checks as TC

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="mma.sync GEMM variant: first hardware run pending (round-1 GPU budget spent)")]


@pytest.fixture()
def mma_impl():
    from tacotron_b200 import kernels as K
    prev = K.set_gemm_impl(1)
    try:
        yield
    finally:
        K.set_gemm_impl(prev)


def _assert(res, tol, floor=1e-6):
    bad = {k: v for k, v in res.items() if not k.startswith("_") and v[0] > tol * (v[1] + floor) + 1e-7}
    assert not bad, "mismatch:\n" + TC.fmt(bad)


def test_gemm_variants_mma(mma_impl):
    _assert(TC.check_gemm(), 2e-5)


def test_model_backward_mma(mma_impl):
    res = TC.check_model_bwd(2, True, "fp32")
    _assert(res, 2e-3, floor=1e-3)
    assert res["_rel_l2"][0] <= 1e-4 and res["_cosine"][0] >= 0.9999


def test_setting_is_restored():
    from tacotron_b200 import kernels as K
    assert K.set_gemm_impl(0) == 0
