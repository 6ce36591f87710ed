"""GPU tests of the opt-in tensor-core variant of taco_gemm (3xTF32 mma.sync, csrc/train.cu::gemm_mma_kernel): the same
checks as the default FFMA kernel (every addressing form against the torch-CPU mirror, then the whole backward against
autograd over the oracle), with the implementation switched by taco_set_gemm_impl(1).  Tolerance 2e-5 of max|ref| for
the kernel (3xTF32 keeps ~21 bits per product), 2e-3 per tensor for whole-model gradients (same bar as the default).

STATUS: green on B200 (profiles/r01_pytest_mma_audio_runxfail.log); since then the tensor-core kernel is what the model's
backward uses in 'tf32' precision mode (Tacotron.backward), the FFMA kernel in 'fp32' mode.
"""
import pytest

from tests import train_checks as TC

pytestmark = pytest.mark.gpu


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


def test_model_backward_mma():
    res = TC.check_model_bwd(2, True, "fp32", gemm_impl=1)
    _assert(res, 2e-3, floor=1e-3)
    assert res["_rel_l2"][0] <= 1e-4 and res["_cosine"][0] >= 0.9999


def test_setting_is_restored():
    from tacotron_b200 import kernels as K
    assert K.set_gemm_impl(0) == 0
