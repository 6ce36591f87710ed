import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """Dump every oracle comparison of the session (what, max-rel, mean-rel, tolerance) next to the GPU logs."""
    try:
        from tests.util import PARITY_LOG
        if PARITY_LOG and os.environ.get("TACO_PARITY_LOG"):
            import json
            os.makedirs(os.path.dirname(os.path.abspath(os.environ["TACO_PARITY_LOG"])), exist_ok=True)
            with open(os.environ["TACO_PARITY_LOG"], "w") as f:
                json.dump([{"what": w, "max_rel": a, "mean_rel": b, "tol": t} for w, a, b, t in PARITY_LOG], f, indent=0)
    except Exception:
        pass
