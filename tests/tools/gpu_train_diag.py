"""One-shot hardware diagnostic of the training path: runs every check of tests/train_checks.py WITHOUT asserting and
writes per-tensor errors to gpurun_out/train_diag.txt (simple kernels first, the cooperative decoder kernel and the
whole model last, so that a fault late in the list does not hide the earlier results).

    gpurun -- python tests/tools/gpu_train_diag.py

(It lives under tests/ because it executes the oracle-backed checks of tests/train_checks.py: test infrastructure.)
"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import train_checks as TC  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
# optional argument: a group letter (A..F) so that each group runs in its own process -- a device fault in one
# group then cannot poison the CUDA context of the others.  No argument = everything in this process.
GROUP = sys.argv[1] if len(sys.argv) > 1 else ""
log = open(os.path.join(OUT, "train_diag.txt"), "a" if GROUP else "w")
GROUPS = {"A": ("smoke", "gemm", "elementwise"), "B": ("bigru",), "C": ("train_forward",), "D": ("decoder_bwd",),
          "E": ("model_bwd", "train_step"), "F": ("timing",)}


def wanted(name):
    return not GROUP or any(name.startswith(p) for p in GROUPS[GROUP])


def say(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    log.write(s + "\n"); log.flush()


CHECKS = [
    ("gemm", lambda: TC.check_gemm()),
    ("elementwise", lambda: TC.check_elementwise()),
    ("bigru_bwd B3 T9", lambda: TC.check_bigru_bwd(3, 9)),
    ("bigru_bwd B2 T1", lambda: TC.check_bigru_bwd(2, 1)),
    ("bigru_bwd B32 T40", lambda: TC.check_bigru_bwd(32, 40)),
    ("train_forward r2 sched fp32", lambda: TC.check_train_forward(2, True, "fp32")),
    ("train_forward r5 teacher fp32", lambda: TC.check_train_forward(5, False, "fp32")),
    ("train_forward r2 sched tf32", lambda: TC.check_train_forward(2, True, "tf32")),
    ("decoder_bwd r2 sched", lambda: TC.check_decoder_bwd(2, True)),
    ("decoder_bwd r5 teacher", lambda: TC.check_decoder_bwd(5, False)),
    ("decoder_bwd r5 sched B32", lambda: TC.check_decoder_bwd(5, True, B=32, Tx=32, T=6)),
    ("model_bwd r2 sched fp32", lambda: TC.check_model_bwd(2, True, "fp32")),
    ("model_bwd r5 teacher fp32", lambda: TC.check_model_bwd(5, False, "fp32")),
    ("model_bwd r5 sched tf32", lambda: TC.check_model_bwd(5, True, "tf32")),
    ("train_step r2 sched fp32 x2", lambda: TC.check_train_step(2, True, "fp32", steps=2)),
]


def timing():
    """C2-shaped training step (B=32, Tx=128, T=200, r=5): device time of forward+backward+optimizer."""
    from oracle import tacotron_oracle as O
    from tacotron_b200.models.tacotron import Config, Tacotron
    cfg = Config(r=5, vocab_size=64, precision="tf32")
    m = Tacotron(cfg, None, train=True)
    inp = O.synthetic_inputs(O.OracleConfig(r=5), 32, 128, 200, seed=0)
    gi = {k: v.cuda() for k, v in inp.items()}
    for _ in range(2):
        m.train_step(gi, lr=1e-4)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    n = 3
    for _ in range(n):
        m.train_step(gi, lr=1e-4)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    say(f"C2 train step: {ms:.2f} ms  -> {32000 / ms * 1e3:.0f} mel frames/s   loss {float(m.loss):.1f}")


if __name__ == "__main__":
    say(f"---- group {GROUP or 'all'}  device:", torch.cuda.get_device_name(0))
    if wanted("smoke"):
        try:                                               # the inference path must be untouched by the rebuilt library
            import __graft_entry__ as ge
            t0 = time.time()
            ge.smoke()
            say(f"== smoke() (inference path vs oracle): OK ({time.time() - t0:.1f}s)")
        except Exception:
            say("== smoke(): EXCEPTION")
            say(traceback.format_exc())
    for name, fn in CHECKS:
        if not wanted(name):
            continue
        t0 = time.time()
        try:
            res = fn()
            say(f"== {name}: worst rel {TC.worst_rel(res, 1e-3):.3e}  ({time.time() - t0:.1f}s)")
            say(TC.fmt(res))
        except Exception:
            say(f"== {name}: EXCEPTION")
            say(traceback.format_exc())
    if wanted("timing"):
        try:
            timing()
        except Exception:
            say("== timing: EXCEPTION")
            say(traceback.format_exc())
    log.close()
