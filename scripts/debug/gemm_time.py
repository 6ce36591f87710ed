"""Times the post-net feed-forward contractions of C2 one by one (fp32x3), eager launches, CUDA events.
    gpurun -- python scripts/debug/gemm_time.py"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from tacotron_b200.models import ops
from tacotron_b200.params import ParamStore

st = ParamStore([("dummy", (4,), "zeros")], "cuda")
rt = ops.Runtime(st, "fp32x3")
g = torch.Generator().manual_seed(0)


def timeit(name, fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) / n * 1e3:8.1f} us", flush=True)


B, T = 32, 1000
for name, Cin, N, taps in [("in-proj 128->768", 128, 768, 1), ("dense 256->1025", 256, 1025, 1), ("proj1 3x1024->256", 1024, 256, 3),
                           ("proj2 3x256->80", 256, 80, 3), ("dense 128->128", 128, 128, 1)]:
    x = torch.randn(B, T, Cin, generator=g).cuda()
    W = (torch.randn(taps, Cin, N, generator=g) / (taps * Cin) ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    Wp = ops._pack(rt, None, W.contiguous(), taps, Cin, N)
    timeit(name, lambda: ops.linear(rt, x, W, Wp, N, taps=taps, bias=b, act=1))
