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


for name, B, T, Cin, N, taps in [("post in-proj 128->768", 32, 1000, 128, 768, 1), ("post dense 256->1025", 32, 1000, 256, 1025, 1),
                                 ("post proj1 3x1024->256", 32, 1000, 1024, 256, 3), ("post proj2 3x256->80", 32, 1000, 256, 80, 3),
                                 ("enc proj1 3x2048->128", 32, 128, 2048, 128, 3), ("enc proj2 3x128->128", 32, 128, 128, 128, 3),
                                 ("enc in-proj 128->768", 32, 128, 128, 768, 1)]:
    x = torch.randn(B, T, Cin, generator=g).cuda()
    W = (torch.randn(taps, Cin, N, generator=g) / (taps * Cin) ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    Wp = ops._pack(rt, None, W.contiguous(), taps, Cin, N)
    timeit(name, lambda: ops.linear(rt, x, W, Wp, N, taps=taps, bias=b, act=1))
