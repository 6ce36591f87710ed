"""The largest weight gradient of the C2 training step (post-net conv projection 3 x 1024 -> 256 over 32 x 1000 frames) on
the tcgen05 kernel, alone in the process -- the target command for ncu (scripts/gpu_profile.sh) and a quick timing.
    python scripts/profile_dw.py [repeats]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tacotron_b200 import kernels as K

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, T, Cin, Cout, taps = 32, 1000, 1024, 256, 3
g = torch.Generator().manual_seed(0)
X = torch.randn(B * T, Cin, generator=g).cuda()
dZ = torch.randn(B * T, Cout, generator=g).cuda()
gW = torch.zeros(taps, Cin, Cout, device="cuda")
K.DW_TC = True
kw = dict(ta=True, beta=1.0, shift=-1, bshift=1, batch=taps, c_bstride=Cin * Cout, period=T)
K.gemm(gW.reshape(taps * Cin, Cout)[:Cin], X, dZ, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    K.gemm(gW.reshape(taps * Cin, Cout)[:Cin], X, dZ, **kw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
flop = 2.0 * B * T * Cin * Cout * taps
print(f"taco_conv_dw 3x1024->256 over {B*T} rows: {ms*1e3:.1f} us = {flop/ms/1e9:.1f} TFLOP/s fp32-equivalent ({3*flop/ms/1e9:.0f} TFLOP/s of TF32 MMAs)")
