"""Compare the tcgen05 (3xTF32) data-gradient route with the taco_gemm route call by call inside a model backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tests import train_checks as TC
from tacotron_b200 import kernels as K

orig = K.conv_dx
log = []
def both(dX, dZ, W, T, beta=0.0):
    ref = dX.clone()
    prev, K.DX_TC = K.DX_TC, False
    orig(ref, dZ, W, T, beta=beta)
    K.DX_TC = True
    got = dX.clone()
    orig(got, dZ, W, T, beta=beta)
    K.DX_TC = prev
    torch.cuda.synchronize()
    err = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    log.append((err, tuple(W.shape), tuple(dZ.shape), dZ.stride(0), T, beta, dZ.data_ptr() % 128))
    dX.copy_(ref)
K.conv_dx = both
res = TC.check_model_bwd(2, True, "fp32x3")
for e in sorted(log, reverse=True)[:12]:
    print("err %.3e W%s dZ%s ld %d T %d beta %s align %d" % e)
print("rel_l2", res["_rel_l2"], "cos", res["_cosine"])
