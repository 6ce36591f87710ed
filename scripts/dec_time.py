"""Time only the decoder section of the C2 step (quick A/B of decoder variants)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tacotron_b200 import Config, Tacotron
B, TX, T, R = 32, 128, 200, 5
m = Tacotron(Config(r=R, vocab_size=64, max_decode_iter=T, precision="tf32"), None, train=False, seed=1)
g = torch.Generator().manual_seed(0)
inp = {"text": torch.randint(1, 64, (B, TX), generator=g, dtype=torch.int32).cuda(), "text_length": torch.full((B,), TX, dtype=torch.int32).cuda()}
m.step_ns = torch.zeros(T, dtype=torch.int64, device="cuda")
for _ in range(3): m.inference(inp, train=False)
torch.cuda.synchronize()
dec = []
for _ in range(5):
    m._marks = []
    m.inference(inp, train=False)
    torch.cuda.synchronize()
    tt = dict(m._marks); m._marks = None
    dec.append(tt["encoder"].elapsed_time(tt["decoder"]))
ns = m.step_ns.cpu().numpy()
print("LL mode", os.environ.get("TACO_DEC_LL", "default"), "decoder ms", statistics.median(dec), "step p50 us", float(statistics.median(((ns[1:]-ns[:-1])/1e3).tolist())))
if os.environ.get("TACO_TRACE"):
    import numpy as np
    ws = m.runtime.dec_ws.view(torch.int64).cpu().numpy()
    n_alloc = 2 + 13 * T
    total = len(ws) - (5 * n_alloc + 32)
    NORD = 12
    n = 2 + NORD * T
    tr = ws[total: total + n]
    names = ["IN", "G1", "C1", "G2", "C2", "G3", "C3", "OUT", "Q", "P1", "ATT", "P2"]
    d = np.diff(tr)[2:]
    d = d[: (len(d) // NORD) * NORD].reshape(-1, NORD)[5:]
    med = np.median(d, axis=0)
    print("per-slot median ns (CTA 0):", {n_: int(v) for n_, v in zip(names, med)}, "sum", int(med.sum()))
    ck = ws[total + n + 16: total + n + 16 + 4 * n].reshape(n, 4)[2:]
    ck = ck[: (len(ck) // NORD) * NORD].reshape(-1, NORD, 4)[5:-1]
    seg = np.stack([ck[:, :, 1] - ck[:, :, 0], ck[:, :, 2] - ck[:, :, 1], ck[:, :, 3] - ck[:, :, 2]], -1)   # inputs+mma, barrier, epilogue
    nxt = np.roll(ck[:, :, 0].reshape(-1), -1).reshape(ck.shape[0], NORD) - ck[:, :, 3]                         # epilogue end -> next slot start
    print("cycles (inputs+mma | barrier | epilogue | to-next):")
    for i, n_ in enumerate(names):
        print(f"  {n_:4s} {int(np.median(seg[:, i, 0])):6d} {int(np.median(seg[:, i, 1])):6d} {int(np.median(seg[:, i, 2])):6d} {int(np.median(nxt[:-1, i])):6d}")

