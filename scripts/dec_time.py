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
    total = (11 * 32 + 48) * 4 * 64             # exchange buffers [4 row groups][nkt][8][8] words (decoder.cu build_ws_layout)
    names = ["IN", "G1", "C1", "G2", "C2", "G3", "C3", "OQP", "A|P2"]
    n = 1 + len(names) * T
    tr = ws[total: total + n]
    d = np.diff(tr)                              # d[i] = time of entry i+1 - entry i; entry 0 = prologue
    d = d[: (len(d) // len(names)) * len(names)].reshape(-1, len(names))[5:]
    med = np.median(d, axis=0)
    print("per-slot median ns (CTA 0):", {n_: int(v) for n_, v in zip(names, med)}, "sum", int(med.sum()))
    ck = ws[total + 16 * T + 16: total + 16 * T + 16 + 4 * 20 * 8].reshape(4, 20, 8)
    inames = ["IN", "G1", "C1", "G2", "C2", "G3", "C3", "OQP", "AP2"]
    print("stage: start->on-chain loads issued | (off-chain half) ->on-chain weights | ->mma done | (finish) ->partials | ->sync | ->epilogue   [cycles, CTA0 thread0, step 11]")
    st = 1
    for i, nm in enumerate(inames):
        c = ck[st, i]
        if c[3] == 0: continue
        f = lambda a, b: int(c[a] - c[b]) if c[a] and c[b] else -1
        nxt = ck[st, i + 1, 3] if i + 1 < len(inames) else ck[st + 1, 0, 3]
        print(f"  {nm:8s} {f(0,3):6d} {f(1,0):6d} {f(2,1):6d} | {f(4,2):6d} {f(5,4):6d} {f(6,5):6d} | total {int(nxt - c[3]):6d}")
    c = ck[1, 8]      # A|P2 stage record of step 11: attention on warp 0 / thread 0
    if c[3]:
        print("attention (cycles from slot start): q arrived+exp %d | bar1 %d | scores %d | bar2 %d | stats+bar3 %d | ctx partial+push %d | merged ctx stored %d"
              % tuple(int(c[i] - c[3]) for i in (0, 1, 2, 4, 5, 6, 7)))
