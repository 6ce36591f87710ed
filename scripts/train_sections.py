"""Where the C2 training step spends its time: CUDA events around every kernel-namespace call of the backward
(a timing proxy over tacotron_b200/kernels.py) and around the forward sections.  Writes gpurun_out/train_sections.txt.

    gpurun -- python scripts/train_sections.py
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tacotron_b200 import _lib as L, kernels as K  # noqa: E402
from tacotron_b200.models import grad, ops  # noqa: E402
from tacotron_b200.models.tacotron import Config, Tacotron  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(OUT, exist_ok=True)


class TimedK:
    """proxy over the kernel namespace: records (label, start event, end event) per call"""
    def __init__(self):
        self.rec = []
        self.section = ""

    def __getattr__(self, name):
        fn = getattr(K, name)
        if name in ("empty", "zeros", "padded_rows"):
            return fn

        def wrapped(*a, **kw):
            label = name
            if name == "gemm":
                C, A, B = a[0], a[1], a[2]
                kind = ("dW" if kw.get("ta") else ("dX" if kw.get("tb") else "fwd"))
                Kd = A.shape[0] if kw.get("ta") else A.shape[1] * kw.get("taps", 1)
                label = f"gemm.{kind} {C.shape[0]}x{C.shape[1]}x{Kd}" + (f" b{kw['batch']}" if kw.get("batch", 1) > 1 else "")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **kw)
            e1.record()
            self.rec.append((self.section, label, e0, e1))
            return r
        return wrapped


def main():
    precision = os.environ.get("PRECISION", "fp32x3")
    cfg = Config(r=5, vocab_size=64, precision=precision)
    # the kernel routes Tacotron.backward() selects for this precision (models/tacotron.py backward())
    K.set_gemm_impl(int(os.environ.get("TACO_GEMM_IMPL", "0" if precision == "fp32" else "1")))
    K.DX_TC = K.DW_TC = K.GEMM_TC = precision != "fp32"
    K.DX_TC_IMPL = L.IMPL_TC if precision == "tf32" else L.IMPL_TC3
    m = Tacotron(cfg, None, train=True)
    g = torch.Generator().manual_seed(0)
    gi = {"text": torch.randint(1, 64, (32, 128), generator=g, dtype=torch.int32).cuda(),
          "text_length": torch.full((32,), 128, dtype=torch.int32).cuda(),
          "mel": torch.randn(32, 200, 400, generator=g).half().float().cuda(),
          "stft": torch.randn(32, 200, 5125, generator=g).half().float().cuda()}
    for _ in range(2):
        m.train_step(gi, lr=1e-4)
    torch.cuda.synchronize()
    lines = []
    # ---- forward sections ----
    m._marks = []
    S = {}
    with ops.saving(S):
        y, out = m.inference(gi, True)
    ev_l0 = torch.cuda.Event(enable_timing=True); ev_l0.record()
    m.add_loss_op(y, out, gi["mel"], gi["stft"])
    ev_l1 = torch.cuda.Event(enable_timing=True); ev_l1.record()
    S.update(text=gi["text"], text_length=gi["text_length"], mel=gi["mel"], stft=gi["stft"])
    S["post/out"] = out
    marks, m._marks = m._marks, None
    # ---- backward with the timing proxy ----
    TK = TimedK()
    m.add_train_op()
    m._opt.zero_grad()
    P, G = m.store, m._gviews
    B, T, OUTW = y.shape
    ev = lambda: torch.cuda.Event(enable_timing=True)
    b0 = ev(); b0.record()
    TK.section = "post dense+loss"
    F = cfg.fft_size
    post = S["post/cbhg/gru_out"].reshape(-1, 256)
    out2 = out.reshape(-1, F)
    dOut = K.padded_rows(out2.shape[0], F, y)
    TK.l1_bwd(dOut, out2, S["stft"].reshape(-1, F))
    dPost = K.empty(post.shape, y)
    grad.dense_bwd(TK, dOut, post, P["post/dense/W"], G["post/dense/W"], G["post/dense/b"], dX=dPost)
    TK.section = "post CBHG"
    dPostIn = grad.cbhg_bwd(TK, P, G, S, "post/cbhg", dPost.view(B, T * cfg.r, 256), 8, (128, 256, 80))
    dY = dPostIn.view(B, T, OUTW)
    TK.l1_bwd(dY.reshape(-1, OUTW), y.reshape(-1, OUTW), S["mel"].reshape(-1, OUTW), beta=1.0)
    TK.section = "decoder"
    dEnc = grad.decoder_bwd(TK, P, G, S, cfg, dY)
    TK.section = "enc CBHG"
    dPre = grad.cbhg_bwd(TK, P, G, S, "enc/cbhg", dEnc, 16, (128, 128, 128))
    TK.section = "enc prenet"
    grad.enc_prenet_bwd(TK, P, G, S, dPre, 2.0)
    b1 = ev(); b1.record()
    m._opt.apply(K, m.store.flat, 1e-4, cfg.cap_grads)
    b2 = ev(); b2.record()
    torch.cuda.synchronize()
    # ---- report ----
    prev = None
    for name, e in marks:
        if prev is not None:
            lines.append(f"forward {name:10s} {prev.elapsed_time(e):8.3f} ms")
        prev = e
    lines.append(f"loss               {ev_l0.elapsed_time(ev_l1):8.3f} ms")
    lines.append(f"backward total     {b0.elapsed_time(b1):8.3f} ms   ({len(TK.rec)} kernel-namespace calls)")
    lines.append(f"sumsq+adam         {b1.elapsed_time(b2):8.3f} ms")
    sec = collections.OrderedDict()
    kind = collections.Counter()
    kcount = collections.Counter()
    calls = []
    for s, label, e0, e1 in TK.rec:
        ms = e0.elapsed_time(e1)
        sec[s] = sec.get(s, 0.0) + ms
        k = label.split(" ")[0]
        kind[k] += ms; kcount[k] += 1
        calls.append((ms, s, label))
    lines.append("-- backward by section (sum of per-call event times) --")
    for s, ms in sec.items():
        lines.append(f"  {s:18s} {ms:8.3f} ms")
    lines.append("-- backward by primitive --")
    for k, ms in kind.most_common():
        lines.append(f"  {k:18s} {ms:8.3f} ms  in {kcount[k]} calls")
    lines.append("-- 25 slowest calls --")
    for ms, s, label in sorted(calls, reverse=True)[:25]:
        lines.append(f"  {ms:8.3f} ms  [{s}] {label}")
    txt = "\n".join(lines)
    print(txt)
    open(os.path.join(OUT, "train_sections.txt"), "w").write(txt + "\n")


if __name__ == "__main__":
    main()
