"""BASELINE config 5 (model part): B=1, prompt padded to 140 chars, 500 mel frames (T=100, r=5), inference;
p50 end-to-end latency through the public API with host buffers (H2D text, D2H output + alignments).
Griffin-Lim (audio.py:77-97) is a 'next' row and is not included."""
import os, sys, time, statistics, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tacotron_b200 import Config, Tacotron
B, TX, T, R = 1, 140, 100, 5
m = Tacotron(Config(r=R, vocab_size=64, max_decode_iter=T, precision="tf32", cuda_graph=True), None, train=False, seed=1)
g = torch.Generator().manual_seed(0)
text_h = torch.randint(1, 64, (B, TX), generator=g, dtype=torch.int32).pin_memory()
len_h = torch.full((B,), 97, dtype=torch.int32).pin_memory()
out_h = torch.empty((B, T, 1025 * R)).pin_memory(); al_h = torch.empty((B, T, TX)).pin_memory()
def once():
    ci = {"text": text_h.cuda(non_blocking=True), "text_length": len_h.cuda(non_blocking=True)}
    y, out = m.inference(ci, train=False)
    out_h.copy_(out, non_blocking=True); al_h.copy_(m.alignments, non_blocking=True)
    torch.cuda.synchronize()
for _ in range(5): once()
ts = []
for _ in range(50):
    t0 = time.perf_counter(); once(); ts.append((time.perf_counter() - t0) * 1e3)
print(json.dumps({"config": "C5 model part: B=1, char 140, 500 frames (T=100, r=5), inference, host in/out", "p50_ms": statistics.median(ts),
                  "p90_ms": sorted(ts)[44], "frames": 500, "frames_per_s": 500 / (statistics.median(ts) / 1e3)}))
