"""One or more C2 inference steps with nothing else in the process -- the target command for ncu.
    python scripts/profile_step.py [steps] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tacotron_b200 import Config, Tacotron

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
precision = sys.argv[2] if len(sys.argv) > 2 else "fp32x3"
B, TX, T, R = 32, 128, 200, 5
m = Tacotron(Config(r=R, vocab_size=64, max_decode_iter=T, precision=precision), None, train=False, seed=1)
g = torch.Generator().manual_seed(0)
inp = {"text": torch.randint(1, 64, (B, TX), generator=g, dtype=torch.int32).cuda(),
       "text_length": torch.full((B,), TX, dtype=torch.int32).cuda()}
for _ in range(steps):
    m.inference(inp, train=False)
torch.cuda.synchronize()
print("done")
