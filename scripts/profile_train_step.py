"""One or more C2 training steps with nothing else in the process -- the target command for ncu.
    python scripts/profile_train_step.py [steps] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tacotron_b200 import Config, Tacotron

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
precision = sys.argv[2] if len(sys.argv) > 2 else "fp32x3"
B, TX, T, R = 32, 128, 200, 5
m = Tacotron(Config(r=R, vocab_size=64, precision=precision), None, train=True, seed=1)
g = torch.Generator().manual_seed(100)
gi = {"text": torch.randint(1, 64, (B, TX), generator=g, dtype=torch.int32).cuda(),
      "text_length": torch.full((B,), TX, dtype=torch.int32).cuda(),
      "mel": torch.randn(B, T, 80 * R, generator=g).half().float().cuda(),
      "stft": torch.randn(B, T, 1025 * R, generator=g).half().float().cuda()}
m.dp = False
for _ in range(steps):
    m.train_step(gi, lr=1e-4)
torch.cuda.synchronize()
print("done")
