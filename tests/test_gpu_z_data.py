"""GPU tests of the input-side rows: the float16 normalisation kernel bit-exact against the numpy
restatement of the reference's in-place statements, the double-buffered device batch iterator, and the two drivers
(tacotron_b200/train.py, test.py) end to end on a tiny synthetic data set: train a few steps, checkpoint, resume,
synthesise from prompts, invert with the GPU Griffin-Lim.

STATUS (round 1): the normalisation kernel and the device batch iterator are green on B200; the driver test is the one
piece without a hardware run (non-strict xfail: XPASS in the round-end run = it works).
"""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import data_oracle as DO

pytestmark = pytest.mark.gpu


def _dataset(root, N=40, Tx=12, T=8, r=2, seed=0):
    rng = np.random.RandomState(seed)
    d = os.path.join(root, "data", "toy") + "/"
    os.makedirs(d, exist_ok=True)
    np.save(d + "texts.npy", rng.randint(1, 20, size=(N, Tx)))
    np.save(d + "text_lens.npy", rng.randint(4, Tx + 1, size=N))
    np.save(d + "stfts.npy", (rng.randn(N, T, 1025 * r) * 2 - 3).astype(np.float16))
    np.save(d + "mels.npy", (rng.randn(N, T, 80 * r) * 2 - 3).astype(np.float16))
    np.save(d + "speech_lens.npy", rng.randint(3, T + 1, size=N))
    with open(d + "meta.pkl", "wb") as f:
        pickle.dump({"r": r, "vocab": {i: chr(96 + i) for i in range(20)}}, f)
    return d


def test_normalize_f16_bit_exact():
    from tacotron_b200 import kernels as K
    rng = np.random.RandomState(1)
    for shape in ((6, 5, 2050), (3, 7, 5125), (1, 1, 161)):                       # even and odd widths / totals
        x = (rng.randn(*shape) * 3 - 2).astype(np.float16)
        x.reshape(-1)[:4] = [65504, -65504, 6e-8, 0]
        mean, std = DO.sample_stats(x, rng.randint(len(x), size=100))
        ref = DO.normalize_explicit(x, mean, std)
        out = torch.empty(shape, dtype=torch.float32, device="cuda")
        K.normalize_f16(out, torch.from_numpy(x).cuda(), torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda())
        assert np.array_equal(out.cpu().numpy(), ref, equal_nan=True)


def test_device_batches_bit_exact(tmp_path):
    from tacotron_b200 import data_input
    d = _dataset(str(tmp_path))
    arrays, names, _, stft_mean, stft_std = data_input.load_from_npy(d, rng=np.random.RandomState(3))
    it = data_input.build_dataset(arrays, names, batch_size=4, buffer_size=16, seed=5)
    for _ in range(12):
        b = next(it)
        torch.cuda.synchronize()
        assert b["stft"].is_cuda and b["stft"].dtype == torch.float32
        text = b["text"].cpu().numpy()
        for k in range(4):
            j = int(np.where((arrays["text"] == text[k]).all(1))[0][0])
            assert np.array_equal(b["stft"][k].cpu().numpy(), DO.normalize_explicit(np.asarray(arrays["stft"][j]), stft_mean, stft_std))


def test_drivers_train_checkpoint_resume_synthesise(tmp_path, monkeypatch):
    from tacotron_b200 import checkpoint, test as synth, train as trainer
    from tacotron_b200.models.tacotron import Config, Tacotron
    _dataset(str(tmp_path))
    monkeypatch.chdir(tmp_path)
    cfg = Config()
    cfg.data_path = "data/toy/"
    cfg.save_path = "toy/tacotron"
    cfg.restore = False
    logs = []
    m = trainer.train(Tacotron, cfg, num_steps=4, log=logs.append, save_every=2)
    assert m.global_step == 4 and os.path.exists("weights/toy/tacotron-4.npz") and os.path.exists("weights/toy/tacotron-4_sample.wav")
    # resume: parameters, Adam slots and the step counter come back
    cfg2 = Config(); cfg2.data_path, cfg2.save_path, cfg2.restore = cfg.data_path, cfg.save_path, True
    m2 = trainer.train(Tacotron, cfg2, num_steps=1, log=logs.append, save_every=1000)
    assert m2.global_step == 5
    # synthesis from prompts with the restored weights
    cfg3 = Config(); cfg3.data_path, cfg3.save_path = cfg.data_path, cfg.save_path
    waves = synth.test(Tacotron, cfg3, ["abc de\n", "hello\n"], log=logs.append)
    assert len(waves) == 2 and all(torch.isfinite(w).all() for w in waves)
    assert os.path.exists("log/toy/tacotron/test/1.wav")
