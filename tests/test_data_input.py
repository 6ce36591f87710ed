"""CPU tests of the input-side rows: float16 normalisation arithmetic, the shuffle-buffer batch
stream, prompt encoding, the double-buffered batch iterator (over the CPU mirror kernel), checkpoint round trip and the
TF-variable importer."""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import data_oracle as DO
from tacotron_b200 import data_input
from tests import mirror_kernels as MK


def _dataset(tmp_path, N=40, T=8, r=2, seed=0):
    rng = np.random.RandomState(seed)
    d = str(tmp_path) + "/"
    texts = rng.randint(1, 20, size=(N, 12)).astype(np.int64)
    np.save(d + "texts.npy", texts)
    np.save(d + "text_lens.npy", rng.randint(4, 13, size=N).astype(np.int64))
    np.save(d + "stfts.npy", (rng.randn(N, T, 1025 * r) * 2 - 3).astype(np.float16))
    np.save(d + "mels.npy", (rng.randn(N, T, 80 * r) * 2 - 3).astype(np.float16))
    np.save(d + "speech_lens.npy", rng.randint(3, T + 1, size=N))
    with open(d + "meta.pkl", "wb") as f:
        pickle.dump({"r": r, "vocab": {i: chr(96 + i) for i in range(1, 20)}}, f)
    return d


def test_float16_normalisation_restatement_is_numpy_inplace():
    """the per-element spelling the device kernel implements == the reference's two in-place numpy statements"""
    rng = np.random.RandomState(1)
    x = (rng.randn(6, 5, 64) * 3 - 2).astype(np.float16)
    x[0, 0, :4] = [65504, -65504, 6e-8, 0]                       # extremes: max half, subnormal, zero
    index = rng.randint(len(x), size=100)
    mean, std = DO.sample_stats(x, index)
    assert mean.dtype == np.float16 and std.dtype == np.float32
    ref = DO.normalize_inplace(x.copy(), mean, std)
    assert ref.dtype == np.float16
    got = DO.normalize_explicit(x, mean, std)
    assert got.dtype == np.float32 and np.array_equal(got, ref.astype(np.float32), equal_nan=True)
    out = torch.empty(x.shape, dtype=torch.float32)
    MK.normalize_f16(out, torch.from_numpy(x), torch.from_numpy(mean), torch.from_numpy(std))
    assert np.array_equal(out.numpy(), got, equal_nan=True)       # the kernel's mirror is bit-exact too


def test_load_from_npy_and_batches(tmp_path):
    d = _dataset(tmp_path)
    assert data_input.load_meta(d)["r"] == 2
    arrays, names, nspk, stft_mean, stft_std = data_input.load_from_npy(d, rng=np.random.RandomState(3))
    assert names == ["text", "text_length", "stft", "mel", "speech_length"] and nspk == 1
    assert arrays["stft"].dtype == np.float16 and stft_mean.dtype == np.float16 and stft_std.dtype == np.float32
    assert (arrays["speech_length"] == 8).all()                  # data_input.py:71-72: padded length for every utterance
    idx = np.random.RandomState(3).randint(40, size=100)
    m2, s2 = DO.sample_stats(np.load(d + "stfts.npy"), idx)
    assert np.array_equal(m2, stft_mean) and np.array_equal(s2, stft_std)
    it = data_input.build_dataset(arrays, names, batch_size=4, buffer_size=16, seed=5, device="cpu", K=MK)
    seen = []
    for _ in range(30):
        b = next(it)
        assert b["stft"].dtype == torch.float32 and b["stft"].shape == (4, 8, 2050) and b["text"].dtype == torch.int32
        # locate the utterances by their text rows and check the normalised spectrograms bit for bit
        for k in range(4):
            j = int(np.where((arrays["text"] == b["text"][k].numpy()).all(1))[0][0])
            ref = DO.normalize_explicit(np.asarray(arrays["stft"][j]), stft_mean, stft_std)
            assert np.array_equal(b["stft"][k].numpy(), ref)
            refm = DO.normalize_explicit(np.asarray(arrays["mel"][j]), arrays["_stats"]["mel_mean"], arrays["_stats"]["mel_std"])
            assert np.array_equal(b["mel"][k].numpy(), refm)
            seen.append(j)
    assert len(set(seen)) == 40                                  # 120 draws through a 16-slot buffer visit everything


def test_loader_and_normalisation_match_the_reference_code(tmp_path):
    """tests/golden/reference_data_input.npz holds what the reference's OWN data_input.load_from_npy (executed from
    /root/reference, tests/golden/make_golden.py) returns for a tiny data set: its in-place float16 normalisation, the
    statistics of its 100-utterance sample (numpy's seeded stream), dtypes and the speech_length override.  The loader
    here keeps the spectrograms as stored; its statistics and the per-batch device normalisation (kernel mirror) must
    give exactly the reference's values."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_data_input.npz"))
    d = str(tmp_path) + "/"
    for k in ("texts", "text_lens", "stfts", "mels", "speech_lens"):
        np.save(d + k + ".npy", z["raw_" + k])
    arrays, names, nspk, stft_mean, stft_std = data_input.load_from_npy(d, rng=np.random.RandomState(int(z["seed"])))
    assert nspk == int(z["num_speakers"]) and names == ["text", "text_length", "stft", "mel", "speech_length"]
    assert stft_mean.dtype == z["ref_stft_mean"].dtype and np.array_equal(stft_mean, z["ref_stft_mean"])
    assert stft_std.dtype == z["ref_stft_std"].dtype and np.array_equal(stft_std, z["ref_stft_std"])
    for k in ("text", "text_length", "speech_length"):
        assert arrays[k].dtype == z["ref_" + k].dtype and np.array_equal(arrays[k], z["ref_" + k])
    for k in ("stft", "mel"):
        st = arrays["_stats"]
        out = torch.empty(arrays[k].shape, dtype=torch.float32)
        MK.normalize_f16(out, torch.from_numpy(np.asarray(arrays[k])), torch.from_numpy(st[f"{k}_mean"]), torch.from_numpy(st[f"{k}_std"]))
        assert np.array_equal(out.numpy(), z["ref_" + k].astype(np.float32))          # == tf.cast(normalised float16, float32)
        assert np.array_equal(DO.normalize_explicit(np.asarray(arrays[k]), st[f"{k}_mean"], st[f"{k}_std"]), z["ref_" + k].astype(np.float32))


def test_shuffle_buffer_semantics():
    n, buf = 10, 4
    g = data_input.shuffled_indices(n, buf, seed=1)
    out = [next(g) for _ in range(200)]
    assert set(out) == set(range(n))
    # element i of the repeated stream cannot be emitted before i - (buffer-1) outputs have happened ... i.e. position >= i - buf + 1
    pos = {}
    count = {k: 0 for k in range(n)}
    for p, v in enumerate(out):
        stream_index = v + n * count[v]
        count[v] += 1
        assert p >= stream_index - (buf - 1)
    # two ranks draw disjoint slices of the same stream
    a = data_input.DeviceBatches.__new__(data_input.DeviceBatches)
    b = data_input.DeviceBatches.__new__(data_input.DeviceBatches)
    for obj, rank in ((a, 0), (b, 1)):
        obj.B, obj.rank, obj.world = 3, rank, 2
        obj.idx = data_input.shuffled_indices(50, 8, seed=2)
    ga = data_input.shuffled_indices(50, 8, seed=2)
    full = [next(ga) for _ in range(12)]
    assert sorted(a._next_indices().tolist()) == sorted(full[0:3]) and sorted(b._next_indices().tolist()) == sorted(full[3:6])
    assert sorted(a._next_indices().tolist()) == sorted(full[6:9]) and sorted(b._next_indices().tolist()) == sorted(full[9:12])


def test_prompt_encoding_quirks():
    ivocab = {0: "<pad>", 1: "a", 2: "b", 3: " "}
    prompts = ["ab a!\n", "b\n"]
    ref_text, ref_len = DO.encode_prompts(prompts, ivocab)
    batches = list(data_input.load_prompts(prompts, ivocab, device="cpu"))
    assert len(batches) == 1
    t, l = batches[0]["text"].numpy(), batches[0]["text_length"].numpy()
    assert t.shape == (2, 140) and np.array_equal(t, ref_text) and np.array_equal(l, ref_len)
    assert l.tolist() == [6, 2]                                   # RAW line lengths (newline and '!' counted), data_input.py:96
    assert t[0, :4].tolist() == [1, 2, 3, 1] and (t[0, 4:] == 0).all()          # '!' dropped from the ids, padded with 0
    many = ["a\n"] * 70
    sizes = [b["text"].shape[0] for b in data_input.load_prompts(many, ivocab, device="cpu")]
    assert sizes == [32, 32, 6]                                   # allow_smaller_final_batch


class _FakeStore:
    def __init__(self, shapes):
        self.shapes = {n: (s, None) for n, s in shapes.items()}
        self.views = {n: torch.zeros(s) for n, s in shapes.items()}


class _FakeModel:
    def __init__(self):
        self.store = _FakeStore({"embedding": (5, 4), "enc/prenet/W1": (4, 3), "post/dense/b": (7,)})
        self.global_step = 0
        self._opt = None

    def load_params(self, p):
        for n in self.store.shapes:
            self.store.views[n].copy_(p[n])

    def add_train_op(self):
        pass


def test_checkpoint_roundtrip_and_rotation(tmp_path):
    from tacotron_b200 import checkpoint
    m = _FakeModel()
    prefix = str(tmp_path / "weights" / "nancy" / "tacotron")
    assert checkpoint.latest_checkpoint(prefix) is None
    for step in (5000, 10000, 15000, 20000):
        for v in m.store.views.values():
            v.fill_(float(step))
        m.global_step = step
        checkpoint.save(m, prefix, stft_mean=np.arange(3, dtype=np.float16), stft_std=np.ones(3, dtype=np.float32))
    files = sorted(os.listdir(tmp_path / "weights" / "nancy"))
    assert files == ["tacotron-10000.npz", "tacotron-15000.npz", "tacotron-20000.npz"]        # max_to_keep=3, train.py:47
    assert checkpoint.latest_checkpoint(prefix).endswith("tacotron-20000.npz")
    m2 = _FakeModel()
    mean, std = checkpoint.restore(m2, checkpoint.latest_checkpoint(prefix))
    assert m2.global_step == 20000 and float(m2.store.views["embedding"][0, 0]) == 20000.0
    assert mean.dtype == np.float16 and np.array_equal(mean, np.arange(3, dtype=np.float16))


def test_import_tf_variables_maps_names():
    from tacotron_b200 import checkpoint
    m = _FakeModel()
    tf_vars = {"embedding/embedding:0": np.full((5, 4), 2.0), "encoder/pre_net/dense/kernel": np.full((4, 3), 3.0),
               "post-process/dense/bias": np.full((7,), 4.0), "global_step": np.int64(12), "beta1_power": np.float32(0.5)}
    extra = checkpoint.import_tf_variables(m, tf_vars)
    assert float(m.store.views["embedding"][0, 0]) == 2.0 and float(m.store.views["enc/prenet/W1"][0, 0]) == 3.0
    assert float(m.store.views["post/dense/b"][0]) == 4.0 and set(extra) == {"global_step", "beta1_power"}
