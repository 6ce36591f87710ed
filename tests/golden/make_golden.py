"""Generates the committed golden fixtures.  Run IN THE BUILD CONTAINER only (needs /root/reference):

    python tests/golden/make_golden.py

1. reshape_frames.npz -- outputs of the REFERENCE's own audio.reshape_frames (audio.py:23-35),
   obtained by importing /root/reference/audio.py with its unavailable top-level imports
   (librosa, tensorflow, tqdm) stubbed out; the function itself is pure numpy and runs unmodified.
   This is the only function on/around the hot path the reference's code base can execute here.
2. oracle_small_r{2,5}.npz -- end-to-end outputs of the CPU oracle (oracle/tacotron_oracle.py) on
   seeded weights/inputs at B=2, Tx=12, T=6.  These pin the oracle against regressions and give the
   GPU tests a fixture that does not depend on re-running the oracle; they are NOT reference
   outputs (TensorFlow 1.2 cannot run here -- parity unpinned, see oracle/tf12.py).
3. reference_wiring_r{2,5}.npz -- the reference's own model code executed over oracle/tf12_shim.py (see reference_wiring).
4. reference_griffinlim.npz -- the reference's own audio.invert_spectrogram executed with librosa's stft/istft stood in
   by the restatements of oracle/audio_oracle.py (see reference_griffinlim).
5. reference_data_input.npz -- the reference's own data_input.load_from_npy executed on a tiny synthetic data set.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def reference_reshape_frames():
    for name in ("librosa", "tensorflow", "tqdm"):
        m = types.ModuleType(name)
        if name == "tqdm":
            m.tqdm = lambda x, **k: x
        sys.modules[name] = m
    sys.path.insert(0, "/root/reference")
    import audio as ref_audio                      # the reference module, unmodified
    out = {}
    rng = np.random.RandomState(0)
    for r in (2, 5):
        ref_audio.r = r                            # module global the function reads (audio.py:17)
        x = rng.randn(7, 361).astype(np.float32)   # [F, n_frames]; 361 = 108000/300 + 1 as in the reference
        fwd = ref_audio.reshape_frames(x)
        inv = ref_audio.reshape_frames(fwd, forward=False)
        out[f"x_r{r}"] = x
        out[f"fwd_r{r}"] = fwd
        out[f"inv_r{r}"] = inv
    # the reference's own self-test (audio.py:106-115), r = 2
    ref_audio.r = 2
    test = np.repeat(np.arange(40)[:, None] + 1, 7, axis=1)
    o = ref_audio.reshape_frames(test.T)
    inv = ref_audio.reshape_frames(o, forward=False)
    assert np.array_equal(test, inv)
    out["selftest_in"] = test
    out["selftest_fwd"] = o
    np.savez_compressed(os.path.join(HERE, "reshape_frames.npz"), **out)
    print("reshape_frames.npz", {k: v.shape for k, v in out.items()})


def oracle_small():
    from oracle import tacotron_oracle as O
    for r in (2, 5):
        cfg = O.OracleConfig(r=r, max_decode_iter=6, vocab_size=20)
        p = O.init_params(cfg, seed=1, trained_like=True)
        inp = O.synthetic_inputs(cfg, 2, 12, 6, seed=0, ragged=True)
        enc_m, dec_m = O.dropout_masks(cfg, 2, 12, 6, seed=2)
        sm = O.sched_mask(cfg, 2, 6, seed=3)
        y_i, o_i, a_i = O.inference(p, inp, cfg, train=False)
        y_t, o_t, a_t = O.inference(p, inp, cfg, train=True, enc_drop_masks=enc_m, dec_drop_masks=dec_m)
        y_s, o_s, a_s = O.inference(p, inp, cfg, train=True, enc_drop_masks=enc_m, dec_drop_masks=dec_m, sample_mask=sm)
        loss, ls, lo = O.loss(y_t, o_t, inp["mel"], inp["stft"])
        np.savez_compressed(
            os.path.join(HERE, f"oracle_small_r{r}.npz"),
            y_infer=y_i.numpy(), out_infer=o_i.numpy(), align_infer=a_i.numpy(),
            y_teacher=y_t.numpy(), out_teacher=o_t.numpy(), align_teacher=a_t.numpy(),
            y_sched=y_s.numpy(), out_sched=o_s.numpy(), align_sched=a_s.numpy(),
            loss_teacher=np.array([float(loss), float(ls), float(lo)]))
        print(f"oracle_small_r{r}.npz written")


def reference_wiring():
    """Execute the reference's OWN models/tacotron.py + models/ops.py (unmodified, imported from /root/reference)
    over oracle/tf12_shim.py and save its outputs: pins the oracle's graph wiring against the reference source."""
    from oracle import tacotron_oracle as O
    from oracle import tf12_shim as shim
    from tacotron_b200.tf_names import tf_name_to_param as tf_name_to_oracle
    for name in ("librosa", "tqdm"):
        m = types.ModuleType(name)
        if name == "tqdm":
            m.tqdm = lambda x, **k: x
        sys.modules[name] = m
    shim.install()
    for k in [k for k in sys.modules if k == "audio" or k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    import models.tacotron as ref_tacotron            # the reference module, unmodified

    for r in (2, 5):
        B, Tx, T = 2, 12, 6
        cfg_o = O.OracleConfig(r=r, max_decode_iter=T, vocab_size=20)
        params = O.init_params(cfg_o, seed=1, trained_like=True)
        inp = O.synthetic_inputs(cfg_o, B, Tx, T, seed=0, ragged=True)
        enc_m, dec_m = O.dropout_masks(cfg_o, B, Tx, T, seed=2)
        sm = O.sched_mask(cfg_o, B, T, seed=3)
        used = set()

        def lookup(name, shape):
            on = tf_name_to_oracle(name)
            used.add(on)
            return params[on]

        out = {}
        for mode in ("infer", "teacher", "sched"):
            config = ref_tacotron.Config()
            config.r, config.vocab_size, config.max_decode_iter = r, 20, T
            config.scheduled_sample = 0.5 if mode == "sched" else 0
            train = mode != "infer"
            drops = []
            if train:                                  # call order: encoder pre-net (2), then 2 per decoder step
                drops = [enc_m[0], enc_m[1]]
                for t in range(T):
                    drops += [dec_m[0][t], dec_m[1][t]]
            shim.S.reset(lookup, dropout_masks=drops, sample_masks=[sm[t] for t in range(T)] if mode == "sched" else None)
            model = ref_tacotron.Tacotron(config, inp, train=train)
            out[f"y_{mode}"] = model.seq2seq_output.numpy()
            out[f"out_{mode}"] = model.output.numpy()
            out[f"align_{mode}"] = model.alignments.numpy()
            if train:
                out[f"loss_{mode}"] = np.array(float(model.loss))
            assert not shim.S.dropout_masks, "dropout masks left over: the reference made fewer dropout calls than assumed"
        assert used == set(params), sorted(set(params) - used)
        out["tf_variable_names"] = np.array(shim.S.created)
        np.savez_compressed(os.path.join(HERE, f"reference_wiring_r{r}.npz"), **out)
        print(f"reference_wiring_r{r}.npz written ({len(shim.S.created)} variables)")


def reference_griffinlim():
    """Execute the reference's OWN audio.invert_spectrogram / audio.griffinlim (audio.py:67-97, unmodified, imported from
    /root/reference) with `librosa.stft` / `librosa.istft` stood in by the restatements of oracle/audio_oracle.py (librosa
    is not installed) and save input + output: pins the WIRING of the inversion (reshape_frames inverse, exp, the
    50-iteration loop, the angle update, the final istft) against the reference source.  The random initial phase
    (audio.py:81) is reproduced from the numpy seed stored in the fixture."""
    from oracle import audio_oracle as A
    for k in [k for k in sys.modules if k in ("audio", "librosa", "tensorflow", "tqdm")]:
        del sys.modules[k]
    lib = types.ModuleType("librosa")
    lib.stft = lambda y, n_fft=2048, hop_length=None, win_length=None, window="hann": A.stft(y, n_fft, hop_length, win_length)
    lib.istft = lambda D, hop_length=None, win_length=None, window="hann": A.istft(D, hop_length, win_length)
    sys.modules["librosa"] = lib
    sys.modules["tensorflow"] = types.ModuleType("tensorflow")
    tq = types.ModuleType("tqdm")
    tq.tqdm = lambda x, **k: x
    sys.modules["tqdm"] = tq
    if not hasattr(np, "complex"):
        np.complex = complex                       # alias removed from numpy >= 1.24; the reference (2017) uses it (audio.py:84,93)
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    import audio as ref_audio                      # the reference module, unmodified
    out = {}
    for r, T in ((2, 8), (5, 4)):
        ref_audio.r = r                            # module global read by reshape_frames (audio.py:17)
        rng = np.random.RandomState(10 + r)
        spec = (rng.randn(T, 1025 * r) * 0.5).astype(np.float32)
        seed = 100 + r
        np.random.seed(seed)
        wave = ref_audio.invert_spectrogram(spec)
        out[f"spec_r{r}"] = spec
        out[f"seed_r{r}"] = np.array(seed)
        out[f"wave_r{r}"] = np.asarray(wave)
    np.savez_compressed(os.path.join(HERE, "reference_griffinlim.npz"), **out)
    print("reference_griffinlim.npz", {k: v.shape for k, v in out.items()})


def reference_data_input():
    """Execute the reference's OWN data_input.load_from_npy / load_prompts arithmetic (data_input.py:42-85, unmodified,
    imported from /root/reference with tensorflow / matplotlib stubbed) on a tiny synthetic data set and save inputs +
    outputs: pins the float16 in-place normalisation, the 100-utterance statistics (numpy's seeded stream) and the
    speech_length override against the reference source."""
    import tempfile
    for k in [k for k in sys.modules if k in ("data_input", "tensorflow", "matplotlib", "matplotlib.pyplot")]:
        del sys.modules[k]
    sys.modules["tensorflow"] = types.ModuleType("tensorflow")
    mpl = types.ModuleType("matplotlib"); mpl.use = lambda *a, **k: None
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    import data_input as ref_di                    # the reference module, unmodified
    rng = np.random.RandomState(0)
    N, Tx, T, r = 12, 8, 4, 2
    raw = {"texts": rng.randint(1, 20, size=(N, Tx)).astype(np.int64), "text_lens": rng.randint(3, Tx + 1, size=N).astype(np.int64),
           "stfts": (rng.randn(N, T, 1025 * r) * 2 - 3).astype(np.float16), "mels": (rng.randn(N, T, 80 * r) * 2 - 3).astype(np.float16),
           "speech_lens": rng.randint(2, T + 1, size=N).astype(np.int64)}
    with tempfile.TemporaryDirectory() as d:
        for k, v in raw.items():
            np.save(os.path.join(d, k + ".npy"), v)
        seed = 7
        np.random.seed(seed)
        inputs, names, num_speakers, stft_mean, stft_std = ref_di.load_from_npy(d + "/")
    out = {"raw_" + k: v for k, v in raw.items()}
    out["seed"] = np.array(seed)
    for n, a in zip(names, inputs):
        out["ref_" + n] = np.asarray(a)
    out["ref_stft_mean"], out["ref_stft_std"] = np.asarray(stft_mean), np.asarray(stft_std)
    out["num_speakers"] = np.array(num_speakers)
    np.savez_compressed(os.path.join(HERE, "reference_data_input.npz"), **out)
    print("reference_data_input.npz", names, {k: (v.shape, str(v.dtype)) for k, v in out.items() if k.startswith("ref_")})


if __name__ == "__main__":
    reference_data_input()
    reference_reshape_frames()
    oracle_small()
    reference_wiring()
    reference_griffinlim()
