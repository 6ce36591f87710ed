"""CPU tests of the spectrogram-inversion row (reference audio.py:23-35, :67-97):
  * the numpy oracle's stft/istft restatement of librosa against torch.stft/istft (an independent implementation of the
    same published conventions);
  * tacotron_b200.audio.reshape_frames (index permutation) against the restatement pinned on the reference's own outputs;
  * the host orchestration of Griffin-Lim (tacotron_b200/audio.py) over the torch-CPU mirror kernels against the oracle."""
import numpy as np
import pytest
import torch

from oracle import audio_oracle as A
from oracle import tf12
from tacotron_b200 import audio
from tests import mirror_kernels as MK


def test_oracle_stft_istft_match_torch():
    rng = np.random.default_rng(0)
    y = rng.standard_normal(300 * 23)
    w = torch.hann_window(1200, periodic=True, dtype=torch.float64)
    D = A.stft(y)
    Dt = torch.stft(torch.from_numpy(y), 2048, 300, 1200, window=w, center=True, pad_mode="reflect", return_complex=True).numpy()
    assert D.shape == Dt.shape == (1025, 24)
    assert np.abs(D - Dt).max() < 1e-10
    R = rng.standard_normal((1025, 24)) + 1j * rng.standard_normal((1025, 24))       # not a consistent STFT
    yt = torch.istft(torch.from_numpy(R), 2048, 300, 1200, window=w, center=True).numpy()
    yo = A.istft(R)
    assert yo.shape == yt.shape == (300 * 23,)
    assert np.abs(yo - yt).max() < 1e-12
    assert np.abs(A.istft(D) - y).max() < 1e-10                                          # perfect reconstruction (NOLA holds)


@pytest.mark.parametrize("r", [1, 2, 5])
def test_reshape_frames_permutation(r):
    rng = np.random.default_rng(r)
    Fd, n = 7, 4 * r * 3 + 5                                                             # trailing partial block is dropped
    sig = rng.standard_normal((Fd, n))
    fwd_ref = tf12.reshape_frames(sig, r, forward=True)
    fwd = audio.reshape_frames(torch.from_numpy(sig), r, forward=True).numpy()
    assert np.array_equal(fwd, fwd_ref)
    inv_ref = tf12.reshape_frames(fwd_ref, r, forward=False)
    inv = audio.reshape_frames(torch.from_numpy(fwd_ref), r, forward=False).numpy()
    assert np.array_equal(inv, inv_ref)
    assert np.array_equal(inv.T, sig[:, :4 * r * (n // (4 * r))])                        # audio.py:106-115 round trip


@pytest.mark.parametrize("r,T,n_iter", [(2, 8, 3), (5, 4, 2), (1, 12, 4)])
def test_griffinlim_host_logic_matches_oracle(r, T, n_iter):
    g = torch.Generator().manual_seed(7)
    B = 2
    spec = (torch.randn(B, T, 1025 * r, generator=g, dtype=torch.float64) * 0.5)
    mean = torch.randn(1025 * r, generator=g, dtype=torch.float64) * 0.1
    std = torch.rand(1025 * r, generator=g, dtype=torch.float64) + 0.5
    n = 4 * r * (T // 4)
    pu = torch.rand(B, n, 1025, generator=g, dtype=torch.float64)
    y = audio.invert_spectrogram(spec, r, n_iter=n_iter, stft_mean=mean, stft_std=std, phase_u=pu, K=MK)
    assert y.shape == (B, 300 * (n - 1))
    for b in range(B):
        ang0 = np.exp(2j * np.pi * pu[b].numpy().T)                                      # [1025, n]
        ref = A.invert_spectrogram((spec[b] * std + mean).numpy(), r, ang0, n_iter=n_iter)
        assert np.abs(y[b].numpy() - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("r", [2, 5])
def test_inversion_wiring_matches_the_reference_code(r):
    """tests/golden/reference_griffinlim.npz holds outputs of the reference's OWN audio.invert_spectrogram (executed from
    /root/reference with librosa's stft/istft stood in by the restatements, tests/golden/make_golden.py): the oracle and
    the device orchestration (over the CPU mirror kernels) must reproduce them -- reshape_frames inverse, exp, the random
    initial phase drawn from numpy's seeded stream, 50 iterations, final istft."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_griffinlim.npz"))
    spec, seed, wave = z[f"spec_r{r}"], int(z[f"seed_r{r}"]), z[f"wave_r{r}"]
    n = 4 * r * (spec.shape[0] // 4)
    u = np.random.RandomState(seed).rand(1025, n)                                       # audio.py:81 with np.random.seed(seed)
    ref = A.invert_spectrogram(spec, r, np.exp(2j * np.pi * u), n_iter=50)
    assert ref.shape == wave.shape and np.abs(ref - wave).max() < 1e-9 * np.abs(wave).max()
    y = audio.invert_spectrogram(torch.from_numpy(spec.astype(np.float64)), r, n_iter=50,
                                 phase_u=torch.from_numpy(np.ascontiguousarray(u.T)), K=MK)
    # (the reference exponentiates its float32 input in float32, audio.py:70; the orchestration here runs in float64 end to
    #  end, so the two differ by float32 rounding of exp(), carried through 50 iterations: 3e-6 relative)
    assert np.abs(y.numpy() - wave).max() < 2e-5 * np.abs(wave).max()


def test_griffinlim_single_spectrogram_api():
    g = torch.Generator().manual_seed(3)
    mag = torch.rand(1025, 8, generator=g, dtype=torch.float64) + 0.1
    pu = torch.rand(8, 1025, generator=g, dtype=torch.float64)
    y = audio.griffinlim(mag, n_iter=2, phase_u=pu, K=MK)
    ref = A.griffinlim(mag.numpy(), np.exp(2j * np.pi * pu.numpy().T), n_iter=2)
    assert np.abs(y.numpy() - ref).max() < 1e-9
