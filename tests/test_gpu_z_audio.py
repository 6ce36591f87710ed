"""GPU parity tests of the spectrogram-inversion row (reference audio.py:67-97): the Griffin-Lim
glue kernels (csrc/audio.cu) each against their torch-CPU mirror, and the whole inversion (cuFFT + kernels) against the
numpy oracle.  fp32 on the device vs fp64 oracle: 1e-4 of the signal's max after a few iterations (the phase update
divides by |rebuilt|, which is ill-conditioned where a bin is nearly cancelled, so long runs are compared through the
property the algorithm optimises -- spectral convergence -- instead of sample by sample).

STATUS: green on B200 (profiles/r01_pytest_mma_audio_runxfail.log).
"""
import math

import numpy as np
import pytest
import torch

from oracle import audio_oracle as A
from tests import mirror_kernels as MK

pytestmark = pytest.mark.gpu


def _K():
    from tacotron_b200 import kernels
    return kernels


def _close(got, ref, tol):
    got, ref = got.detach().cpu(), ref.detach().cpu()
    if got.is_complex():
        got, ref = torch.view_as_real(got), torch.view_as_real(ref)
    got, ref = got.double(), ref.double()
    err = (got - ref).abs().max().item()
    assert err <= tol * (ref.abs().max().item() + 1e-6), (err, ref.abs().max().item())


@pytest.mark.parametrize("r,T", [(2, 8), (5, 6)])
def test_gl_kernels_match_mirror(r, T):
    K = _K()
    g = torch.Generator().manual_seed(1)
    B, F = 2, 1025
    n = 4 * r * (T // 4)
    L = 300 * (n - 1)
    spec = torch.randn(B, T, F * r, generator=g) * 0.5
    scale, shift = torch.rand(F * r, generator=g) + 0.5, torch.randn(F * r, generator=g) * 0.1
    pu = torch.rand(B, n, F, generator=g)
    # gl_init (with and without de-normalisation)
    for sc, sh in ((scale, shift), (None, None)):
        mag_c, full_c = torch.empty(B, n, F), torch.empty(B, n, F, dtype=torch.complex64)
        MK.gl_init(full_c, mag_c, spec, pu, r, sc, sh)
        mag_g, full_g = torch.empty(B, n, F, device="cuda"), torch.empty(B, n, F, dtype=torch.complex64, device="cuda")
        K.gl_init(full_g, mag_g, spec.cuda(), pu.cuda(), r, None if sc is None else sc.cuda(), None if sh is None else sh.cuda())
        _close(mag_g, mag_c, 1e-5); _close(full_g, full_c, 1e-5)
    # gl_ola
    fr = torch.randn(B, n, 2048, generator=g)
    y_c = torch.empty(B, L); MK.gl_ola(y_c, fr, 300, 1200)
    y_g = torch.empty(B, L, device="cuda"); K.gl_ola(y_g, fr.cuda(), 300, 1200)
    _close(y_g, y_c, 1e-5)
    # gl_frame
    f_c = torch.empty(B, n, 2048); MK.gl_frame(f_c, y_c, 300, 1200)
    f_g = torch.empty(B, n, 2048, device="cuda"); K.gl_frame(f_g, y_c.cuda(), 300, 1200)
    _close(f_g, f_c, 1e-5)
    # gl_phase (including exact zeros)
    reb = torch.complex(torch.randn(B, n, F, generator=g), torch.randn(B, n, F, generator=g))
    reb[0, 0, :5] = 0
    p_c = torch.empty(B, n, F, dtype=torch.complex64); MK.gl_phase(p_c, mag_c, reb)
    p_g = torch.empty(B, n, F, dtype=torch.complex64, device="cuda"); K.gl_phase(p_g, mag_c.cuda(), reb.cuda())
    _close(p_g, p_c, 1e-5)


@pytest.mark.parametrize("r,T,n_iter", [(2, 8, 3), (5, 8, 2)])
def test_invert_spectrogram_matches_oracle(r, T, n_iter):
    from tacotron_b200 import audio
    g = torch.Generator().manual_seed(7)
    B = 2
    spec = torch.randn(B, T, 1025 * r, generator=g) * 0.5
    mean, std = torch.randn(1025 * r, generator=g) * 0.1, torch.rand(1025 * r, generator=g) + 0.5
    n = 4 * r * (T // 4)
    pu = torch.rand(B, n, 1025, generator=g)
    y = audio.invert_spectrogram(spec.cuda(), r, n_iter=n_iter, stft_mean=mean.cuda(), stft_std=std.cuda(), phase_u=pu.cuda())
    torch.cuda.synchronize()
    for b in range(B):
        ang0 = np.exp(2j * np.pi * pu[b].double().numpy().T)
        ref = A.invert_spectrogram((spec[b] * std + mean).double().numpy(), r, ang0, n_iter=n_iter)
        err = np.abs(y[b].cpu().double().numpy() - ref).max()
        assert err <= 1e-4 * np.abs(ref).max(), (err, np.abs(ref).max())


def test_cuda_graph_replay_equals_eager():
    """GriffinLimGraph replays exactly the eager launch sequence: identical samples for identical inputs, twice"""
    from tacotron_b200 import audio
    g = torch.Generator().manual_seed(5)
    B, T, r = 2, 8, 2
    n = 4 * r * (T // 4)
    mean, std = (torch.randn(1025 * r, generator=g) * 0.1).cuda(), (torch.rand(1025 * r, generator=g) + 0.5).cuda()
    glg = audio.GriffinLimGraph(B, T, r, n_iter=4)
    for seed in (1, 2):
        gg = torch.Generator().manual_seed(seed)
        spec = (torch.randn(B, T, 1025 * r, generator=gg) * 0.5).cuda()
        pu = torch.rand(B, n, 1025, generator=gg).cuda()
        eager = audio.invert_spectrogram(spec, r, n_iter=4, stft_mean=mean, stft_std=std, phase_u=pu)
        graphed = glg(spec, stft_mean=mean, stft_std=std, phase_u=pu).clone()
        torch.cuda.synchronize()
        assert torch.equal(eager, graphed)


def test_fifty_iterations_converge_like_the_oracle():
    """C5-shaped: 500 frames (T=100, r=5), 50 iterations; compare the spectral convergence ||  |STFT(y)| - mag ||_F / ||mag||_F."""
    from tacotron_b200 import audio
    g = torch.Generator().manual_seed(11)
    r, T = 5, 100
    n = 4 * r * (T // 4)
    # a smooth, speech-like log-magnitude surface (random spectrograms are not invertible and converge nowhere)
    t = torch.linspace(0, 1, n)[:, None]; k = torch.linspace(0, 1, 1025)[None, :]
    logmag = -3.0 * k + 1.5 * torch.sin(2 * math.pi * (3 * t + 8 * k * (1 + 0.3 * torch.sin(2 * math.pi * 2 * t)))) - 1.0
    spec = audio.reshape_frames(logmag.t().contiguous(), r, forward=True)[None]          # [1, T, 1025*r]
    pu = torch.rand(1, n, 1025, generator=g)
    y = audio.invert_spectrogram(spec.cuda().contiguous(), r, n_iter=50, phase_u=pu.cuda())[0].cpu().double().numpy()
    ref = A.invert_spectrogram(spec[0].double().numpy(), r, np.exp(2j * np.pi * pu[0].double().numpy().T), n_iter=50)
    mag = np.exp(logmag.double().numpy().T)

    def sc(w):
        return np.linalg.norm(np.abs(A.stft(w)) - mag) / np.linalg.norm(mag)
    s_gpu, s_ref = sc(y), sc(ref)
    assert abs(s_gpu - s_ref) <= 0.1 * s_ref + 1e-3, (s_gpu, s_ref)
