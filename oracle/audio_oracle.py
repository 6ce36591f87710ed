"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  SURVEY.md section 8(f) rank 1: spectrogram inversion.

numpy restatement of the reference's ``audio.invert_spectrogram`` / ``audio.griffinlim`` (audio.py:67-97) and of the
two librosa calls they are written in.  librosa is an un-vendored, unpinned dependency of the reference (README: "librosa";
no requirements file) and is not installed here, so its published algorithm (librosa >= 0.5, the first release whose
``istft`` takes ``win_length``/``window`` as the reference passes them, audio.py:85-86,94) is restated:

  stft (center=True, pad_mode='reflect'):  y padded by n_fft//2 on both sides by reflection; frame t = y_pad[t*hop : t*hop+n_fft];
        D[:, t] = rfft(w * frame),  w = periodic hann(win_length) zero-padded symmetrically to n_fft.
  istft (center=True): y[t*hop : t*hop+n_fft] += w * irfft(D[:, t]); y /= sum_t w^2 shifted by t*hop wherever that sum
        exceeds tiny; the n_fft//2 border samples are trimmed.

PARITY UNPINNED w.r.t. librosa itself (cannot run here).  Pinned instead (tests/test_audio.py): the stft/istft
restatements against torch.stft / torch.istft, which implement the same published conventions independently (1e-14);
`reshape_frames` against outputs of the reference's own function (tests/golden/reshape_frames.npz); and the WIRING of the
inversion against outputs of the reference's own `audio.invert_spectrogram`, executed unmodified from /root/reference with
librosa's two calls stood in by the restatements below (tests/golden/reference_griffinlim.npz; reproduced exactly).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np

from .tf12 import reshape_frames

N_FFT, WIN_LENGTH, HOP_LENGTH = 2048, 1200, 300        # audio.py:10-12


def hann_padded(win_length=WIN_LENGTH, n_fft=N_FFT):
    """scipy.signal.get_window('hann', win_length, fftbins=True) centred in n_fft zeros (librosa util.pad_center)."""
    n = np.arange(win_length)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)
    lpad = (n_fft - win_length) // 2
    out = np.zeros(n_fft)
    out[lpad:lpad + win_length] = w
    return out


def stft(y, n_fft=N_FFT, hop=HOP_LENGTH, win_length=WIN_LENGTH):
    """librosa.stft(y, n_fft, hop_length, win_length, window='hann')  ->  complex [1+n_fft/2, n_frames]"""
    w = hann_padded(win_length, n_fft)
    yp = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop
    D = np.empty((1 + n_fft // 2, n_frames), dtype=np.complex128)
    for t in range(n_frames):
        D[:, t] = np.fft.rfft(w * yp[t * hop:t * hop + n_fft])
    return D


def window_sumsquare(n_frames, n_fft=N_FFT, hop=HOP_LENGTH, win_length=WIN_LENGTH):
    w2 = hann_padded(win_length, n_fft) ** 2
    n = n_fft + hop * (n_frames - 1)
    x = np.zeros(n)
    for t in range(n_frames):
        x[t * hop:t * hop + n_fft] += w2
    return x


def istft(D, hop=HOP_LENGTH, win_length=WIN_LENGTH):
    """librosa.istft(D, hop_length, win_length, window='hann')  ->  real [hop*(n_frames-1)]"""
    n_fft = 2 * (D.shape[0] - 1)
    n_frames = D.shape[1]
    w = hann_padded(win_length, n_fft)
    y = np.zeros(n_fft + hop * (n_frames - 1))
    for t in range(n_frames):
        y[t * hop:t * hop + n_fft] += w * np.fft.irfft(D[:, t], n=n_fft)
    ss = window_sumsquare(n_frames, n_fft, hop, win_length)
    nz = ss > np.finfo(np.float64).tiny
    y[nz] /= ss[nz]
    return y[n_fft // 2:len(y) - n_fft // 2]


def griffinlim(magnitude, angles0, n_iter=50, n_fft=N_FFT, win_length=WIN_LENGTH, hop=HOP_LENGTH):
    """audio.py:77-97 with the random initial phase (np.random.rand, :81) made an explicit input.
    magnitude [1+n_fft/2, frames] >= 0; angles0 complex unit phasors of the same shape."""
    mag = np.abs(magnitude).astype(np.complex128)
    angles = angles0.astype(np.complex128)
    for _ in range(n_iter):
        inverse = istft(mag * angles, hop, win_length)                         # :85
        rebuilt = stft(inverse, n_fft, hop, win_length)                        # :86
        angles = np.exp(1j * np.angle(rebuilt))                                # :87
    return istft(mag * angles, hop, win_length)                                # :93-94


def invert_spectrogram(spec, r, angles0, n_iter=50):
    """audio.py:67-75: spec [T, 1025*r] log-magnitudes in the r-frames-per-step layout -> waveform."""
    s = reshape_frames(spec, r, forward=False)                                 # [frames, 1025]
    return griffinlim(np.exp(s.T), angles0, n_iter=n_iter)
