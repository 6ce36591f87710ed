"""Spectrogram inversion on the GPU -- host-side mirror of the reference's ``audio.py`` inversion surface
(``reshape_frames`` :23-35, ``invert_spectrogram`` :67-75, ``griffinlim`` :77-97), SURVEY.md section 8(f) rank 1:
it sits inside BASELINE config 5's end-to-end latency (test.py:64) and would dominate it on the CPU.

Design: a batch of utterances stays on the device for all iterations.  The two transforms per iteration are the library's
own shared-memory FFT kernels (csrc/audio.cu ``taco_rfft2048`` / ``taco_irfft2048``: radix-2 Stockham, one CTA per frame --
round 1 called cuFFT through ``torch.fft`` here); everything between them is four fused kernels of libtaco_b200.so:

    gl_init   reshape_frames(forward=False) + de-normalisation + exp + initial phase  ->  magnitude, first spectrum
    gl_ola    window, overlap-add, window-sum-square normalisation, centre trim        (the tail of librosa.istft)
    gl_frame  reflect padding, framing, window                                         (the head of librosa.stft)
    gl_phase  mag * rebuilt/|rebuilt|                                                  (audio.py:84,87)

so one iteration is irfft2048 -> gl_ola -> gl_frame -> rfft2048 -> gl_phase: 5 launches of this library, ~20 MB of HBM traffic per utterance of
500 frames.  All arithmetic goes through the kernel namespace K (tacotron_b200/kernels.py); tests run the same
orchestration over tests/mirror_kernels.py on CPU tensors against the numpy oracle (oracle/audio_oracle.py).
"""
from __future__ import annotations

import torch

n_fft = 2048            # audio.py:10
win_length = 1200       # audio.py:11
hop_length = 300        # audio.py:12
maximum_audio_length = 108000   # audio.py:13
F_BINS = 1 + n_fft // 2


def _K():
    from . import kernels
    return kernels


def reshape_frames(signal, r, forward=False):
    """audio.reshape_frames (audio.py:23-35) as a pure index permutation on a torch tensor (any device).
    forward=False: [T, F*r] -> [4r*(T//4), F]; forward=True: [F, n] -> [4*(n//4r), F*r]."""
    if forward:
        Fd, n = signal.shape
        nb = n // (4 * r)
        s = signal[:, :nb * 4 * r].reshape(Fd, nb, r, 4)                     # frame = 4r*b + 4c + tl
        return s.permute(1, 3, 2, 0).reshape(nb * 4, r * Fd)                 # row = 4b + tl, col = c*F + k
    T, W = signal.shape
    Fd = W // r
    nb = T // 4
    s = signal[:nb * 4].reshape(nb, 4, r, Fd)                                # [b, tl, c, k]
    return s.permute(0, 2, 1, 3).reshape(nb * 4 * r, Fd)                     # frame = 4r*b + 4c + tl


def griffinlim_batch(spec, r, n_iter=50, scale=None, shift=None, phase_u=None, K=None, generator=None):
    """spec [B, T, 1025*r] (log magnitudes in the decoder's r-frames-per-step layout, optionally normalised:
    magnitude = exp(spec*scale + shift), the driver's ``out*stft_std + stft_mean`` of test.py:64) -> waveforms [B, L],
    L = hop*(4r*(T//4) - 1).  phase_u: uniform [0,1) initial phases [B, n, 1025] (audio.py:81), drawn if None."""
    K = K or _K()
    B, T, W = spec.shape
    assert W == F_BINS * r, (W, r)
    n = 4 * r * (T // 4)
    assert n >= 2, "need at least one block of 4 decoder steps"
    L = hop_length * (n - 1)
    assert L > n_fft // 2, f"{n} frames are fewer than the reflect padding of the STFT needs (>= 5)"
    rdt = spec.dtype
    cdt = torch.complex64 if rdt == torch.float32 else torch.complex128
    if phase_u is None:
        phase_u = torch.rand((B, n, F_BINS), dtype=rdt, device=spec.device, generator=generator)
    mag = torch.empty((B, n, F_BINS), dtype=rdt, device=spec.device)
    full = torch.empty((B, n, F_BINS), dtype=cdt, device=spec.device)
    y = torch.empty((B, L), dtype=rdt, device=spec.device)
    frw = torch.empty((B, n, n_fft), dtype=rdt, device=spec.device)
    _gl_iterations(K, full, mag, spec, phase_u, r, scale, shift, y, frw, n_iter)
    return y


def _gl_iterations(K, full, mag, spec, phase_u, r, scale, shift, y, frw, n_iter, rebuilt=None):
    """frw [B, n, 2048] doubles as the istft frame buffer: irfft writes it, gl_ola consumes it, gl_frame refills it."""
    if rebuilt is None:
        rebuilt = torch.empty_like(full)
    K.gl_init(full, mag, spec, phase_u, r, scale, shift)
    for it in range(n_iter + 1):
        K.irfft2048(frw, full)                                               # inverse real FFT of every frame
        K.gl_ola(y, frw, hop_length, win_length)
        if it == n_iter:
            break
        K.gl_frame(frw, y, hop_length, win_length)
        K.rfft2048(rebuilt, frw)                                             # real FFT of every re-framed, windowed frame
        K.gl_phase(full, mag, rebuilt)


class GriffinLimGraph:
    """The same 5-launches-per-iteration loop captured ONCE into a CUDA graph for a fixed (B, T, r, n_iter) and replayed:
    251 launches (and their Python / ctypes overhead, ~0.1 ms per iteration of host time) collapse into one
    graph launch.  Inputs are copied into static device buffers; the result is the static `y` (valid until the next call).
    Opt-in: `invert_spectrogram(...)` stays eager."""

    def __init__(self, B, T, r, n_iter=50, device="cuda", with_affine=True):
        K = _K()
        self.B, self.T, self.r, self.n_iter = B, T, r, n_iter
        n = 4 * r * (T // 4)
        L = hop_length * (n - 1)
        assert L > n_fft // 2
        dev = torch.device(device)
        f32 = dict(dtype=torch.float32, device=dev)
        self.spec = torch.zeros((B, T, F_BINS * r), **f32)
        self.phase_u = torch.zeros((B, n, F_BINS), **f32)
        self.scale = torch.ones((F_BINS * r,), **f32) if with_affine else None
        self.shift = torch.zeros((F_BINS * r,), **f32) if with_affine else None
        self.mag = torch.empty((B, n, F_BINS), **f32)
        self.full = torch.empty((B, n, F_BINS), dtype=torch.complex64, device=dev)
        self.y = torch.empty((B, L), **f32)
        self.frw = torch.empty((B, n, n_fft), **f32)
        self.rebuilt = torch.empty_like(self.full)
        args = (K, self.full, self.mag, self.spec, self.phase_u, r, self.scale, self.shift, self.y, self.frw, n_iter, self.rebuilt)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                     # warm-up outside capture: lazy module loads
            _gl_iterations(*args)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            _gl_iterations(*args)

    def __call__(self, spec, stft_mean=None, stft_std=None, phase_u=None, generator=None):
        self.spec.copy_(spec.reshape(self.spec.shape), non_blocking=True)
        if self.scale is not None:
            if stft_std is not None:
                self.scale.copy_(stft_std, non_blocking=True); self.shift.copy_(stft_mean, non_blocking=True)
            else:
                self.scale.fill_(1.0); self.shift.zero_()
        if phase_u is None:
            self.phase_u.uniform_(0.0, 1.0, generator=generator)             # audio.py:81 (in place, on the device)
        else:
            self.phase_u.copy_(phase_u.reshape(self.phase_u.shape), non_blocking=True)
        self.graph.replay()
        return self.y


def griffinlim(spectrogram, n_iter=50, phase_u=None, K=None):
    """audio.griffinlim for ONE magnitude spectrogram [1025, frames] (audio.py:77) -> waveform [hop*(frames-1)]."""
    Fd, n = spectrogram.shape
    assert Fd == F_BINS
    spec = torch.log(spectrogram.t().contiguous().clamp_min(1e-30))[None]    # reuse the batch path with r = 1 layout
    # r=1: reshape_frames(forward=False) is the identity on whole blocks of 4 frames
    assert n % 4 == 0, "frame count must be a multiple of 4 (the reference always produces 4r-frame blocks)"
    pu = None if phase_u is None else phase_u[None]
    return griffinlim_batch(spec, 1, n_iter=n_iter, phase_u=pu, K=K)[0]


def invert_spectrogram(spec, r, n_iter=50, stft_mean=None, stft_std=None, phase_u=None, K=None):
    """audio.invert_spectrogram (audio.py:67-75) with the driver's de-normalisation folded in.
    spec [T, 1025*r] or [B, T, 1025*r] on the device -> waveform(s)."""
    single = spec.dim() == 2
    s = spec[None] if single else spec
    scale, shift = (stft_std, stft_mean) if stft_std is not None else (None, None)
    if phase_u is not None and single:
        phase_u = phase_u[None]
    y = griffinlim_batch(s.contiguous(), r, n_iter=n_iter, scale=scale, shift=shift, phase_u=phase_u, K=K)
    return y[0] if single else y
