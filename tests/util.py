"""Shared helpers for the parity tests (oracle = checker only)."""
import torch

from oracle import tacotron_oracle as O


def ocfg(r=5, T=200, vocab=64):
    return O.OracleConfig(r=r, max_decode_iter=T, vocab_size=vocab)


def make_model(cfg_o, params, precision="fp32"):
    from tacotron_b200 import Config, Tacotron
    cfg = Config(r=cfg_o.r, vocab_size=cfg_o.vocab_size, max_decode_iter=cfg_o.max_decode_iter, precision=precision)
    m = Tacotron(cfg, None, train=False)
    m.load_params(params)
    return m


def to_cuda(inp):
    return {k: v.cuda().contiguous() for k, v in inp.items()}


PARITY_LOG = []          # (what, max-rel, mean-rel, tol) of every comparison made in this process (dumped by conftest.py)


def relerr(a, b):
    """max |a-b| / max|b| and mean-relative error, a = device result, b = oracle."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    d = (a - b).abs()
    return float(d.max() / b.abs().max().clamp_min(1e-30)), float(d.mean() / b.abs().mean().clamp_min(1e-30))


def assert_close(a, b, tol, what=""):
    """max|a-b| <= tol * max|b|  AND  mean|a-b| <= tol * mean|b| (the second bound keeps a result that is only right
    on its few largest elements -- most of a log-spectrogram is small -- from passing on the global scale alone)."""
    mx, mean = relerr(a, b)
    PARITY_LOG.append((what, mx, mean, tol))
    assert mx <= tol, f"{what}: max-rel err {mx:.3e} (mean-rel {mean:.3e}) > tol {tol:.1e}"
    assert mean <= tol, f"{what}: mean-rel err {mean:.3e} (max-rel {mx:.3e}) > tol {tol:.1e}"
    return mx
