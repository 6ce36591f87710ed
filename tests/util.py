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


def relerr(a, b):
    """max |a-b| / max|b| and mean-relative error, a = device result, b = oracle."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    d = (a - b).abs()
    return float(d.max() / b.abs().max().clamp_min(1e-30)), float(d.mean() / b.abs().mean().clamp_min(1e-30))


def assert_close(a, b, tol, what=""):
    mx, mean = relerr(a, b)
    assert mx <= tol, f"{what}: max-rel err {mx:.3e} (mean-rel {mean:.3e}) > tol {tol:.1e}"
    return mx
