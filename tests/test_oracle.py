"""CPU tests of the oracle itself (-m "not gpu").  The oracle restates TF-1.2 semantics
(SURVEY.md Appendix A); these tests pin the easy-to-get-wrong ones with closed-form cases and
check the oracle against the committed golden fixtures."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import tf12
from oracle import tacotron_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_reshape_frames_matches_reference_golden():
    """Fixture produced by the REFERENCE's audio.reshape_frames (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, "reshape_frames.npz"))
    for r in (2, 5):
        fwd = tf12.reshape_frames(g[f"x_r{r}"], r)
        assert np.array_equal(fwd, g[f"fwd_r{r}"])
        inv = tf12.reshape_frames(fwd, r, forward=False)
        assert np.array_equal(inv, g[f"inv_r{r}"])
    # the reference's own self-test, audio.py:106-115
    out = tf12.reshape_frames(g["selftest_in"].T, 2)
    assert np.array_equal(out, g["selftest_fwd"])
    assert np.array_equal(tf12.reshape_frames(out, 2, forward=False), g["selftest_in"])


def test_reshape_frames_layout():
    # r=5, step 0 -> frames 0,4,8,12,16 ; step 1 -> 1,5,9,13,17 ; step 4 -> 20,24,...
    x = np.arange(361, dtype=np.float32)[None, :]
    f = tf12.reshape_frames(x, 5)
    assert f.shape == (72, 5)
    assert f[0].tolist() == [0, 4, 8, 12, 16]
    assert f[1].tolist() == [1, 5, 9, 13, 17]
    assert f[4].tolist() == [20, 24, 28, 32, 36]


@pytest.mark.parametrize("k", [1, 2, 3, 4, 7, 8, 16])
def test_conv_same_padding_asymmetry(k):
    """Even kernels pad one more on the RIGHT: an impulse at s shows up at t = s + pad_left - j."""
    T = 12
    x = torch.zeros(1, T, 1)
    x[0, 5, 0] = 1.0
    W = torch.arange(1, k + 1, dtype=torch.float32).view(k, 1, 1)
    y = tf12.conv1d_same(x, W)[0, :, 0]
    pl = (k - 1) // 2
    for j in range(k):
        t = 5 + pl - j
        if 0 <= t < T:
            assert y[t] == W[j, 0, 0]
    torch.testing.assert_close(tf12.conv1d_same(x, W), tf12.conv1d_same_loops(x, W))


def test_conv_same_random_vs_loops():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 9, 3, generator=g)
    for k in (1, 2, 3, 5, 8):
        W = torch.randn(k, 3, 4, generator=g)
        b = torch.randn(4, generator=g)
        torch.testing.assert_close(tf12.conv1d_same(x, W, b), tf12.conv1d_same_loops(x, W, b), rtol=1e-5, atol=1e-5)


def test_max_pool_looks_forward():
    x = torch.tensor([1.0, 5.0, 2.0, 3.0, -1.0]).view(1, 5, 1)
    y = tf12.max_pool_2_1_same(x)[0, :, 0]
    assert y.tolist() == [5.0, 5.0, 3.0, 3.0, -1.0]


def test_batch_norm_is_affine_with_eps():
    x = torch.tensor([[[2.0]]])
    y = tf12.batch_norm_inference(x, torch.tensor([3.0]), torch.tensor([0.5]), torch.tensor([0.0]), torch.tensor([1.0]))
    assert abs(float(y) - (3.0 * 2.0 / math.sqrt(1.001) + 0.5)) < 1e-6


def test_gru_cell_hand_computed():
    # n=1, x=1: Wg rows = (x, h), cols = (r, u)
    x = torch.tensor([[0.5]]); h = torch.tensor([[0.25]])
    Wg = torch.tensor([[0.1, 0.2], [0.3, 0.4]]); bg = torch.tensor([1.0, 1.0])
    Wc = torch.tensor([[0.7], [-0.6]]); bc = torch.tensor([0.05])
    r = 1 / (1 + math.exp(-(0.5 * 0.1 + 0.25 * 0.3 + 1)))
    u = 1 / (1 + math.exp(-(0.5 * 0.2 + 0.25 * 0.4 + 1)))
    c = math.tanh(0.5 * 0.7 + (r * 0.25) * (-0.6) + 0.05)          # reset BEFORE the candidate product
    hn = u * 0.25 + (1 - u) * c
    assert abs(float(tf12.gru_cell(x, h, Wg, bg, Wc, bc)) - hn) < 1e-6


def test_bidirectional_gru_backward_is_time_reversed():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 5, 3, generator=g)
    mk = lambda: (torch.randn(5, 4, generator=g) * 0.3, torch.ones(4), torch.randn(5, 2, generator=g) * 0.3, torch.zeros(2))
    fw, bw = mk(), mk()
    out = tf12.bidirectional_gru(x, fw, bw)
    out_rev = tf12.bidirectional_gru(x.flip(1), bw, fw)
    torch.testing.assert_close(out[..., 2:], out_rev[..., :2].flip(1))


def test_attention_mask_and_softmax():
    B, Tx, D = 2, 6, 4
    g = torch.Generator().manual_seed(2)
    mem = torch.randn(B, Tx, D, generator=g)
    length = torch.tensor([6, 3], dtype=torch.int32)
    W_mem = torch.randn(D, 5, generator=g)
    values, keys, mask = tf12.attention_prepare(mem, length, W_mem)
    assert torch.all(values[1, 3:] == 0) and torch.all(keys[1, 3:] == 0)
    a = tf12.bahdanau_alignments(torch.randn(B, 7, generator=g), keys, mask, torch.randn(7, 5, generator=g),
                                 torch.randn(5, generator=g))
    assert torch.all(a[1, 3:] == 0)
    torch.testing.assert_close(a.sum(-1), torch.ones(B))


def test_adam_and_clip_tf_form():
    g = [torch.tensor([3.0, 4.0])]
    clipped, gn = tf12.clip_by_global_norm(g, 2.5)
    assert abs(float(gn) - 5.0) < 1e-6
    torch.testing.assert_close(clipped[0], torch.tensor([1.5, 2.0]))
    p, m, v = tf12.adam_tf(torch.tensor([1.0]), torch.tensor([0.5]), torch.zeros(1), torch.zeros(1), lr=0.1, step=1)
    # step 1: m=0.05, v=0.00025, lr_t = 0.1*sqrt(0.001)/0.1 ; update = lr_t*m/(sqrt(v)+eps)
    lr_t = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    assert abs(float(p) - (1.0 - lr_t * 0.05 / (math.sqrt(0.00025) + 1e-8))) < 1e-6


def test_param_inventory_matches_survey():
    cfg = O.OracleConfig(vocab_size=40)
    n = sum(int(np.prod(s)) for _, s, _ in O.param_shapes(cfg))
    assert n == 7113377                                   # SURVEY.md section 8a (incl. 7328 BN moving stats)


@pytest.mark.parametrize("r", [2, 5])
def test_oracle_matches_golden_small(r):
    g = np.load(os.path.join(GOLD, f"oracle_small_r{r}.npz"))
    cfg = O.OracleConfig(r=r, max_decode_iter=6, vocab_size=20)
    p = O.init_params(cfg, seed=1, trained_like=True)
    inp = O.synthetic_inputs(cfg, 2, 12, 6, seed=0, ragged=True)
    enc_m, dec_m = O.dropout_masks(cfg, 2, 12, 6, seed=2)
    sm = O.sched_mask(cfg, 2, 6, seed=3)
    y, o, a = O.inference(p, inp, cfg, train=False)
    np.testing.assert_allclose(y.numpy(), g["y_infer"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(o.numpy(), g["out_infer"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(a.numpy(), g["align_infer"], rtol=1e-4, atol=1e-6)
    y, o, a = O.inference(p, inp, cfg, train=True, enc_drop_masks=enc_m, dec_drop_masks=dec_m, sample_mask=sm)
    np.testing.assert_allclose(y.numpy(), g["y_sched"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(o.numpy(), g["out_sched"], rtol=1e-4, atol=1e-5)


def test_fp64_twin_agrees():
    cfg = O.OracleConfig(r=2, max_decode_iter=5, vocab_size=20)
    p = O.init_params(cfg, 1, True)
    inp = O.synthetic_inputs(cfg, 2, 8, 5, ragged=True)
    y, o, a = O.inference(p, inp, cfg, train=False)
    y64, o64, a64 = O.inference({k: v.double() for k, v in p.items()}, inp, cfg, train=False)
    assert float((y.double() - y64).abs().max()) < 1e-5
    assert float((o.double() - o64).abs().max()) < 1e-5


def test_decoder_teacher_forcing_has_no_go_frame():
    """A.8: decoder input at step t is mel[:, t] (not shifted): changing mel[:, t] must change y[:, t]."""
    cfg = O.OracleConfig(r=2, max_decode_iter=4, vocab_size=20)
    p = O.init_params(cfg, 1, True)
    inp = O.synthetic_inputs(cfg, 1, 8, 4)
    y0, _, _ = O.inference(p, inp, cfg, train=True)
    inp2 = dict(inp); inp2["mel"] = inp["mel"].clone(); inp2["mel"][:, 2, -1] += 1.0   # last frame of group 2
    y1, _, _ = O.inference(p, inp2, cfg, train=True)
    assert torch.equal(y0[:, :2], y1[:, :2]) and not torch.equal(y0[:, 2], y1[:, 2])


@pytest.mark.parametrize("r", [2, 5])
def test_oracle_matches_reference_code_wiring(r):
    """tests/golden/reference_wiring_r*.npz was produced by EXECUTING the reference's own models/tacotron.py +
    models/ops.py (imported from /root/reference, unmodified) over oracle/tf12_shim.py with these weights, inputs,
    dropout masks and sampling draws.  The oracle's independent restatement of the graph must agree: this pins the
    wiring (call order, slices, wrapper nesting, helper semantics, loss) against the reference source."""
    g = np.load(os.path.join(GOLD, f"reference_wiring_r{r}.npz"))
    cfg = O.OracleConfig(r=r, max_decode_iter=6, vocab_size=20)
    p = O.init_params(cfg, seed=1, trained_like=True)
    inp = O.synthetic_inputs(cfg, 2, 12, 6, seed=0, ragged=True)
    enc_m, dec_m = O.dropout_masks(cfg, 2, 12, 6, seed=2)
    sm = O.sched_mask(cfg, 2, 6, seed=3)
    y, o, a = O.inference(p, inp, cfg, train=False)
    for got, key in ((y, "y_infer"), (o, "out_infer"), (a, "align_infer")):
        np.testing.assert_allclose(got.numpy(), g[key], rtol=2e-5, atol=2e-6)
    y, o, a = O.inference(p, inp, cfg, train=True, enc_drop_masks=enc_m, dec_drop_masks=dec_m)
    for got, key in ((y, "y_teacher"), (o, "out_teacher"), (a, "align_teacher")):
        np.testing.assert_allclose(got.numpy(), g[key], rtol=2e-5, atol=2e-6)
    assert abs(float(O.loss(y, o, inp["mel"], inp["stft"])[0]) - float(g["loss_teacher"])) / float(g["loss_teacher"]) < 1e-6
    y, o, a = O.inference(p, inp, cfg, train=True, enc_drop_masks=enc_m, dec_drop_masks=dec_m, sample_mask=sm)
    for got, key in ((y, "y_sched"), (o, "out_sched"), (a, "align_sched")):
        np.testing.assert_allclose(got.numpy(), g[key], rtol=2e-5, atol=2e-6)
    assert len(g["tf_variable_names"]) == len(p)


def test_oracle_gradients_match_finite_differences():
    """Target for the (future) CUDA backward: autograd over the oracle, checked against central differences in fp64."""
    cfg = O.OracleConfig(r=2, max_decode_iter=3, vocab_size=12)
    p = {k: v.double() for k, v in O.init_params(cfg, 1, True).items()}
    inp = O.synthetic_inputs(cfg, 2, 8, 3, seed=0, ragged=True)
    inp = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in inp.items()}
    enc_m, dec_m = O.dropout_masks(cfg, 2, 8, 3, seed=2)
    sm = O.sched_mask(cfg, 2, 3, seed=3)
    kw = dict(enc_drop_masks=enc_m, dec_drop_masks=dec_m, sample_mask=sm)
    total, g = O.loss_and_grads(p, inp, cfg, **kw)
    assert set(g) == {k for k in p if not k.endswith(("bn_mean", "bn_var"))}
    gen = torch.Generator().manual_seed(0)
    for name in ("dec/gru2/Wc", "enc/cbhg/bank/W3", "dec/attn/W_q", "post/cbhg/gru_bw/Wg", "dec/prenet/W1", "embedding"):
        idx = tuple(int(torch.randint(0, s, (1,), generator=gen)) for s in p[name].shape)
        eps = 1e-6
        def f(delta):
            q = dict(p); t = p[name].clone(); t[idx] += delta; q[name] = t
            y, o, _ = O.inference(q, inp, cfg, train=True, **kw)
            return float(O.loss(y, o, inp["mel"], inp["stft"])[0])
        fd = (f(eps) - f(-eps)) / (2 * eps)
        an = float(g[name][idx])
        assert abs(fd - an) <= 1e-4 * max(1.0, abs(fd)), (name, idx, fd, an)   # L1 loss: piecewise linear, kinks are measure-zero


def test_oracle_train_step_reduces_nothing_but_runs():
    cfg = O.OracleConfig(r=2, max_decode_iter=3, vocab_size=12)
    p = O.init_params(cfg, 1, True)
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v = {k: torch.zeros_like(t) for k, t in p.items()}
    inp = O.synthetic_inputs(cfg, 2, 8, 3, seed=0)
    l0, gn = O.train_step(p, m, v, inp, cfg, lr=cfg.init_lr, step=1)
    l1, _ = O.train_step(p, m, v, inp, cfg, lr=cfg.init_lr, step=2)
    assert float(gn) > 0 and float(l1) < float(l0)          # same batch twice: Adam must make progress
