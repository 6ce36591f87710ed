"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/tf12.py header).

Restatement of the reference model graph (models/tacotron.py, models/ops.py) over
the TF-1.2 primitives in oracle/tf12.py.  Each function cites the reference lines it
follows.  Parameters are an explicit dict name -> tensor (layouts = TF variable
layouts), dropout keep-masks and the scheduled-sampling mask are explicit inputs so
the CUDA path and the oracle see identical bits.

Also serves as the "reference CPU path" stand-in timed by bench.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

from . import tf12


# ----------------------------------------------------------------------------
# Config: models/tacotron.py:12-33 (class attributes), plus the fields the
# drivers inject at run time (r, vocab_size: train.py:20-22) made explicit.
# ----------------------------------------------------------------------------
@dataclass
class OracleConfig:
    r: int = 5
    vocab_size: int = 64
    max_decode_iter: int = 200
    attention_units: int = 256
    decoder_units: int = 256
    mel_features: int = 80
    embed_dim: int = 256
    fft_size: int = 1025
    char_dropout_prob: float = 0.5
    audio_dropout_prob: float = 0.5
    scheduled_sample: float = 0.5
    cap_grads: float = 5.0
    init_lr: float = 0.0005
    enc_K: int = 16
    enc_c: tuple = (128, 128, 128)
    post_K: int = 8
    post_c: tuple = (128, 256, 80)
    gru_units: int = 128
    num_highway_layers: int = 4


# ----------------------------------------------------------------------------
# Parameter construction with TF-default initialisers (SURVEY.md section 8a inventory).
# ----------------------------------------------------------------------------
def _glorot(gen, shape, fan_in, fan_out, dtype):
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * limit).to(dtype)


def cbhg_param_shapes(prefix, cin, K, c, gru_units, n_hw):
    """Ordered (name, shape, kind) list of one CBHG's variables (models/ops.py:48-132)."""
    out = []
    for k in range(1, K + 1):
        out.append((f"{prefix}/bank/W{k}", (k, cin, c[0]), "conv"))
        out.append((f"{prefix}/bank/b{k}", (c[0],), "zeros"))
    ch = K * c[0]
    for nm in ("gamma", "beta", "mean", "var"):
        out.append((f"{prefix}/bank/bn_{nm}", (ch,), nm))
    prev = ch
    for i, co in enumerate(c[1:], start=1):
        out.append((f"{prefix}/proj{i}/W", (3, prev, co), "conv"))
        out.append((f"{prefix}/proj{i}/b", (co,), "zeros"))
        for nm in ("gamma", "beta", "mean", "var"):
            out.append((f"{prefix}/proj{i}/bn_{nm}", (co,), nm))
        prev = co
    hw_units = 128                                            # ops.highway(units=128) default, ops.py:27
    d = cin                                                   # residual output has the input width (ops.py:92)
    for l in range(n_hw):
        if d != hw_units:                                     # ops.py:29-30
            out.append((f"{prefix}/highway{l}/Wd", (d, hw_units), "dense"))
            out.append((f"{prefix}/highway{l}/bd", (hw_units,), "zeros"))
            d = hw_units
        out.append((f"{prefix}/highway{l}/WT", (hw_units, hw_units), "dense"))
        out.append((f"{prefix}/highway{l}/bT", (hw_units,), "zeros"))
        out.append((f"{prefix}/highway{l}/WH", (hw_units, hw_units), "dense"))
        out.append((f"{prefix}/highway{l}/bH", (hw_units,), "zeros"))
    for dname in ("gru_fw", "gru_bw"):
        out.append((f"{prefix}/{dname}/Wg", (hw_units + gru_units, 2 * gru_units), "dense"))
        out.append((f"{prefix}/{dname}/bg", (2 * gru_units,), "ones"))
        out.append((f"{prefix}/{dname}/Wc", (hw_units + gru_units, gru_units), "dense"))
        out.append((f"{prefix}/{dname}/bc", (gru_units,), "zeros"))
    return out


def param_shapes(cfg: OracleConfig):
    """Ordered list of every variable on the hot path: (name, shape, init-kind)."""
    mel_out = cfg.mel_features * cfg.r
    U = cfg.decoder_units
    A = cfg.attention_units
    P = [("embedding", (cfg.vocab_size, cfg.embed_dim), "dense")]
    P += [("enc/prenet/W1", (cfg.embed_dim, 256), "dense"), ("enc/prenet/b1", (256,), "zeros"),
          ("enc/prenet/W2", (256, 128), "dense"), ("enc/prenet/b2", (128,), "zeros")]
    P += cbhg_param_shapes("enc/cbhg", 128, cfg.enc_K, cfg.enc_c, cfg.gru_units, cfg.num_highway_layers)
    P += [("dec/attn/W_mem", (2 * cfg.gru_units, A), "dense"),
          ("dec/attn/W_q", (mel_out, A), "dense"),
          ("dec/attn/v", (A,), "attn_v"),
          ("dec/attn/W_a", (mel_out + 2 * cfg.gru_units, A), "dense")]
    P += [("dec/prenet/W1", (cfg.mel_features, 256), "dense"), ("dec/prenet/b1", (256,), "zeros"),
          ("dec/prenet/W2", (256, 128), "dense"), ("dec/prenet/b2", (128,), "zeros")]
    P += [("dec/in_proj/W", (128 + A, U), "dense"), ("dec/in_proj/b", (U,), "zeros")]
    for i in (1, 2, 3):
        P += [(f"dec/gru{i}/Wg", (2 * U, 2 * U), "dense"), (f"dec/gru{i}/bg", (2 * U,), "ones"),
              (f"dec/gru{i}/Wc", (2 * U, U), "dense"), (f"dec/gru{i}/bc", (U,), "zeros")]
    P += [("dec/out_proj/W", (U, mel_out), "dense"), ("dec/out_proj/b", (mel_out,), "zeros")]
    P += cbhg_param_shapes("post/cbhg", cfg.mel_features, cfg.post_K, cfg.post_c, cfg.gru_units,
                           cfg.num_highway_layers)
    P += [("post/dense/W", (2 * cfg.gru_units, cfg.fft_size), "dense"),
          ("post/dense/b", (cfg.fft_size,), "zeros")]
    return P


NON_TRAINABLE_SUFFIXES = ("bn_mean", "bn_var")


def init_params(cfg: OracleConfig, seed=1, trained_like=False, dtype=torch.float32):
    """TF-default initialisation: glorot-uniform kernels (conv fan = k*C), zero biases,
    GRU gate bias 1.0, BN gamma=1 beta=0 mean=0 var=1.  trained_like=True perturbs biases
    / BN affine / moving stats so that they are non-trivial."""
    gen = torch.Generator().manual_seed(seed)
    params = {}
    for name, shape, kind in param_shapes(cfg):
        if kind == "dense":
            t = _glorot(gen, shape, shape[0], shape[1], dtype)
        elif kind == "conv":
            k, ci, co = shape
            t = _glorot(gen, shape, k * ci, k * co, dtype)
        elif kind == "attn_v":
            # BahdanauAttention attention_v: get_variable default = glorot-uniform on a [U] vector
            limit = math.sqrt(6.0 / (shape[0] + shape[0]))   # fan_in = fan_out = U
            t = ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * limit).to(dtype)
        elif kind in ("zeros", "beta", "mean"):
            t = torch.zeros(shape, dtype=dtype)
            if trained_like:
                t = (0.1 * torch.randn(shape, generator=gen, dtype=torch.float64)).to(dtype)
        elif kind in ("ones", "gamma"):
            t = torch.ones(shape, dtype=dtype)
            if trained_like:
                t = (1.0 + 0.1 * torch.randn(shape, generator=gen, dtype=torch.float64)).to(dtype)
        elif kind == "var":
            t = torch.ones(shape, dtype=dtype)
            if trained_like:
                t = (1.0 + 0.2 * torch.rand(shape, generator=gen, dtype=torch.float64)).to(dtype)
        else:
            raise ValueError(kind)
        params[name] = t
    return params


# ----------------------------------------------------------------------------
# models/tacotron.py:38-44  Tacotron.pre_net
# ----------------------------------------------------------------------------
def pre_net(x, p, prefix, rate, masks=None):
    m1, m2 = masks if masks is not None else (None, None)
    l1 = tf12.dense(x, p[f"{prefix}/W1"], p[f"{prefix}/b1"], torch.relu)
    l1 = tf12.dropout(l1, m1, rate)
    l2 = tf12.dense(l1, p[f"{prefix}/W2"], p[f"{prefix}/b2"], torch.relu)
    l2 = tf12.dropout(l2, m2, rate)
    return l2


# ----------------------------------------------------------------------------
# models/ops.py:27-46  highway
# ----------------------------------------------------------------------------
def highway(x, p, prefix):
    if f"{prefix}/Wd" in p:                                    # ops.py:29-30 (input width != units)
        x = tf12.dense(x, p[f"{prefix}/Wd"], p[f"{prefix}/bd"])
    T = tf12.dense(x, p[f"{prefix}/WT"], p[f"{prefix}/bT"], torch.sigmoid)
    H = tf12.dense(x, p[f"{prefix}/WH"], p[f"{prefix}/bH"], torch.relu)
    return H * T + x * (1 - T)


# ----------------------------------------------------------------------------
# models/ops.py:48-132  CBHG  (speaker_embed=None path only; multi-speaker is out of scope)
# ----------------------------------------------------------------------------
def cbhg(x, p, prefix, K, n_proj=2, n_hw=4, trace=None):
    def bn(t, pre):
        return tf12.batch_norm_inference(t, p[f"{pre}/bn_gamma"], p[f"{pre}/bn_beta"],
                                         p[f"{pre}/bn_mean"], p[f"{pre}/bn_var"])
    # ops.py:54-62 conv bank, concat on channels
    bank = torch.cat([tf12.conv1d_same(x, p[f"{prefix}/bank/W{k}"], p[f"{prefix}/bank/b{k}"], torch.relu)
                      for k in range(1, K + 1)], -1)
    bank = bn(bank, f"{prefix}/bank")                         # ops.py:64
    bank = tf12.max_pool_2_1_same(bank)                        # ops.py:66-71
    proj = bank
    for i in range(1, n_proj + 1):                             # ops.py:76-87
        act = None if i == n_proj else torch.relu
        proj = tf12.conv1d_same(proj, p[f"{prefix}/proj{i}/W"], p[f"{prefix}/proj{i}/b"], act)
        proj = bn(proj, f"{prefix}/proj{i}")
    res = proj + x                                             # ops.py:92
    h = res
    for l in range(n_hw):                                      # ops.py:97-107
        h = highway(h, p, f"{prefix}/highway{l}")
    fw = tuple(p[f"{prefix}/gru_fw/{n}"] for n in ("Wg", "bg", "Wc", "bc"))
    bw = tuple(p[f"{prefix}/gru_bw/{n}"] for n in ("Wg", "bg", "Wc", "bc"))
    out = tf12.bidirectional_gru(h, fw, bw)                    # ops.py:118-128
    if trace is not None:
        trace[f"{prefix}/bank_pool"] = bank
        trace[f"{prefix}/res"] = res
        trace[f"{prefix}/highway_out"] = h
        trace[f"{prefix}/out"] = out
    return out


# ----------------------------------------------------------------------------
# models/tacotron.py:46-105, 136-138  decoder (create_decoder + dynamic_decode)
# ----------------------------------------------------------------------------
def decoder(encoded, text_length, p, cfg: OracleConfig, mode, T, mel=None,
            drop_masks=None, sample_mask=None, trace=None):
    """mode: 'infer' (InferenceHelper, ops.py:5-25), 'teacher' (TrainingHelper),
    'sched' (ScheduledOutputTrainingHelper with explicit Bernoulli mask [T,B], 1 = feed own output).
    drop_masks: None or (m1 [T,B,256], m2 [T,B,128]) keep-masks for the decoder pre-net.
    Returns y [B,T,80r], alignments [B,T,Tx]."""
    B, Tx, _ = encoded.shape
    dt = encoded.dtype
    U = cfg.decoder_units
    mf, r = cfg.mel_features, cfg.r
    values, keys, mask = tf12.attention_prepare(encoded, text_length, p["dec/attn/W_mem"])   # tacotron.py:48-52
    h = [torch.zeros(B, U, dtype=dt) for _ in range(3)]         # cell.zero_state, tacotron.py:94
    attn = torch.zeros(B, cfg.attention_units, dtype=dt)
    if mode == "infer":
        x = torch.zeros(B, mf * r, dtype=dt)                    # ops.py:10
    else:
        x = mel[:, 0]                                           # TrainingHelper.initialize: inputs[:,0] (A.8)
    ys, als = [], []
    for t in range(T):
        dm = None if drop_masks is None else (drop_masks[0][t], drop_masks[1][t])
        # cell_input_fn, tacotron.py:64-71: pre_net on the LAST of the r frames, concat attention
        pn = pre_net(x[:, (r - 1) * mf:], p, "dec/prenet", cfg.audio_dropout_prob, dm)
        z = tf12.dense(torch.cat([pn, attn], -1), p["dec/in_proj/W"], p["dec/in_proj/b"])  # InputProjectionWrapper
        inp = z
        for i in range(3):                                      # MultiRNNCell
            h[i] = tf12.gru_cell(inp, h[i], p[f"dec/gru{i+1}/Wg"], p[f"dec/gru{i+1}/bg"],
                                 p[f"dec/gru{i+1}/Wc"], p[f"dec/gru{i+1}/bc"])
            inp = h[i]
        res = z + inp                                           # ResidualWrapper around the 3-stack
        y = tf12.dense(res, p["dec/out_proj/W"], p["dec/out_proj/b"])   # OutputProjectionWrapper
        # AttentionWrapper.call (A.6): query = cell_output = y
        a = tf12.bahdanau_alignments(y, keys, mask, p["dec/attn/W_q"], p["dec/attn/v"])
        ctx = torch.bmm(a[:, None, :], values)[:, 0]
        attn = torch.cat([y, ctx], -1) @ p["dec/attn/W_a"]
        ys.append(y)
        als.append(a)
        # helper.next_inputs
        if mode == "infer":
            x = y
        else:
            nxt = mel[:, t + 1] if t + 1 < T else torch.zeros_like(y)
            if mode == "sched":
                s = sample_mask[t].to(torch.bool)[:, None]      # per batch element (A.9)
                x = torch.where(s, y, nxt)
            else:
                x = nxt
    yy = torch.stack(ys, 1)
    aa = torch.stack(als, 1)
    if trace is not None:
        trace["dec/keys"] = keys
        trace["dec/values"] = values
    return yy, aa


# ----------------------------------------------------------------------------
# models/tacotron.py:107-154  Tacotron.inference
# ----------------------------------------------------------------------------
def inference(p, inputs, cfg: OracleConfig, train=False, enc_drop_masks=None, dec_drop_masks=None,
              sample_mask=None, T=None, trace=None):
    """inputs: dict with 'text' [B,Tx] int, 'text_length' [B] int, and for train 'mel' [B,T,80r].
    train=False: dropout off, InferenceHelper, exactly T (= cfg.max_decode_iter) steps.
    train=True: dropout via the explicit keep masks, teacher forcing; scheduled sampling if
    sample_mask is given (tacotron.py:82-87)."""
    text = inputs["text"].to(torch.int64)
    emb = p["embedding"][text]                                                      # :111-114
    pre = pre_net(emb, p, "enc/prenet", cfg.char_dropout_prob, enc_drop_masks)      # :128
    encoded = cbhg(pre, p, "enc/cbhg", cfg.enc_K, trace=trace)                       # :131
    if train:
        mel = inputs["mel"]
        T = mel.shape[1] if T is None else T
        mode = "sched" if sample_mask is not None else "teacher"
    else:
        mel = None
        T = cfg.max_decode_iter if T is None else T
        mode = "infer"
    y, align = decoder(encoded, inputs["text_length"], p, cfg, mode, T, mel=mel,
                       drop_masks=dec_drop_masks, sample_mask=sample_mask, trace=trace)   # :135-138
    B = y.shape[0]
    post_in = y.reshape(B, -1, cfg.mel_features)                                     # :144-145
    post = cbhg(post_in, p, "post/cbhg", cfg.post_K, trace=trace)                    # :147
    out = tf12.dense(post, p["post/dense/W"], p["post/dense/b"])                     # :148
    out = out.reshape(B, -1, cfg.fft_size * cfg.r)                                   # :151
    if trace is not None:
        trace["emb"] = emb
        trace["enc/prenet_out"] = pre
        trace["encoded"] = encoded
    return y, out, align


# ----------------------------------------------------------------------------
# models/tacotron.py:156-165  add_loss_op
# ----------------------------------------------------------------------------
def loss(seq2seq_output, output, mel, linear):
    s = (seq2seq_output - mel).abs().sum()
    o = (output - linear).abs().sum()
    return s + o, s, o


# ----------------------------------------------------------------------------
# Synthetic inputs: identical bits for oracle and GPU path.
# ----------------------------------------------------------------------------
def synthetic_inputs(cfg: OracleConfig, B, Tx, T, seed=0, ragged=False, with_targets=True):
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(1, cfg.vocab_size, (B, Tx), generator=g, dtype=torch.int32)
    if ragged:
        lo = max(1, Tx // 2)
        tl = torch.randint(lo, Tx + 1, (B,), generator=g, dtype=torch.int32)
        text = text * (torch.arange(Tx)[None, :] < tl[:, None]).to(torch.int32)
    else:
        tl = torch.full((B,), Tx, dtype=torch.int32)
    inp = {"text": text, "text_length": tl}
    if with_targets:
        mel = torch.randn(B, T, cfg.mel_features * cfg.r, generator=g).half().float()
        stft = torch.randn(B, T, cfg.fft_size * cfg.r, generator=g).half().float()
        inp.update(mel=mel, stft=stft, speech_length=torch.full((B,), T, dtype=torch.int32))
    return inp


def dropout_masks(cfg: OracleConfig, B, Tx, T, seed=2):
    g = torch.Generator().manual_seed(seed)
    bern = lambda *s: (torch.rand(*s, generator=g) >= 0.5).to(torch.uint8)
    enc = (bern(B, Tx, 256), bern(B, Tx, 128))
    dec = (bern(T, B, 256), bern(T, B, 128))
    return enc, dec


def sched_mask(cfg: OracleConfig, B, T, seed=3):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(T, B, generator=g) < cfg.scheduled_sample).to(torch.uint8)


# ----------------------------------------------------------------------------
# models/tacotron.py:167-185  add_train_op -- gradients / clip / Adam, as the target the CUDA backward
# (not built yet) will be checked against.  torch.autograd over the forward above; TF semantics A.12.
# ----------------------------------------------------------------------------
def loss_and_grads(p, inputs, cfg: OracleConfig, enc_drop_masks=None, dec_drop_masks=None, sample_mask=None):
    """Returns (loss, {name: grad}) for the trainable variables.  Scheduled sampling back-propagates through the
    sampled outputs (no stop_gradient, A.9); BN moving statistics get no gradient (A.4)."""
    q = {k: v.detach().clone().requires_grad_(not k.endswith(NON_TRAINABLE_SUFFIXES)) for k, v in p.items()}
    y, out, _ = inference(q, inputs, cfg, train=True, enc_drop_masks=enc_drop_masks, dec_drop_masks=dec_drop_masks,
                          sample_mask=sample_mask)
    total, _, _ = loss(y, out, inputs["mel"].to(y.dtype), inputs["stft"].to(y.dtype))
    names = [k for k, v in q.items() if v.requires_grad]
    grads = torch.autograd.grad(total, [q[k] for k in names])
    return total.detach(), dict(zip(names, grads))


def train_step(p, m, v, inputs, cfg: OracleConfig, lr, step, **kw):
    """One reference training step: grads -> clip_by_global_norm(cap_grads) -> TF Adam.  p/m/v: dicts; step is 1-based."""
    total, g = loss_and_grads(p, inputs, cfg, **kw)
    names = list(g)
    clipped, gnorm = tf12.clip_by_global_norm([g[n] for n in names], float(cfg.cap_grads))
    for n, gc in zip(names, clipped):
        p[n], m[n], v[n] = tf12.adam_tf(p[n], gc, m[n], v[n], lr, step)
    return total, gnorm
