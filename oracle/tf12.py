"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see below).

Restatement, in plain PyTorch-CPU tensor arithmetic, of the TensorFlow 1.2
primitives that barronalex/Tacotron's hot path is written in.  TensorFlow 1.2 is
an un-vendored dependency of the reference (README.md:18 "Tensorflow 1.2"; no
requirements file / lock), it cannot be installed here (Python 3.12, no network),
and the reference ships no golden vectors or tests for the model path
(SURVEY.md section 4 / 8c).  Hence: **parity unpinned** -- the semantics below follow
the published TF r1.2 sources (tf.layers, rnn_cell_impl.py, contrib/seq2seq
attention_wrapper.py / helper.py / basic_decoder.py) as restated in SURVEY.md
Appendix A, and are anchored on the reference's own call sites (cited per
function).

What IS pinned (tests/golden/make_golden.py, tests/test_oracle.py):
  * audio.reshape_frames against outputs of the reference's own function;
  * the model graph WIRING (call order, slices, wrapper nesting, helper semantics, loss) against outputs
    obtained by executing the reference's own models/tacotron.py + models/ops.py, unmodified, over
    oracle/tf12_shim.py (a stand-in for the TF-1.2 API whose primitives delegate to the functions below).
What is NOT pinned: the numerics/semantics of the TF-1.2 primitives themselves (this file) -- no
TensorFlow 1.2 run, checkpoint or golden vector is reachable from this container.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product path (tacotron_b200/) never
does: it fails loudly when the CUDA library is missing.

Every function takes/returns torch CPU tensors; dtype (float32 or float64) is
whatever the caller passes in, so the same code doubles as the fp64 twin used
to calibrate tolerances.
"""
from __future__ import annotations

import math
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# A.1  tf.layers.dense            call sites: models/ops.py:30,32,39  models/tacotron.py:40,42,148
# ----------------------------------------------------------------------------
def dense(x, W, b=None, activation=None):
    """y = act(x @ W + b); W is [in, out]; applied on the last axis."""
    y = x @ W
    if b is not None:
        y = y + b
    if activation is not None:
        y = activation(y)
    return y


# ----------------------------------------------------------------------------
# A.2  tf.layers.conv1d(padding='same', strides=1), NWC   call sites: models/ops.py:54,80
# ----------------------------------------------------------------------------
def conv1d_same(x, W, b=None, activation=None):
    """x [B,T,Cin], W [k,Cin,Cout] (TF layout), cross-correlation, zero padding
    pad_left=(k-1)//2, pad_right=(k-1)-pad_left (even k pads one more on the RIGHT):
        y[t,o] = b[o] + sum_j sum_c x[t - pad_left + j, c] * W[j,c,o]
    """
    k = W.shape[0]
    pl = (k - 1) // 2
    pr = (k - 1) - pl
    xp = F.pad(x.transpose(1, 2), (pl, pr))                # [B,Cin,T+k-1]
    w = W.permute(2, 1, 0).contiguous()                    # [Cout,Cin,k]
    y = F.conv1d(xp, w).transpose(1, 2)                    # [B,T,Cout]
    if b is not None:
        y = y + b
    if activation is not None:
        y = activation(y)
    return y


def conv1d_same_loops(x, W, b=None):
    """Literal-loop twin of conv1d_same (small cases only); used by the oracle's
    own self-tests to pin the padding asymmetry."""
    B, T, Cin = x.shape
    k, _, Cout = W.shape
    pl = (k - 1) // 2
    y = torch.zeros(B, T, Cout, dtype=x.dtype)
    for t in range(T):
        for j in range(k):
            s = t - pl + j
            if 0 <= s < T:
                y[:, t, :] += x[:, s, :] @ W[j]
    if b is not None:
        y = y + b
    return y


# ----------------------------------------------------------------------------
# A.3  tf.layers.max_pooling1d(pool_size=2, strides=1, padding='same')   models/ops.py:66-71
# ----------------------------------------------------------------------------
def max_pool_2_1_same(x):
    """y[t] = max(x[t], x[t+1]) for t < T-1 ; y[T-1] = x[T-1] (pad right, padding never wins)."""
    y = x.clone()
    y[:, :-1, :] = torch.maximum(x[:, :-1, :], x[:, 1:, :])
    return y


# ----------------------------------------------------------------------------
# A.4  tf.layers.batch_normalization(x) with training=False (always)   models/ops.py:64,87
# ----------------------------------------------------------------------------
BN_EPS = 1e-3


def batch_norm_inference(x, gamma, beta, mean, var):
    return gamma * (x - mean) / torch.sqrt(var + BN_EPS) + beta


# ----------------------------------------------------------------------------
# A.5  tf.contrib.rnn.GRUCell (TF<=1.x form)   models/ops.py:118-119  models/tacotron.py:54
# ----------------------------------------------------------------------------
def gru_cell(x, h, Wg, bg, Wc, bc):
    """[r,u] = sigmoid([x,h] @ Wg + bg);  c = tanh([x, r*h] @ Wc + bc);  h' = u*h + (1-u)*c."""
    n = h.shape[-1]
    ru = torch.sigmoid(torch.cat([x, h], -1) @ Wg + bg)
    r, u = ru[..., :n], ru[..., n:]
    c = torch.tanh(torch.cat([x, r * h], -1) @ Wc + bc)
    return u * h + (1 - u) * c


def bidirectional_gru(x, fw, bw):
    """tf.nn.bidirectional_dynamic_rnn without sequence_length (models/ops.py:120-128):
    the bw cell runs over the time-reversed FULL padded sequence, its outputs are
    reversed back; zero initial states; output = concat(fw, bw) on the last axis.
    fw / bw are (Wg, bg, Wc, bc)."""
    B, T, _ = x.shape
    n = fw[3].shape[0]
    hf = torch.zeros(B, n, dtype=x.dtype)
    hb = torch.zeros(B, n, dtype=x.dtype)
    of = torch.empty(B, T, n, dtype=x.dtype)
    ob = torch.empty(B, T, n, dtype=x.dtype)
    for t in range(T):
        hf = gru_cell(x[:, t], hf, *fw)
        of[:, t] = hf
        tb = T - 1 - t
        hb = gru_cell(x[:, tb], hb, *bw)
        ob[:, tb] = hb
    return torch.cat([of, ob], -1)


# ----------------------------------------------------------------------------
# A.11  tf.layers.dropout(rate, training)     models/tacotron.py:41,43
# ----------------------------------------------------------------------------
def dropout(x, keep_mask, rate):
    """Training-mode dropout with an EXPLICIT keep mask (1 = keep): kept values are
    scaled by 1/(1-rate).  keep_mask=None means inference (identity)."""
    if keep_mask is None:
        return x
    return x * keep_mask.to(x.dtype) * (1.0 / (1.0 - rate))


# ----------------------------------------------------------------------------
# A.6  BahdanauAttention(num_units, memory, memory_sequence_length, normalize=False)
#      models/tacotron.py:48-52
# ----------------------------------------------------------------------------
def attention_prepare(memory, memory_length, W_mem):
    """values = memory with rows j >= length zeroed ; keys = values @ W_mem (no bias)."""
    B, Tx, _ = memory.shape
    mask = (torch.arange(Tx)[None, :] < memory_length[:, None].to(torch.int64))   # [B,Tx]
    values = memory * mask[..., None].to(memory.dtype)
    keys = values @ W_mem
    return values, keys, mask


def bahdanau_alignments(query, keys, mask, W_q, v):
    """score_j = sum_d v_d * tanh(keys_jd + (query @ W_q)_d); masked scores = -inf; softmax."""
    pq = query @ W_q                                              # [B,U]
    score = (v * torch.tanh(keys + pq[:, None, :])).sum(-1)       # [B,Tx]
    score = torch.where(mask, score, torch.full_like(score, -math.inf))
    return torch.softmax(score, dim=-1)


# ----------------------------------------------------------------------------
# A.12  clip_by_global_norm + TF Adam       models/tacotron.py:170-184
# ----------------------------------------------------------------------------
def clip_by_global_norm(grads, clip):
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).to(grads[0].dtype)
    scale = clip / torch.maximum(gn, torch.tensor(clip, dtype=gn.dtype))
    return [g * scale for g in grads], gn


def adam_tf(p, g, m, v, lr, step, b1=0.9, b2=0.999, eps=1e-8):
    """TF form: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t * m / (sqrt(v) + eps). step is 1-based."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
    p = p - lr_t * m / (torch.sqrt(v) + eps)
    return p, m, v


# ----------------------------------------------------------------------------
# audio.reshape_frames (audio.py:23-35) -- defines the r-frames-per-step layout.  numpy only.
# ----------------------------------------------------------------------------
def reshape_frames(signal, r, forward=True):
    """Restatement of audio.reshape_frames with r explicit (the reference reads a
    module global).  forward: [F, n_frames] -> [n_frames//r, r*F]; blocks of 4r
    frames, step t of block b holds frames 4r*b + (t mod 4) + 4c, c = 0..r-1."""
    import numpy as np
    if forward:
        Fdim, n = signal.shape
        nb = n // (4 * r)                       # the trailing partial block is dropped (splits[:-1])
        blocks = []
        for b in range(nb):
            s = signal[:, 4 * r * b: 4 * r * (b + 1)]              # [F, 4r]
            chunks = [s[:, 4 * c: 4 * (c + 1)] for c in range(r)]  # r x [F,4]
            blocks.append(np.concatenate(chunks, axis=0))          # [rF, 4]
        return np.concatenate(blocks, axis=1).T                    # [4*nb, rF]
    else:
        steps, width = signal.shape
        Fdim = width // r
        sig = np.reshape(signal, (-1, Fdim))                        # [steps*r, F]
        nb = sig.shape[0] // (4 * r)
        out = []
        for b in range(nb):
            s = sig[4 * r * b: 4 * r * (b + 1)]                    # [4r, F]: rows (t, c) -> t*r + c
            # np.split(s, 4r/r = 4 pieces of r rows) then concat on axis 1 -> [r, 4F]
            pieces = [s[r * i: r * (i + 1)] for i in range(4)]
            out.append(np.concatenate(pieces, axis=1))              # [r, 4F]
        new = np.concatenate(out, axis=0)
        return np.reshape(new, (-1, Fdim))
