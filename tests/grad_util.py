"""Test helper: an oracle forward that also returns the activations the train-mode CUDA forward saves
(names = the S[...] keys of tacotron_b200/models/grad.py), built from oracle/tf12.py primitives."""
from __future__ import annotations

import torch

from oracle import tf12
from oracle import tacotron_oracle as O


def _bn(p, pre, t, S):
    """folded inference affine, exactly as the CUDA forward applies it (ops.bn_affine): y = t*scale + shift"""
    scale = p[f"{pre}/bn_gamma"] / torch.sqrt(p[f"{pre}/bn_var"] + tf12.BN_EPS)
    shift = p[f"{pre}/bn_beta"] - p[f"{pre}/bn_mean"] * scale
    S[f"{pre}/bn_affine"] = (scale, shift)
    return t * scale + shift


def cbhg_saving(x, p, prefix, K, S):
    B, T, Cin = x.shape
    S[f"{prefix}/x_in"] = x
    bank = torch.cat([tf12.conv1d_same(x, p[f"{prefix}/bank/W{k}"], p[f"{prefix}/bank/b{k}"], torch.relu)
                      for k in range(1, K + 1)], -1)
    bank_bn = _bn(p, f"{prefix}/bank", bank, S)
    S[f"{prefix}/bank_bn"] = bank_bn
    pool = tf12.max_pool_2_1_same(bank_bn)
    S[f"{prefix}/bank_pool"] = pool
    p1 = _bn(p, f"{prefix}/proj1", tf12.conv1d_same(pool, p[f"{prefix}/proj1/W"], p[f"{prefix}/proj1/b"], torch.relu), S)
    S[f"{prefix}/proj1"] = p1
    res = _bn(p, f"{prefix}/proj2", tf12.conv1d_same(p1, p[f"{prefix}/proj2/W"], p[f"{prefix}/proj2/b"]), S) + x
    S[f"{prefix}/res"] = res
    h = res
    for l in range(4):
        pre = f"{prefix}/highway{l}"
        if f"{pre}/Wd" in p:
            h = tf12.dense(h, p[f"{pre}/Wd"], p[f"{pre}/bd"])
        S[f"{prefix}/hw{l}_in"] = h
        Pm = torch.cat([tf12.dense(h, p[f"{pre}/WH"], p[f"{pre}/bH"]), tf12.dense(h, p[f"{pre}/WT"], p[f"{pre}/bT"])], -1)
        S[f"{prefix}/hw{l}_P"] = Pm
        h = torch.relu(Pm[..., :128]) * torch.sigmoid(Pm[..., 128:]) + h * (1 - torch.sigmoid(Pm[..., 128:]))
    S[f"{prefix}/hw_out"] = h
    fw = tuple(p[f"{prefix}/gru_fw/{n}"] for n in ("Wg", "bg", "Wc", "bc"))
    bw = tuple(p[f"{prefix}/gru_bw/{n}"] for n in ("Wg", "bg", "Wc", "bc"))
    Wx = torch.cat([fw[0][:128], fw[2][:128], bw[0][:128], bw[2][:128]], 1)
    bx = torch.cat([fw[1], fw[3], bw[1], bw[3]])
    S[f"{prefix}/xp"] = h @ Wx + bx
    out = tf12.bidirectional_gru(h, fw, bw)
    S[f"{prefix}/gru_out"] = out
    return out


def decoder_saving(encoded, text_length, p, cfg, T, mel, drop_masks, sample_mask, S):
    """oracle decoder (teacher / scheduled sampling) that also records the three GRU state sequences [3,T,B,256]."""
    B, Tx, _ = encoded.shape
    dt = encoded.dtype
    U, mf, r = cfg.decoder_units, cfg.mel_features, cfg.r
    values, keys, mask = tf12.attention_prepare(encoded, text_length, p["dec/attn/W_mem"])
    h = [torch.zeros(B, U, dtype=dt) for _ in range(3)]
    attn = torch.zeros(B, cfg.attention_units, dtype=dt)
    x = mel[:, 0]
    ys, als = [], []
    Hs = torch.zeros(3, T, B, U, dtype=dt)
    for t in range(T):
        dm = (drop_masks[0][t], drop_masks[1][t])
        pn = O.pre_net(x[:, (r - 1) * mf:], p, "dec/prenet", cfg.audio_dropout_prob, dm)
        z = tf12.dense(torch.cat([pn, attn], -1), p["dec/in_proj/W"], p["dec/in_proj/b"])
        inp = z
        for i in range(3):
            h[i] = tf12.gru_cell(inp, h[i], p[f"dec/gru{i+1}/Wg"], p[f"dec/gru{i+1}/bg"], p[f"dec/gru{i+1}/Wc"], p[f"dec/gru{i+1}/bc"])
            Hs[i, t] = h[i]
            inp = h[i]
        y = tf12.dense(z + inp, p["dec/out_proj/W"], p["dec/out_proj/b"])
        a = tf12.bahdanau_alignments(y, keys, mask, p["dec/attn/W_q"], p["dec/attn/v"])
        ctx = torch.bmm(a[:, None, :], values)[:, 0]
        attn = torch.cat([y, ctx], -1) @ p["dec/attn/W_a"]
        ys.append(y); als.append(a)
        nxt = mel[:, t + 1] if t + 1 < T else torch.zeros_like(y)
        if sample_mask is not None:
            x = torch.where(sample_mask[t].to(torch.bool)[:, None], y, nxt)
        else:
            x = nxt
    S["dec/values"], S["dec/keys"] = values, keys
    S["dec/y"] = torch.stack(ys, 1).contiguous()
    S["dec/align"] = torch.stack(als, 1).contiguous()
    S["dec/H"] = Hs
    return S["dec/y"]


def saving_forward(p, inputs, cfg, enc_masks, dec_masks, sample_mask):
    """Returns (S, y, out) for the train-mode forward."""
    S = {"text": inputs["text"], "text_length": inputs["text_length"], "mel": inputs["mel"], "stft": inputs["stft"]}
    text = inputs["text"].to(torch.int64)
    ks = 1.0 / (1.0 - cfg.char_dropout_prob)
    t1 = torch.relu(p["embedding"] @ p["enc/prenet/W1"] + p["enc/prenet/b1"])
    l1 = t1[text] * enc_masks[0].to(t1.dtype) * ks
    l2 = torch.relu(l1 @ p["enc/prenet/W2"] + p["enc/prenet/b2"]) * enc_masks[1].to(t1.dtype) * ks
    S["enc/prenet/t1"], S["enc/prenet/l1"], S["enc/prenet/l2"] = t1, l1, l2
    encoded = cbhg_saving(l2, p, "enc/cbhg", cfg.enc_K, S)
    T = inputs["mel"].shape[1]
    S["dec/keep1"], S["dec/keep2"] = dec_masks
    if sample_mask is not None:
        S["dec/sample_mask"] = sample_mask
    y = decoder_saving(encoded, inputs["text_length"], p, cfg, T, inputs["mel"], dec_masks, sample_mask, S)
    B = y.shape[0]
    post = cbhg_saving(y.reshape(B, -1, cfg.mel_features), p, "post/cbhg", cfg.post_K, S)
    out = tf12.dense(post, p["post/dense/W"], p["post/dense/b"]).reshape(B, -1, cfg.fft_size * cfg.r)
    S["post/out"] = out
    return S, y, out
