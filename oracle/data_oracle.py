"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  SURVEY.md section 8(f) rank 3: the data formats on the input side of the path.

numpy restatement of the arithmetic in the reference's ``data_input.load_from_npy`` (data_input.py:42-85) and the
dtype handling of ``build_dataset`` (data_input.py:20-40): spectrograms are STORED as float16 (preprocess.py:179-180),
normalised IN PLACE in that dtype with statistics of a 100-utterance sample, and cast to float32 when batched.

    mean = np.mean(x[index], axis=(0,1))                 float16 result (numpy keeps the input dtype)      :56-57
    std  = np.std(x[index], axis=(0,1), dtype=float32)                                                    :58-59
    x -= mean   -> float16( float32(x) - float32(mean) )                                                   :61-62
    x /= std    -> float16( float32(x) / std )                                                             :63-64
    batch = tf.cast(x, tf.float32)                                                                         :38-39

Bit-exact target for the device kernel taco_normalize_f16 (tacotron_b200/data_input.py).  Pinned twice
(tests/test_data_input.py): the per-element restatement equals the same numpy in-place statements the reference makes,
and it equals the arrays the reference's OWN data_input.load_from_npy returns (executed unmodified from /root/reference
by tests/golden/make_golden.py -> tests/golden/reference_data_input.npz), statistics and dtypes included.
"""
from __future__ import annotations

import numpy as np


def sample_stats(x, index):
    """x float16 [N,T,W]; index int [100] -> (mean float16 [W], std float32 [W])   (data_input.py:56-59)"""
    mean = np.mean(x[index], axis=(0, 1))
    std = np.std(x[index], axis=(0, 1), dtype=np.float32)
    return mean, std


def normalize_inplace(x, mean, std):
    """the reference's two in-place statements, verbatim semantics (data_input.py:61-64)"""
    x -= mean
    x /= std
    return x


def normalize_explicit(x, mean, std):
    """the same arithmetic spelled out per element: what the device kernel implements"""
    d = (x.astype(np.float32) - mean.astype(np.float32)).astype(np.float16)
    q = (d.astype(np.float32) / std.astype(np.float32)).astype(np.float16)
    return q.astype(np.float32)                                   # tf.cast(.., tf.float32), data_input.py:38-39


def encode_prompts(prompts, ivocab, max_text_len=140):
    """data_input.load_prompts (data_input.py:92-99): characters not in the vocabulary are dropped from the id sequence,
    but text_length is the length of the RAW prompt line (newline and dropped characters included) -- kept as is."""
    vocab = {v: k for k, v in ivocab.items()}
    text = [[vocab[w] for w in p.strip() if w in vocab] for p in prompts]
    text_length = np.array([len(p) for p in prompts], dtype=np.int32)
    out = np.zeros((len(prompts), max_text_len), dtype=np.int32)
    for i, t in enumerate(text):
        out[i, :len(t)] = t
    return out, text_length
