"""Optimizer step of Tacotron.add_train_op (models/tacotron.py:167-185) on the flat parameter bucket:

    [data parallel] all-reduce(SUM) of the flat gradient     (the loss is a SUM over utterances
                                                               (tacotron.py:158-160), so SUM over ranks reproduces the
                                                               single-process step on the concatenated batch)
    global norm -> clip_by_global_norm(cap_grads)              tacotron.py:182
    TF Adam (beta1 .9, beta2 .999, eps 1e-8, lr fed per step)  tacotron.py:170-184, SURVEY A.12

Parameters, gradients and both moments are single flat fp32 buffers (tacotron_b200/params.py), so the exchange is
ONE NCCL all-reduce of 28.4 MB over NVLink and the update is one fused kernel; the clip scale is read from a device
scalar, so the step has no host synchronisation.  All arithmetic goes through the kernel namespace `K`
(tacotron_b200/kernels.py); this file only sequences it, which is why the same code runs over the CPU mirror with the
gloo backend in tests/test_dp_gloo.py.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

BETA1, BETA2, EPSILON = 0.9, 0.999, 1e-8


class FlatAdam:
    def __init__(self, flat):
        self.g = torch.zeros_like(flat)
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self.sumsq = torch.zeros(1, dtype=flat.dtype, device=flat.device)
        self.step = 0

    def views(self, store_like):
        """name -> view of the flat gradient with the layout of `store_like` (offsets + shapes)."""
        out = {}
        for n, o in store_like.offsets.items():
            shape = store_like.shapes[n][0]
            numel = 1
            for s in shape:
                numel *= s
            out[n] = self.g[o:o + numel].view(shape)
        return out

    def zero_grad(self):
        self.g.zero_()

    def apply(self, K, flat, lr, clip, allreduce=True):
        """flat <- Adam(flat, clip(all_reduce(g)));  returns the device scalar holding sum(g^2) (pre-clip)."""
        if allreduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.g, op=dist.ReduceOp.SUM)
        K.sumsq(self.sumsq, self.g)
        self.step += 1
        lr_t = lr * math.sqrt(1.0 - BETA2 ** self.step) / (1.0 - BETA1 ** self.step)
        K.adam_step(flat, self.g, self.m, self.v, lr_t, BETA1, BETA2, EPSILON, float(clip), self.sumsq)
        return self.sumsq
