"""Data-parallel training step, world_size 2 over gloo on CPU.

Each rank runs the hand-written backward (tacotron_b200/models/grad.py over the CPU mirror kernels) on ITS half of the
batch, the flat gradient bucket is all-reduced with SUM (tacotron_b200/optim.py -- the same code path NCCL takes on
GPUs), then global-norm clip + Adam.  Because the loss is a sum over utterances the result must equal the
single-process step on the whole batch, and both ranks must end with identical parameters."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class _Layout:
    """offsets/shapes of a flat bucket (the part of ParamStore that FlatAdam.views reads)"""
    def __init__(self, params):
        self.offsets, self.shapes = {}, {}
        off = 0
        for n, v in params.items():
            self.offsets[n] = off
            self.shapes[n] = (tuple(v.shape), None)
            off += (v.numel() + 3) // 4 * 4
        self.total = off


def _step(rank, world, lo, hi):
    """one training step on utterances [lo, hi) of the fixed synthetic batch; returns the updated flat parameters"""
    from oracle import tacotron_oracle as O
    from tacotron_b200.models import grad
    from tacotron_b200.optim import FlatAdam
    from tests import grad_util, mirror_kernels as MK
    torch.set_num_threads(2)
    dt = torch.float64
    B, Tx, T = 2, 8, 5
    cfg = O.OracleConfig(r=2, vocab_size=20)
    p = O.init_params(cfg, seed=1, trained_like=True, dtype=dt)
    inp = O.synthetic_inputs(cfg, B, Tx, T, seed=0, ragged=True)
    inp = {k: (v.to(dt) if v.dtype.is_floating_point else v) for k, v in inp.items()}
    enc_m, dec_m = O.dropout_masks(cfg, B, Tx, T, seed=2)
    sm = O.sched_mask(cfg, B, T, seed=3)
    inp = {k: v[lo:hi].contiguous() for k, v in inp.items()}
    enc_m = tuple(m[lo:hi].contiguous() for m in enc_m)
    dec_m = tuple(m[:, lo:hi].contiguous() for m in dec_m)
    sm = sm[:, lo:hi].contiguous()
    lay = _Layout(p)
    flat = torch.zeros(lay.total, dtype=dt)
    views = {}
    for n, v in p.items():
        views[n] = flat[lay.offsets[n]:lay.offsets[n] + v.numel()].view(v.shape)
        views[n].copy_(v)
    opt = FlatAdam(flat)
    G = opt.views(lay)
    S, _, _ = grad_util.saving_forward(views, inp, cfg, enc_m, dec_m, sm)
    opt.zero_grad()
    grad.model_bwd(MK, views, G, S, cfg)
    opt.apply(MK, flat, 1e-3, float(cfg.cap_grads))
    return flat, float(opt.sumsq[0])


def _worker(rank, world, port, q):
    os.environ.update(WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat, ss = _step(rank, world, rank, rank + 1)
    q.put((rank, flat.numpy(), ss))
    dist.barrier()
    dist.destroy_process_group()


def test_dp2_equals_single_process_step():
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        os.environ.pop(k, None)
    ref, ss_ref = _step(0, 1, 0, 2)                      # single process, whole batch, no process group
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in ps:
        p.join(120); assert p.exitcode == 0
    f0, f1 = torch.from_numpy(res[0][1]), torch.from_numpy(res[1][1])
    assert torch.equal(f0, f1), "ranks diverged after the replicated Adam step"
    assert abs(res[0][2] - ss_ref) <= 1e-9 * ss_ref, (res[0][2], ss_ref)      # global norm of the reduced gradient
    assert torch.allclose(f0, ref, rtol=0, atol=1e-12), (f0 - ref).abs().max()
