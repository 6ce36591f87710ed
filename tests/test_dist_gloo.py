"""world_size-2 gloo test (CPU) of the multi-rank plumbing used by bench.py for N > 1: environment-driven init,
barrier, max-over-ranks timing, utterance sharding.  The data path itself has no collective (replicas only)."""
import os
import socket

import torch
import torch.multiprocessing as mp

from tacotron_b200.utils import dist as D


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    ws, r, _ = D.init("gloo")
    D.barrier()
    slow = D.max_over_ranks(10.0 + 5.0 * r)                 # rank 1 is the slow one
    lo, hi = D.shard_utterances(33, ws, r)
    out.put((r, ws, slow, lo, hi, D.aggregate_throughput(32000, ws, slow)))
    D.barrier()
    torch.distributed.destroy_process_group()


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(60); assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1] and all(r[1] == 2 for r in res)
    assert all(abs(r[2] - 15.0) < 1e-9 for r in res)         # both ranks see the max
    assert (res[0][3], res[0][4]) == (0, 17) and (res[1][3], res[1][4]) == (17, 33)   # disjoint cover of 33 utterances
    assert abs(res[0][5] - 2 * 32000 / 0.015) < 1e-3


def test_single_process_is_noop():
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        os.environ.pop(k, None)
    assert D.world() == (1, 0, 0)
    assert D.max_over_ranks(3.5) == 3.5
    assert D.shard_utterances(32, 1, 0) == (0, 32)
