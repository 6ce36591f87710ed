"""Multi-GPU plumbing for the replicated inference path (inference shards by utterance with
no data-path collective).  One process per GPU; the only cross-rank traffic is the timing reduction of the
benchmark harness (max over ranks) and a start/stop barrier.  Works with NCCL on GPUs and gloo on CPU (tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    ws, rank, local = world()
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return ws, rank, local


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(value: float) -> float:
    """Whole-job time of a step = the slowest rank's device time."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shard_utterances(n_utterances: int, world_size: int, rank: int):
    """Contiguous utterance shard [lo, hi) of this rank (utterances are independent in every kernel)."""
    base, rem = divmod(n_utterances, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def aggregate_throughput(units_per_rank_step: int, world_size: int, ms_per_step_max: float) -> float:
    """Weak-scaling aggregate: every rank processes `units_per_rank_step`; the job advances at the slowest rank."""
    return world_size * units_per_rank_step / (ms_per_step_max / 1e3)
