"""Partition of lookup (or particle) ids over ranks, and the all-reduce of the result -- the host
logic of the one-process-per-GPU path (bench.py under torchrun; tests/test_multirank_cpu.py runs the
same code on CPU with gloo).

Lookup i depends only on i (cuda/Simulation.cu:53-56: seed = fast_forward_LCG(1070, 2*i)) and
history particle p only on p (openmp-threading/Simulation.c:167), so any partition of the id
range is exact: the verification sums of the parts add up to the sum of the whole.  The grid
is replicated; the only exchange is the sum of {verification, n_lookups}.
"""
from __future__ import annotations

from typing import Tuple


def strong_shard(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of [0, total) into `world` near-equal ranges -> (first_id, count)."""
    if world < 1 or not 0 <= rank < world or total < 0:
        raise ValueError("bad shard request")
    lo = total * rank // world
    hi = total * (rank + 1) // world
    return lo, hi - lo


def weak_shard(per_rank: int, rank: int, world: int) -> Tuple[int, int]:
    """Every rank owns `per_rank` distinct ids: rank r gets [r*per_rank, (r+1)*per_rank)."""
    if world < 1 or not 0 <= rank < world or per_rank < 0:
        raise ValueError("bad shard request")
    return rank * per_rank, per_rank


def allreduce_result(verification: int, n_lookups: int, device=None) -> Tuple[int, int]:
    """Sum {verification, n_lookups} over the default torch.distributed group (NCCL on GPUs,
    gloo on CPU).  Without an initialised group this is the identity."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return verification, n_lookups
    t = torch.tensor([verification, n_lookups], dtype=torch.int64, device=device)
    dist.all_reduce(t)
    return int(t[0].item()), int(t[1].item())


class ResultReducer:
    """The per-step collective of the multi-rank path, without a host round trip per step: the
    rank's {verification, n_lookups} go host (pinned) -> device asynchronously and are summed over
    the default group with one all-reduce (NCCL on GPUs, gloo on CPU) that stays enqueued on the
    current stream; `result()` is the only call that waits.  Without a process group it is a plain
    holder, so single-rank callers use the same code path."""

    def __init__(self, device=None):
        import torch
        self._torch = torch
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.accum = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.pair = torch.zeros(2, dtype=torch.int64)
        if self.device.type == "cuda":
            self.pair = self.pair.pin_memory()

    def submit(self, verification: int, n_lookups: int) -> None:
        import torch.distributed as dist
        self.pair[0], self.pair[1] = verification, n_lookups
        self.accum.copy_(self.pair, non_blocking=True)
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(self.accum)

    def result(self) -> Tuple[int, int]:
        return int(self.accum[0].item()), int(self.accum[1].item())
