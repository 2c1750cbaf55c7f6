"""Multi-GPU host logic: the refinement path shards by batch of (source, driving) pairs with no
data-path collective; the only collective is an all-reduce of the
reconstruction-L1 / timing statistics, mirroring the loss reduce of train.py:74-77 and the L1 of
reconstruction.py:68.  One process per GPU (NCCL); the same code runs over gloo in CPU tests.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kwargs)
    return rank, world, local


def shard_range(total_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `total_pairs` frame pairs: rank r takes [start, end).
    The first `total_pairs % world` ranks get one extra pair; empty shards are legal."""
    base, extra = divmod(total_pairs, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def reduce_stats(l1_sum: float, n_elements: float, elapsed_s: float, n_pairs: float, device=None,
                 extra_max=()) -> dict:
    """All-reduce {sum|out - driving|, element count, pair count} (SUM) and elapsed seconds (MAX);
    `extra_max`: further per-rank timings reduced with MAX alongside `elapsed_s`."""
    device = device or ("cuda" if (dist.is_initialized() and dist.get_backend() == "nccl") else "cpu")
    sums = torch.tensor([l1_sum, n_elements, n_pairs], dtype=torch.float64, device=device)
    tmax = torch.tensor([elapsed_s, *extra_max], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    l1, n, pairs = (float(x) for x in sums)
    t = float(tmax[0])
    return {"l1_mean": l1 / max(n, 1.0), "pairs": pairs, "elapsed_s": t, "pairs_per_s": pairs / t if t > 0 else 0.0,
            "extra_max": [float(x) for x in tmax[1:]]}
