"""Data-parallel plumbing (new functionality: the reference has no collective at all).

One process per GPU; env shards are independent, so the only cross-rank traffic per SGD step is
  * one all-reduce (sum) of the flat gradient buffer,
  * three doubles of advantage statistics, and per training iteration
  * the normalizer batch moments and the valid-sample count,
all through torch.distributed (NCCL on GPUs, gloo in the CPU tests).  The functions here are the pure host-side
logic: they only need tensors and a process group, so the CPU test-suite exercises them with gloo, world_size 2.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment. Returns (rank, local_rank, world_size)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def world_size(group=None) -> int:
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def allreduce_sum_(t: Tensor, group=None) -> Tensor:
    if world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def pooled_moments_(batch_mean: Tensor, batch_var: Tensor, rows_per_rank: int, group=None) -> int:
    """Turn per-rank (mean, UNBIASED var) over `rows_per_rank` rows into the moments of the concatenation of all
    ranks' rows (exact pooled formula: sum and centred second moment add up), in place. Returns the global row count.

    This is what makes G ranks x N envs update the running normalizers (running_mean_std.py:72-77) exactly like one
    process holding all G*N envs would."""
    g = world_size(group)
    if g == 1:
        return rows_per_rank
    n = float(rows_per_rank)
    m = batch_mean.double()
    pack = torch.stack([m * n, batch_var.double() * (n - 1.0) + m * m * n])   # [sum x, sum x^2]
    allreduce_sum_(pack, group)
    total = rows_per_rank * g
    gmean = pack[0] / total
    gm2 = pack[1] - gmean * gmean * total
    batch_mean.copy_(gmean.to(batch_mean.dtype))
    batch_var.copy_((gm2 / (total - 1.0)).to(batch_var.dtype))
    return total
