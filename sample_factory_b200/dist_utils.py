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


# loss-statistics rows (ops.LS): how each column combines across ranks.  Every rank's row holds rank-local sums divided by
# the GLOBAL valid count, so the means add up; num_valid / adv_mean / adv_std are already global; extrema take max / min;
# value_mean is a mean over the rank's whole minibatch (valid or not; equal sizes) -> mean over the ranks.
def _ls_masks():
    from . import ops

    keep = sum(1 << ops.LS[k] for k in ("num_valid", "adv_mean", "adv_std"))
    mx = sum(1 << ops.LS[k] for k in ("kl_old_max", "ratio_max"))
    mn = 1 << ops.LS["ratio_min"]
    avg = 1 << ops.LS["value_mean"]
    return keep, mx, mn, avg


class PeerComm:
    """One NVLink peer-memory communicator (csrc/comm.cu): a zero-filled comm buffer in this rank's HBM
    [header | fp64 scratch | flat fp32 gradient], exported through CUDA IPC and mapped by every peer.  torch.distributed
    is used once, to exchange the 64-byte handles; the collectives themselves are libsfb200 kernels."""

    SCRATCH_BYTES = 1 << 20

    def __init__(self, device: torch.device, grad_numel: int, group=None):
        from . import ops

        assert dist.is_initialized()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if "expandable_segments:True" in os.environ.get("PYTORCH_CUDA_ALLOC_CONF", ""):
            raise RuntimeError("PeerComm needs cudaMalloc-backed allocations (CUDA IPC): disable expandable_segments")
        self.header = ops.dp_header_bytes()
        total = self.header + self.SCRATCH_BYTES + 4 * grad_numel
        # an allocation of its own (>= 20 MB blocks are never shared with other tensors by the caching allocator)
        self.buf = torch.zeros(max(total, 21 << 20), dtype=torch.uint8, device=device)
        torch.cuda.synchronize(device)
        handle, offset = ops.ipc_export(self.buf)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (handle, offset), group=group)
        self._imported = []
        ptrs = []
        for r, (h, off) in enumerate(gathered):
            if r == self.rank:
                ptrs.append(self.buf.data_ptr())
            else:
                p = ops.ipc_import(h, off)
                self._imported.append((p, off))
                ptrs.append(p)
        self.comm = ops.dp_create(self.rank, self.world, ptrs, self.SCRATCH_BYTES)
        g0 = self.header + self.SCRATCH_BYTES
        self.grad = self.buf[g0: g0 + 4 * grad_numel].view(torch.float32)
        self.workspace = torch.zeros(1024, dtype=torch.float32, device=device)
        dist.barrier(group=group)      # every rank's buffer is zeroed and mapped before the first collective

    def close(self) -> None:
        from . import ops

        if getattr(self, "comm", None) is not None:
            ops.dp_destroy(self.comm)
            for p, off in self._imported:
                ops.ipc_close(p, off)
            self.comm = None


def peer_comm_wanted() -> bool:
    """SFB200_DP_COMM=nccl keeps every exchange on torch.distributed (NCCL); the default is the NVLink peer kernels."""
    return os.environ.get("SFB200_DP_COMM", "peer") != "nccl"
