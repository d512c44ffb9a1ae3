"""Multi-GPU execution: problems are independent, so the batch is cut into contiguous shards, one process per GPU
(``torch.distributed``), and the only communication is the final gather of the results.

Nothing here touches the data path: there is no halo, no reduction, no model state.  ``backend="nccl"`` on GPUs (NVLink /
NVSwitch), ``"gloo"`` in the CPU tests, where the per-shard solve is injected.
"""

from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from .pack import ProblemBatch


def shard_bounds(B: int, world_size: int, rank: int):
    """Contiguous static partition [lo, hi) of B problems — cost is uniform within a configuration (L N^3 each)."""
    base, rem = divmod(B, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: ProblemBatch, world_size: int, rank: int) -> ProblemBatch:
    lo, hi = shard_bounds(batch.B, world_size, rank)
    return batch.subset(slice(lo, hi))


def solve_sharded(batch: ProblemBatch, rtsolver_options: Optional[dict] = None, *,
                  solve_fn: Optional[Callable] = None, gather: bool = True):
    """Solve this rank's shard and (optionally) all-gather the values so that every rank holds the full result.

    Must be called by every rank of an initialised ``torch.distributed`` process group.  ``solve_fn(shard, options)``
    returns the HostOutputs of a shard; the default is the CUDA path on ``cuda:LOCAL_RANK``.
    """
    import torch
    import torch.distributed as dist

    import os

    rank, world = dist.get_rank(), dist.get_world_size()
    shard = shard_batch(batch, world, rank)
    device = int(os.environ.get("LOCAL_RANK", rank))  # the GPU this rank solves on AND stages the gather on
    if solve_fn is None:
        from .model import solve_batch

        out = solve_batch(shard, rtsolver_options, device=device)
    else:
        out = solve_fn(shard, rtsolver_options)
    if not gather:
        return out.values, out.status
    # ragged all_gather: pad every shard to the largest one
    sizes = [shard_bounds(batch.B, world, r) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    tail = out.values.shape[1:]
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", device) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)  # NCCL picks the current device: one GPU per rank
    buf = torch.zeros((nmax,) + tail, dtype=torch.float64, device=dev)
    buf[:out.values.shape[0]] = torch.from_numpy(out.values).to(dev)
    st = torch.zeros(nmax, dtype=torch.int32, device=dev)
    st[:len(out.status)] = torch.from_numpy(out.status).to(dev)
    vals = [torch.empty_like(buf) for _ in range(world)]
    sts = [torch.empty_like(st) for _ in range(world)]
    dist.all_gather(vals, buf)
    dist.all_gather(sts, st)
    values = np.concatenate([v[:hi - lo].cpu().numpy() for v, (lo, hi) in zip(vals, sizes)], axis=0)
    status = np.concatenate([s[:hi - lo].cpu().numpy() for s, (lo, hi) in zip(sts, sizes)], axis=0)
    return values, status
