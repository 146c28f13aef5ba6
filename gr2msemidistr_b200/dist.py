"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the exchange.

The hot path shards without any data-path collective: every rank holds the whole
forcing table and evaluates a slice of the parameter sets; the only exchange is an all-gather of
the objective values (8 bytes per set) -- NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np

from .sceua import split_counts


def shard_range(n: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of n items owned by `rank`."""
    return n * rank // world, n * (rank + 1) // world


def make_allgather(device=None):
    """allgather(local_values, counts) -> concatenated float64 vector, over the default process group."""
    import torch
    import torch.distributed as dist

    def allgather(local, counts):
        dev = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")
        m = max(counts)                                   # ragged slices are padded to a common length
        mine = torch.zeros(m, dtype=torch.float64, device=dev)
        mine[:len(local)] = torch.as_tensor(np.ascontiguousarray(local, np.float64), device=dev)
        out = torch.empty(m * len(counts), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, mine)
        out = out.view(len(counts), m).cpu().numpy()
        return np.concatenate([out[r, :c] for r, c in enumerate(counts)])

    return allgather


def shard_spec():
    """(rank, world, allgather) for api.Optim_GR2MSemiDistr(shard=...), or None when not distributed."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    return dist.get_rank(), dist.get_world_size(), make_allgather()


__all__ = ["shard_range", "split_counts", "make_allgather", "shard_spec"]
