"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the exchange.

The hot path shards without any data-path collective: every rank holds the whole
forcing table and evaluates a slice of the parameter sets; the only exchange is an all-gather of
the objective values (8 bytes per set) -- NCCL on GPUs, gloo in the CPU tests.  Replicating the table
itself is the one place an exchange pays: N ranks each pushing all of it through the host's PCIe
complex is slower than each pushing 1/N and all-gathering over NVLink (replicate_forcing).
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


def replicate_forcing(P, E, device=None):
    """Whole [ncell, ntime] tables on every rank's device: each rank uploads the rows of ncell/world
    subbasins (from pinned host memory if P, E are pinned torch tensors) and the ranks all-gather them.
    Returns torch tensors (P_dev, E_dev); without a process group it is a plain upload."""
    import torch
    import torch.distributed as dist
    P, E = torch.as_tensor(P), torch.as_tensor(E)
    ncell, ntime = P.shape
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        dev = device or ("cuda" if torch.cuda.is_available() else "cpu")
        return P.to(dev, non_blocking=True), E.to(dev, non_blocking=True)
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")
    m = -(-ncell // world)                                # rows per rank; the last slice is padded
    lo, hi = min(rank * m, ncell), min((rank + 1) * m, ncell)
    out = []
    for X in (P, E):
        mine = torch.zeros((m, ntime), dtype=torch.float64, device=dev)
        mine[:hi - lo].copy_(X[lo:hi], non_blocking=True)
        full = torch.empty((m * world, ntime), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(full, mine)
        out.append(full[:ncell])
    return out[0], out[1]


def shard_spec():
    """(rank, world, allgather) for api.Optim_GR2MSemiDistr(shard=...), or None when not distributed."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    return dist.get_rank(), dist.get_world_size(), make_allgather()


__all__ = ["shard_range", "split_counts", "make_allgather", "shard_spec", "replicate_forcing"]
