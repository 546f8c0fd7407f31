"""Sharding of the candidate list over the GPUs of one box: one process per GPU.

Every (child, parent set) fit is independent (the reference already exploits this with
``foreach %dopar%``, R/build_score_cache_mle.R:636-681), so there is no collective inside the data
path: each rank holds the full data set, scores an interleaved share of the candidate list, and the
score vectors are gathered once with ``all_gather`` (NCCL over NVLink under ``torchrun``; gloo in
the CPU tests).

Partition: the list is cut into blocks of ``block`` consecutive candidates (neighbours share the
child and the leading parents, which is what gives column reuse in L1/L2) and block ``b`` goes to
rank ``b % world``.  The cache order groups candidates by child (= family) and parent count, so a
contiguous split would hand one GPU all the (free) Gaussian fits; interleaving gives every rank the
same mix.  Per-fit arithmetic does not depend on the partition, so the gathered cache is bit-identical
to the single-GPU one.
"""
from __future__ import annotations

import numpy as np


def partition(rows: int, world: int, rank: int, block: int = 1024) -> np.ndarray:
    """Indices (ascending) of the candidates rank `rank` of `world` scores."""
    if world <= 1:
        return np.arange(rows, dtype=np.int64)
    nblocks = (rows + block - 1) // block
    mine = np.arange(rank, nblocks, world, dtype=np.int64)
    idx = (mine[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).ravel()
    return idx[idx < rows]


def slab_len(rows: int, world: int, block: int = 1024) -> int:
    """Length every rank pads its result slab to (the largest share)."""
    return max(len(partition(rows, world, r, block)) for r in range(world)) if world > 1 else rows


def gather(local: dict, rows: int, world: int, rank: int, block: int = 1024, device=None, group=None) -> dict:
    """All-gather the per-rank result slabs and put them back in cache order.

    `local` maps field name -> 1-D numpy array over this rank's share (float64 or int32).
    Uses ``torch.distributed`` (must be initialised when world > 1); tensors travel on `device`
    ("cuda:N" for NCCL, None / "cpu" for gloo).
    """
    if world <= 1:
        return {k: np.asarray(v) for k, v in local.items()}
    import torch
    import torch.distributed as dist

    L = slab_len(rows, world, block)
    out = {}
    for key in sorted(local):
        v = np.asarray(local[key])
        dtype = torch.float64 if v.dtype.kind == "f" else torch.int32
        slab = torch.zeros(L, dtype=dtype, device=device)
        slab[:len(v)] = torch.from_numpy(np.ascontiguousarray(v)).to(dtype).to(slab.device)
        full = torch.empty(world * L, dtype=dtype, device=device)
        dist.all_gather_into_tensor(full, slab, group=group)
        full = full.cpu().numpy().reshape(world, L)
        res = np.empty(rows, dtype=v.dtype)
        for r in range(world):
            idx = partition(rows, world, r, block)
            res[idx] = full[r, :len(idx)]
        out[key] = res
    return out


def score_sharded(score_fn, children0: np.ndarray, parents: np.ndarray, world: int, rank: int, block: int = 1024,
                  device=None, group=None) -> dict:
    """Score this rank's share with `score_fn(children0, parents) -> dict of arrays` and gather."""
    rows = len(children0)
    idx = partition(rows, world, rank, block)
    local = score_fn(children0[idx], parents[idx])
    return gather(local, rows, world, rank, block, device=device, group=group)
