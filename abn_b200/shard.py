"""Sharding of the candidate list over the GPUs of one box: one process per GPU.

Every (child, parent set) fit is independent (the reference already exploits this with
``foreach %dopar%``, R/build_score_cache_mle.R:636-681), so there is no collective inside the data
path: each rank holds the full data set, scores an interleaved share of the candidate list, and the
score vectors are gathered with ONE ``all_gather`` (NCCL over NVLink under ``torchrun``; gloo in the
CPU tests).

Partition: the list is cut into blocks of ``block`` consecutive candidates (neighbours share the
child and the leading parents, which is what gives column reuse in L1/L2) and block ``b`` goes to
rank ``b % world``.  The cache order groups candidates by child (= family) and parent count, so a
contiguous split would hand one GPU all the (free) Gaussian fits; interleaving gives every rank the
same mix.  Per-fit arithmetic does not depend on the partition, so the gathered cache is bit-identical
to the single-GPU one.

Result slab (per rank, on the device the kernels wrote it on): a ``(5, L)`` float64 tensor, 40 bytes per
candidate - rows 0..3 are mlik / aic / bic / mdl, row 4 reinterpreted as ``2 L`` int32 holds status
(first L) and iters (last L); ``L`` = blocks per rank (rounded up) x block.  ``abn_b200_score`` with
``ABN_FLAG_RESULT_ON_DEVICE`` writes straight into it, ``gather_slab`` all-gathers the slabs, puts the
blocks back in cache order with one permute on the device and makes one device-to-host copy.
"""
from __future__ import annotations

import numpy as np

FIELDS = ("mlik", "aic", "bic", "mdl")


def partition(rows: int, world: int, rank: int, block: int = 1024) -> np.ndarray:
    """Indices (ascending) of the candidates rank `rank` of `world` scores."""
    if world <= 1:
        return np.arange(rows, dtype=np.int64)
    nblocks = (rows + block - 1) // block
    mine = np.arange(rank, nblocks, world, dtype=np.int64)
    idx = (mine[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).ravel()
    return idx[idx < rows]


def slab_len(rows: int, world: int, block: int = 1024) -> int:
    """Length every rank pads its result slab to: blocks per rank (rounded up) x block."""
    if world <= 1:
        return rows
    nblocks = (rows + block - 1) // block
    return ((nblocks + world - 1) // world) * block


def new_slab(rows: int, world: int, block: int = 1024, device=None):
    """An empty result slab for one rank (torch tensor, float64 (5, L))."""
    import torch
    return torch.empty((5, slab_len(rows, world, block)), dtype=torch.float64, device=device)


def slab_from_dict(local: dict, rows: int, world: int, block: int = 1024, device=None):
    """Pack host results (dict of numpy arrays over this rank's share) into a slab - the host-staged path the
    CPU tests and callers without device-resident results use."""
    import torch
    slab = torch.zeros((5, slab_len(rows, world, block)), dtype=torch.float64)
    n = len(next(iter(local.values())))
    for f, key in enumerate(FIELDS):
        if key in local:
            slab[f, :n] = torch.from_numpy(np.ascontiguousarray(local[key], dtype=np.float64))
    ints = slab[4].view(torch.int32)
    L = slab.shape[1]
    if "status" in local:
        ints[:n] = torch.from_numpy(np.ascontiguousarray(local["status"], dtype=np.int32))
    if "iters" in local:
        ints[L:L + n] = torch.from_numpy(np.ascontiguousarray(local["iters"], dtype=np.int32))
    return slab.to(device) if device is not None else slab


_staging = {}


def _pinned_staging(rows: int, ring: int = 2):
    """Page-locked host buffers for the gathered results, (4, rows) float64 + (2, rows) int32, kept per size in a
    ring of `ring`: a device-to-host copy into fresh pageable memory costs 15-30 ms for 84 MB (page faults + staged
    DMA), into a resident pinned buffer 2 ms. The arrays gather_slab returns are views of these buffers and stay
    valid until `ring` further calls with the same `rows` have been made."""
    import torch
    e = _staging.get(rows)
    if e is None:
        if len(_staging) >= 4:
            _staging.clear()
        e = {"next": 0, "bufs": [(torch.empty((4, rows), dtype=torch.float64, pin_memory=True),
                                  torch.empty((2, rows), dtype=torch.int32, pin_memory=True)) for _ in range(ring)]}
        _staging[rows] = e
    b = e["bufs"][e["next"]]
    e["next"] = (e["next"] + 1) % len(e["bufs"])
    return b


def gather_slab(slab, rows: int, world: int, block: int = 1024, group=None, to_host: bool = True, pinned: bool = False):
    """ONE all_gather of the per-rank slabs + un-interleave on the slab's device.

    Returns a dict of arrays in cache order (numpy if to_host, else torch tensors on the device). pinned = True
    (CUDA slabs only): the host arrays are views of reused page-locked staging buffers, see _pinned_staging."""
    import torch
    import torch.distributed as dist

    L = slab.shape[1]
    if world > 1:
        flat = torch.empty(world * 5 * L, dtype=slab.dtype, device=slab.device)
        dist.all_gather_into_tensor(flat, slab.contiguous().reshape(-1), group=group)
        full = flat.reshape(world, 5, L)
    else:
        full = slab.reshape(1, 5, L)
    nb = L // block if world > 1 else 1
    blk = block if world > 1 else L
    # rank r, local block i  ->  global block i * world + r
    dbl = full[:, :4, :].reshape(world, 4, nb, blk).permute(1, 2, 0, 3).reshape(4, nb * world * blk)[:, :rows]
    ints = full[:, 4, :].contiguous().view(torch.int32).reshape(world, 2, nb, blk).permute(1, 2, 0, 3)
    ints = ints.reshape(2, nb * world * blk)[:, :rows]
    if not to_host:
        out = {k: dbl[f] for f, k in enumerate(FIELDS)}
        out["status"], out["iters"] = ints[0], ints[1]
        return out
    if pinned and slab.is_cuda:
        dbl_p, ints_p = _pinned_staging(rows)  # (an existing ring is reused whatever its length)
        dbl_p.copy_(dbl, non_blocking=True)
        ints_p.copy_(ints, non_blocking=True)
        torch.cuda.current_stream(slab.device).synchronize()
        dbl_h, ints_h = dbl_p.numpy(), ints_p.numpy()
    else:
        dbl_h = dbl.contiguous().cpu().numpy()
        ints_h = ints.contiguous().cpu().numpy()
    out = {k: dbl_h[f] for f, k in enumerate(FIELDS)}
    out["status"], out["iters"] = ints_h[0], ints_h[1]
    return out


def gather(local: dict, rows: int, world: int, rank: int, block: int = 1024, device=None, group=None) -> dict:
    """All-gather host results (dict of numpy arrays over this rank's share) and put them back in cache
    order. One collective for all fields; returns only the keys that were passed in."""
    if world <= 1:
        return {k: np.asarray(v) for k, v in local.items()}
    slab = slab_from_dict(local, rows, world, block, device)
    out = gather_slab(slab, rows, world, block, group=group)
    res = {}
    for k, v in local.items():
        res[k] = out[k].astype(np.asarray(v).dtype, copy=False)
    return res


def score_sharded(score_fn, children0: np.ndarray, parents: np.ndarray, world: int, rank: int, block: int = 1024,
                  device=None, group=None) -> dict:
    """Score this rank's share with `score_fn(children0, parents) -> dict of arrays` and gather."""
    rows = len(children0)
    idx = partition(rows, world, rank, block)
    local = score_fn(children0[idx], parents[idx])
    return gather(local, rows, world, rank, block, device=device, group=group)
