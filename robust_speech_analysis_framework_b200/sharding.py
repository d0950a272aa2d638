"""Clip-level sharding across GPUs.

Every recording's 25 features depend on that recording only (mshds_extractor.py:408-448 has no cross-file state), so the
path partitions into independent units: one process per GPU, greedy longest-processing-time assignment of clips by
duration, no data-path collective; only the [n_local, 25] float64 feature matrix (200 B per clip) is gathered.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np


def lpt_assign(lengths: Sequence[int], world_size: int) -> list[list[int]]:
    """Greedy LPT: clips sorted by length (desc, stable) go to the currently least-loaded rank (ties: lowest rank).  Clips of
    equal length are interchangeable, so within every group of equal lengths the indices are then dealt out in input order,
    rank by rank, keeping every rank's count of that length: same loads, but runs of consecutive clips stay together (a corpus
    of equal-length clips shards into contiguous blocks that pack without a copy)."""
    import heapq

    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    order = np.lexsort((np.arange(n), -lengths))                  # by length desc, then index
    heap = [(0, r) for r in range(world_size)]
    owner = np.empty(n, dtype=np.int64)
    for i in order:
        load, r = heapq.heappop(heap)
        owner[i] = r
        heapq.heappush(heap, (load + int(lengths[i]), r))
    # regroup equal lengths: order[] lists every length group contiguously, indices ascending inside a group
    sorted_len = lengths[order]
    bounds = np.flatnonzero(np.diff(sorted_len)) + 1
    for g in np.split(order, bounds):
        if len(g) < 2:
            continue
        counts = np.bincount(owner[g], minlength=world_size)
        owner[g] = np.repeat(np.arange(world_size), counts)
    return [np.flatnonzero(owner == r).tolist() for r in range(world_size)]


def pack_subset(pcm: np.ndarray, offsets: np.ndarray, idx: Sequence[int]):
    """Packed (pcm, offsets) of the chosen clips, in the given order.  Runs of consecutive clips are copied as one slice; a
    single run is returned as a view of `pcm` (no copy)."""
    idx = np.asarray(idx, dtype=np.int64)
    if len(idx) == 0:
        return np.zeros(0, dtype=pcm.dtype), np.zeros(1, dtype=np.int64)
    offsets = np.asarray(offsets, dtype=np.int64)
    lens = offsets[idx + 1] - offsets[idx]
    offs = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    breaks = np.flatnonzero(np.diff(idx) != 1) + 1
    starts = np.concatenate(([0], breaks))
    ends = np.concatenate((breaks, [len(idx)]))
    if len(starts) == 1:
        return pcm[offsets[idx[0]]: offsets[idx[-1] + 1]], offs
    data = np.concatenate([pcm[offsets[idx[a]]: offsets[idx[b - 1] + 1]] for a, b in zip(starts, ends)])
    return data, offs


def extract_sharded(pcm: np.ndarray, offsets: np.ndarray, compute: Callable[[np.ndarray, np.ndarray], tuple],
                    rank: int, world_size: int, gather: bool = True, dst: int = 0):
    """Each rank computes its LPT share with `compute(pcm, offsets) -> (features [m,25], status [m])`; rank `dst` gets the
    full matrices in input order (other ranks get None).  Uses torch.distributed when world_size > 1."""
    n = len(offsets) - 1
    lengths = np.diff(offsets)
    parts = lpt_assign(lengths, world_size)
    mine = parts[rank]
    sub_pcm, sub_off = pack_subset(pcm, offsets, mine)
    feats, status = compute(sub_pcm, sub_off) if mine else (np.zeros((0, 25)), np.zeros(0, np.uint32))
    feats = np.asarray(feats, dtype=np.float64).reshape(len(mine), 25)
    status = np.asarray(status, dtype=np.uint32).reshape(len(mine))
    if world_size == 1 or not gather:
        out = np.full((n, 25), np.nan)
        st = np.zeros(n, np.uint32)
        out[mine] = feats
        st[mine] = status
        return out, st
    import torch
    import torch.distributed as dist

    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    maxn = max(len(p) for p in parts)
    buf = torch.full((maxn, 26), float("nan"), dtype=torch.float64)
    if mine:
        buf[: len(mine), :25] = torch.from_numpy(feats)
        buf[: len(mine), 25] = torch.from_numpy(status.astype(np.float64))
    buf = buf.to(dev)
    gathered = [torch.empty_like(buf) for _ in range(world_size)]
    dist.all_gather(gathered, buf)
    if rank != dst:
        return None, None
    out = np.full((n, 25), np.nan)
    st = np.zeros(n, np.uint32)
    for r in range(world_size):
        g = gathered[r].cpu().numpy()
        k = len(parts[r])
        if k:
            out[parts[r]] = g[:k, :25]
            st[parts[r]] = g[:k, 25].astype(np.uint32)
    return out, st
