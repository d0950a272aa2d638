"""Clip-level sharding across GPUs.

Every recording's 25 features depend on that recording only (mshds_extractor.py:408-448 has no cross-file state), so the
path partitions into independent units: one process per GPU, greedy longest-processing-time assignment of clips by
duration, no data-path collective; only the [n_local, 25] float64 feature matrix (200 B per clip) is gathered.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np


def lpt_assign(lengths: Sequence[int], world_size: int) -> list[list[int]]:
    """Greedy LPT: clips sorted by length (desc, stable) go to the currently least-loaded rank."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    parts: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += int(lengths[i])
    for p in parts:
        p.sort()
    return parts


def pack_subset(pcm: np.ndarray, offsets: np.ndarray, idx: Sequence[int]):
    """Packed (pcm, offsets) of the chosen clips, in the given order."""
    chunks = [pcm[offsets[i]: offsets[i + 1]] for i in idx]
    offs = np.cumsum([0] + [len(c) for c in chunks]).astype(np.int64)
    data = np.concatenate(chunks) if chunks else np.zeros(0, dtype=pcm.dtype)
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
