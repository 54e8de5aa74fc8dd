"""Multi-GPU plumbing: independent stereo pairs shard across ranks (one process per GPU).

The algorithm has no exchange step, so the only collectives are the work split
(rank 0 -> everyone), the result gather (everyone -> rank 0) and a max-reduction of timings.
`torch.distributed` provides them (NCCL on GPUs, gloo in the CPU tests); the per-rank compute is
whatever `match_fn` does -- in production `Engine.match_batch`.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np


def shard_bounds(n_pairs: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous blocks, sizes differing by at most one: rank r owns [lo, hi)."""
    base, extra = divmod(n_pairs, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def run_sharded(match_fn: Callable[[np.ndarray, np.ndarray], np.ndarray], lefts, rights, height: int, width: int,
                device=None):
    """Rank 0 passes the full batch (lefts/rights [n][H][W][3] uint8); other ranks pass None.
    Returns the [n][H][W] float32 disparity maps on rank 0 (None elsewhere), in input order."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    hdr = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == 0:
        hdr[0] = int(lefts.shape[0])
    dist.broadcast(hdr, src=0)                       # job descriptor
    n = int(hdr.item())
    bounds = shard_bounds(n, world)
    lo, hi = bounds[rank]
    img = height * width * 3
    # scatter inputs: rank 0 sends every other rank its block (point-to-point, grouped)
    mine = torch.empty((hi - lo, 2, img), dtype=torch.uint8, device=dev)
    if rank == 0:
        both = torch.stack([torch.from_numpy(np.ascontiguousarray(lefts)).reshape(n, img),
                            torch.from_numpy(np.ascontiguousarray(rights)).reshape(n, img)], dim=1).to(dev)
        reqs = [dist.isend(both[a:b].contiguous(), dst=r) for r, (a, b) in enumerate(bounds) if r != 0 and b > a]
        mine.copy_(both[lo:hi])
        for q in reqs:
            q.wait()
    elif hi > lo:
        dist.recv(mine, src=0)
    out = None
    if hi > lo:
        host = mine.cpu().numpy()
        out = match_fn(host[:, 0].reshape(-1, height, width, 3), host[:, 1].reshape(-1, height, width, 3))
    # gather results on rank 0 in rank order == input order
    if rank == 0:
        result = np.empty((n, height, width), np.float32)
        if hi > lo:
            result[lo:hi] = out
        for r, (a, b) in enumerate(bounds):
            if r == 0 or b == a:
                continue
            buf = torch.empty((b - a, height, width), dtype=torch.float32, device=dev)
            dist.recv(buf, src=r)
            result[a:b] = buf.cpu().numpy()
        return result
    if hi > lo:
        dist.send(torch.from_numpy(np.ascontiguousarray(out)).to(dev), dst=0)
    return None
