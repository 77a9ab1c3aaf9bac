"""Frame sharding over GPUs / ranks: frames are independent units (DoFrame,
enc:1433-1447), so the only multi-GPU mechanism of the normal path is a partition of the frame list;
no collective touches the data.  The same greedy rule runs inside libgsc_host.so (gsch_encode_pcm).
"""
from __future__ import annotations

import os
from typing import List, Sequence


def shard_frames(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-first partition of frame indexes over `world` ranks; every rank's list is
    ascending (frames are written in index order, enc:1213-1214).  Deterministic on every rank."""
    if world < 1:
        raise ValueError("world must be >= 1")
    order = sorted(range(len(lengths)), key=lambda k: (-int(lengths[k]), k))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for k in order:
        d = min(range(world), key=lambda r: (load[r], r))
        out[d].append(k)
        load[d] += int(lengths[k])
    return [sorted(s) for s in out]


def rank_env():
    """(rank, local_rank, world) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def gather_frames(local_results, my_frames: Sequence[int], n_frames: int, dist=None):
    """All ranks' per-frame results in frame order on every rank (all_gather_object; host objects,
    used for the .gsc writer which runs on the host).  dist=None: single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        pairs = list(zip(my_frames, local_results))
    else:
        box = [None] * dist.get_world_size()
        dist.all_gather_object(box, list(zip(my_frames, local_results)))
        pairs = [p for part in box for p in part]
    out = [None] * n_frames
    for k, r in pairs:
        if out[k] is not None:
            raise RuntimeError(f"frame {k} was encoded twice")
        out[k] = r
    missing = [k for k, r in enumerate(out) if r is None]
    if missing:
        raise RuntimeError(f"frames {missing[:8]} were not encoded by any rank")
    return out


def reduce_scalar(x: float, op: str, dist=None, device=None) -> float:
    """max / sum of a host scalar over the ranks (timing and unit counts of bench.py)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())
