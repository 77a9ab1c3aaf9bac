"""Oversized single frame (BASELINE.json configs[3]): the points of ONE frame are
split over the ranks, the codebook is replicated, and every Lloyd iteration all-reduces the
K x (D+1) partial sums (sums of the member rows + member count) -- the only collective of the
whole encoder.  With the NCCL backend the buffer is device memory filled by gsc_split_step and the
reduction runs over NVLink; the same code runs over gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable

import numpy as np


def lloyd_split(partial: Callable, centroids: np.ndarray, iters: int, dist=None, update: Callable | None = None):
    """Generic driver.  `partial(cen)` -> tensor [K][D+1] of THIS rank's sums and counts for the
    current centroids (host array or device tensor); the tensor is all-reduced in place.
    `update(acc)` (optional) consumes the reduced tensor on the device; without it the division
    happens here: mean = Single(sum / count) in Double, empty clusters keep their centroid."""
    cen = np.array(centroids, dtype=np.float32, copy=True)
    for _ in range(iters):
        acc = partial(cen)
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        if update is not None:
            cen = update(acc)
            continue
        a = (acc.detach().cpu().numpy() if hasattr(acc, "detach") else np.asarray(acc)).astype(np.float64)
        cnt = a[:, -1]
        nz = cnt > 0
        cen[nz] = (a[nz, :-1] / cnt[nz, None]).astype(np.float32)
    return cen


def lloyd_split_gpu(ctx, X_shard: np.ndarray, centroids: np.ndarray, iters: int, dist=None):
    """The device path: this rank's shard lives in `ctx` (libgsc_cuda), the K x (D+1) partial sums
    are written straight into a torch CUDA tensor and all-reduced by NCCL over NVLink; the division
    runs on the device.  Returns (centroids, labels of this shard) after a final assignment."""
    import torch
    K, D = np.asarray(centroids).shape
    ctx.split_begin(X_shard, centroids)
    acc = torch.empty((K, D + 1), dtype=torch.float64, device=torch.device("cuda", ctx.device))
    multi = dist is not None and dist.is_initialized() and dist.get_world_size() > 1
    for _ in range(iters):
        ctx.split_step(acc.data_ptr())              # returns with the library stream idle
        if multi:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize(acc.device)
        ctx.split_update(acc.data_ptr())
    return ctx.split_end()
