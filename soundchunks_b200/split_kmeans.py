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
    happens here in float32: mean = sum / count, empty clusters keep their centroid."""
    cen = np.array(centroids, dtype=np.float32, copy=True)
    for _ in range(iters):
        acc = partial(cen)
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        if update is not None:
            cen = update(acc)
            continue
        a = acc.detach().cpu().numpy() if hasattr(acc, "detach") else np.asarray(acc)
        cnt = a[:, -1]
        nz = cnt > 0
        cen[nz] = (a[nz, :-1] / cnt[nz, None]).astype(np.float32)
    return cen
