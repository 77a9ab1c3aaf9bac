"""Drop-in for the reference's `-py` reducer (encoder/cluster.py, driven by extern.pas:350-437).

Same command line and file protocol:
    python cluster.py -i <file> -n <num_clusters> [-t <threshold>] [-d]
reads `<file>` (one row per point: "<idx> v0 v1 ..."), writes `<file>.membership` (one `%d` label per
line) and `<file>.cluster_centres` (`%f` rows).  The clustering itself runs on the GPU through
libgsc_cuda.so: yakmo-style k-means++ seeding (gsc_yakmo) followed by Lloyd iterations (gsc_lloyd)
until no centroid coordinate moves by more than the threshold (at most 100 iterations).  The
centres written are the class means of the final labels, which is what the reference's
`NearestCentroid().fit(data, labels)` computes (cluster.py:26-30).  No CPU fallback.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from soundchunks_b200.binding import Context
else:
    from .binding import Context


def reduce_points(data: np.ndarray, n_clusters: int, tol: float = 1e-4, max_iter: int = 100, verbose: bool = False,
                  ctx: Context | None = None):
    """-> (labels int32 [N], centres float64 [n_used][D])"""
    X = np.ascontiguousarray(data, dtype=np.float32)
    N, D = X.shape
    if D not in (4, 8, 16):
        raise ValueError(f"feature dimension {D} is not supported by libgsc_cuda (4, 8, 16)")
    K = min(n_clusters, N)
    own = ctx is None
    ctx = ctx or Context(-1)
    try:
        cen, labels, _ = ctx.yakmo(X, K)                     # seeding + one mean update
        cen = np.nan_to_num(cen, nan=0.0)
        for it in range(max_iter):
            new, labels = ctx.lloyd(X, cen, 1)
            moved = float(np.max(np.abs(new - cen))) if len(cen) else 0.0
            cen = new
            if verbose:
                print(f"iteration {it}: max centroid move {moved:.3g}")
            if moved <= tol:
                break
    finally:
        if own:
            ctx.close()
    used = np.unique(labels)
    remap = np.full(K, -1, np.int32)
    remap[used] = np.arange(len(used), dtype=np.int32)
    labels = remap[labels]
    centres = np.zeros((len(used), D), np.float64)
    np.add.at(centres, labels, X.astype(np.float64))
    centres /= np.bincount(labels, minlength=len(used))[:, None]
    return labels.astype(np.int32), centres


def main(argv=None) -> int:
    parser = argparse.ArgumentParser()
    parser.add_argument("-i", help="filename")
    parser.add_argument("-n", help="num_clusters", type=int)
    parser.add_argument("-t", help="threshold", type=float, default=0.0001)
    parser.add_argument("-d", help="debug", action="store_true")
    args = parser.parse_args(argv)
    data = np.loadtxt(args.i, ndmin=2)
    data = np.delete(data, 0, 1)
    labels, centres = reduce_points(data, args.n, args.t, verbose=args.d)
    np.savetxt(args.i + ".membership", labels, fmt="%d")
    np.savetxt(args.i + ".cluster_centres", centres, fmt="%f")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
