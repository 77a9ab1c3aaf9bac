"""BASELINE.json configs[3]: one oversized frame (default 1M chunks, K = 4096, 10 Lloyd iterations) split by
points over the ranks, NCCL all-reduce of the K x 9 partial sums.  Launch with torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/split_frame.py [--points 1048576] [--iters 10] [--check]
--check compares with the single-GPU gsc_lloyd of the whole frame (1e-4 relative, BASELINE.json)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import soundchunks_b200 as sc
from soundchunks_b200.split_kmeans import lloyd_split_gpu
from soundchunks_b200.synth import synth_audio

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=1 << 20)
ap.add_argument("--K", type=int, default=4096)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--check", action="store_true")
a = ap.parse_args()
rank, lr, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = sc.Context(lr)
# the same synthetic frame on every rank (seeded), features from the library itself
secs = a.points * 4 / (2 * 48000) + 0.01
pcm = np.ascontiguousarray(synth_audio(secs, 48000, 2, seed=77)[:, : a.points * 4 // 2])
feat = ctx.make_chunks(pcm, 4, 12, 6)[2]
N = len(feat)
rng = np.random.default_rng(5)
c0 = feat[np.sort(rng.choice(N, a.K, replace=False))].copy()
lo, hi = rank * N // world, (rank + 1) * N // world
torch.cuda.synchronize()
if world > 1:
    dist.barrier(device_ids=[lr])
t0 = time.perf_counter()
cen, labels = lloyd_split_gpu(ctx, feat[lo:hi], c0, a.iters, dist if world > 1 else None)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t = torch.tensor([dt], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
out = {"config": "oversized frame split", "points": N, "K": a.K, "iters": a.iters, "n_gpus": world,
       "seconds": float(t.item()), "allreduce_bytes_per_iter": a.K * 9 * 8,
       "tflops_dense": 2.0 * N * a.K * 8 * (a.iters + 1) / float(t.item()) / 1e12}
if a.check:
    ref_cen, ref_lab = ctx.lloyd(feat, c0, a.iters)
    # per centroid: largest coordinate difference relative to the centroid's largest coordinate
    rel = np.max(np.abs(cen - ref_cen), axis=1) / np.maximum(np.max(np.abs(ref_cen), axis=1), 1e-12)
    out["centroid_rel_diff"] = {"median": float(np.median(rel)), "p99": float(np.quantile(rel, 0.99)), "max": float(rel.max()),
                                "frac_within_1e-4": float(np.mean(rel <= 1e-4))}
    out["label_mismatch_frac"] = float(np.mean(labels != ref_lab[lo:hi]))
    # Double accumulation makes the means independent of the summation order: expect identical results
    out["bit_identical"] = bool(np.array_equal(cen.view(np.uint32), ref_cen.view(np.uint32)))
    assert out["centroid_rel_diff"]["max"] <= 1e-4 and out["label_mismatch_frac"] < 1e-4, out
if rank == 0:
    print(json.dumps(out), flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
