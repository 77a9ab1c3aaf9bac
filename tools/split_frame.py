"""BASELINE.json configs[3]: one oversized frame (default 1M chunks, K = 4096, 10 Lloyd iterations) split by points
over the ranks; the collective (ncclAllReduce of the K x 9 Double partial sums) runs INSIDE libgsc_cuda
(gsc_split_lloyd).  One rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/split_frame.py [--points 1048576] [--iters 10] [--check]
torch.distributed is only the transport of the 128-byte NCCL id and of the max-over-ranks of the times.
--check compares with the single-GPU gsc_lloyd of the whole frame: BASELINE.json's 1e-4 relative is asserted; bit
identity is reported.  (The partial sums are Doubles of float terms and almost always exact, so the order of the
additions -- atomics on one GPU, NCCL's reduction tree over ranks -- rarely shows: 2 and 4 ranks came out bit
identical, 8 ranks differed in 2 of 32,768 coordinates by one float ulp with every label equal.)"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run_split(ctx, rank, world, points, K, iters, bcast, maxred, check=False, seed_split=True):
    """-> result dict (same on every rank).  bcast(bytes) -> bytes from rank 0; maxred(float) -> max over ranks."""
    import soundchunks_b200 as sc
    from soundchunks_b200.synth import synth_audio
    secs = points * 4 / (2 * 48000) + 0.01
    pcm = np.ascontiguousarray(synth_audio(secs, 48000, 2, seed=77)[:, : points * 4 // 2])   # the same frame on every rank
    lo, hi = rank * points // world, (rank + 1) * points // world
    # this rank's shard of the frame's chunks (chunk n = i*C + ch: a shard is a range of sample positions)
    feat = ctx.make_chunks(np.ascontiguousarray(pcm[:, (lo // 2) * 4:((hi + 1) // 2) * 4]), 4, 12, 6)[2][: hi - lo]
    if world > 1:
        uid = bcast(sc.Context.split_unique_id() if rank == 0 else b"\0" * 128)
        ctx.split_comm_init(world, rank, uid)
    c0 = ctx.split_seed(feat, K)                                  # k-means++ on rank 0's shard, broadcast
    ctx.split_lloyd(feat, c0, 1)                                  # warm-up: allocations, NCCL channels
    cen, labels, ms = ctx.split_lloyd(feat, c0, iters)
    t = maxred(ms["ms_loop"]) * 1e-3
    ar = maxred(ms["ms_allreduce"]) * 1e-3
    out = {"config": "BASELINE.json configs[3]: one frame of %d chunks (D=8), K=%d, %d Lloyd iterations + final assignment, "
                     "points split over %d rank(s), all-reduce of K x 9 doubles per iteration inside the library (NCCL)" % (points, K, iters, world),
           "n_gpus": world, "seconds": t, "ms_per_iteration": 1e3 * t / (iters + 1), "allreduce_share": ar / t if t > 0 else None,
           "allreduce_bytes_per_iter": K * 9 * 8,
           "tflops_dense": 2.0 * points * K * 8 * (iters + 1) / t / 1e12,
           "tflops_dense_per_gpu": 2.0 * points * K * 8 * (iters + 1) / t / 1e12 / world}
    if check:
        full = ctx.make_chunks(pcm, 4, 12, 6)[2][:points]
        ref_cen, ref_lab = ctx.lloyd(full, c0, iters)
        out["bit_identical_to_single_gpu"] = bool(np.array_equal(cen.view(np.uint32), ref_cen.view(np.uint32))
                                                  and np.array_equal(labels, ref_lab[lo:hi]))
        rel = np.max(np.abs(cen - ref_cen), axis=1) / np.maximum(np.max(np.abs(ref_cen), axis=1), 1e-12)
        out["check"] = {"centroid_coords_differing": int((cen.view(np.uint32) != ref_cen.view(np.uint32)).sum()),
                        "centroid_max_rel_diff": float(rel.max()), "labels_differing": int((labels != ref_lab[lo:hi]).sum()),
                        "rank": rank}
        if not out["bit_identical_to_single_gpu"]:      # one iteration at a time: where does it start?
            for it in (1, 2, 3):
                c_s, _, _ = ctx.split_lloyd(feat, c0, it)
                c_r, _ = ctx.lloyd(full, c0, it)
                out["check"][f"coords_differing_after_{it}"] = int((c_s.view(np.uint32) != c_r.view(np.uint32)).sum())
        assert rel.max() <= 1e-4, out                   # BASELINE.json's contract; bit-identity is reported
    if world > 1:
        ctx.split_comm_destroy()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1 << 20)
    ap.add_argument("--K", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import soundchunks_b200 as sc
    rank, lr, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

    def bcast(b):
        obj = [b]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    def maxred(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    ctx = sc.Context(lr)
    out = run_split(ctx, rank, world, a.points, a.K, a.iters, bcast, maxred, a.check)
    if rank == 0:
        print(json.dumps(out), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
