"""CPU, world_size 2 over gloo: the N>1 host logic -- frame partition, result gathering in frame
order, max/sum reductions of bench.py -- and the collective step of the oversized-frame k-means
split (allreduce of K x (D+1) partial sums) with the oracle standing in for the device kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from soundchunks_b200.sharding import gather_frames, reduce_scalar, shard_frames


def test_shard_frames_partition():
    rng = np.random.default_rng(0)
    lens = rng.integers(20000, 200000, 37).tolist()
    for world in (1, 2, 4, 8):
        sh = shard_frames(lens, world)
        assert sorted(k for s in sh for k in s) == list(range(37))
        assert all(s == sorted(s) for s in sh)
        loads = [sum(lens[k] for k in s) for s in sh]
        assert max(loads) - min(loads) <= max(lens)          # greedy longest-first bound
    assert shard_frames([], 2) == [[], []]
    assert shard_frames([5], 4)[0] == [0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import gsc_oracle as O
        from soundchunks_b200.split_kmeans import lloyd_split
        from soundchunks_b200.synth import synth_audio
        # --- frame sharding + gather in frame order ---
        lens = [4000, 12000, 8000, 8000, 2000]
        mine = shard_frames(lens, world)[rank]
        frames = [np.ascontiguousarray(synth_audio(n / 44100, 44100, 1, seed=k)[:, : n // 4 * 4]) for k, n in enumerate(lens)]
        local = [O.encode_frame(frames[k], chunk_bit_depth=8, chunks_per_frame=256) for k in mine]
        allr = gather_frames(local, mine, len(lens), dist)
        blob = b"".join(O.write_frame(r, 1, 4, 8, 44100) for r in allr)
        t = reduce_scalar(1.0 + rank, "max", dist)
        n = reduce_scalar(float(len(mine)), "sum", dist)
        # --- oversized frame: points split over the ranks, allreduce of the partial sums ---
        pcm = synth_audio(0.6, 44100, 2, seed=99)
        raw, attr, atten, feat, dst = O.make_chunks(np.ascontiguousarray(pcm[:, : pcm.shape[1] // 4 * 4]), 4, 12, 6)
        c0 = np.nan_to_num(O.yakmo(feat, 128)[0])
        lo, hi = rank * len(feat) // world, (rank + 1) * len(feat) // world

        def partial(cen):      # stand-in for gsc_split_step: assign + per-cluster sums of this rank's points
            lab, _ = O.assign(feat[lo:hi], cen)
            acc = np.zeros((128, 9), np.float64)
            np.add.at(acc[:, :8], lab, feat[lo:hi].astype(np.float64))
            acc[:, 8] = np.bincount(lab, minlength=128)
            return torch.from_numpy(acc)

        cen = lloyd_split(partial, c0, 6, dist)
        q.put((rank, blob, t, n, cen, mine))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    from oracle import gsc_oracle as O
    from soundchunks_b200.synth import synth_audio
    O.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, blob0, t0, n0, cen0, m0), (_, blob1, t1, n1, cen1, m1) = out
    assert sorted(m0 + m1) == [0, 1, 2, 3, 4] and not set(m0) & set(m1)
    assert blob0 == blob1 and t0 == t1 == 2.0 and n0 == n1 == 5.0
    # same stream as a single process encoding the frames in order
    lens = [4000, 12000, 8000, 8000, 2000]
    ref = b"".join(O.write_frame(O.encode_frame(np.ascontiguousarray(synth_audio(n / 44100, 44100, 1, seed=k)[:, : n // 4 * 4]),
                                                chunk_bit_depth=8, chunks_per_frame=256), 1, 4, 8, 44100)
                   for k, n in enumerate(lens))
    assert blob0 == ref
    # split Lloyd == single-process Lloyd within BASELINE.json's 1e-4 relative
    pcm = synth_audio(0.6, 44100, 2, seed=99)
    raw, attr, atten, feat, dst = O.make_chunks(np.ascontiguousarray(pcm[:, : pcm.shape[1] // 4 * 4]), 4, 12, 6)
    c0 = np.nan_to_num(O.yakmo(feat, 128)[0])
    cref, _ = O.lloyd(feat, c0, 6)
    assert np.array_equal(cen0, cen1)
    assert np.max(np.abs(cen0 - cref) / np.maximum(np.abs(cref), 1e-6)) <= 1e-4
    assert np.array_equal(cen0.view(np.uint32), cref.view(np.uint32))   # Double accumulation: order-independent
