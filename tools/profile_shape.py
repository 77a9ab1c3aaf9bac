"""Bench-shaped workload for ncu captures: n frames of the bench generator (44.1 kHz stereo) through one
gsc_encode_frames call.  usage: python tools/profile_shape.py frames seconds K bits max_passes [lloyd_iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soundchunks_b200 as sc
import bench

nf, sec, K, bits, mp = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
lloyd = int(sys.argv[6]) if len(sys.argv) > 6 else 0
frames = bench.make_frames(nf, sec, seed=1234)
with sc.Context(0) as ctx:
    kw = dict(kmeans_mode=1, lloyd_iters=lloyd) if lloyd else {}
    res = ctx.encode_frames(frames, chunk_bit_depth=bits, chunks_per_frame=K, max_passes=mp, **kw)
    blob, sizes = ctx.fetch_stream(nf, 44100)
    st = ctx.stats()
print("frames", nf, "N", res[0].N, "passes", [r.passes for r in res][:8], {k: round(v, 2) for k, v in st["stage_ms"].items()}, "bytes", len(blob))
