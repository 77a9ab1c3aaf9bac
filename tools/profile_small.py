"""Small, short workload for ncu captures (one gsc_encode_frames call).
usage: python tools/profile_small.py [n_frames] [seconds] [K] [max_passes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soundchunks_b200 as sc
from soundchunks_b200.synth import synth_frames

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
K = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
mp = int(sys.argv[4]) if len(sys.argv) > 4 else 3
frames = synth_frames(nf, sec, 48000, 2, seed=1234)
with sc.Context(0) as ctx:
    res = ctx.encode_frames(frames, chunk_bit_depth=12, chunks_per_frame=K, max_passes=mp)
    st = ctx.stats()
print("frames", nf, "N", res[0].N, "passes", [r.passes for r in res], {k: round(v, 2) for k, v in st["stage_ms"].items()})
