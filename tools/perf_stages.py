"""Stage timings of gsc_encode_frames on synthetic frames.
usage: python tools/perf_stages.py [n_frames] [K,bits ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soundchunks_b200 as sc
from soundchunks_b200.synth import synth_frames

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfgs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] or [(4096, 12), (256, 8)]
t = time.time()
base = synth_frames(8, 4.0, 48000, 2, seed=1234)
print("synth s", round(time.time() - t, 2))
frames = [base[i % 8] for i in range(nf)]
ctx = sc.Context(0)
print("fp32 peak TF", round(ctx.fp32_peak_tflops(), 2))
for K, bits in cfgs:
    t = time.time()
    res = ctx.encode_frames(frames, chunk_bit_depth=bits, chunks_per_frame=K)
    dt = time.time() - t
    st = ctx.stats()
    print("K", K, "frames", nf, "wall", round(dt, 3), "audio-s/s", round(nf * 4.0 / dt, 1), "passes",
          [r.passes for r in res][:8], "R", [r.R for r in res][:4], "overfull", [r.overfull for r in res][:4])
    print({k: round(v, 2) for k, v in st["stage_ms"].items()})
    sd = ctx.seed_counters(min(nf, 2)).astype(float)
    for r in sd:
        st_ = max(r[3], 1)
        print(f"   seeding cycles/step: pick {r[0] / st_:.0f} distance {r[1] / st_:.0f} summaries+chain {r[2] / st_:.0f} (chain {r[7] / st_:.0f}); "
              f"per step: exact windows {r[4] / st_:.1f} blocks visited {r[5] / st_:.1f} windows re-summarised {r[6] / st_:.1f}")
    cn = ctx.online_counters(min(nf, 8)).astype(float)
    for r in cn[:4]:
        b, p, e, rd, ov, ca = r[:6]
        print(f"   batches {b:.0f} points {p:.0f} exhaustive {e:.0f} rounds/batch {rd / max(b, 1):.2f} full lists {ov:.0f} cands/pt {ca / max(p, 1):.2f}")
