"""Small cases for compute-sanitizer / debugging, each checked against the CPU oracle.

  python tools/mini_stage.py online K seconds passes   # gsc_knn_scan_reduce (k_online) from oracle seeds
  python tools/mini_stage.py seed   K seconds          # gsc_yakmo (k_seed + mean update)
  python tools/mini_stage.py frame  K seconds bits     # gsc_encode_frames + gsc_fetch_stream (every kernel of the pipeline)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soundchunks_b200 as sc
from soundchunks_b200.synth import synth_audio
from oracle import gsc_oracle as O

what = sys.argv[1]
K = int(sys.argv[2])
sec = float(sys.argv[3])
a = synth_audio(sec, 44100, 1, 11)
pcm = np.ascontiguousarray(a[:, : a.shape[1] // 4 * 4])
raw, attr, atten, feat, dst = O.make_chunks(pcm, 4, 12, 6)
ok = True
with sc.Context(0) as ctx:
    if what == "online":
        mp = int(sys.argv[4])
        c0, _, _ = O.yakmo(feat, K)
        ref = O.knn_scan_reduce(feat, c0, 3, mp)
        got = ctx.knn_scan_reduce(feat, c0, 3, mp)
        ok = np.array_equal(got[1], ref[1]) and got[2] == ref[2] and got[3] == ref[3] and \
            np.array_equal(got[0].view(np.uint32), ref[0].view(np.uint32))
        print("online K", K, "N", len(feat), "passes", got[2], ref[2], "err", got[3], ref[3])
    elif what == "seed":
        rc, rl, rs = O.yakmo(feat, K)
        gc, gl, gs = ctx.yakmo(feat, K)
        ok = np.array_equal(gs, rs) and np.array_equal(gc.view(np.uint32), rc.view(np.uint32))
        print("seed K", K, "N", len(feat), "seeds equal", np.array_equal(gs, rs))
    else:
        bits = int(sys.argv[4])
        fr = [pcm, np.ascontiguousarray(pcm[:, ::-1])]
        res = ctx.encode_frames(fr, chunk_bit_depth=bits, chunks_per_frame=K)
        blob, sizes = ctx.fetch_stream(len(fr), 44100)
        e2, ns = ctx.fetch_quality(len(fr))
        want = b""
        for f, r in zip(fr, res):
            ref = O.encode_frame(f, chunk_bit_depth=bits, chunks_per_frame=K, band_all=1)
            ok &= (r.passes == ref.passes and r.err == ref.err and np.array_equal(r.index, ref.index)
                   and np.array_equal(r.dict, ref.dict) and np.array_equal(r.attr, ref.attr))
            want += O.write_frame(ref, 1, 4, bits, 44100)
        ok &= (blob == want)
        print("frame K", K, "N", res[0].N, "passes", [r.passes for r in res], "stream bytes", len(blob))
print("RESULT", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
