"""Tiny online k-means case for compute-sanitizer / debugging: python tools/mini_online.py [K] [seconds] [passes]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soundchunks_b200 as sc
from soundchunks_b200.synth import synth_audio
from oracle import gsc_oracle as O

K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
mp = int(sys.argv[3]) if len(sys.argv) > 3 else 7
a = synth_audio(sec, 44100, 1, 11)
pcm = np.ascontiguousarray(a[:, : a.shape[1] // 4 * 4])
raw, attr, atten, feat, dst = O.make_chunks(pcm, 4, 12, 6)
c0, _, _ = O.yakmo(feat, K)
ref = O.knn_scan_reduce(feat, c0, 3, mp)
with sc.Context(0) as ctx:
    got = ctx.knn_scan_reduce(feat, c0, 3, mp)
print("passes", got[2], ref[2], "labels equal", np.array_equal(got[1], ref[1]), "first diff",
      int(np.argmax(got[1] != ref[1])) if not np.array_equal(got[1], ref[1]) else -1, "err", got[3], ref[3])
