"""Where does the host-buffer call spend its time?  python tools/e2e_probe.py [frames]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soundchunks_b200 as sc
from soundchunks_b200.binding import FrameDesc, FrameResultC
sys.argv = [sys.argv[0]] + sys.argv[1:]
import bench

F = int(sys.argv[1]) if len(sys.argv) > 1 else 296
frames = bench.make_frames(F, 4.0, seed=1234)
p = sc.default_params(chunk_bit_depth=12, chunks_per_frame=4096)
ctx = sc.Context(0)
for rep in range(2):
    t0 = time.perf_counter()
    n = len(frames)
    desc = (FrameDesc * n)()
    res = (FrameResultC * n)()
    bufs = []
    for i, f in enumerate(frames):
        Cn, S = f.shape
        N = ((S - 1) // 4 + 1) * Cn
        desc[i] = FrameDesc(f.ctypes.data, S, Cn, S)
        d = np.zeros((4096, 4), np.int16); a = np.zeros(4096, np.uint8); ix = np.zeros(N, np.int32); at = np.zeros(N, np.uint8)
        bufs.append((d, a, ix, at))
        res[i].dict, res[i].datten, res[i].index, res[i].attr = d.ctypes.data, a.ctypes.data, ix.ctypes.data, at.ctypes.data
    t1 = time.perf_counter()
    rc = ctx.L.gsc_encode_frames(C.c_void_p(ctx.h), desc, n, C.byref(p), res)
    t2 = time.perf_counter()
    st = ctx.stats()
    print(f"rep {rep}: python prep {t1 - t0:.3f} s, C call {t2 - t1:.3f} s, rc {rc}, GPU stage totals (both lanes added) {st['stage_ms']['total']:.0f} ms")
