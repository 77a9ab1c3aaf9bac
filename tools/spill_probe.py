"""Does the online k-means (k_online, gsc_online.cuh) depend on the register allocation?

Builds register-capped variants of the library (ptxas then spills in k_online), runs gsc_knn_scan_reduce on
full-size frames (4 s stereo, N = 88,200, K = 4096) for a ladder of pass counts with every variant, and compares
labels / centroids / pass count / error sum bit for bit with the shipped build and with the CPU oracle
(enc:699-765).  Prints the first pass count and point at which a variant diverges.

  python tools/spill_probe.py build            # here (no GPU): nvcc the variants into .scratch/variants/
  python tools/spill_probe.py oracle [n]       # CPU: oracle results for n inputs -> .scratch/variants/oracle_*.npz
  python tools/spill_probe.py run              # on the GPU box: compare
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "tools", "_variants")
CAPS = [224, 192, 160]
LADDER = [1, 2, 3, 5, 8, 16, 100]


def inputs(n):
    from soundchunks_b200.synth import synth_frames
    return synth_frames(n, 4.0, 44100, 2, seed=777)


def main():
    cmd = sys.argv[1] if len(sys.argv) > 1 else "run"
    os.makedirs(VDIR, exist_ok=True)
    if cmd == "build":
        from soundchunks_b200 import build as b
        for cap in CAPS:
            sp = b.build_variant(os.path.join(VDIR, f"libgsc_cuda_r{cap}.so"), cap)
            print("cap", cap, "spilling k_online shapes:", len(sp))
        sp = b.build_variant(os.path.join(VDIR, "libgsc_cuda_checks.so"), 0, checks=True)
        print("checks build, spilling k_online shapes:", len(sp))
        return
    if cmd == "oracle":
        from concurrent.futures import ThreadPoolExecutor
        from oracle import gsc_oracle as O
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
        fr = inputs(n)

        def one(i):
            raw, attr, atten, feat, dst = O.make_chunks(fr[i], 4, 12, 6)
            c0, _, _ = O.yakmo(feat, 4096)
            out = {"feat": feat, "c0": c0}
            for P in LADDER:
                cen, lab, it, err = O.knn_scan_reduce(feat, c0, 3, P)
                out[f"cen{P}"], out[f"lab{P}"], out[f"it{P}"], out[f"err{P}"] = cen, lab, it, err
            np.savez_compressed(os.path.join(VDIR, f"oracle_{i}.npz"), **out)
            return i
        with ThreadPoolExecutor(n) as ex:
            print(list(ex.map(one, range(n))))
        return
    if cmd == "child":
        import soundchunks_b200 as sc
        tag = sys.argv[2]
        files = sorted(f for f in os.listdir(VDIR) if f.startswith("oracle_"))
        with sc.Context(0) as ctx:
            for f in files:
                g = np.load(os.path.join(VDIR, f))
                for P in LADDER:
                    cen, lab, it, err = ctx.knn_scan_reduce(g["feat"], g["c0"], 3, P)
                    ok = (np.array_equal(lab, g[f"lab{P}"]) and it == int(g[f"it{P}"]) and err == float(g[f"err{P}"])
                          and np.array_equal(cen.view(np.uint32), g[f"cen{P}"].view(np.uint32)))
                    msg = f"{tag} {f} passes<={P}: it={it} err={err!r} {'OK' if ok else 'MISMATCH'}"
                    if not ok:
                        d = np.nonzero(lab != g[f"lab{P}"])[0]
                        msg += f" first label diff at {d[0] if len(d) else -1} ({len(d)} differ), oracle it={int(g[f'it{P}'])} err={float(g[f'err{P}'])!r}"
                    cn = ctx.online_counters(1)[0]
                    if cn[8]:
                        msg += f" CHECKS FAILED: {int(cn[8])} (first at gsc_online.cuh:{int(cn[9])})"
                    elif tag == "checks":
                        msg += " checks clean"
                    print(msg, flush=True)
        return
    # run: shipped build + every variant, each in its own process (the library path is fixed at import)
    libs = [("shipped", None)] + [(f[len("libgsc_cuda_"):-3], os.path.join(VDIR, f)) for f in sorted(os.listdir(VDIR))
                                   if f.startswith("libgsc_cuda_") and f.endswith(".so")]
    only = sys.argv[2:]                      # optional: tags to run (e.g. `run checks`)
    for tag, so in libs:
        if only and tag not in only:
            continue
        env = dict(os.environ)
        if so:
            if not os.path.exists(so):
                print(tag, "missing", so)
                continue
            env["GSC_CUDA_SO"] = so
        subprocess.run([sys.executable, os.path.abspath(__file__), "child", tag], env=env, check=False)


if __name__ == "__main__":
    main()
