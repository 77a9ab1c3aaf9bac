"""Build libgsc_cuda.so in-tree with nvcc for sm_100a (no JIT cache, the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libgsc_cuda.so")
SOURCES = ["gsc_api.cu"]
DEPS = ["gsc_api.cu", "gsc_kernels.cuh", "gsc_online.cuh", "gsc_device.cuh", os.path.join("..", "..", "include", "gsc_cuda.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return SO
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-fmad=false", "-Xcompiler", "-fPIC,-O2,-fvisibility=default", "-shared", "-o", SO]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libgsc_cuda.so")
    return SO


def build_host(force: bool = False) -> str:
    """libgsc_host.so + gsc_encode / gsc_decode (host/Makefile, g++ only)."""
    host = os.path.join(os.path.dirname(HERE), "host")
    out = os.path.join(host, "_build", "libgsc_host.so")
    build(force=False)
    r = subprocess.run(["make", "-C", host, "-s"] + (["-B"] if force else []), capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("building the host library failed")
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
    print(build_host(force=True))
