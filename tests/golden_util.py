"""Helpers shared by the golden-fixture tests (CPU: oracle vs fixtures; GPU: CUDA path vs fixtures).

Two kinds of fixtures live in tests/golden/ (both written by make_golden.py from the reference's own test audio):
  *.npz (not full_*)  short excerpts with every stage output stored as arrays
  full_*.npz          whole frames as the reference's planner cuts them (BASELINE.json configs[0..2]) + two
                      synthetic lattice-noise frames that overfill the epsilon band; PCM, scalars and the
                      SHA-256 of every stage output
"""
import glob
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EXCERPTS = sorted(p for p in glob.glob(os.path.join(HERE, "golden", "*.npz")) if not os.path.basename(p).startswith("full_"))
FULL = sorted(glob.glob(os.path.join(HERE, "golden", "full_*.npz")))


def ids(paths):
    return [os.path.basename(p)[:-4] for p in paths]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def canon(a):
    """float32 array with every NaN replaced by the canonical quiet NaN (x86 and CUDA produce different payloads)."""
    a = np.array(a, np.float32, copy=True)
    a.view(np.uint32)[np.isnan(a)] = 0x7FC00000
    return a


def load_full(path):
    g = np.load(path)
    return (np.ascontiguousarray(g["pcm"]), int(g["sample_rate"]), int(g["bits"]), int(g["K"]),
            json.loads(str(g["scalars"])), json.loads(str(g["hashes"])), g["band_hist"])


def check_stage(name, got, hashes):
    assert sha(got) == hashes[name], f"stage '{name}' differs from the golden fixture"
