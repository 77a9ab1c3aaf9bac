"""soundchunks_b200 -- B200 (sm_100a) implementation of the SoundChunks encoder hot path.

The product is libgsc_cuda.so (csrc/, C ABI in include/gsc_cuda.h); this package is the thin
Python host side over it: ctypes binding, the encoder pipeline mirror (frame planning, .gsc
bitstream writer stay on the host) and the cluster.py-compatible shim.
"""
from .binding import (Context, FrameResult, GscError, LegacyAnn, Params, default_params, device_count,  # noqa: F401
                      legacy_yakmo, load_library, EXPORTS, SO_PATH)
from . import host  # noqa: E402,F401  (libgsc_host.so: planner, .gsc writer/decoder, multi-GPU scheduler)
