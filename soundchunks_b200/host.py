"""ctypes binding of libgsc_host.so (include/gsc_host.h): the host side of the encoder --
WAV I/O, frame planning, .gsc writer/decoder, the multi-GPU frame scheduler.

The same library backs the `gsc_encode` / `gsc_decode` command-line tools (host/).  Encoding
needs a B200: there is no CPU fallback, gsch_encode_* fail with libgsc_cuda's error text.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from .binding import FrameResult, FrameResultC, GscError, load_library

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_DIR = os.path.join(_ROOT, "host")
HOST_SO = os.path.join(HOST_DIR, "_build", "libgsc_host.so")
ENCODE_BIN = os.path.join(HOST_DIR, "_build", "gsc_encode")
DECODE_BIN = os.path.join(HOST_DIR, "_build", "gsc_decode")

HOST_EXPORTS = [
    "gsch_last_error", "gsch_default_options", "gsch_parse_option", "gsch_load_wav", "gsch_save_wav", "gsch_free",
    "gsch_padded_samples", "gsch_plan_frames", "gsch_write_frame", "gsch_decode", "gsch_reconstruct_frame",
    "gsch_psy_a_delta", "gsch_encode_pcm", "gsch_encode_file", "gsch_decode_file",
]


class Options(C.Structure):
    _fields_ = [("bitrate", C.c_int32), ("precision", C.c_int32), ("low_cut", C.c_double), ("high_cut", C.c_double),
                ("vfr", C.c_double), ("frame_length_ms", C.c_double), ("chunk_bit_depth", C.c_int32),
                ("chunk_size", C.c_int32), ("chunks_per_frame", C.c_int32), ("chunk_blend", C.c_int32),
                ("verbose", C.c_int32), ("kmeans_mode", C.c_int32), ("lloyd_iters", C.c_int32),
                ("max_passes", C.c_int32), ("devices", C.c_int32), ("frames_per_call", C.c_int32)]


class Report(C.Structure):
    _fields_ = [("frames", C.c_int32), ("channels", C.c_int32), ("sample_rate", C.c_int32),
                ("chunks_per_frame", C.c_int32), ("devices", C.c_int32), ("samples", C.c_int64),
                ("gsc_bytes", C.c_int64), ("bitrate_kbps", C.c_double), ("psy_a_delta", C.c_double),
                ("encode_seconds", C.c_double), ("overfull", C.c_int64)]


_hlib = None


def load_host_library() -> C.CDLL:
    global _hlib
    if _hlib is not None:
        return _hlib
    if not os.path.exists(HOST_SO):
        raise GscError(f"{HOST_SO} is missing: build it with `python -m soundchunks_b200.build`")
    load_library()          # libgsc_cuda.so first (rpath also finds it)
    L = C.CDLL(HOST_SO)
    L.gsch_last_error.restype = C.c_char_p
    L.gsch_padded_samples.restype = C.c_int64
    L.gsch_padded_samples.argtypes = [C.c_int64, C.POINTER(Options)]
    L.gsch_plan_frames.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_int, C.POINTER(Options), C.c_void_p, C.c_int]
    L.gsch_write_frame.restype = C.c_int64
    L.gsch_write_frame.argtypes = [C.POINTER(FrameResultC), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64]
    L.gsch_decode.restype = C.c_int64
    L.gsch_decode.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.gsch_reconstruct_frame.argtypes = [C.POINTER(FrameResultC), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64]
    L.gsch_psy_a_delta.restype = C.c_double
    L.gsch_psy_a_delta.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.gsch_encode_pcm.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_int, C.POINTER(Options),
                                  C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(Report)]
    L.gsch_free.argtypes = [C.c_void_p]
    L.gsch_parse_option.argtypes = [C.POINTER(Options), C.c_char_p]
    L.gsch_load_wav.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int)]
    L.gsch_save_wav.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int64, C.c_int]
    L.gsch_encode_file.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(Options), C.POINTER(Report)]
    L.gsch_decode_file.argtypes = [C.c_char_p, C.c_char_p]
    _hlib = L
    return L


def _err(L) -> str:
    return L.gsch_last_error().decode()


def default_options(*cli: str, **kw) -> Options:
    """Options with the reference's defaults (enc:1486-1509); `cli` are `-xx<value>` arguments."""
    L = load_host_library()
    o = Options()
    L.gsch_default_options(C.byref(o))
    for a in cli:
        if L.gsch_parse_option(C.byref(o), a.encode()):
            raise GscError(f"unknown option {a}")
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _planar(pcm) -> np.ndarray:
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    return pcm[None, :] if pcm.ndim == 1 else pcm


def pad(pcm, opts: Options) -> np.ndarray:
    pcm = _planar(pcm)
    L = load_host_library()
    S = int(L.gsch_padded_samples(pcm.shape[1], C.byref(opts)))
    if S == pcm.shape[1]:
        return pcm
    return np.ascontiguousarray(np.concatenate([pcm, np.zeros((pcm.shape[0], S - pcm.shape[1]), np.int16)], axis=1))


def plan_frames(pcm, sample_rate: int, opts: Options):
    """enc:1294-1429 on padded planar PCM -> (starts, chunks_per_frame after the bit-rate solve)."""
    L = load_host_library()
    pcm = _planar(pcm)
    Cn, S = pcm.shape
    o = Options.from_buffer_copy(opts)
    cap = S // max(1, o.chunk_size) + 2
    starts = np.zeros(cap, np.int64)
    n = L.gsch_plan_frames(pcm.ctypes.data, S, Cn, S, sample_rate, C.byref(o), starts.ctypes.data, cap)
    if n < 0:
        raise GscError(_err(L))
    return starts[:n].copy(), int(o.chunks_per_frame)


def _frame_struct(fr):
    d = np.ascontiguousarray(fr.dict, np.int16)
    a = np.ascontiguousarray(fr.datten, np.uint8)
    i = np.ascontiguousarray(fr.index, np.int32)
    t = np.ascontiguousarray(fr.attr, np.uint8)
    r = FrameResultC(d.ctypes.data, a.ctypes.data, i.ctypes.data, t.ctypes.data, fr.N, fr.R, fr.divider, fr.passes,
                     fr.err, fr.overfull, 0)
    return r, (d, a, i, t)


def write_frame(fr, channels: int, cs: int, bits: int, sample_rate: int) -> bytes:
    """enc:980-1107 for one frame result (a FrameResult of the binding or anything with its fields)."""
    L = load_host_library()
    r, keep = _frame_struct(fr)
    n = L.gsch_write_frame(C.byref(r), channels, cs, bits, sample_rate, None, 0)
    buf = np.zeros(n, np.uint8)
    L.gsch_write_frame(C.byref(r), channels, cs, bits, sample_rate, buf.ctypes.data, n)
    del keep
    return buf.tobytes()


def decode(gsc: bytes):
    """dec:37-220 -> (planar int16 [C][S], sample_rate)"""
    L = load_host_library()
    g = np.frombuffer(gsc, np.uint8)
    ch, sr = C.c_int(0), C.c_int(0)
    n = L.gsch_decode(g.ctypes.data, len(g), None, 0, C.byref(ch), C.byref(sr))
    if n < 0:
        raise ValueError("malformed .gsc stream")
    out = np.zeros((max(ch.value, 1), max(n, 1)), np.int16)
    L.gsch_decode(g.ctypes.data, len(g), out.ctypes.data, n, C.byref(ch), C.byref(sr))
    return out[:ch.value, :n].copy(), sr.value


def reconstruct_frame(fr, channels: int, samples: int, cs: int, bits: int) -> np.ndarray:
    L = load_host_library()
    r, keep = _frame_struct(fr)
    out = np.zeros((channels, samples), np.int16)
    L.gsch_reconstruct_frame(C.byref(r), channels, samples, cs, bits, out.ctypes.data, samples)
    del keep
    return out


def psy_a_delta(a, b) -> float:
    L = load_host_library()
    a = np.ascontiguousarray(a, np.int16).ravel()
    b = np.ascontiguousarray(b, np.int16).ravel()
    return float(L.gsch_psy_a_delta(a.ctypes.data, b.ctypes.data, len(a)))


def encode_pcm(pcm, sample_rate: int, opts: Optional[Options] = None, **kw):
    """Whole encode (Load .. SaveGSC minus file I/O) on the visible GPUs -> (gsc bytes, report dict)."""
    L = load_host_library()
    pcm = _planar(pcm)
    o = opts if opts is not None else default_options(**kw)
    blob, n, rep = C.c_void_p(), C.c_int64(0), Report()
    if L.gsch_encode_pcm(pcm.ctypes.data, pcm.shape[1], pcm.shape[0], pcm.shape[1], sample_rate, C.byref(o),
                         C.byref(blob), C.byref(n), C.byref(rep)):
        raise GscError(_err(L))
    data = C.string_at(blob, n.value)
    L.gsch_free(blob)
    return data, {f: getattr(rep, f) for f, _ in Report._fields_}
