"""ctypes binding of the CPU oracle (oracle/gsc_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(soundchunks_b200) never imports this module.  PARITY UNPINNED, see
gsc_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgsc_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "gsc_oracle.c")
    hdr = os.path.join(_HERE, "gsc_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class _Params(C.Structure):
    _fields_ = [("chunk_size", C.c_int), ("chunk_bit_depth", C.c_int),
                ("chunks_per_frame", C.c_int), ("precision", C.c_int),
                ("max_passes", C.c_int), ("kmeans_mode", C.c_int),
                ("lloyd_iters", C.c_int), ("batch", C.c_int),
                ("frame_length_ms", C.c_double), ("vfr", C.c_double),
                ("band_all", C.c_int), ("reserved", C.c_int)]


class _FrameOut(C.Structure):
    _fields_ = [("N", C.c_int), ("R", C.c_int), ("divider", C.c_int), ("passes", C.c_int),
                ("err", C.c_double),
                ("dict", C.POINTER(C.c_int16)), ("datten", C.POINTER(C.c_uint8)),
                ("index", C.POINTER(C.c_int32)), ("attr", C.POINTER(C.c_uint8)),
                ("overfull", C.c_int)]


@dataclass
class FrameResult:
    N: int
    R: int
    divider: int
    passes: int
    err: float
    dict: np.ndarray      # int16 [R][cs]
    datten: np.ndarray    # uint8 [R]
    index: np.ndarray     # int32 [N]
    attr: np.ndarray      # uint8 [N]  bit1 neg, bit0 rev
    overfull: int


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.gsc_ref_log_cr.restype = C.c_double
        L.gsc_ref_log_cr.argtypes = [C.c_double]
        L.gsc_ref_float_sample.restype = C.c_double
        L.gsc_ref_float_sample.argtypes = [C.c_int16]
        L.gsc_ref_make16.restype = C.c_int16
        L.gsc_ref_make16.argtypes = [C.c_double]
        L.gsc_ref_quant.restype = C.c_int16
        L.gsc_ref_quant.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_double]
        L.gsc_ref_dequant.restype = C.c_double
        L.gsc_ref_dequant.argtypes = [C.c_int16, C.c_int, C.c_int, C.c_int, C.c_double]
        L.gsc_ref_attenuation.restype = C.c_int
        L.gsc_ref_knnfit.restype = C.c_float
        L.gsc_ref_psy_a_delta.restype = C.c_double
        L.gsc_ref_snr_db.restype = C.c_double
        L.gsc_ref_write_frame.restype = C.c_int64
        L.gsc_ref_decode.restype = C.c_int64
        _lib = L
    return _lib


def default_params(**kw) -> _Params:
    p = _Params()
    lib().gsc_ref_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _pcm(pcm):
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    if pcm.ndim == 1:
        pcm = pcm[None, :]
    return pcm


# ---- scalar helpers -------------------------------------------------------
def quant(x, bits, atten, neg, law):
    return int(lib().gsc_ref_quant(float(x), bits, atten, int(neg), float(law)))


def dequant(q, bits, atten, neg, law):
    return float(lib().gsc_ref_dequant(int(q), bits, atten, int(neg), float(law)))


def attenuation(x, law):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return int(lib().gsc_ref_attenuation(len(x), _p(x, C.c_double), C.c_double(law)))


# ---- per frame ------------------------------------------------------------
def find_attenuation_divider(pcm, cs=4, bits=12, return_v=False):
    pcm = _pcm(pcm)
    Cn, S = pcm.shape
    v = np.zeros(64, np.float64)
    d = lib().gsc_ref_find_attenuation_divider(_p(pcm, C.c_int16), C.c_int64(S), Cn, S, cs, bits,
                                               _p(v, C.c_double))
    return (d, v) if return_v else d


def make_chunks(pcm, cs=4, bits=12, divider=6):
    """-> raw f64[N][cs], attr u8[N], atten u8[N], feat f32[N][2cs], dst i16[N][cs]"""
    pcm = _pcm(pcm)
    Cn, S = pcm.shape
    N = ((S - 1) // cs + 1) * Cn
    raw = np.zeros((N, cs), np.float64)
    attr = np.zeros(N, np.uint8)
    atten = np.zeros(N, np.uint8)
    feat = np.zeros((N, 2 * cs), np.float32)
    dst = np.zeros((N, cs), np.int16)
    n = lib().gsc_ref_make_chunks(_p(pcm, C.c_int16), C.c_int64(S), Cn, S, cs, bits, divider,
                                  _p(raw, C.c_double), _p(attr, C.c_uint8), _p(atten, C.c_uint8),
                                  _p(feat, C.c_float), _p(dst, C.c_int16))
    assert n == N
    return raw, attr, atten, feat, dst


def yakmo(X, K, init_type=1, max_iter=0):
    """-> centroids f32[K][D], labels i32[N], seeds i32[K]"""
    X = np.ascontiguousarray(X, dtype=np.float32)
    N, D = X.shape
    cen = np.zeros((K, D), np.float32)
    labels = np.zeros(N, np.int32)
    seeds = np.zeros(K, np.int32)
    lib().gsc_ref_yakmo(_p(X, C.c_float), N, D, K, init_type, max_iter, _p(cen, C.c_float),
                        _p(labels, C.c_int32), _p(seeds, C.c_int32))
    return cen, labels, seeds


def knn_scan_reduce(X, centroids, precision=3, max_passes=100, batch=1):
    """-> centroids f32[K][D], labels i32[N], passes, err"""
    X = np.ascontiguousarray(X, dtype=np.float32)
    cen = np.array(centroids, dtype=np.float32, order="C", copy=True)
    N, D = X.shape
    K = cen.shape[0]
    labels = np.zeros(N, np.int32)
    err = C.c_double(0)
    it = lib().gsc_ref_knn_scan_reduce_batched(_p(X, C.c_float), N, D, _p(cen, C.c_float), K,
                                               precision, max_passes, batch,
                                               _p(labels, C.c_int32), C.byref(err))
    return cen, labels, it, err.value


def knn_scan_reduce_kdtree(X, centroids, precision=3, max_passes=100, stats=False):
    """enc:699-765 through the ANN-style kd-tree (rebuilt per pass, live rows, stale planes).
    -> centroids, labels, passes, err[, (points visited, answers that are not the exact nearest centroid)]"""
    X = np.ascontiguousarray(X, dtype=np.float32)
    cen = np.array(centroids, dtype=np.float32, order="C", copy=True)
    N, D = X.shape
    labels = np.zeros(N, np.int32)
    err = C.c_double(0)
    st = (C.c_long * 2)(0, 0)
    it = lib().gsc_ref_knn_scan_reduce_kdtree(_p(X, C.c_float), N, D, _p(cen, C.c_float), cen.shape[0], precision,
                                              max_passes, _p(labels, C.c_int32), C.byref(err), st if stats else None)
    return (cen, labels, it, err.value, (st[0], st[1])) if stats else (cen, labels, it, err.value)


def lloyd(X, centroids, iters):
    X = np.ascontiguousarray(X, dtype=np.float32)
    cen = np.array(centroids, dtype=np.float32, order="C", copy=True)
    N, D = X.shape
    labels = np.zeros(N, np.int32)
    lib().gsc_ref_lloyd(_p(X, C.c_float), N, D, _p(cen, C.c_float), cen.shape[0], iters,
                        _p(labels, C.c_int32))
    return cen, labels


def assign(X, centroids):
    X = np.ascontiguousarray(X, dtype=np.float32)
    cen = np.ascontiguousarray(centroids, dtype=np.float32)
    N, D = X.shape
    labels = np.zeros(N, np.int32)
    dist = np.zeros(N, np.float32)
    lib().gsc_ref_assign(_p(X, C.c_float), N, D, _p(cen, C.c_float), cen.shape[0],
                         _p(labels, C.c_int32), _p(dist, C.c_float))
    return labels, dist


def build_dictionary(labels, raw, attr, K, bits=12, divider=6):
    """-> dict(means, order, counts, dict, datten, dattr, entry)"""
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    attr = np.ascontiguousarray(attr, dtype=np.uint8)
    N, cs = raw.shape
    out = dict(means=np.zeros((K, cs), np.float32), order=np.zeros(K, np.int32),
               counts=np.zeros(K, np.int32), dict=np.zeros((K, cs), np.int16),
               datten=np.zeros(K, np.uint8), dattr=np.zeros(K, np.uint8),
               entry=np.zeros(N, np.int32))
    lib().gsc_ref_build_dictionary(_p(labels, C.c_int32), _p(raw, C.c_double), _p(attr, C.c_uint8),
                                   N, cs, K, bits, divider, _p(out["means"], C.c_float),
                                   _p(out["order"], C.c_int32), _p(out["counts"], C.c_int32),
                                   _p(out["dict"], C.c_int16), _p(out["datten"], C.c_uint8),
                                   _p(out["dattr"], C.c_uint8), _p(out["entry"], C.c_int32))
    return out


def passthrough_dictionary(raw, bits=12, divider=6):
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    N, cs = raw.shape
    d = np.zeros((N, cs), np.int16)
    a = np.zeros(N, np.uint8)
    t = np.zeros(N, np.uint8)
    lib().gsc_ref_passthrough_dictionary(_p(raw, C.c_double), N, cs, bits, divider,
                                         _p(d, C.c_int16), _p(a, C.c_uint8), _p(t, C.c_uint8))
    return d, a, t


def knnfit_variants(dic, datten, bits=12, divider=6):
    dic = np.ascontiguousarray(dic, dtype=np.int16)
    datten = np.ascontiguousarray(datten, dtype=np.uint8)
    R, cs = dic.shape
    V = np.zeros((4 * R, cs), np.float32)
    lib().gsc_ref_knnfit_variants(_p(dic, C.c_int16), _p(datten, C.c_uint8), R, cs, bits, divider,
                                  _p(V, C.c_float))
    return V


def knnfit(dic, datten, raw, bits=12, divider=6):
    """-> dict(best, use, band, best_all, dbl_diff, epsilon)"""
    dic = np.ascontiguousarray(dic, dtype=np.int16)
    datten = np.ascontiguousarray(datten, dtype=np.uint8)
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    R, cs = dic.shape
    N = raw.shape[0]
    out = dict(best=np.zeros(N, np.int32), use=np.zeros(R, np.int32), band=np.zeros(N, np.int32),
               best_all=np.zeros(N, np.int32), dbl_diff=np.zeros(N, np.int32))
    eps = lib().gsc_ref_knnfit(_p(dic, C.c_int16), _p(datten, C.c_uint8), R, cs, bits, divider,
                               _p(raw, C.c_double), N, _p(out["best"], C.c_int32),
                               _p(out["use"], C.c_int32), _p(out["band"], C.c_int32),
                               _p(out["best_all"], C.c_int32), _p(out["dbl_diff"], C.c_int32))
    out["epsilon"] = float(eps)
    return out


def knnfit_kdtree(dic, datten, raw, bits=12, divider=6):
    """KNNFit through the ANN-style kd-tree (64 nearest rows, epsilon band) -> dict(best, use)"""
    dic = np.ascontiguousarray(dic, dtype=np.int16)
    datten = np.ascontiguousarray(datten, dtype=np.uint8)
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    R, cs = dic.shape
    N = raw.shape[0]
    out = dict(best=np.zeros(N, np.int32), use=np.zeros(R, np.int32))
    lib().gsc_ref_knnfit_kdtree(_p(dic, C.c_int16), _p(datten, C.c_uint8), R, cs, bits, divider, _p(raw, C.c_double), N,
                                _p(out["best"], C.c_int32), _p(out["use"], C.c_int32))
    return out


def finalize_dictionary(use):
    use = np.ascontiguousarray(use, dtype=np.int32)
    R = len(use)
    remap = np.zeros(R, np.int32)
    order = np.zeros(R, np.int32)
    n = lib().gsc_ref_finalize_dictionary(_p(use, C.c_int32), R, _p(remap, C.c_int32),
                                          _p(order, C.c_int32))
    return n, remap, order[:n]


def fpc_sort_desc(keys):
    keys = np.ascontiguousarray(keys, dtype=np.int32)
    perm = np.arange(len(keys), dtype=np.int32)
    lib().gsc_ref_fpc_sort_desc(_p(keys, C.c_int32), _p(perm, C.c_int32), len(keys))
    return perm


# ---- whole frame / file -----------------------------------------------------
def encode_frame(pcm, params=None, **kw) -> FrameResult:
    pcm = _pcm(pcm)
    Cn, S = pcm.shape
    p = params if params is not None else default_params(**kw)
    fo = _FrameOut()
    rc = lib().gsc_ref_encode_frame(_p(pcm, C.c_int16), C.c_int64(S), Cn, S, C.byref(p), C.byref(fo))
    assert rc == 0
    cs = p.chunk_size
    res = FrameResult(
        N=fo.N, R=fo.R, divider=fo.divider, passes=fo.passes, err=fo.err,
        dict=np.ctypeslib.as_array(fo.dict, (max(fo.R, 1), cs))[:fo.R].copy(),
        datten=np.ctypeslib.as_array(fo.datten, (max(fo.R, 1),))[:fo.R].copy(),
        index=np.ctypeslib.as_array(fo.index, (fo.N,)).copy(),
        attr=np.ctypeslib.as_array(fo.attr, (fo.N,)).copy(),
        overfull=fo.overfull)
    lib().gsc_ref_free_frame(C.byref(fo))
    return res


def _frame_struct(fr: FrameResult):
    """Keep numpy buffers alive via the returned tuple."""
    fo = _FrameOut()
    d = np.ascontiguousarray(fr.dict, np.int16)
    a = np.ascontiguousarray(fr.datten, np.uint8)
    i = np.ascontiguousarray(fr.index, np.int32)
    t = np.ascontiguousarray(fr.attr, np.uint8)
    fo.N, fo.R, fo.divider, fo.passes, fo.err = fr.N, fr.R, fr.divider, fr.passes, fr.err
    fo.dict, fo.datten = _p(d, C.c_int16), _p(a, C.c_uint8)
    fo.index, fo.attr = _p(i, C.c_int32), _p(t, C.c_uint8)
    return fo, (d, a, i, t)


def plan_frames(pcm, sample_rate, params=None, **kw):
    """pcm planar, already padded to a multiple of chunk_size. -> starts list"""
    pcm = _pcm(pcm)
    Cn, S = pcm.shape
    p = params if params is not None else default_params(**kw)
    cap = 1 + int(np.ceil(S / max(1, p.chunk_size)))
    cap = min(cap, 1 << 20)
    starts = np.zeros(cap, np.int64)
    n = lib().gsc_ref_plan_frames(_p(pcm, C.c_int16), C.c_int64(S), Cn, C.c_int64(S), sample_rate,
                                  C.byref(p), _p(starts, C.c_int64), cap)
    return starts[:n].copy()


def write_frame(fr: FrameResult, channels, cs, bits, sample_rate) -> bytes:
    fo, keep = _frame_struct(fr)
    n = lib().gsc_ref_write_frame(C.byref(fo), channels, cs, bits, sample_rate, None, C.c_int64(0))
    buf = np.zeros(n, np.uint8)
    lib().gsc_ref_write_frame(C.byref(fo), channels, cs, bits, sample_rate, _p(buf, C.c_uint8),
                              C.c_int64(n))
    del keep
    return buf.tobytes()


def decode(gsc: bytes):
    """-> (pcm planar int16 [C][S], sample_rate)"""
    g = np.frombuffer(gsc, np.uint8)
    ch = C.c_int(0)
    sr = C.c_int(0)
    n = lib().gsc_ref_decode(_p(g, C.c_uint8), C.c_int64(len(g)), None, C.c_int64(0), C.byref(ch),
                             C.byref(sr))
    if n < 0:
        raise ValueError("bad .gsc stream")
    out = np.zeros(n * max(ch.value, 1), np.int16)
    lib().gsc_ref_decode(_p(g, C.c_uint8), C.c_int64(len(g)), _p(out, C.c_int16), C.c_int64(n),
                         C.byref(ch), C.byref(sr))
    return out.reshape(n, ch.value).T.copy(), sr.value


def reconstruct_frame(fr: FrameResult, channels, S, cs, bits):
    fo, keep = _frame_struct(fr)
    out = np.zeros((channels, S), np.int16)
    lib().gsc_ref_reconstruct_frame(C.byref(fo), channels, S, cs, bits, _p(out, C.c_int16),
                                    C.c_int64(S))
    del keep
    return out


def psy_a_delta(a, b):
    a = np.ascontiguousarray(a, np.int16).ravel()
    b = np.ascontiguousarray(b, np.int16).ravel()
    return float(lib().gsc_ref_psy_a_delta(_p(a, C.c_int16), _p(b, C.c_int16), C.c_int64(len(a))))


def snr_db(ref, tst):
    a = np.ascontiguousarray(ref, np.int16).ravel()
    b = np.ascontiguousarray(tst, np.int16).ravel()
    return float(lib().gsc_ref_snr_db(_p(a, C.c_int16), _p(b, C.c_int16), C.c_int64(len(a))))


def encode_pcm(pcm, sample_rate, params=None, threads=1, **kw):
    """Whole-file encode (enc:2016-2024 Load..SaveGSC minus file I/O).
    pcm planar int16 [C][S]. -> (gsc bytes, list[FrameResult], starts, padded pcm)"""
    from concurrent.futures import ThreadPoolExecutor
    pcm = _pcm(pcm)
    p = params if params is not None else default_params(**kw)
    cs = p.chunk_size
    Cn, S0 = pcm.shape
    S = ((S0 - 1) // cs + 1) * cs                      # enc:1319
    if S != S0:
        pcm = np.concatenate([pcm, np.zeros((Cn, S - S0), np.int16)], axis=1)
    pcm = np.ascontiguousarray(pcm)
    starts = plan_frames(pcm, sample_rate, p)
    ends = list(starts[1:]) + [S]

    def one(k):
        return encode_frame(np.ascontiguousarray(pcm[:, starts[k]:ends[k]]), p)

    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            frames = list(ex.map(one, range(len(starts))))
    else:
        frames = [one(k) for k in range(len(starts))]
    blob = b"".join(write_frame(f, Cn, cs, p.chunk_bit_depth, sample_rate) for f in frames)
    return blob, frames, starts, pcm
