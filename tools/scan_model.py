"""CPU model of the window-parallel evaluation of yakmo's sequential float prefix sum.

Test infrastructure / design model for DESIGN.md 4.2 "Next" -- nothing in the product path imports this.

    r[j] = fl32(r[j-1] + a[j]),  r[-1] = +0            (yakmo init(): obj += up[j]; r[j] = obj)

`gsc_seq_prefix` (soundchunks_b200/csrc/gsc_kernels.cuh) evaluates this chain exactly with a block-wide scan, one
2048-element window after the other.  The model below evaluates the WINDOWS independently:

  1. every window is summarised, under a PREDICTED exponent E of the running sum, as
        (f0, f1)   increment of the integer significand S for start parity 0 / 1
        neg, pos   bounds of how far the partial sums can fall below / rise above the start
        bad        an element that cannot be handled in binade E (non-finite, >= 4 * 2^E)
  2. the summaries are chained in window order with the exact running sum: a summary is used iff the actual
     exponent equals the predicted one and  S - neg >= 2^23,  S + pos < 2^24  (then every partial sum stays in the
     binade and the composition is the sequential result); otherwise that window is evaluated element by element;
  3. windows that used their summary are replayed from their start value (in parallel on the GPU).

Within a binade every partial sum is S * ulp with 2^23 <= S < 2^24, and adding a_j gives
S + floor(a_j/ulp) + round bit, where the round bit depends on the fraction of a_j/ulp and, for an exact tie, on the
parity of S (round half to even): element j is a function parity -> increment, and those compose associatively.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
WIN = 2048


def seq_prefix(a: np.ndarray) -> np.ndarray:
    """The reference chain, one float32 addition after the other."""
    r = np.empty(len(a), F32)
    run = F32(0.0)
    for j, v in enumerate(a.astype(F32)):
        run = F32(run + v)
        r[j] = run
    return r


def _exp_field(x: F32) -> int:
    return int((np.array(x, F32).view(np.uint32) >> 23) & 0xFF)


def _classify(v: F32, e0: int):
    """(I, inc, tie, bad) of one element for a running sum in the binade with exponent field e0."""
    av = abs(float(v))
    huge = float(2.0 ** (e0 + 2 - 127))
    if not (av < huge):                      # also catches NaN / inf
        return 0, 0, 0, True
    aq = av * float(2.0 ** (150 - e0))       # a / ulp, exact (power-of-two scaling), < 2^25
    fl = float(np.floor(aq))
    g = aq - fl
    ni = int(fl)
    if v >= 0:
        return ni, int(g > 0.5), int(g == 0.5), False
    if g == 0.0:
        return -ni, 0, 0, False
    return -(ni + 1), int(g < 0.5), int(g == 0.5), False   # floor(-x) = -(n+1), fraction 1 - g


def _step(s: int, i: int, inc: int, tie: int) -> int:
    t = s + i
    return t + (inc | (tie & t & 1))


def summarise(a: np.ndarray, e0: int):
    """Window summary under the predicted exponent field e0: (f0, f1, neg, pos, bad)."""
    s0, s1, neg, pos = 0, 1, 0, 0
    for v in a.astype(F32):
        i, inc, tie, bad = _classify(v, e0)
        if bad:
            return 0, 0, 0, 0, True
        s0 = _step(s0, i, inc, tie)
        s1 = _step(s1, i, inc, tie)
        if i < 0:
            neg += -i
        else:
            pos += i + 1
    return s0, s1 - 1, neg, pos, False


def _replay(a: np.ndarray, out: np.ndarray, s: int, e0: int) -> None:
    for j, v in enumerate(a.astype(F32)):
        i, inc, tie, _ = _classify(v, e0)
        s = _step(s, i, inc, tie)
        out[j] = np.array((e0 << 23) | (s & 0x7FFFFF), np.uint32).view(F32)


def window_prefix(a: np.ndarray, predicted: list[int] | None = None, win: int = WIN):
    """-> (r, exponents, used): r == seq_prefix(a) bit for bit; exponents[w] = exponent field of the running sum at
    the start of window w (the prediction for the next call); used[w] = the summary of window w was accepted."""
    a = a.astype(F32)
    n = len(a)
    nw = (n + win - 1) // win
    r = np.empty(n, F32)
    summaries = []
    for w in range(nw):                      # step 1: independent of each other
        e = predicted[w] if predicted is not None and w < len(predicted) else -1
        summaries.append(summarise(a[w * win:(w + 1) * win], e) if 24 <= e <= 250 else None)
    run = F32(0.0)
    exps, used, starts = [], [], []
    for w in range(nw):                      # step 2: the chain
        seg = a[w * win:(w + 1) * win]
        bits = int(np.array(run, F32).view(np.uint32))
        e_act = (bits >> 23) & 0xFF
        exps.append(e_act)
        sm = summaries[w]
        ok = False
        if sm is not None and (bits >> 31) == 0 and e_act == predicted[w]:
            f0, f1, neg, pos, bad = sm
            s = (bits & 0x7FFFFF) | 0x800000
            ok = (not bad) and s - neg >= (1 << 23) and s + pos < (1 << 24)
        used.append(ok)
        if ok:
            starts.append(s)
            s_end = s + (f1 if (s & 1) else f0)
            run = np.array((e_act << 23) | (s_end & 0x7FFFFF), np.uint32).view(F32)[()]
        else:
            starts.append(None)
            for j, v in enumerate(seg):      # exact, element by element (the block-wide rounds on the GPU)
                run = F32(run + v)
                r[w * win + j] = run
    for w in range(nw):                      # step 3: independent of each other
        if used[w]:
            _replay(a[w * win:(w + 1) * win], r[w * win:(w + 1) * win], starts[w], exps[w])
    return r, exps, used
