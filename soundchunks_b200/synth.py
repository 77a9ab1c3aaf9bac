"""Seeded synthetic audio of the shape SURVEY.md 8(d) C5 names.

Per channel: 12 partials with log-uniform frequencies 40..12000 Hz, independent
slow (0.2..2 Hz) raised-cosine amplitude envelopes, plus pink-ish noise
(first-order low-passed white, about -30 dBFS), peak-normalised to -3 dBFS,
channels decorrelated by independent phases.  Pure tones / white noise are
avoided on purpose (k-means degenerates on them).
"""
from __future__ import annotations

import numpy as np


def synth_audio(seconds: float, sample_rate: int = 48000, channels: int = 2, seed: int = 1234) -> np.ndarray:
    """-> planar int16 [channels][samples]"""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sample_rate))
    t = np.arange(n, dtype=np.float64) / sample_rate
    out = np.zeros((channels, n), np.float64)
    for ch in range(channels):
        freqs = np.exp(rng.uniform(np.log(40.0), np.log(12000.0), 12))
        env_f = rng.uniform(0.2, 2.0, 12)
        ph = rng.uniform(0, 2 * np.pi, 12)
        env_ph = rng.uniform(0, 2 * np.pi, 12)
        amp = rng.uniform(0.2, 1.0, 12) / np.sqrt(freqs / 40.0)
        x = np.zeros(n, np.float64)
        for k in range(12):
            env = 0.5 * (1.0 - np.cos(2 * np.pi * env_f[k] * t + env_ph[k]))
            x += amp[k] * env * np.sin(2 * np.pi * freqs[k] * t + ph[k])
        w = rng.standard_normal(n)
        # first-order low-pass (pink-ish), vectorised as an IIR via lfilter-free recursion on blocks
        a = 0.97
        from scipy.signal import lfilter
        pn = lfilter([1 - a], [1, -a], w)
        pn *= (10 ** (-30 / 20)) / (np.sqrt(np.mean(pn ** 2)) + 1e-30)
        x = x / (np.max(np.abs(x)) + 1e-30) + pn
        out[ch] = x
    peak = np.max(np.abs(out)) + 1e-30
    out *= (10 ** (-3 / 20)) / peak
    return np.clip(np.round(out * 32767.0), -32768, 32767).astype(np.int16)


def synth_frames(n_frames: int, frame_seconds: float = 4.0, sample_rate: int = 48000, channels: int = 2,
                 seed: int = 1234, chunk_size: int = 4):
    """n independent frames (each its own seed) -> list of planar int16 [C][S], S a multiple of chunk_size."""
    S = int(round(frame_seconds * sample_rate)) // chunk_size * chunk_size
    return [np.ascontiguousarray(synth_audio(frame_seconds, sample_rate, channels, seed + 7919 * i)[:, :S])
            for i in range(n_frames)]
