"""GPU: the benchmark shape itself against the oracle.

bench.py runs 4 s stereo 44.1 kHz frames (N = 88,200 chunks), ChunkCount = 4096, 12-bit chunks, up to 100 online
passes, hundreds of frames per launch on two internal streams.  These tests compare exactly that -- whole frames
through gsc_encode_frames in a two-lane batch -- with oracle.encode_frame (enc:1433-1447), and check that copies of
one frame spread over a batch come out identical (a result that depended on timing or on the SM a frame landed on
would show up here)."""
import hashlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same(r, ref):
    return ((r.N, r.R, r.divider, r.passes, r.err, r.overfull) == (ref.N, ref.R, ref.divider, ref.passes, ref.err, ref.overfull)
            and np.array_equal(r.dict, ref.dict) and np.array_equal(r.datten, ref.datten)
            and np.array_equal(r.index, ref.index) and np.array_equal(r.attr, ref.attr))


def test_bench_shape_two_lane_batch_matches_oracle(ctx, oracle):
    from soundchunks_b200.synth import synth_frames
    distinct = synth_frames(4, 4.0, 44100, 2, seed=4242)
    frames = [distinct[i % 4] for i in range(32)]               # 32 frames -> two lanes (even / odd frames)
    assert frames[0].shape == (2, 176400)
    res = ctx.encode_frames(frames, chunk_bit_depth=12, chunks_per_frame=4096, max_passes=100)
    blob, sizes = ctx.fetch_stream(32, 44100)
    assert res[0].N == 88200 and max(r.passes for r in res) > 20
    # frames 0 (lane A), 1 and 3 (lane B) against the oracle, one host thread each
    with ThreadPoolExecutor(3) as ex:
        refs = list(ex.map(lambda i: oracle.encode_frame(frames[i], chunk_bit_depth=12, chunks_per_frame=4096,
                                                         max_passes=100, band_all=1), [0, 1, 3]))
    offs = np.concatenate([[0], np.cumsum(sizes)])
    for i, ref in zip([0, 1, 3], refs):
        assert _same(res[i], ref), f"frame {i} differs from the oracle"
        want = oracle.write_frame(ref, 2, 4, 12, 44100)
        assert blob[offs[i]:offs[i + 1]] == want
    # every copy of a frame, whatever lane / SM / neighbours it had, is identical to the first
    for i in range(4, 32):
        assert _same(res[i], res[i % 4]), f"copy {i} of frame {i % 4} differs"
        assert blob[offs[i]:offs[i + 1]] == blob[offs[i % 4]:offs[i % 4 + 1]]


def test_copies_under_load_are_identical(ctx):
    """296 one-second frames (two waves of CTAs on 148 SMs, both lanes busy): 8 distinct inputs x 37 copies."""
    from soundchunks_b200.synth import synth_frames
    distinct = synth_frames(8, 1.0, 44100, 2, seed=99)
    frames = [distinct[i % 8] for i in range(296)]
    blob, sizes = ctx.encode_to_stream(frames, 44100, chunk_bit_depth=12, chunks_per_frame=4096)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    first = [hashlib.sha256(blob[offs[i]:offs[i + 1]]).digest() for i in range(8)]
    for i in range(8, 296):
        assert hashlib.sha256(blob[offs[i]:offs[i + 1]]).digest() == first[i % 8], f"copy {i} differs"


def test_library_log_is_the_oracles_log(ctx, oracle):
    """gsc_log_cr on the device == the same routine on the host, value by value, over the whole double range."""
    rng = np.random.default_rng(3)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 200000)), 1.0 + rng.uniform(-1e-6, 1e-6, 50000),
                        rng.uniform(1e-13, 64.0, 250000), [1.0, 2.0, 0.5, 1e-12, 5e-324, 1.7976931348623157e308]])
    got = ctx.log_array(x)
    L = oracle.lib()
    want = np.array([L.gsc_ref_log_cr(float(v)) for v in x[:20000]] + [L.gsc_ref_log_cr(float(v)) for v in x[-6:]])
    assert np.array_equal(got[:20000].view(np.uint64), want[:20000].view(np.uint64))
    assert np.array_equal(got[-6:].view(np.uint64), want[-6:].view(np.uint64))
    # and against glibc within one ulp everywhere (a gross error in the untested part would show here)
    ref = np.log(x)
    assert np.max(np.abs(got.view(np.int64) - ref.view(np.int64))) <= 1
