"""CPU: the oracle against known answers worked out from the Pascal text and against the committed
golden fixtures (tests/golden/*.npz, the reference's own test audio; make_golden.py made them)."""
import hashlib
import os

import numpy as np
import pytest

from tests.golden_util import EXCERPTS as GOLD, FULL, check_stage, load_full, sha


def test_fixtures_present():
    assert len(GOLD) >= 5 and len(FULL) >= 16
    # BASELINE.json configs[0..2] as written: every frame of my_test/test.wav, >= 8 lame_test tracks at K = 4096 /
    # 12 bits, the opus_test stand-ins at K = 256 / 8 bits; wide epsilon bands and an overfull band are covered
    names = [os.path.basename(p) for p in FULL]
    assert sum(n.startswith("full_test_f") for n in names) == 3
    assert sum(n.endswith("_k4096_12.npz") for n in names) >= 12 and sum(n.endswith("_k256_8.npz") for n in names) >= 4
    sc = [load_full(p)[4] for p in FULL]
    assert max(s["band_max"] for s in sc) > 64 and any(s["overfull"] > 0 for s in sc)
    assert any(4 < s["band_max"] <= 64 and s["overfull"] == 0 for s in sc)


# ---- known answers (hand-derived from enc:1638-1698, dec:6,88-96) ---------------------------
def test_quantiser_known_answers(oracle):
    O = oracle
    # coeff(a) = 1 + sum_{i<=a} i*Law ; obd(12) = 2047, obd(8) = 127           enc:1654-1659
    assert O.quant(0.5, 12, 0, False, 1.0 / 6) == 1024          # round_half_even(1023.5) = 1024
    assert O.quant(0.25, 8, 0, False, 1.0) == 32                # round_half_even(31.75)
    assert O.quant(2.5 / 127, 8, 0, False, 1.0) == 2            # 2.5 -> 2 (half to even), not 3
    assert O.quant(1.0, 12, 0, False, 0.5) == 2046              # clamp to obd-1               enc:1661
    assert O.quant(-1.0, 12, 0, False, 0.5) == -2046
    assert O.quant(0.5, 12, 2, False, 0.5) == 2046              # coeff = 1 + 0.5 + 1 = 2.5 -> 2559 -> clamp
    assert O.quant(0.1, 12, 2, True, 0.5) == -512               # 0.1*2047*2.5 = 511.75 -> 512, negated
    assert O.dequant(1024, 12, 0, False, 0.5) == 1024 / 2047.0
    assert O.dequant(-512, 12, 2, True, 0.5) == 512 / (2047.0 * 2.5)
    assert O.dequant(2046, 8, 0, False, 0.5) == 1.0             # clamp to +-1                 enc:1679
    # attenuation: largest a with hi*coeff(a) <= 32767                                        enc:1687-1697
    assert O.attenuation([0.0, 0.0, 0.0, 0.0], 1.0 / 6) == 15
    assert O.attenuation([1.0, 0.0, 0.0, 0.0], 1.0 / 6) == 0


def test_attenuation_boundaries(oracle):
    O = oracle
    # hi = ceil(|x*32767|); r counts up while hi*(1 + sum_{i<=r} i*law) <= 32767
    law = 1.0
    # x = 0.25 -> hi = 8192: coeff(1)=2 -> 16384 ok, coeff(2)=4 -> 32768 > 32767 stop at r=2 -> a=1
    assert O.attenuation([0.25, 0, 0, 0], law) == 1
    # x = 8191/32767 -> hi = 8191: coeff(2)=4 -> 32764 ok, coeff(3)=7 -> stop -> a=2
    assert O.attenuation([8191 / 32767.0, 0, 0, 0], law) == 2
    assert O.attenuation([0.5, -0.25, 0.0, 0.1], law) == 0      # hi = 16384: coeff(1) = 2 -> 32768 > 32767


def test_decoder_constants():
    # CAttrMul = round(32768 * 32767 / 2047)                                                  dec:6
    assert int(np.rint(32768.0 * (32767.0 / 2047.0))) == 524528


def test_fpc_quicksort_order(oracle):
    # FreePascal fgl QuickSort is unstable: the tie order is part of the stream (enc:865, 974)
    keys = np.array([3, 1, 3, 2, 1, 3, 0, 2], np.int32)
    perm = oracle.fpc_sort_desc(keys)
    assert sorted(perm.tolist()) == list(range(8))
    assert np.all(np.diff(keys[perm]) <= 0)
    n, remap, order = oracle.finalize_dictionary(np.array([2, 0, 5, 0, 2], np.int32))
    assert n == 3 and order.tolist()[0] == 2 and remap[1] == -1 and remap[3] == -1


# ---- golden fixtures ---------------------------------------------------------------------------
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_reproduces_golden(oracle, path):
    O = oracle
    g = np.load(path)
    pcm, sr, bits, K = g["pcm"], int(g["sample_rate"]), int(g["bits"]), int(g["K"])
    div, v = O.find_attenuation_divider(pcm, 4, bits, return_v=True)
    assert div == int(g["divider"]) and np.array_equal(v, g["divider_v"])
    raw, attr, atten, feat, dst = O.make_chunks(pcm, 4, bits, div)
    assert np.array_equal(attr, g["attr"]) and np.array_equal(atten, g["atten"])
    assert np.array_equal(feat[:, :4].view(np.uint32), g["feat_dct"].view(np.uint32))
    assert np.array_equal(feat[:, 4:].view(np.uint32), g["feat_cep"].view(np.uint32))
    big = K > 1024                      # keep the CPU suite short: the big case checks the cheap stages + the stream
    if not big:
        cen0, lab0, seeds = O.yakmo(feat, K)
        assert np.array_equal(seeds, g["seeds"])
        assert np.array_equal(np.nan_to_num(cen0).view(np.uint32), np.nan_to_num(g["cen0"]).view(np.uint32))
        cen, labels, passes, err = O.knn_scan_reduce(feat, cen0, 3, 100)
        assert passes == int(g["passes"]) and err == float(g["err"])
        assert np.array_equal(labels, g["labels"]) and np.array_equal(cen.view(np.uint32), g["cen"].view(np.uint32))
    d = O.build_dictionary(g["labels"], raw, attr, K, bits, div)
    assert np.array_equal(d["dict"], g["dict_q"]) and np.array_equal(d["datten"], g["dict_atten"])
    assert np.array_equal(d["counts"], g["dict_counts"])
    fit = O.knnfit(d["dict"], d["datten"], raw, bits, div)
    assert np.array_equal(fit["best_all"], g["best"]) and np.array_equal(fit["band"], g["band"])
    # stream: writer + decoder on the stored frame result
    fr = O.FrameResult(len(g["frame_index"]), int(g["frame_R"]), div, int(g["passes"]), float(g["err"]), g["frame_dict"],
                       g["frame_datten"], g["frame_index"], g["frame_attr"], int(g["frame_overfull"]))
    blob = O.write_frame(fr, pcm.shape[0], 4, bits, sr)
    assert len(blob) == int(g["gsc_len"])
    assert hashlib.sha256(blob).digest() == g["gsc_sha256"].tobytes()
    dec, sr2 = O.decode(blob)
    assert sr2 == sr and hashlib.sha256(dec.tobytes()).digest() == g["decoded_sha256"].tobytes()
    assert O.snr_db(pcm, dec) == float(g["snr_db"])


def _oracle_full(O, path):
    pcm, sr, bits, K, scal, hs, hist = load_full(path)
    div, v = O.find_attenuation_divider(pcm, 4, bits, return_v=True)
    assert div == scal["divider"]
    check_stage("divider_v", v, hs)
    raw, attr, atten, feat, dst = O.make_chunks(pcm, 4, bits, div)
    for n, a in (("attr", attr), ("atten", atten), ("feat", feat)):
        check_stage(n, a, hs)
    fr = O.encode_frame(pcm, chunk_bit_depth=bits, chunks_per_frame=K, band_all=1)
    assert (fr.N, fr.R, fr.passes, fr.err, fr.overfull) == (scal["N"], scal["R"], scal["passes"], scal["err"], scal["overfull"])
    for n, a in (("frame_dict", fr.dict), ("frame_datten", fr.datten), ("frame_index", fr.index), ("frame_attr", fr.attr)):
        check_stage(n, a, hs)
    blob = O.write_frame(fr, pcm.shape[0], 4, bits, sr)
    assert hashlib.sha256(blob).hexdigest() == hs["gsc"] and len(blob) == scal["gsc_len"]
    dec, _ = O.decode(blob)
    assert sha(dec) == hs["decoded"] and O.snr_db(pcm, dec) == scal["snr_db"] and O.psy_a_delta(pcm, dec) == scal["psy_a_delta"]
    return os.path.basename(path)


def test_oracle_reproduces_full_frames(oracle):
    """The oracle on every full-frame fixture (one host thread per frame: ~1 minute on 8 cores)."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(os.cpu_count() or 4) as ex:
        done = list(ex.map(lambda p: _oracle_full(oracle, p), FULL))
    assert len(done) == len(FULL)


def test_round_trip_properties(oracle):
    """encode -> write -> decode on a synthetic frame: shapes, index range, passthrough identity."""
    from soundchunks_b200.synth import synth_audio
    O = oracle
    pcm = np.ascontiguousarray(synth_audio(0.2, 44100, 2, seed=3)[:, :8000])
    fr = O.encode_frame(pcm, chunk_bit_depth=12, chunks_per_frame=256)
    assert fr.N == 2 * 2000 and 1 <= fr.R <= 256 and fr.index.max() < fr.R and fr.attr.max() <= 3
    dec, _ = O.decode(O.write_frame(fr, 2, 4, 12, 44100))
    assert dec.shape == pcm.shape and O.snr_db(pcm, dec) > 15.0
    # N <= K: every chunk is its own entry (enc:891-912): the stream is the 12-bit quantisation of the input
    tiny = np.ascontiguousarray(pcm[:, :400])
    ft = O.encode_frame(tiny, chunk_bit_depth=12, chunks_per_frame=256)
    assert ft.passes == 0 and ft.R <= ft.N == 200
    dt, _ = O.decode(O.write_frame(ft, 2, 4, 12, 44100))
    assert O.snr_db(tiny, dt) > 40.0
    # empty / silent input: one entry, all indexes 0
    z = np.zeros((1, 4096), np.int16)
    fz = O.encode_frame(z, chunk_bit_depth=8, chunks_per_frame=256)
    dz, _ = O.decode(O.write_frame(fz, 1, 4, 8, 44100))
    assert not dz.any()
