"""Generates tests/golden/*.npz: short excerpts of the reference's own test audio
(/root/reference/{my_test,lame_test,opus_test}) and what the CPU oracle computes on them, stage by
stage.  Run HERE (the reference tree is not on the GPU box):  python tests/golden/make_golden.py

The reference has no golden vectors of its own and cannot be built, so these files
pin (a) the oracle against accidental change and (b) the CUDA path on the reference's real inputs.
"""
import hashlib
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gsc_oracle as O  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, wav, start s, seconds, bits, K
    ("test_k256_12", "my_test/test.wav", 2.0, 1.0, 12, 256),
    ("iron_k256_8", "lame_test/iron.wav", 10.0, 0.8, 8, 256),
    ("hihat_k512_12", "lame_test/hihat.wav", 0.5, 1.0, 12, 512),
    ("stereo48_k256_8", "opus_test/mo_b_44_2.wav", 3.0, 0.5, 8, 256),
    ("velvet_k4096_12", "lame_test/velvet.wav", 5.0, 1.5, 12, 4096),
]


def load_wav(path):
    b = open(path, "rb").read()
    ch = struct.unpack("<H", b[0x16:0x18])[0]
    sr = struct.unpack("<i", b[0x18:0x1c])[0]
    d = np.frombuffer(b[44:44 + (len(b) - 44) // (2 * ch) * 2 * ch], np.int16).reshape(-1, ch).T
    return np.ascontiguousarray(d), sr


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def canon(a):
    """float32 array with every NaN replaced by the canonical quiet NaN (x86 and CUDA produce different payloads)."""
    a = np.array(a, np.float32, copy=True)
    a.view(np.uint32)[np.isnan(a)] = 0x7FC00000
    return a


def run_case(pcm, sr, bits, K):
    """Every stage of the oracle on one frame; returns a dict of arrays."""
    cs = 4
    div, v = O.find_attenuation_divider(pcm, cs, bits, return_v=True)
    raw, attr, atten, feat, dst = O.make_chunks(pcm, cs, bits, div)
    cen0, lab0, seeds = O.yakmo(feat, K)
    cen, labels, passes, err = O.knn_scan_reduce(feat, cen0, 3, 100)
    d = O.build_dictionary(labels, raw, attr, K, bits, div)
    fit = O.knnfit(d["dict"], d["datten"], raw, bits, div)
    fr = O.encode_frame(pcm, chunk_bit_depth=bits, chunks_per_frame=K)
    blob = O.write_frame(fr, pcm.shape[0], cs, bits, sr)
    dec, _ = O.decode(blob)
    return dict(divider=np.int32(div), divider_v=v, attr=attr, atten=atten, feat_dct=feat[:, :4].copy(),
                feat_cep=feat[:, 4:].copy(), seeds=seeds, cen0=cen0, passes=np.int32(passes), err=np.float64(err),
                cen=cen, labels=labels, dict_means=d["means"], dict_q=d["dict"], dict_atten=d["datten"],
                dict_counts=d["counts"], best=fit["best_all"], band=fit["band"],
                frame_R=np.int32(fr.R), frame_dict=fr.dict, frame_datten=fr.datten, frame_index=fr.index,
                frame_attr=fr.attr, frame_overfull=np.int32(fr.overfull),
                gsc_sha256=np.frombuffer(hashlib.sha256(blob).digest(), np.uint8), gsc_len=np.int64(len(blob)),
                decoded_sha256=np.frombuffer(hashlib.sha256(dec.tobytes()).digest(), np.uint8),
                snr_db=np.float64(O.snr_db(pcm, dec)))


# Full frames as the reference's planner cuts them (enc:1374-1425), BASELINE.json configs[0..2] as written:
#   (wav, frame indexes, bits, K).  Only the PCM, scalars and the SHA-256 of every stage output are stored.
FULL = [
    ("my_test/test.wav", [0, 1, 2], 12, 4096),          # configs[0]: all three frames
    ("lame_test/iron.wav", [3], 12, 4096),              # configs[1]: one full frame per track
    ("lame_test/hihat.wav", [0], 12, 4096),             #   (near-silent stretches: wide epsilon bands)
    ("lame_test/velvet.wav", [1], 12, 4096),
    ("lame_test/castanets.wav", [0], 12, 4096),
    ("lame_test/applaud.wav", [1], 12, 4096),
    ("lame_test/fatboy.wav", [0], 12, 4096),
    ("lame_test/pipes.wav", [2], 12, 4096),
    ("lame_test/60.wav", [0], 12, 4096),
    ("lame_test/testsignal2.wav", [0], 12, 4096),
    ("lame_test/spahm.wav", [0], 12, 4096),
    ("opus_test/mo_b_44_2.wav", [0, 2], 8, 256),        # configs[2] stand-ins (music_orig.wav is not shipped)
    ("opus_test/mo_62_32.wav", [0, 7], 8, 256),
]


def run_full(pcm, sr, bits, K):
    """Stage outputs of the oracle on one full frame -> (scalars dict, {stage: sha256 hex})."""
    cs = 4
    div, v = O.find_attenuation_divider(pcm, cs, bits, return_v=True)
    raw, attr, atten, feat, dst = O.make_chunks(pcm, cs, bits, div)
    cen0, lab0, seeds = O.yakmo(feat, K)
    cen, labels, passes, err = O.knn_scan_reduce(feat, cen0, 3, 100)
    d = O.build_dictionary(labels, raw, attr, K, bits, div)
    fit = O.knnfit(d["dict"], d["datten"], raw, bits, div)
    use_all = np.bincount(fit["best_all"] >> 2, minlength=K).astype(np.int32)
    fr = O.encode_frame(pcm, chunk_bit_depth=bits, chunks_per_frame=K, band_all=1)
    fr64 = O.encode_frame(pcm, chunk_bit_depth=bits, chunks_per_frame=K, band_all=0)
    blob = O.write_frame(fr, pcm.shape[0], cs, bits, sr)
    blob64 = O.write_frame(fr64, pcm.shape[0], cs, bits, sr)
    dec, _ = O.decode(blob)
    band = fit["band"]
    scal = dict(divider=int(div), passes=int(passes), err=float(err), R=int(fr.R), overfull=int(fr.overfull),
                N=int(len(feat)), band_max=int(band.max()), band_gt4=int((band > 4).sum()),
                band_rule_differs=int((fit["best"] != fit["best_all"]).sum()),
                dbl_diff=int(fit["dbl_diff"].sum()), snr_db=float(O.snr_db(pcm, dec)),
                psy_a_delta=float(O.psy_a_delta(pcm, dec)),
                # encoder-side reconstruction (enc:487-522), what the encoder prints as PsyADelta (enc:2026)
                psy_a_delta_enc=float(O.psy_a_delta(pcm, O.reconstruct_frame(fr, pcm.shape[0], pcm.shape[1], cs, bits))),
                gsc_len=len(blob))
    hs = dict(divider_v=sha(v), attr=sha(attr), atten=sha(atten), feat=sha(feat), seeds=sha(seeds), cen0=sha(canon(cen0)),
              labels=sha(labels), cen=sha(canon(cen)), dict_means=sha(d["means"]), dict_q=sha(d["dict"]),
              dict_atten=sha(d["datten"]), dict_counts=sha(d["counts"]), best_all=sha(fit["best_all"]),
              use_all=sha(use_all), band=sha(band), frame_dict=sha(fr.dict), frame_datten=sha(fr.datten),
              frame_index=sha(fr.index), frame_attr=sha(fr.attr), gsc=hashlib.sha256(blob).hexdigest(),
              gsc_bucket64=hashlib.sha256(blob64).hexdigest(), decoded=sha(dec))
    return scal, hs, np.bincount(np.minimum(band, 65), minlength=66).astype(np.int64)


def lattice_noise(seed=5, n=44100):
    """Synthetic frame that fills the epsilon band: two thirds of it is noise on the integer lattice {-2..2} (a few
    hundred distinct, tightly packed chunks -> far more than 64 variant rows within epsilon of a query, the case
    where ANN's 64-row bucket truncates, enc:917), one third a loud burst so that the k-means has work to do."""
    rng = np.random.default_rng(seed)
    x = rng.integers(-2, 3, n).astype(np.float64)
    x[:n // 3] += 2000 * np.sin(np.arange(n // 3) * 0.05) * np.hanning(n // 3) + rng.normal(0, 300, n // 3)
    q = np.clip(np.round(x), -32768, 32767).astype(np.int16)[None, :]
    return np.ascontiguousarray(q[:, :n // 4 * 4])


def full_cases():
    """-> list of (name, wav, frame index, bits, K, pcm, sample_rate)"""
    out = [("full_latticenoise_k4096_12", "synthetic lattice_noise(seed=5)", 0, 12, 4096, lattice_noise(), 44100, 1),
           ("full_latticenoise_k256_8", "synthetic lattice_noise(seed=6)", 0, 8, 256, lattice_noise(6), 44100, 1)]
    for wav, idxs, bits, K in FULL:
        pcm, sr = load_wav(os.path.join(REF, wav))
        S0 = pcm.shape[1]
        S = ((S0 - 1) // 4 + 1) * 4
        if S != S0:
            pcm = np.concatenate([pcm, np.zeros((pcm.shape[0], S - S0), np.int16)], axis=1)
        pcm = np.ascontiguousarray(pcm)
        starts = list(O.plan_frames(pcm, sr, chunk_bit_depth=bits, chunks_per_frame=K))
        ends = starts[1:] + [S]
        for k in idxs:
            k = min(k, len(starts) - 1)
            stem = os.path.splitext(os.path.basename(wav))[0]
            out.append((f"full_{stem}_f{k}_k{K}_{bits}", wav, k, bits, K,
                        np.ascontiguousarray(pcm[:, starts[k]:ends[k]]), sr, len(starts)))
    return out


def main_full():
    import json
    from concurrent.futures import ThreadPoolExecutor
    cases = full_cases()

    def one(c):
        name, wav, k, bits, K, pcm, sr, nfr = c
        scal, hs, hist = run_full(pcm, sr, bits, K)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), pcm=pcm, sample_rate=np.int32(sr), bits=np.int32(bits),
                            K=np.int32(K), source=np.array(f"{wav} frame {k} of {nfr} (reference planner)"),
                            scalars=np.array(json.dumps(scal)), hashes=np.array(json.dumps(hs)), band_hist=hist)
        return name, pcm.shape, scal
    with ThreadPoolExecutor(8) as ex:
        for name, shape, scal in ex.map(one, cases):
            print(name, shape, scal, flush=True)


def main():
    if "--full-only" not in sys.argv:
        main_excerpts()
    if "--excerpts-only" not in sys.argv:
        main_full()


def main_excerpts():
    for name, wav, t0, secs, bits, K in CASES:
        pcm, sr = load_wav(os.path.join(REF, wav))
        a = int(t0 * sr) // 4 * 4
        n = int(secs * sr) // 4 * 4
        ex = np.ascontiguousarray(pcm[:, a:a + n])
        res = run_case(ex, sr, bits, K)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), pcm=ex, sample_rate=np.int32(sr), bits=np.int32(bits),
                            K=np.int32(K), source=np.array(f"{wav} [{a}:{a + n}]"), **res)
        print(name, ex.shape, "divider", int(res["divider"]), "passes", int(res["passes"]), "R", int(res["frame_R"]),
              "overfull", int(res["frame_overfull"]), "snr", round(float(res["snr_db"]), 2), "bytes", int(res["gsc_len"]))


if __name__ == "__main__":
    main()
