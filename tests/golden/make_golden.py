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


def main():
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
