"""CPU: the host side (libgsc_host.so: planner, .gsc writer, decoder, reconstruction, CLI) against
the oracle, and the C-ABI surface of both shared libraries against their headers.  No compute
call into libgsc_cuda.so happens here (there is no GPU in this container)."""
import glob
import os
import re
import subprocess

import numpy as np
import pytest

from tests.golden_util import EXCERPTS as GOLD

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def host(lib_built):
    from soundchunks_b200 import build, host as h
    build.build_host()
    h.load_host_library()
    return h


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:gsc|gsch|yakmo|ann_kdtree)_[a-z0-9_]+)\s*\(", txt)))


def test_libgsc_cuda_exports_every_declared_symbol(lib_built):
    import ctypes
    import soundchunks_b200 as sc
    names = _declared("gsc_cuda.h")
    assert len(names) >= 35 and set(names) == set(sc.EXPORTS)
    lib = ctypes.CDLL(lib_built)
    for n in names:
        assert getattr(lib, n) is not None, n
    # the 9 symbols encoder.exe imports today (extern.pas:112-123)
    for n in ("yakmo_create", "yakmo_destroy", "yakmo_load_train_data", "yakmo_train_on_data", "yakmo_get_centroids",
              "ann_kdtree_create", "ann_kdtree_destroy", "ann_kdtree_search", "ann_kdtree_pri_search_multi"):
        assert n in names


def test_libgsc_host_exports_every_declared_symbol(host):
    names = _declared("gsc_host.h")
    assert set(names) == set(host.HOST_EXPORTS)
    lib = host.load_host_library()
    for n in names:
        assert getattr(lib, n) is not None, n


def test_build_rejects_spilling_online_shapes():
    """build.py treats a k_online shape that ptxas compiled with register spills as a build error."""
    from soundchunks_b200.build import online_spills
    log = ("ptxas info    : Compiling entry function '_Z8k_onlineILi8ELi16ELi256EEvPK8GscFrame' for 'sm_100a'\n"
           "ptxas info    : Function properties for _Z8k_onlineILi8ELi16ELi256EEvPK8GscFrame\n"
           "    24 bytes stack frame, 36 bytes spill stores, 32 bytes spill loads\n"
           "ptxas info    : Compiling entry function '_Z8k_onlineILi8ELi8ELi128EEvPK8GscFrame' for 'sm_100a'\n"
           "    0 bytes stack frame, 0 bytes spill stores, 0 bytes spill loads\n"
           "ptxas info    : Compiling entry function '_Z6k_seedILi8EEvPK8GscFrame' for 'sm_100a'\n"
           "    104 bytes stack frame, 172 bytes spill stores, 152 bytes spill loads\n")
    assert online_spills(log) == ["_Z8k_onlineILi8ELi16ELi256EEvPK8GscFrame"]


def test_no_gpu_fails_loudly(lib_built):
    """Without a usable sm_100 device every entry point must fail with a message, never compute on the CPU."""
    import soundchunks_b200 as sc
    if sc.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(sc.GscError, match="no usable sm_100"):
        sc.Context(0)
    lib = sc.load_library()
    assert not lib.yakmo_create(16, 1, 0, 1, 0, 0, 0)
    assert b"sm_100" in lib.gsc_last_error()
    from soundchunks_b200 import host as h
    with pytest.raises(sc.GscError, match="sm_100"):
        h.encode_pcm(np.zeros((1, 4096), np.int16), 44100)


def test_product_never_imports_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "soundchunks_b200", "**", "*"), recursive=True) + \
            glob.glob(os.path.join(ROOT, "host", "*.cpp")) + glob.glob(os.path.join(ROOT, "include", "*.h")):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
            txt = open(path, errors="ignore").read()
            assert "gsc_oracle" not in txt and "gsc_ref_" not in txt and "import oracle" not in txt, path


def test_option_parsing(host):
    o = host.default_options()
    assert (o.bitrate, o.precision, o.chunk_bit_depth, o.chunk_size, o.chunks_per_frame) == (-1, 3, 8, 4, 4096)
    assert o.frame_length_ms == 4000 and o.vfr == 1.0                       # enc:1486-1509
    o = host.default_options("-cbd12", "-cpf100", "-pr5", "-fl500", "-vfr2", "-br250", "-v")
    assert o.chunk_bit_depth == 12 and o.chunks_per_frame == 256            # clamp 256..4096, enc:1993
    assert o.precision == 5 and o.frame_length_ms == 500 and o.vfr == 1.0 and o.bitrate == 250 and o.verbose == 1
    assert host.default_options("-cb1").chunk_blend == 1 and host.default_options("-cbd12").chunk_blend == 0
    from soundchunks_b200 import GscError
    with pytest.raises(GscError):
        host.default_options("-zz9")


@pytest.mark.parametrize("channels,sr,seconds,cli", [(1, 44100, 9.7, ()), (2, 48000, 5.3, ("-fl1000",)),
                                                     (2, 44100, 6.0, ("-vfr0.3", "-fl1500")), (1, 32000, 2.0, ("-fl4000",))])
def test_planner_matches_oracle(host, oracle, channels, sr, seconds, cli):
    from soundchunks_b200.synth import synth_audio
    pcm = synth_audio(seconds, sr, channels, seed=17)[:, : int(seconds * sr) - 3]   # ragged length: padding path
    o = host.default_options("-cbd12", *cli)
    p = host.pad(pcm, o)
    assert p.shape[1] % 4 == 0 and p.shape[1] - pcm.shape[1] < 4 and not p[:, pcm.shape[1]:].any()
    starts, cpf = host.plan_frames(p, sr, o)
    ref = oracle.plan_frames(p, sr, oracle.default_params(chunk_bit_depth=12, frame_length_ms=o.frame_length_ms, vfr=o.vfr))
    assert np.array_equal(starts, ref) and cpf == 4096
    assert starts[0] == 0 and np.all(np.diff(starts) > 0) and np.all(starts % 4 == 0)


def test_bitrate_solver(host):
    from soundchunks_b200.synth import synth_audio
    pcm = synth_audio(8.0, 44100, 1, seed=2)
    o = host.default_options("-br128")
    _, cpf = host.plan_frames(host.pad(pcm, o), 44100, o)
    assert 1 <= cpf < 4096
    # enc:1337-1351 recomputed independently
    S, C, fc = pcm.shape[1], 1, 2
    want = None
    for k in range(4096, 0, -1):
        band = (S * C * (np.log2(k) + 3 + 1 + 1)) / (8 * 4)
        frame = (k * 4) * 8 / 8 + k * 4 / 8 + 16
        if np.rint(band * 0.8 + fc * frame) <= np.ceil((S / 44100) * (128 * 1024 / 8)) or k <= 1:
            want = k
            break
    assert cpf == want


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_writer_decoder_match_oracle_on_golden(host, oracle, path):
    import hashlib
    g = np.load(path)
    pcm, sr, bits = g["pcm"], int(g["sample_rate"]), int(g["bits"])
    fr = oracle.FrameResult(len(g["frame_index"]), int(g["frame_R"]), int(g["divider"]), int(g["passes"]), float(g["err"]),
                            g["frame_dict"], g["frame_datten"], g["frame_index"], g["frame_attr"], int(g["frame_overfull"]))
    blob = host.write_frame(fr, pcm.shape[0], 4, bits, sr)
    assert hashlib.sha256(blob).digest() == g["gsc_sha256"].tobytes()
    dec, sr2 = host.decode(blob)
    assert sr2 == sr and hashlib.sha256(dec.tobytes()).digest() == g["decoded_sha256"].tobytes()
    two, _ = host.decode(blob + blob)                                   # frames concatenate (enc:1208-1214)
    assert np.array_equal(two, np.concatenate([dec, dec], axis=1))
    rec = host.reconstruct_frame(fr, pcm.shape[0], pcm.shape[1], 4, bits)
    assert np.array_equal(rec, oracle.reconstruct_frame(fr, pcm.shape[0], pcm.shape[1], 4, bits))
    assert host.psy_a_delta(pcm, rec) == oracle.psy_a_delta(pcm, rec)
    with pytest.raises(ValueError):
        host.decode(blob[:40])


def test_cli_tools(host, oracle, tmp_path):
    g = np.load(GOLD[0])
    fr = oracle.FrameResult(len(g["frame_index"]), int(g["frame_R"]), int(g["divider"]), int(g["passes"]), float(g["err"]),
                            g["frame_dict"], g["frame_datten"], g["frame_index"], g["frame_attr"], 0)
    blob = oracle.write_frame(fr, 1, 4, int(g["bits"]), int(g["sample_rate"]))
    p = tmp_path / "a.gsc"
    p.write_bytes(blob)
    r = subprocess.run([host.DECODE_BIN, str(p), str(tmp_path / "a.wav")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L = host.load_host_library()
    import ctypes as C
    pcm, ch, n, sr = C.c_void_p(), C.c_int(0), C.c_int64(0), C.c_int(0)
    assert L.gsch_load_wav(str(tmp_path / "a.wav").encode(), C.byref(pcm), C.byref(ch), C.byref(n), C.byref(sr)) == 0
    got = np.ctypeslib.as_array(C.cast(pcm, C.POINTER(C.c_int16)), (ch.value, n.value)).copy()
    L.gsch_free(pcm)
    dec, _ = oracle.decode(blob)
    assert sr.value == int(g["sample_rate"]) and np.array_equal(got, dec)
    u = subprocess.run([host.ENCODE_BIN], capture_output=True, text=True)
    assert u.returncode == 0 and "-cbd" in u.stdout and "-cpf" in u.stdout
