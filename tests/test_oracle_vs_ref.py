"""CPU-only: pins the oracle restatement against the REFERENCE'S OWN CODE.

(1) oracle/_ref/ref_vs_oracle: the reference's lib/<block>_impl.cc sources, compiled where they lie
    against header-only stand-ins for GNU Radio/Boost/Eigen/VOLK (oracle/build_ref.sh), driven block by
    block on seeded inputs; every output, tag, consume count and published message must equal the
    oracle's bit for bit (~1000 checks: background ring, tx interleave, stale-TX skip, transpose
    back-pressure, estimator ties/edges/gates, fft_peak_detect branches, cp remover, target_simulator).
(2) tests/golden/*.npz: outputs of that reference build (tests/golden/make_golden.py) -- the oracle must
    reproduce them exactly.  These fixtures travel to the GPU box where /root/reference is absent."""
import glob
import os
import subprocess

import numpy as np
import pytest

from mimo_ofdm_jrc import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "ref_vs_oracle")
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def test_reference_sources_agree_with_oracle():
    if not os.path.exists(REF_EXE):
        if os.path.isdir("/root/reference"):
            subprocess.run(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")], check=True, capture_output=True)
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    r = subprocess.run([REF_EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "ORACLE PINNED AGAINST REFERENCE SOURCES" in r.stdout and " 0 failed" in r.stdout


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_reference_golden_vectors(orc, path):
    g = np.load(path)
    T, R, S, N, IR, IA = (int(v) for v in g["cfg"])
    est = synth.default_estimator_params(N, T * R, IR, IA)
    m, _, d = orc.chain_batch(g["rx"], g["tx"], N, T, R, S, IR, IA, est)
    assert np.array_equal(d["range_idx"], g["range_idx"]) and np.array_equal(d["angle_idx"], g["angle_idx"])
    assert np.array_equal(d["peak_power"], g["peak_power"]) and np.array_equal(d["snr_db"], g["snr_db"])
    assert np.array_equal(m[0], g["map0"])
    assert np.array_equal(m.reshape(len(d), -1).max(axis=1), g["map_max"])
    # the two KATs of SURVEY.md section 4 embedded in the fixtures
    exp0 = synth.expected_peak(10.0, 0.0, N, IR, T * R, IA)
    assert (int(g["range_idx"][0]), int(g["angle_idx"][0])) == exp0
    assert len(GOLDEN) >= 2


def test_capture_line_golden_is_what_the_reference_block_writes(orc):
    """tests/golden/c1_capture_line.txt = the reference block's own capture_radar_data() output for the golden frame
    (regenerated here from /root/reference through oracle/_ref; a process of its own, see make_golden.py), and the values in
    it are the oracle's channel estimate printed with 7 significant digits."""
    lib = os.path.join(ROOT, "oracle", "_ref", "libjrc_ref.so")
    if not os.path.exists(lib):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    import sys
    gdir = os.path.join(ROOT, "tests", "golden")
    frame, path = os.path.join(gdir, "c1_frame0.c64"), "/tmp/jrc_ref_capture_test.csv"
    if os.path.exists(path):
        os.remove(path)
    code = ("import ctypes as C, sys\n"
            "buf = open(sys.argv[2], 'rb').read()\n"
            "lib = C.CDLL(sys.argv[1]); lib.ref_capture.restype = None\n"
            "lib.ref_capture.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_char_p]\n"
            "b = C.create_string_buffer(buf, len(buf)); a = C.addressof(b)\n"
            "lib.ref_capture(a, a + 8 * 4 * 4 * 64, 64, 4, 2, 4, 0, 0, sys.argv[3].encode())\n")
    subprocess.run([sys.executable, "-c", code, lib, frame, path], check=True, capture_output=True)
    golden = open(os.path.join(gdir, "c1_capture_line.txt")).read()
    assert open(path).read().split(", ", 1)[1] == golden
    # same numbers as the oracle's conj-MAC
    f = np.fromfile(frame, dtype=np.complex64).reshape(6, 4 * 64)
    rad = orc.Radar(64, 4, 2, 4, 0, False, False, 1, 1, False)
    H = rad.work(list(f[:4]), list(f[4:])).ravel()
    fields = golden.split(":", 1)[1].strip().rstrip(";").split(";")
    assert golden.startswith("4, 2, 64:") and len(fields) == H.size and golden.endswith(";\n\n")
    got = np.array([complex(*map(float, x.strip("()").split(","))) for x in fields])
    assert np.allclose(got.real, H.real, rtol=1e-6, atol=1e-12) and np.allclose(got.imag, H.imag, rtol=1e-6, atol=1e-12)
