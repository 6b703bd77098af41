#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE's own block code.

Run in the build container (needs /root/reference and oracle/_ref/libjrc_ref.so, i.e.
`bash oracle/build_ref.sh`).  The outputs are produced by ref_chain_batch(): the reference's
mimo_ofdm_radar, matrix_transpose and range_angle_estimator work() functions compiled from
/root/reference/lib, with the oracle's fft_vcc arithmetic standing in for the two stock GNU Radio
FFT blocks (the reference tree has no FFT on this path).  The fixtures travel to the GPU box, where
/root/reference does not exist.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "python", "mimo_ofdm_jrc"))
import synth  # noqa: E402  (imported as a plain module: the product package needs no GPU for synth)
from oracle import orc  # noqa: E402


def ref_chain(rx, tx, cfg, est):
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libjrc_ref.so"))
    lib.ref_chain_batch.argtypes = [C.POINTER(orc.ChainCfg), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
    rx, tx = orc.c64(rx), orc.c64(tx)
    n = rx.shape[0]
    Nr, Na = cfg["N"] * cfg["IR"], cfg["T"] * cfg["R"] * cfg["IA"]
    rb, ab = orc.f32(est["range_bins"]), orc.f32(est["angle_bins"])
    c = orc.ChainCfg(cfg["N"], cfg["T"], cfg["R"], cfg["S"], 0, cfg["IR"], cfg["IA"], 0, rb.ctypes.data, ab.ctypes.data,
                     est["noise_discard_range_m"], est["noise_discard_angle_deg"], est["snr_threshold"], est["power_threshold"])
    m = np.empty((n, Nr, Na), np.float32)
    d = np.zeros(n, orc.DET_DTYPE)
    lib.ref_chain_batch(C.byref(c), rx.ctypes.data, tx.ctypes.data, 1, n, 0, m.ctypes.data, None, d.ctypes.data)
    return m, d


def main():
    cases = {"c1_shipped": dict(T=4, R=2, S=4, N=64, IR=8, IA=16), "c2_bench": dict(T=4, R=2, S=4, N=64, IR=16, IA=8)}
    for name, cfg in cases.items():
        rng = np.random.default_rng(2022)
        tx = synth.tx_symbols(cfg["T"], cfg["S"], cfg["N"])
        n = 8
        r, a, amp = synth.random_scene(rng, n, 2, cfg["N"], amp_db_span=10.0)
        r[0, 0], a[0, 0] = 10.0, 0.0          # SURVEY.md section 4 KATs ride along
        r[1, 0], a[1, 0] = 25.0, 30.0
        rx = synth.rx_symbols(tx, cfg["R"], r, a, amp, snr_db=20.0, rng=rng)
        est = synth.default_estimator_params(cfg["N"], cfg["T"] * cfg["R"], cfg["IR"], cfg["IA"])
        m, d = ref_chain(rx, tx, cfg, est)
        assert (d["flags"] == 1).all()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), rx=rx, tx=tx, ranges=r, azimuths=a,
                            range_idx=d["range_idx"], angle_idx=d["angle_idx"], peak_power=d["peak_power"],
                            snr_db=d["snr_db"], map0=m[0], map_max=m.reshape(n, -1).max(axis=1),
                            map_sum=m.reshape(n, -1).astype(np.float64).sum(axis=1),
                            cfg=np.array([cfg[k] for k in ("T", "R", "S", "N", "IR", "IA")]))
        print(name, list(zip(d["range_idx"].tolist(), d["angle_idx"].tolist())))
        if name == "c1_shipped":
            # the reference block's own capture_radar_data() line for frame 0 (lib/mimo_ofdm_radar_impl.cc:348-377; Eigen's
            # FullPrecision = 7 significant digits for float, Eigen 3.3), without its time stamp, and the frame as raw
            # complex64 (tx[T][S][N] then rx[R][S][N]) for the C++ block test
            frame = os.path.join(HERE, "c1_frame0.c64")
            t0, r0 = orc.c64(tx), orc.c64(rx[0])
            np.concatenate([t0.ravel(), r0.ravel()]).astype(np.complex64).tofile(frame)
            path = "/tmp/jrc_ref_capture.csv"
            if os.path.exists(path):
                os.remove(path)
            # (in a process of its own, ctypes only: with NumPy's bundled runtime libraries loaded, std::put_time inside the
            #  reference's current_date_time2() crashes)
            import subprocess
            code = (
                "import ctypes as C, sys\n"
                "buf = open(sys.argv[2], 'rb').read()\n"
                "lib = C.CDLL(sys.argv[1]); lib.ref_capture.restype = None\n"
                "lib.ref_capture.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_char_p]\n"
                "T, R, S, N = (int(a) for a in sys.argv[4:8])\n"
                "b = C.create_string_buffer(buf, len(buf)); a = C.addressof(b)\n"
                "lib.ref_capture(a, a + 8 * T * S * N, N, T, R, S, 0, 0, sys.argv[3].encode())\n")
            subprocess.run([sys.executable, "-c", code, os.path.join(ROOT, "oracle", "_ref", "libjrc_ref.so"), frame, path,
                            str(cfg["T"]), str(cfg["R"]), str(cfg["S"]), str(cfg["N"])], check=True)
            line = open(path).read()
            open(os.path.join(HERE, "c1_capture_line.txt"), "w").write(line.split(", ", 1)[1])

if __name__ == "__main__":
    main()
