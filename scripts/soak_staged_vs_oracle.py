"""Soak run of the one-kernel-per-block GPU path against the CPU oracle: random scenes, every record field that the
path promises bit for bit (range_idx, angle_idx, peak_power, noise_power, n_noise, gate flag) -- in particular the
reference-order noise sum, which the device runs speculatively (jrc_common.cuh seq_sum_sq_warp).
    python scripts/soak_staged_vs_oracle.py [seconds] -> one JSON line"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import mimo_ofdm_jrc as jrc
from mimo_ofdm_jrc import synth
from oracle import orc
from test_gpu_parity import CFGS

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
plans = [("C2", 96), ("C1", 96), ("sq8", 128), ("odd", 256), ("C3s", 24)]
t_end, seed, tot = time.time() + budget, 0, {k: 0 for k, _ in plans}
while time.time() < t_end:
    for name, n in plans:
        cfg = CFGS[name]
        seed += 1
        rng = np.random.default_rng(10_000 + seed)
        est = synth.default_estimator_params(cfg["N"], cfg["T"] * cfg["R"], cfg["IR"], cfg["IA"])
        est["snr_threshold"] = float(rng.uniform(5, 40))
        tx = synth.tx_symbols(cfg["T"], cfg["S"], cfg["N"])
        nt = int(rng.integers(1, 5))
        r, a, amp = synth.random_scene(rng, n, nt, cfg["N"], amp_db_span=float(rng.uniform(0, 25)))
        amp = amp * float(10 ** rng.uniform(-3, 3))                 # window sums over twelve decades of power
        rx = synth.rx_symbols(tx, cfg["R"], r, a, amp, snr_db=float(rng.uniform(-5, 35)), rng=rng, chunk=64)
        rc = jrc.radar_chain(cfg["N"], cfg["T"], cfg["R"], cfg["S"], cfg["IR"], cfg["IA"], estimator=est)
        _, d = rc.run(torch.from_numpy(rx).cuda(), torch.from_numpy(tx).cuda(), want_map=False, path=jrc.PATH_STAGED)
        rc.sync()
        d = rc.dets_to_numpy(d)
        _, _, do = orc.chain_batch(rx, tx, cfg["N"], cfg["T"], cfg["R"], cfg["S"], cfg["IR"], cfg["IA"], est)
        for f in ("range_idx", "angle_idx", "peak_power", "noise_power", "n_noise"):
            assert np.array_equal(d[f], do[f]), (name, seed, f, np.flatnonzero(d[f] != do[f])[:5])
        ok = np.abs(d["snr_db"] - est["snr_threshold"]) > 1e-4       # (device log10f against libm on the provisional SNR)
        assert np.array_equal((d["flags"] & 1)[ok], (do["flags"] & 1)[ok]), (name, seed, "gate")
        tot[name] += n
        del rc
print(json.dumps({"seconds": budget, "scenes": seed, "records_bit_identical": True, "cpis": tot}))
