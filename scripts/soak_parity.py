"""Soak run of the detection-list claim: random scenes (1-4 targets, amplitude spreads up to 20 dB, SNR 0..30 dB so that
many records sit near the 15 dB gate, plus tie scenes whose map repeats after Nr/2 rows) through the fast paths and the
one-kernel-per-block path on the GPU; range_idx / angle_idx / n_noise / gate flag must be equal on EVERY CPI and every
record redone in the reference's order equal in every bit (tests/test_gpu_parity.py::check_detections).
    python scripts/soak_parity.py [seconds] [gate] -> one JSON line"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import mimo_ofdm_jrc as jrc
from mimo_ofdm_jrc import synth
from test_gpu_parity import check_detections, CFGS

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
gate_mode = len(sys.argv) > 2 and sys.argv[2] == "gate"       # every scene gets a threshold on one of its own SNRs
plans = [("C2", 2048), ("C1", 2048), ("C3s", 256), ("C3", 24), ("C5", 12), ("sq8", 1024)]
chains = {}
tot = {k: dict(cpis=0, exact=0, marked=0, ties_in_kernel=0, near_gate=0) for k, _ in plans}
t_end, seed = time.time() + budget, 0
while time.time() < t_end:
    for name, n in plans:
        cfg = CFGS[name]
        seed += 1
        rng = np.random.default_rng(seed)
        est = synth.default_estimator_params(cfg["N"], cfg["T"] * cfg["R"], cfg["IR"], cfg["IA"])
        tx = synth.tx_symbols(cfg["T"], cfg["S"], cfg["N"])
        kind = seed % 4
        if kind == 3:      # ties: two echoes half the unambiguous range apart
            half = synth.C_LIGHT / (2 * 125e6) * (cfg["N"] / 2)
            r1 = rng.uniform(0.05, 0.4, n) * 2 * half
            r = np.stack([r1, r1 + half], axis=1)
            az = rng.uniform(-50, 50, n)
            a = np.stack([az, az], axis=1)
            amp = np.ones((n, 2)); amp[n // 2:, 1] += 1e-6 * rng.standard_normal(n - n // 2)
            rx = synth.rx_symbols(tx, cfg["R"], r, a, amp, chunk=64)
        else:
            nt = int(rng.integers(1, 5))
            r, a, amp = synth.random_scene(rng, n, nt, cfg["N"], amp_db_span=float(rng.uniform(0, 20)))
            rx = synth.rx_symbols(tx, cfg["R"], r, a, amp, snr_db=float(rng.uniform(0, 30)), rng=rng, chunk=64)
        if name not in chains:
            chains[name] = jrc.radar_chain(cfg["N"], cfg["T"], cfg["R"], cfg["S"], cfg["IR"], cfg["IA"], estimator=est)
        rc = chains[name]
        if gate_mode:
            # gate threshold on a record's own SNR (first pass), so that records land inside the gate margin
            _, dq = rc.run(torch.from_numpy(rx).cuda(), torch.from_numpy(tx).cuda(), want_map=True)
            rc.sync()
            snr = rc.dets_to_numpy(dq)["snr_db"]
            est = dict(est, snr_threshold=float(np.float32(snr[int(rng.integers(0, n))]) + np.float32(rng.choice([0.0, 1e-5, -1e-5, 3e-4]))))
            rc.chain.set_thresholds(est["snr_threshold"], est["power_threshold"])
        drx, dtx = torch.from_numpy(rx).cuda(), torch.from_numpy(tx).cuda()
        _, d1 = rc.run(drx, dtx, want_map=True)
        rc.sync()
        _, d2 = rc.run(drx, dtx, want_map=False, path=jrc.PATH_STAGED)
        rc.sync()
        d1, d2 = rc.dets_to_numpy(d1), rc.dets_to_numpy(d2)
        tot[name]["exact"] += check_detections(jrc, d1, d2, f"soak {name} seed {seed}")
        tot[name]["cpis"] += n
        tot[name]["near_gate"] += int((np.abs(d2["snr_db"] - est["snr_threshold"]) < 0.5).sum())
for name, rc in chains.items():
    st = rc.chain.exact_stats()
    tot[name]["marked"], tot[name]["ties_in_kernel"] = st["marked"], st["ties_in_kernel"]
print(json.dumps({"seconds": budget, "gate_mode": gate_mode, "scenes": seed, "detection_lists_identical": True, "per_config": tot}))
