"""Small fused / staged / stream / tc runs for compute-sanitizer (memcheck, racecheck, synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "python"))
import numpy as np
import mimo_ofdm_jrc as jrc
from mimo_ofdm_jrc import synth
for (T, R, S, N, IR, IA) in ((4, 2, 4, 64, 8, 16), (4, 2, 4, 64, 16, 8)):
    est = synth.default_estimator_params(N, T * R, IR, IA)
    rng = np.random.default_rng(0)
    tx = synth.tx_symbols(T, S, N)
    r, a, amp = synth.random_scene(rng, 700, 2, N, amp_db_span=10)
    rx = synth.rx_symbols(tx, R, r, a, amp, snr_db=20.0, rng=rng)
    ch = jrc.Chain(N, T, R, S, 0, IR, IA)
    ch.set_estimator(**est)
    m, d = ch.run_host(rx, tx)
    print(os.environ.get("JRC_FUSED_KERNEL", "cta"), (IR, IA), "path", ch.last_path, "peaks", d["range_idx"][:3], d["flags"].mean())
