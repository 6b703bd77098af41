"""Streaming latency mode (BASELINE configs[3]): one CPI per call through the host-buffer C ABI
(jrc_chain_run_host, what radar_chain::general_work does per frame), pinned and pageable host
buffers, with and without reading the |.|^2 map back.  Prints p50/p99 per-CPI latency as JSON."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "python"))
import numpy as np, torch
import mimo_ofdm_jrc as jrc
from mimo_ofdm_jrc import synth

def run(cfg, n=10000, pinned=True, with_map=True):
    T, R, S, N, IR, IA = (cfg[k] for k in ("T", "R", "S", "N", "IR", "IA"))
    est = synth.default_estimator_params(N, T * R, IR, IA)
    ch = jrc.Chain(N, T, R, S, 0, IR, IA)
    ch.set_estimator(**est)
    rng = np.random.default_rng(0)
    tx = synth.tx_symbols(T, S, N)
    r, a, amp = synth.random_scene(rng, 64, 1, N)
    rx = synth.rx_symbols(tx, R, r, a, amp, snr_db=20.0, rng=rng)
    mk = (lambda x: torch.from_numpy(x).pin_memory()) if pinned else torch.from_numpy
    rxs = [mk(np.ascontiguousarray(rx[i])) for i in range(64)]
    txb = mk(tx)
    m = torch.empty((ch.Nr, ch.Na), dtype=torch.float32)
    d = torch.zeros(32, dtype=torch.uint8)
    if pinned: m, d = m.pin_memory(), d.pin_memory()
    lat = np.empty(n)
    for i in range(n + 200):
        t0 = time.perf_counter_ns()
        ch.run_host_ptr(rxs[i & 63].data_ptr(), txb.data_ptr(), True, 1, i, m.data_ptr() if with_map else None, d.data_ptr())
        if i >= 200: lat[i - 200] = (time.perf_counter_ns() - t0) * 1e-3
    det = d.numpy().view(jrc.DET_DTYPE)[0]
    assert det["flags"] & 1 and det["cpi"] == n + 199
    return dict(p50_us=float(np.percentile(lat, 50)), p99_us=float(np.percentile(lat, 99)), mean_us=float(lat.mean()),
                cpis_per_s=float(1e6 / lat.mean()))

if __name__ == "__main__":
    out = {}
    for name, cfg in (("C1 shipped 512x128", dict(T=4, R=2, S=4, N=64, IR=8, IA=16)), ("C2 1024x64", dict(T=4, R=2, S=4, N=64, IR=16, IA=8))):
        for pinned in (True, False):
            for with_map in (True, False):
                out[f"{name} | {'pinned' if pinned else 'pageable'} | {'map+det' if with_map else 'det only'}"] = run(cfg, pinned=pinned, with_map=with_map)
    print(json.dumps(out, indent=1))
