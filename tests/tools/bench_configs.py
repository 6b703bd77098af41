"""Throughput + parity spot-check of every BASELINE config on one GPU (documentation run, the
headline bench is bench.py).  configs[2] (4x8, 256 sc, 4096x256) and configs[4] (8x16, 2048 sc)
have no fused specialisation yet and run on the staged kernels."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "python"))
import numpy as np, torch
import mimo_ofdm_jrc as jrc
from mimo_ofdm_jrc import synth
from oracle import orc

CONFIGS = {
    "configs[0] shipped sim 4x2, 64 sc, 512x128": dict(T=4, R=2, S=4, N=64, IR=8, IA=16, n=4096, targets=1),
    "configs[1] 64 sc, 8 ch, 1024x64": dict(T=4, R=2, S=4, N=64, IR=16, IA=8, n=4096, targets=2),
    "configs[2] 4x8, 256 sc, 4096x256, multi-target": dict(T=4, R=8, S=4, N=256, IR=16, IA=8, n=888, targets=5),
    "configs[4] 8x16, 2048 sc, 2048x128": dict(T=8, R=16, S=8, N=2048, IR=1, IA=1, n=222, targets=3),
}

def b_alg(c):
    return (c["T"] + c["R"]) * c["S"] * c["N"] * 8 + (c["N"] * c["IR"]) * (c["T"] * c["R"] * c["IA"]) * 4 + 32

ONLY = [a for a in sys.argv[1:] if not a.startswith('-')]   # e.g. "configs[2]"
REPS = 2 if '--short' in sys.argv else 20
out = {}
for name, c in CONFIGS.items():
    if ONLY and not any(name.startswith(o) for o in ONLY): continue
    T, R, S, N, IR, IA, n = (c[k] for k in ("T", "R", "S", "N", "IR", "IA", "n"))
    rng = np.random.default_rng(5)
    tx = synth.tx_symbols(T, S, N)
    r, a, amp = synth.random_scene(rng, n, c["targets"], N, amp_db_span=15.0 if c["targets"] > 1 else 0.0)
    rx = synth.rx_symbols(tx, R, r, a, amp, snr_db=20.0, rng=rng, chunk=16)
    est = synth.default_estimator_params(N, T * R, IR, IA)
    rc = jrc.radar_chain(N, T, R, S, IR, IA, estimator=est)
    drx, dtx = torch.from_numpy(rx).cuda(), torch.from_numpy(tx).cuda()
    dmap = torch.empty((n, rc.Nr, rc.Na), dtype=torch.float32, device="cuda")
    ddet = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    ext = torch.cuda.ExternalStream(rc.chain.stream)
    torch.cuda.synchronize()
    def step():
        rc.run(drx, dtx, map_out=dmap, dets_out=ddet, sync_inputs=False)
    with torch.cuda.stream(ext):
        for _ in range(3): step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = REPS
        e0.record(ext)
        for _ in range(reps): step()
        e1.record(ext)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    d = rc.dets_to_numpy(ddet)
    k = min(n, 6)
    mo, _, do = orc.chain_batch(rx[:k], tx, N, T, R, S, IR, IA, est)
    mg = dmap[:k].cpu().numpy()
    err = float((np.abs(mg - mo).reshape(k, -1).max(axis=1) / mo.reshape(k, -1).max(axis=1)).max())
    same = bool(np.array_equal(d["range_idx"][:k], do["range_idx"]) and np.array_equal(d["angle_idx"][:k], do["angle_idx"]))
    rate = n / (ms * 1e-3)
    out[name] = dict(path={jrc.PATH_FUSED: "fused", jrc.PATH_TILED: "tiled", jrc.PATH_STAGED: "staged"}[rc.chain.last_path], cpis=n, ms=ms, cpis_per_s=rate,
                     complex_msps=rate * R * S * N / 1e6, alg_gbs=rate * b_alg(c) / 1e9, map_err_of_peak=err,
                     peaks_equal_oracle=same, gate_pass_fraction=float((d["flags"] & 1).mean()))
    del drx, dtx, dmap, ddet, rc
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
