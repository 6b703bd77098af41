"""Small runs of every fast path for compute-sanitizer (memcheck, racecheck, synccheck):
fused (k_fused64x8 incl. the in-kernel tie resolution and k_est_exact via a threshold on a CPI's own SNR), configs[2]
(k_slice256), configs[4] (k_wide_*), a generic tiled shape, the submit/wait slots.
    compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "python"))
import numpy as np
import mimo_ofdm_jrc as jrc
from mimo_ofdm_jrc import synth

def scene(T, R, S, N, n, targets=2, equal=False):
    rng = np.random.default_rng(0)
    tx = synth.tx_symbols(T, S, N)
    r, a, amp = synth.random_scene(rng, n, targets, N, amp_db_span=0.0 if equal else 10)
    return synth.rx_symbols(tx, R, r, a, amp, snr_db=None if equal else 20.0, rng=rng, chunk=8), tx

for (T, R, S, N, IR, IA, n) in ((4, 2, 4, 64, 8, 16, 300), (4, 2, 4, 64, 16, 8, 300), (4, 8, 4, 256, 16, 8, 3), (8, 16, 8, 2048, 1, 1, 3),
                                (4, 2, 4, 128, 4, 16, 8)):
    est = synth.default_estimator_params(N, T * R, IR, IA)
    rx, tx = scene(T, R, S, N, n)
    ch = jrc.Chain(N, T, R, S, 0, IR, IA)
    ch.set_estimator(**est)
    m, d = ch.run_host(rx, tx)
    # a threshold on CPI 0's own SNR sends it through k_est_exact; equal-amplitude targets exercise the tie logic
    ch.set_thresholds(np.float32(d["snr_db"][0]), est["power_threshold"])
    m, d2 = ch.run_host(rx, tx)
    rx2, _ = scene(T, R, S, N, min(n, 16), equal=True)
    ch.set_thresholds(est["snr_threshold"], est["power_threshold"])
    m, d3 = ch.run_host(rx2, tx)
    print((N, T * R, IR, IA), "path", ch.last_path, "peaks", d["range_idx"][:3], "exact", int((d2["flags"] & 2 != 0).sum()), ch.exact_stats())
# streaming slots
T, R, S, N, IR, IA = 4, 2, 4, 64, 16, 8
rx, tx = scene(T, R, S, N, 8)
ch = jrc.Chain(N, T, R, S, 0, IR, IA)
ch.set_estimator(**synth.default_estimator_params(N, T * R, IR, IA))
maps = [np.empty((1, N * IR, T * R * IA), np.float32) for _ in range(4)]
dets = [np.zeros(1, jrc.DET_DTYPE) for _ in range(4)]
tk = [ch.submit_ptr(jrc.cabi.np_ptr(np.ascontiguousarray(rx[i:i + 1])), jrc.cabi.np_ptr(tx), True, 1, i, jrc.cabi.np_ptr(maps[i]), jrc.cabi.np_ptr(dets[i])) for i in range(4)]
for t in tk:
    ch.wait(t)
print("submit/wait", [int(x["range_idx"][0]) for x in dets])
# raw time samples in front of the chain (k_ofdm_demod64 / k_ofdm_demod_batch)
import torch
for (T, R, S, N, IR, IA, n) in ((4, 2, 4, 64, 16, 8, 64), (4, 8, 4, 256, 4, 2, 4)):
    rx, tx = scene(T, R, S, N, n)
    cp = N // 4
    td = np.fft.ifft(np.fft.ifftshift(rx.astype(np.complex128), axes=-1), axis=-1)
    td = np.concatenate([td[..., N - cp:], td], axis=-1).astype(np.complex64)
    ch = jrc.Chain(N, T, R, S, 0, IR, IA)
    ch.set_estimator(**synth.default_estimator_params(N, T * R, IR, IA))
    dtd, dtx = torch.from_numpy(td).cuda(), torch.from_numpy(tx).cuda()
    m = torch.empty((n, ch.Nr, ch.Na), dtype=torch.float32, device="cuda")
    d = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ch.run_batch_time_ptr(dtd.data_ptr(), R * S * (N + cp), S * (N + cp), cp, dtx.data_ptr(), 0, S * N, n, 0, m.data_ptr(), None, d.data_ptr())
    ch.sync()
    print("time samples", (N, T * R), "path", ch.last_path, float(m.max()))
