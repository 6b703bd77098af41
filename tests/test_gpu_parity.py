"""GPU parity tests (run on the B200 box): the CUDA path through the C ABI against the CPU
oracle on identical seeded inputs.

Bars: byte/index work and every per-block stage call bit-exact; the fused and tiled chains' maps within
1e-4 of the map peak (BASELINE.json north_star); their detection lists -- peak range/angle indices, the size of
the noise window and the gate flag -- IDENTICAL to the oracle's on every CPI, no exclusions: a decision that FFT
rounding could turn is redone in the reference's order (csrc/jrc_exact.cuh), and such a record (DET_EXACT) is
bit-identical to the oracle's in every field."""
import numpy as np
import pytest

from mimo_ofdm_jrc import synth

pytestmark = pytest.mark.gpu

CFGS = {
    "C1": dict(T=4, R=2, S=4, N=64, IR=8, IA=16),      # shipped flowgraph, 512 x 128
    "C2": dict(T=4, R=2, S=4, N=64, IR=16, IA=8),      # BASELINE configs[1], 1024 x 64
    "C2b": dict(T=2, R=4, S=2, N=64, IR=16, IA=8),     # "2 TX x 4 RX" wording of configs[0]
    "sq8": dict(T=4, R=2, S=4, N=64, IR=8, IA=8),
    "sq16": dict(T=8, R=1, S=8, N=64, IR=16, IA=16),
    "C3s": dict(T=4, R=8, S=4, N=256, IR=4, IA=2),     # 32 virtual channels (staged path)
    "odd": dict(T=2, R=2, S=3, N=32, IR=2, IA=4),
    "C3": dict(T=4, R=8, S=4, N=256, IR=16, IA=8),     # BASELINE configs[2]: 4096 x 256 map
    "C5": dict(T=8, R=16, S=8, N=2048, IR=1, IA=1),    # BASELINE configs[4]: 8 x 16 array, 2048 sc
}


def scene(cfg, n_cpi, seed, n_targets=1, snr_db=20.0, amp_db_span=0.0, tx_per_cpi=False):
    rng = np.random.default_rng(seed)
    tx = synth.tx_symbols(cfg["T"], cfg["S"], cfg["N"])
    r, a, amp = synth.random_scene(rng, n_cpi, n_targets, cfg["N"], amp_db_span=amp_db_span)
    rx = synth.rx_symbols(tx, cfg["R"], r, a, amp, snr_db=snr_db, rng=rng)
    if tx_per_cpi:
        tx = np.broadcast_to(tx, (n_cpi,) + tx.shape).copy()
    return rx, tx, (r, a)


def est_for(cfg):
    return synth.default_estimator_params(cfg["N"], cfg["T"] * cfg["R"], cfg["IR"], cfg["IA"])


def gpu_chain(jrc, cfg, est, **kw):
    ch = jrc.Chain(cfg["N"], cfg["T"], cfg["R"], cfg["S"], 0, cfg["IR"], cfg["IA"], **kw)
    ch.set_estimator(**est)
    return ch


def oracle(orc, rx, tx, cfg, est, **kw):
    return orc.chain_batch(rx, tx, cfg["N"], cfg["T"], cfg["R"], cfg["S"], cfg["IR"], cfg["IA"], est, **kw)


def check_detections(jrc, d, do, what, rtol_peak=5e-6):
    """Detection lists equal on EVERY CPI; records redone in the reference's order equal in every bit."""
    n = len(do)
    for f in ("range_idx", "angle_idx", "n_noise"):
        assert np.array_equal(d[f], do[f]), (what, f, np.flatnonzero(d[f] != do[f])[:8])
    assert np.array_equal(d["flags"] & jrc.DET_PASSED, do["flags"] & 1), (what, "gate")
    ex = (d["flags"] & jrc.DET_EXACT) != 0
    for f in ("peak_power", "noise_power", "snr_db"):
        assert np.array_equal(d[f][ex], do[f][ex], equal_nan=True), (what, f, "DET_EXACT records")
    ok = ~ex & (do["range_idx"] >= 0)
    np.testing.assert_allclose(d["peak_power"][ok], do["peak_power"][ok], rtol=rtol_peak)
    # the window samples carry the FFT rounding of the whole map (~3e-7 of the PEAK amplitude each): relative to a noise
    # floor far below the peak that is 3e-7 * sqrt(peak / noise) per sample, a fraction of it after averaging
    rel = 1e-5 + 3e-8 * np.sqrt(do["peak_power"][ok] / do["noise_power"][ok])
    assert (np.abs(d["noise_power"][ok] - do["noise_power"][ok]) <= rel * do["noise_power"][ok]).all()
    assert (np.abs(d["snr_db"][ok] - do["snr_db"][ok]) <= 4.35 * rel + 2e-5).all()
    print(f"[{what}] {n} CPIs: detection lists identical; {int(ex.sum())} redone in the reference's order")
    return int(ex.sum())


def top2_margin(m):
    flat = np.partition(m.reshape(m.shape[0], -1), -2, axis=1)
    return (flat[:, -1] - flat[:, -2]) / flat[:, -1]


@pytest.mark.parametrize("name", ["C1", "C2", "C2b", "sq8", "sq16"])
def test_fused_chain_vs_oracle(jrc, orc, name):
    cfg = CFGS[name]
    est = est_for(cfg)
    rx, tx, _ = scene(cfg, 96, seed=7, n_targets=2, amp_db_span=12.0, tx_per_cpi=True)
    ch = gpu_chain(jrc, cfg, est)
    m, d = ch.run_host(rx, tx)
    assert ch.last_path == jrc.PATH_FUSED
    mo, _, do = oracle(orc, rx, tx, cfg, est)
    peak = mo.reshape(96, -1).max(axis=1)
    err = np.abs(m - mo).reshape(96, -1).max(axis=1) / peak
    assert err.max() <= 1e-4, err.max()               # north_star tolerance
    assert err.max() <= 5e-6, err.max()               # what float32 should actually deliver
    assert np.array_equal(d["cpi"], np.arange(96))
    check_detections(jrc, d, do, f"fused {name}")


@pytest.mark.parametrize("name", list(CFGS))
def test_staged_chain_is_bit_exact(jrc, orc, name):
    """One kernel per reference block, oracle float order -> identical bits (full-size BASELINE
    configs[2] and configs[4] included, multi-target scenes)."""
    import torch
    cfg = CFGS[name]
    est = est_for(cfg)
    n = 3 if name in ("C3", "C5") else 12
    rx, tx, _ = scene(cfg, n, seed=3, n_targets=5 if name == "C3" else 3, amp_db_span=20.0)
    mo, cmo, do = oracle(orc, rx, tx, cfg, est, want_cmap=True)
    rc = jrc.radar_chain(cfg["N"], cfg["T"], cfg["R"], cfg["S"], cfg["IR"], cfg["IA"], estimator=est)
    drx, dtx = torch.from_numpy(rx).cuda(), torch.from_numpy(tx).cuda()
    Nr, Na = rc.Nr, rc.Na
    m = torch.empty((n, Nr, Na), dtype=torch.float32, device="cuda")
    cm = torch.empty((n, Nr, Na), dtype=torch.complex64, device="cuda")
    dets = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    per_ant = cfg["S"] * cfg["N"]
    torch.cuda.synchronize()      # the handle runs on its own stream: inputs must have landed
    rc.chain.run_batch_ptr(drx.data_ptr(), cfg["R"] * per_ant, per_ant, dtx.data_ptr(), 0, per_ant, n, 0,
                           m.data_ptr(), cm.data_ptr(), dets.data_ptr(), jrc.PATH_STAGED)
    rc.sync()
    assert rc.chain.last_path == jrc.PATH_STAGED
    assert np.array_equal(cm.cpu().numpy(), cmo)
    assert np.array_equal(m.cpu().numpy(), mo)
    d = rc.dets_to_numpy(dets)
    for f in ("range_idx", "angle_idx", "peak_power", "noise_power", "n_noise", "cpi"):
        assert np.array_equal(d[f], do[f]), f
    np.testing.assert_allclose(d["snr_db"], do["snr_db"], rtol=1e-6)   # device log10f vs libm
    assert np.array_equal(d["flags"], do["flags"])


def test_fused_equals_staged_detections_large_batch(jrc):
    """Full BASELINE configs[1] batch (4096 CPIs): the two independent GPU paths agree, and the
    single-target peaks sit on the analytic bins (size-independent property)."""
    import torch
    cfg = CFGS["C2"]
    est = est_for(cfg)
    n = 4096
    rx, tx, (r, a) = scene(cfg, n, seed=11, n_targets=1, snr_db=25.0)
    rc = jrc.radar_chain(cfg["N"], cfg["T"], cfg["R"], cfg["S"], cfg["IR"], cfg["IA"], estimator=est)
    drx, dtx = torch.from_numpy(rx).cuda(), torch.from_numpy(tx).cuda()
    m1, d1 = rc.run(drx, dtx, path=jrc.PATH_FUSED)
    rc.sync()
    m2, d2 = rc.run(drx, dtx, path=jrc.PATH_STAGED)
    rc.sync()
    d1, d2 = rc.dets_to_numpy(d1), rc.dets_to_numpy(d2)
    pk = m2.reshape(n, -1).max(dim=1).values
    err = ((m1 - m2).abs().reshape(n, -1).max(dim=1).values / pk).max().item()
    assert err <= 5e-6, err
    n_exact = check_detections(jrc, d1, d2, "fused vs staged, 4096 CPIs of configs[1]")
    assert n_exact <= 16, n_exact                     # the reference-order pass is the exception (expected ~1e-4 .. 1e-3 of the CPIs)
    exp = np.array([synth.expected_peak(r[i, 0], a[i, 0], 64, 16, 8, 8) for i in range(n)])
    close = (np.abs(d1["range_idx"] - exp[:, 0]) <= 1) & (np.abs(d1["angle_idx"] - exp[:, 1]) <= 1)
    assert close.mean() > 0.99
    assert (d1["flags"] & 1).mean() > 0.99


def test_fused_linearity_and_tx_sharing(jrc):
    """|chain(2x)|^2 = 4 |chain(x)|^2 exactly (power-of-two scaling commutes with rounding), and a
    shared TX frame gives the same bits as per-CPI copies of it."""
    cfg = CFGS["C2"]
    est = est_for(cfg)
    rx, tx, _ = scene(cfg, 16, seed=5)
    ch = gpu_chain(jrc, cfg, est)
    m1, d1 = ch.run_host(rx, tx)
    m2, d2 = ch.run_host(rx * np.float32(2), tx)
    assert np.array_equal(m2, m1 * np.float32(4))
    assert np.array_equal(d1["range_idx"], d2["range_idx"]) and np.array_equal(d1["angle_idx"], d2["angle_idx"])
    m3, d3 = ch.run_host(rx, np.broadcast_to(tx, (16,) + tx.shape).copy())
    assert np.array_equal(m3, m1) and np.array_equal(d3, d1)


def test_fused_background_removal_matches_block_sequence(jrc, orc):
    cfg = CFGS["C1"]
    est = est_for(cfg)
    n = 20
    rx, tx, _ = scene(cfg, n, seed=9, n_targets=2, amp_db_span=6.0, tx_per_cpi=True)
    rng = np.random.default_rng(1)
    clutter = synth.rx_symbols(tx[0], cfg["R"], [[7.0, 31.0]], [[-20.0, 40.0]], [[3.0, 2.0]])[0]
    rx = (rx + clutter[None]).astype(np.complex64)
    ch = gpu_chain(jrc, cfg, est, background_removal=True, background_recording=True, record_len=4)
    m, d = ch.run_host(rx, tx)
    # oracle: block-by-block with the ring-buffer state machine
    rad = orc.Radar(cfg["N"], cfg["T"], cfg["R"], cfg["S"], 0, True, True, 4, cfg["IR"], False)
    for c in range(n):
        pad = rad.work(list(tx[c].reshape(cfg["T"], -1)), list(rx[c].reshape(cfg["R"], -1)))
        y = orc.fft_vcc(pad, False, False)
        cm = orc.fft_vcc(orc.matrix_transpose(y, 8, cfg["IA"]), True, True)
        mo = orc.mag_squared(cm)
        assert np.abs(m[c] - mo).max() <= 1e-5 * mo.max() + 1e-3, c
        do = orc.range_angle_estimate(cm, **est)
        assert (d[c]["range_idx"], d[c]["angle_idx"]) == (do["range_idx"], do["angle_idx"]), c
        assert (d[c]["flags"] & jrc.DET_PASSED) == (do["flags"] & 1), c


@pytest.mark.parametrize("name", ["C1", "C3s", "odd"])
def test_stage_calls_bit_exact(jrc, orc, name):
    cfg = CFGS[name]
    T, R, S, N, IR, IA = (cfg[k] for k in ("T", "R", "S", "N", "IR", "IA"))
    V, pre = T * R, 5
    rng = np.random.default_rng(21)
    def frame(extra=0):
        return [(rng.standard_normal((pre + S + extra) * N) + 1j * rng.standard_normal((pre + S + extra) * N)).astype(np.complex64)
                for _ in range(T + R)]
    for interleave in (False, True):
        blk = jrc.mimo_ofdm_radar(N, T, R, S, pre, True, True, 3, IR, interleave, "/tmp/jrc_chan.csv")
        ref = orc.Radar(N, T, R, S, pre, True, True, 3, IR, interleave)
        for i in range(5):
            f = frame(extra=2)
            out, tags, consumed = blk.work(f[:T], f[T:])
            assert np.array_equal(out, ref.work(f[:T], f[T:])), (interleave, i)
            assert tags[0]["value"] == V and tags[0]["offset"] == i * V
            if i == 2:
                blk.set_background_record(False); ref.set_background_record(False)
    # stale TX frame skipping
    blk = jrc.mimo_ofdm_radar(N, T, R, S, pre, False, False, 1, IR, False, "/tmp/jrc_chan.csv")
    ref = orc.Radar(N, T, R, S, pre, False, False, 1, IR, False)
    f1, f2 = frame(), frame()
    txcat = [np.concatenate([f1[i], f2[i]]) for i in range(T)]
    out, tags, consumed = blk.work(txcat, f2[T:], tx_tag_lens=[pre + S, pre + S], rx_tag_lens=[pre + S])
    assert np.array_equal(out, ref.work(f2[:T], f2[T:])) and consumed["tx"][0] == 2 * (pre + S)
    out, tags, consumed = blk.work(f1[:T], f1[T:], rx_tag_lens=[])
    assert out is None and tags == []
    # fft_vcc both directions, transpose, mag^2
    Nr, Na = N * IR, V * IA
    x = (rng.standard_normal((V, Nr)) + 1j * rng.standard_normal((V, Nr))).astype(np.complex64)
    y = jrc.fft_vcc(Nr, False, shift=False).work(x)
    assert np.array_equal(y, orc.fft_vcc(x, False, False))
    tr = jrc.matrix_transpose(Nr, V, IA).work(y)
    assert np.array_equal(tr, orc.matrix_transpose(y, V, IA))
    cm = jrc.fft_vcc(Na, True, shift=True).work(tr)
    assert np.array_equal(cm, orc.fft_vcc(tr, True, True))
    assert np.array_equal(jrc.complex_to_mag_squared(Na).work(cm), orc.mag_squared(cm))
    assert np.array_equal(jrc.fft_vcc(Na, False, shift=True).work(tr), orc.fft_vcc(tr, False, True))
    # estimator on the same complex map: every field identical, snr included (host libm)
    est = est_for(cfg)
    blk = jrc.range_angle_estimator(Na, est["range_bins"], est["angle_bins"], est["noise_discard_range_m"],
                                    est["noise_discard_angle_deg"], -1e9, 0.0, "/tmp/jrc_radar_log.csv", False)
    det = blk.work(cm)
    do = orc.range_angle_estimate(cm, **dict(est, snr_threshold=np.float32(-1e9)))
    for fld in ("range_idx", "angle_idx", "peak_power", "noise_power", "snr_db", "n_noise", "flags"):
        assert det[fld] == do[fld], fld
    assert len(blk.messages) == 1 and blk.messages[0][0][0] == "range"
    with pytest.raises(RuntimeError):
        jrc.matrix_transpose(Nr, V + 1, 1).work(y[:V])


def test_estimator_edge_maps(jrc, orc):
    est = synth.default_estimator_params(64, 8, 8, 16)
    blk = jrc.range_angle_estimator(128, est["range_bins"], est["angle_bins"], est["noise_discard_range_m"],
                                    est["noise_discard_angle_deg"], 15.0, 0.0, "/tmp/jrc_radar_log.csv", True)
    rng = np.random.default_rng(8)
    base = (0.01 * (rng.standard_normal((512, 128)) + 1j * rng.standard_normal((512, 128)))).astype(np.complex64)
    for (pr, pa) in [(100, 64), (100, 61), (500, 30), (0, 0), (511, 127), (256, 8)]:
        m = base.copy()
        m[pr, pa] = 5.0
        m[(pr + 77) % 512, (pa + 5) % 128] = 5.0      # exact tie: first in row-major order wins
        det = blk.work(m)
        do = orc.range_angle_estimate(m, **est)
        for fld in ("range_idx", "angle_idx", "peak_power", "noise_power", "snr_db", "n_noise", "flags"):
            assert det[fld] == do[fld], (pr, pa, fld)
    blk.set_snr_threshold(1e9)
    assert blk.work(m)["flags"] == 0
    lines = [l for l in open("/tmp/jrc_radar_log.csv").read().splitlines() if l.strip()]
    assert any("NEW RECORD" in l for l in lines) and len(lines[-1].split(",")) == 5
    # the consumer's reader (mimo_precoder's radar-aided steering) sees the last gated detection
    rec = jrc.radar_log_read_last("/tmp/jrc_radar_log.csv")
    assert rec is not None and abs(rec[4] - est["angle_bins"][8]) < 1e-3 and abs(rec[3] - est["range_bins"][256]) < 1e-3


def test_nlog10_heatmap_feed(jrc):
    """blocks_nlog10_ff between complex_to_mag_squared and gui_heatmap_plot."""
    rng = np.random.default_rng(4)
    x = (rng.random((512, 128)) ** 8).astype(np.float32)
    x[0, :4] = [0.0, 1e-30, 1.0, 1e-18]
    y = jrc.nlog10_ff(10.0, 128, 0.0).work(x)
    ref = (10.0 * np.log10(np.maximum(x, np.float32(1e-18)).astype(np.float64))).astype(np.float32)
    np.testing.assert_allclose(y, ref, rtol=2e-6, atol=2e-5)
    assert y[0, 0] == y[0, 1] and abs(y[0, 0] + 180.0) < 1e-4 and y[0, 2] == 0.0     # clamp at 1e-18
    y2 = jrc.nlog10_ff(20.0, 128, 3.0).work(x)
    np.testing.assert_allclose(y2, 2.0 * ref + 3.0, rtol=2e-6, atol=5e-5)


def test_peak1d_and_zero_pad(jrc, orc):
    rng = np.random.default_rng(13)
    n = 40000                      # alignment flowgraph packet: 5000 * 8
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    x[31000] = 40 + 9j
    blk = jrc.fft_peak_detect(1000000, 8.0, 10.0, 25, [0.0], False)
    assert blk.work(x) == orc.fft_peak_detect(x, 1000000, 8.0, 10.0, 25)
    x[17] = 100.0                  # protected sample must be ignored
    x[1234] = 40 + 9j              # equal magnitude earlier in the packet wins
    assert blk.work(x) == orc.fft_peak_detect(x, 1000000, 8.0, 10.0, 25)
    assert blk.work(x)[0] == 1234
    blk.set_threshold(90.0)
    assert blk.work(x)[0] == -1 and orc.fft_peak_detect(x, 1000000, 8.0, 90.0, 25)[0] == -1
    assert blk.work(x[:0])[0] == -1
    zp = jrc.zero_pad(False, 7, 240)
    y = zp.work(x[:720], seed=3)
    assert y.size == 720 + 247 and np.array_equal(y[7:727], x[:720])
    big = jrc.zero_pad(False, 50000, 50000).work(x[:16], seed=4)
    p = np.concatenate([big[:50000], big[-50000:]])
    assert abs(p.real.std() - 1e-2) < 2e-4 and abs(p.imag.std() - 1e-2) < 2e-4 and abs(p.mean()) < 2e-4
    assert not np.array_equal(zp.work(x[:720]), zp.work(x[:720]))     # fresh seed per call
    assert zp.work(x[:0], seed=1).size == 247


import glob as _glob
import os as _os
_GOLDEN = sorted(_glob.glob(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "*.npz")))


@pytest.mark.parametrize("path", _GOLDEN, ids=[_os.path.basename(p) for p in _GOLDEN])
def test_fused_chain_vs_reference_golden_vectors(jrc, path):
    """Fixtures produced by the reference's own block code (tests/golden/make_golden.py)."""
    g = np.load(path)
    T, R, S, N, IR, IA = (int(v) for v in g["cfg"])
    est = synth.default_estimator_params(N, T * R, IR, IA)
    ch = jrc.Chain(N, T, R, S, 0, IR, IA)
    ch.set_estimator(**est)
    m, d = ch.run_host(g["rx"], g["tx"])
    assert ch.last_path == jrc.PATH_FUSED
    assert np.array_equal(d["range_idx"], g["range_idx"]) and np.array_equal(d["angle_idx"], g["angle_idx"])
    np.testing.assert_allclose(d["peak_power"], g["peak_power"], rtol=5e-6)
    np.testing.assert_allclose(d["snr_db"], g["snr_db"], atol=2e-3)
    assert np.abs(m[0] - g["map0"]).max() <= 1e-4 * g["map0"].max()
    assert np.abs(m[0] - g["map0"]).max() <= 5e-6 * g["map0"].max()
    np.testing.assert_allclose(m.reshape(len(d), -1).max(axis=1), g["map_max"], rtol=5e-6)
    # staged path: the same bits as the reference build
    import torch
    rc = jrc.radar_chain(N, T, R, S, IR, IA, estimator=est)
    drx, dtx = torch.from_numpy(g["rx"]).cuda(), torch.from_numpy(g["tx"]).cuda()
    m2, d2 = rc.run(drx, dtx, path=jrc.PATH_STAGED)
    rc.sync()
    d2 = rc.dets_to_numpy(d2)
    assert np.array_equal(m2[0].cpu().numpy(), g["map0"])
    assert np.array_equal(d2["range_idx"], g["range_idx"]) and np.array_equal(d2["peak_power"], g["peak_power"])


def test_ofdm_demod_front_end(jrc, orc):
    """cyclic-prefix removal (+ the RX OFDM FFT) in front of the radar path."""
    rng = np.random.default_rng(31)
    for (N, cp, nsym) in ((64, 16, 12), (256, 64, 9), (2048, 512, 3), (64, 0, 5)):
        x = (rng.standard_normal(nsym * (N + cp) + 7) + 1j * rng.standard_normal(nsym * (N + cp) + 7)).astype(np.complex64)
        blk = jrc.ofdm_cyclic_prefix_remover(N, cp)
        assert blk.calculate_output_stream_length(x.size) == nsym
        td = blk.work(x)
        assert np.array_equal(td, orc.cp_remove(x, nsym, N, cp))
        fd = blk.work(x, demod=True)
        assert np.array_equal(fd, orc.fft_vcc(orc.cp_remove(x, nsym, N, cp), True, True))


TILED_CFGS = {
    "C3": (CFGS["C3"], 3, 5),                                         # BASELINE configs[2], full size
    "C5": (CFGS["C5"], 2, 3),                                         # BASELINE configs[4], full size
    "C3s": (CFGS["C3s"], 24, 3),                                      # 1024 x 64 map, 32 channels
    "n128": (dict(T=4, R=2, S=4, N=128, IR=4, IA=16), 24, 2),         # 512 x 128: radix 8.8.8 / 8.8.2
    "n64r4": (dict(T=2, R=4, S=2, N=64, IR=4, IA=8), 40, 2),          # 256 x 64, not a k_fused64x8 shape
    "wide": (dict(T=8, R=8, S=2, N=64, IR=2, IA=8), 16, 2),           # 128 x 512: radix 8.8.2 / 8.8.8
    "n2048a": (dict(T=4, R=4, S=2, N=512, IR=16, IA=128), 2, 2),      # 8192 x 2048: the largest supported FFTs
    "r64": (dict(T=2, R=2, S=2, N=64, IR=1, IA=16), 32, 1),           # 64 x 64: shortest range FFT, no zero-pad
    "r1k": (dict(T=4, R=4, S=2, N=128, IR=8, IA=32), 8, 2),           # 1024 x 512: radix 8.8.8.2 pruned / 8.8.8 pruned
    "r2k": (dict(T=8, R=8, S=2, N=256, IR=8, IA=8), 4, 3),            # 2048 x 512: 8.8.8.4 pruned, 64 channels
    "a1k": (dict(T=16, R=16, S=1, N=64, IR=2, IA=4), 4, 2),           # 128 x 1024 from 256 channels: angle 8.8.8.2 unpruned
    "a512": (dict(T=8, R=16, S=1, N=64, IR=4, IA=4), 4, 2),           # 256 x 512 from 128 channels: angle 8.8.8 unpruned
}


@pytest.mark.parametrize("name", list(TILED_CFGS))
def test_tiled_chain_vs_oracle(jrc, orc, name):
    """Configurations without a k_fused64x8 specialisation run the tiled kernels (radix-8 FFTs, the
    transpose / |.|^2 / arg-max fused into the angle FFT): maps within 1e-4 of the map peak, detections
    equal wherever the oracle's top-1/top-2 margin exceeds the float32 FFT error."""
    cfg, n, n_targets = TILED_CFGS[name]
    est = est_for(cfg)
    rx, tx, _ = scene(cfg, n, seed=13, n_targets=n_targets, amp_db_span=12.0)
    ch = gpu_chain(jrc, cfg, est)
    m, d = ch.run_host(rx, tx)
    assert ch.last_path == jrc.PATH_TILED
    mo, _, do = oracle(orc, rx, tx, cfg, est)
    peak = mo.reshape(n, -1).max(axis=1)
    err = np.abs(m - mo).reshape(n, -1).max(axis=1) / peak
    assert err.max() <= 1e-4, err.max()               # north_star tolerance
    assert err.max() <= 1e-5, err.max()               # what float32 should actually deliver
    assert np.array_equal(d["cpi"], np.arange(n))
    check_detections(jrc, d, do, f"tiled {name}", rtol_peak=1e-5)
    # detections without a caller-provided map use the handle's scratch map
    _, d2 = ch.run_host(rx, tx, want_map=False)
    for f in ("range_idx", "angle_idx", "peak_power", "noise_power", "flags"):
        assert np.array_equal(d2[f], d[f]), f


def test_gate_decision_at_the_threshold(jrc, orc):
    """SNR thresholds placed within +-1e-4 dB of a CPI's own SNR (and exactly on it): the fused and tiled paths take the
    oracle's gate decision, because a gate inside the error bound of the fast noise estimate is redone with the
    reference's sequential window sum (DET_EXACT)."""
    for name, n in (("C2", 8), ("C1", 4), ("C3s", 4)):
        cfg = CFGS[name]
        est = est_for(cfg)
        rx, tx, _ = scene(cfg, n, seed=41, n_targets=2, amp_db_span=6.0, snr_db=10.0, tx_per_cpi=True)
        _, _, do0 = oracle(orc, rx, tx, cfg, est)
        ch = gpu_chain(jrc, cfg, est)
        n_exact = 0
        for j in range(n):
            for delta in (-1e-4, -2e-5, -1e-6, 0.0, 1e-6, 2e-5, 1e-4):
                thr = np.float32(np.float64(do0["snr_db"][j]) + delta)
                est_j = dict(est, snr_threshold=thr)
                ch.set_thresholds(thr, est["power_threshold"])
                _, d = ch.run_host(rx, tx)
                _, _, do = oracle(orc, rx, tx, cfg, est_j)
                assert np.array_equal(d["flags"] & jrc.DET_PASSED, do["flags"] & 1), (name, j, delta)
                if abs(delta) <= 2e-5:            # well inside every configuration's error bound: redone, hence identical bits
                    assert d["flags"][j] & jrc.DET_EXACT, (name, j, delta)
                    for f in ("range_idx", "angle_idx", "peak_power", "noise_power", "snr_db", "n_noise"):
                        assert d[f][j] == do[f][j], (name, j, delta, f)
                n_exact += int(((d["flags"] & jrc.DET_EXACT) != 0).sum())
        # a threshold far from every SNR leaves the fast path's records alone
        ch.set_thresholds(np.float32(-50.0), est["power_threshold"])
        _, d = ch.run_host(rx, tx)
        assert ((d["flags"] & jrc.DET_EXACT) != 0).sum() <= 1 and (d["flags"] & jrc.DET_PASSED).all()
        print(f"[gate {name}] {n_exact} records redone over {7 * n} threshold placements")


def test_exact_ties_in_the_map(jrc, orc):
    """Two echoes half the unambiguous range apart, same angle, same amplitude, no noise: the second is the first times
    (-1)^k (the carrier phase of the extra delay is a whole number of turns), so the map repeats after Nr/2 rows and every
    CPI has two maxima that are equal up to rounding.  The arg-max is then taken in the reference's order on the
    candidates -- first maximum in row-major order wins -- on the fused path (inside k_fused64x8) and on the tiled paths
    (k_est_exact's cluster: the candidate cells bin by bin, with 8, 8 and 64 range inputs and 1, 1 and 4 channels per
    lane for C3s, C3 and C5).  The second half of the CPIs has amplitudes 1e-6 apart."""
    for name, n in (("C2", 12), ("C3s", 12), ("C3", 6), ("C5", 4)):
        cfg = CFGS[name]
        est = est_for(cfg)
        tx = synth.tx_symbols(cfg["T"], cfg["S"], cfg["N"])
        rng = np.random.default_rng(5)
        half_range = synth.C_LIGHT / (2 * 125e6) * (cfg["N"] / 2)
        r1 = rng.uniform(0.05, 0.4, n) * 2 * half_range
        r = np.stack([r1, r1 + half_range], axis=1)
        az = rng.uniform(-40, 40, n)
        a = np.stack([az, az], axis=1)
        amp = np.ones((n, 2))
        amp[n // 2:, 1] = 1.0 + 1e-6 * rng.standard_normal(n - n // 2)       # equal and almost equal heights
        rx = synth.rx_symbols(tx, cfg["R"], r, a, amp)
        ch = gpu_chain(jrc, cfg, est)
        _, d = ch.run_host(rx, tx)
        mo, _, do = oracle(orc, rx, tx, cfg, est)
        top = np.sort(mo.reshape(n, -1), axis=1)[:, -2:]
        assert ((top[:, 1] - top[:, 0]) <= 4e-6 * top[:, 1]).sum() >= n // 2      # the scene does what it says
        check_detections(jrc, d, do, f"ties {name}", rtol_peak=1e-5)
        st = ch.exact_stats()
        print(f"[ties {name}] {st}")
        assert st["marked"] + st["ties_in_kernel"] >= n // 2, st          # and those CPIs went through the candidate logic


@pytest.mark.parametrize("n", [1040, 1024, 250])
def test_target_simulator_is_bit_exact(jrc, orc, n):
    """target_simulator on the GPU.  Filters with the reference's float arithmetic on
    the host, FFT/IFFT on the device with the CPU restatement's arithmetic (radix-2 float32 for powers of two,
    float64 direct DFT otherwise): identical bits, for 1-3 targets, both folding rules, self coupling."""
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    pos = [0.0, 0.00625, 0.0125, 0.01875]                     # lambda/2 spacing at 24 GHz
    for (rg, vel, rcs, az, sc, acc) in [([12.0], [0.0], [1.0], [10.0], False, False),
                                        ([7.5, 30.0], [3.0, -8.0], [1.0, 20.0], [-25.0, 40.0], True, False),
                                        ([7.5, 30.0, 18.0], [0.0, 0.0, 5.0], [1.0, 20.0, 3.0], [-25.0, 40.0, 0.0], True, True)]:
        blk = jrc.target_simulator(rg, vel, rcs, az, pos, 125000000, 24e9, -10.0, False, sc, accumulate=acc)
        out, tags = blk.work(x)
        ref = orc.target_simulator(x, rg, vel, rcs, az, pos, 125000000, 24e9, sc, -10.0, acc)
        assert out.shape == ref.shape == (4, n)
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), (n, len(rg), np.abs(out - ref).max())
        assert tags[0] == ("rx_time", (0, 0.0), "stat_targ_sim") and len(tags) == 4
        out2, tags2 = blk.work(x)                                # second packet: same samples, later rx_time
        assert np.array_equal(out2, out) and tags2[0][1][0] == 0 and abs(tags2[0][1][1] - n / 125e6) < 1e-9
    # random phase shift: every target's echo rotated by a unit-modulus factor
    blk = jrc.target_simulator([12.0], [0.0], [1.0], [10.0], pos, 125000000, 24e9, -10.0, True, False)
    o1, _ = blk.work(x)
    base = orc.target_simulator(x, [12.0], [0.0], [1.0], [10.0], pos, 125000000, 24e9, False, -10.0, False)
    ratio = o1[0][np.abs(base[0]) > 1e-9] / base[0][np.abs(base[0]) > 1e-9]
    assert np.allclose(np.abs(ratio), 1.0, atol=1e-4) and np.allclose(ratio, ratio[0], atol=1e-3)


@pytest.mark.parametrize("name", ["C2", "C3s"])
def test_degenerate_inputs(jrc, orc, name):
    """Empty batch, all-zero CPIs and NaN CPIs on the fused (C2) and the tiled (C3s) path: no launch for an empty
    batch; a zero map keeps the scan's initial peak (0, 0) like the reference (strict '>' never fires) and is
    gated out; a NaN CPI is reported with the -1 sentinel and no detection, and does not disturb its neighbours."""
    cfg = CFGS[name]
    est = est_for(cfg)
    ch = gpu_chain(jrc, cfg, est)
    rx, tx, _ = scene(cfg, 6, seed=23, tx_per_cpi=True)
    l0 = ch.launch_count
    m, d = ch.run_host(rx[:0], tx[:0])
    assert m.shape[0] == 0 and d.shape[0] == 0 and ch.launch_count == l0
    rx[1] = 0
    rx[3] = np.nan
    m, d = ch.run_host(rx, tx)
    mo, _, do = oracle(orc, rx, tx, cfg, est)
    good = [0, 2, 4, 5]
    assert np.array_equal(d["range_idx"][good], do["range_idx"][good]) and np.array_equal(d["angle_idx"][good], do["angle_idx"][good])
    assert np.array_equal(d["flags"][good], do["flags"][good]) and d["flags"][good].all()
    assert not m[1].any() and (d["range_idx"][1], d["angle_idx"][1], d["flags"][1] & jrc.DET_PASSED) == (0, 0, 0)
    assert (do["range_idx"][1], do["angle_idx"][1], do["flags"][1]) == (0, 0, 0)
    assert np.isnan(m[3]).all() and (d["flags"][3] & jrc.DET_PASSED) == 0 and d["range_idx"][3] == -1 and do["flags"][3] == 0


def test_tiled_path_with_background_removal_and_requests(jrc, orc):
    """Tiled path: per-CPI TX frames, the background ring in front of the range FFT, explicit path requests."""
    cfg = CFGS["C3s"]
    est = est_for(cfg)
    n = 10
    rx, tx, _ = scene(cfg, n, seed=29, n_targets=2, amp_db_span=6.0, tx_per_cpi=True)
    ch = gpu_chain(jrc, cfg, est, background_removal=True, background_recording=True, record_len=3)
    m, d = ch.run_host(rx, tx)
    assert ch.last_path == jrc.PATH_TILED
    rad = orc.Radar(cfg["N"], cfg["T"], cfg["R"], cfg["S"], 0, True, True, 3, cfg["IR"], False)
    V = cfg["T"] * cfg["R"]
    for c in range(n):
        pad = rad.work(list(tx[c].reshape(cfg["T"], -1)), list(rx[c].reshape(cfg["R"], -1)))
        mo = orc.mag_squared(orc.fft_vcc(orc.matrix_transpose(orc.fft_vcc(pad, False, False), V, cfg["IA"]), True, True))
        assert np.abs(m[c] - mo).max() <= 1e-5 * mo.max() + 1e-3, c
    import torch
    rc = jrc.radar_chain(cfg["N"], cfg["T"], cfg["R"], cfg["S"], cfg["IR"], cfg["IA"], estimator=est)
    drx, dtx = torch.from_numpy(rx).cuda(), torch.from_numpy(tx).cuda()
    m1, d1 = rc.run(drx, dtx, path=jrc.PATH_TILED)
    m2, d2 = rc.run(drx, dtx, path=jrc.PATH_STAGED)
    rc.sync()
    assert (m1 - m2).abs().max().item() <= 1e-5 * m2.max().item()
    with pytest.raises(jrc.JrcError):
        rc.run(drx, dtx, path=jrc.PATH_FUSED)            # no fused specialisation for 32 channels
    rc2 = jrc.radar_chain(32, 2, 2, 3, 2, 4, estimator=synth.default_estimator_params(32, 4, 2, 4))   # 64 x 16 map
    rx2, tx2, _ = scene(CFGS["odd"], 4, seed=1)
    with pytest.raises(jrc.JrcError):
        rc2.run(torch.from_numpy(rx2).cuda(), torch.from_numpy(tx2).cuda(), path=jrc.PATH_TILED)   # Na = 16 < 64


@pytest.mark.parametrize("name", ["C3", "C5"])
def test_shape_specific_paths_with_background_removal_and_per_cpi_tx(jrc, orc, name):
    """k_slice256 (configs[2]) and the k_wide_* pair (configs[4]) behind the background ring (channel estimates instead of
    symbols as their input) and with per-CPI TX frames (the TX tensor map of the TMA loads then has a CPI axis)."""
    cfg = CFGS[name]
    est = est_for(cfg)
    n = 5
    rx, tx, _ = scene(cfg, n, seed=31, n_targets=2, amp_db_span=6.0, tx_per_cpi=True)
    V = cfg["T"] * cfg["R"]
    # per-CPI TX frames, no background: against the oracle chain
    ch0 = gpu_chain(jrc, cfg, est)
    m0, d0 = ch0.run_host(rx, tx)
    assert ch0.last_path == jrc.PATH_TILED
    mo, _, do = oracle(orc, rx, tx, cfg, est)
    err = np.abs(m0 - mo).reshape(n, -1).max(axis=1) / mo.reshape(n, -1).max(axis=1)
    assert err.max() <= 5e-6, err.max()
    check_detections(jrc, d0, do, f"per-CPI TX {name}")
    # background removal + recording: block by block against the oracle's radar block
    ch = gpu_chain(jrc, cfg, est, background_removal=True, background_recording=True, record_len=3)
    m, d = ch.run_host(rx, tx)
    assert ch.last_path == jrc.PATH_TILED
    rad = orc.Radar(cfg["N"], cfg["T"], cfg["R"], cfg["S"], 0, True, True, 3, cfg["IR"], False)
    for c in range(n):
        pad = rad.work(list(tx[c].reshape(cfg["T"], -1)), list(rx[c].reshape(cfg["R"], -1)))
        mo_c = orc.mag_squared(orc.fft_vcc(orc.matrix_transpose(orc.fft_vcc(pad, False, False), V, cfg["IA"]), True, True))
        assert np.abs(m[c] - mo_c).max() <= 1e-5 * mo_c.max() + 1e-3, c


def test_tiled_path_chunks_long_batches(jrc, orc):
    """A batch whose scratch (range spectra + map) exceeds the 1 GiB budget runs in chunks: records of CPIs on both
    sides of the chunk borders equal the oracle's, the CPI ids are continuous."""
    import torch
    cfg = CFGS["C3"]
    est = est_for(cfg)
    n = 450                                               # 252 CPIs per chunk when the map is scratch as well
    rx, tx, _ = scene(cfg, n, seed=31)
    rc = jrc.radar_chain(cfg["N"], cfg["T"], cfg["R"], cfg["S"], cfg["IR"], cfg["IA"], estimator=est)
    l0 = rc.chain.launch_count
    _, d = rc.run(torch.from_numpy(rx).cuda(), torch.from_numpy(tx).cuda(), want_map=False, cpi0=1000)
    rc.sync()
    d = rc.dets_to_numpy(d)
    assert rc.chain.last_path == jrc.PATH_TILED and rc.chain.launch_count - l0 >= 2 * 4     # two chunks of 4 kernels
    idx = [0, 250, 251, 252, 253, 449]
    _, _, do = oracle(orc, rx[idx], tx, cfg, est)
    assert np.array_equal(d["range_idx"][idx], do["range_idx"]) and np.array_equal(d["angle_idx"][idx], do["angle_idx"])
    assert np.array_equal(d["cpi"], 1000 + np.arange(n)) and (d["flags"] & 1).all()


def test_tiled_linearity_and_shard_consistency(jrc):
    """Size-independent properties of the tiled path at BASELINE configs[2] size: |chain(2x)|^2 = 4 |chain(x)|^2 exactly
    (power-of-two scaling commutes with every rounding), and processing a batch as two shards (the multi-GPU split)
    gives the same bits as processing it whole."""
    cfg = CFGS["C3"]
    est = est_for(cfg)
    n = 6
    rx, tx, _ = scene(cfg, n, seed=37, n_targets=3, amp_db_span=10.0)
    ch = gpu_chain(jrc, cfg, est)
    m1, d1 = ch.run_host(rx, tx)
    assert ch.last_path == jrc.PATH_TILED
    m2, d2 = ch.run_host(rx * np.float32(2), tx)
    assert np.array_equal(m2, m1 * np.float32(4))
    assert np.array_equal(d1["range_idx"], d2["range_idx"]) and np.array_equal(d1["angle_idx"], d2["angle_idx"])
    ma, da = ch.run_host(rx[:2], tx, cpi0=0)
    mb, db = ch.run_host(rx[2:], tx, cpi0=2)
    assert np.array_equal(np.concatenate([ma, mb]), m1)
    assert np.array_equal(np.concatenate([da, db]), d1)


def test_capture_radar_data_line_equals_the_reference_blocks(jrc):
    """mimo_ofdm_radar::capture_radar_data (lib/mimo_ofdm_radar_impl.cc:348-377): the CSV line for the golden frame
    against the line the reference block itself wrote (tests/golden/c1_capture_line.txt, make_golden.py); only the time
    stamp differs.  (The C++ block is checked the same way by tests/cpp/test_blocks.cc.)"""
    gdir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")
    frame = np.fromfile(_os.path.join(gdir, "c1_frame0.c64"), dtype=np.complex64).reshape(6, 4 * 64)
    golden = open(_os.path.join(gdir, "c1_capture_line.txt")).read()
    path = "/tmp/jrc_py_capture.csv"
    if _os.path.exists(path):
        _os.remove(path)
    blk = jrc.mimo_ofdm_radar(64, 4, 2, 4, 0, False, False, 1, 1, False, path)
    out, tags, consumed = blk.work(list(frame[:4]), list(frame[4:]))
    assert out is not None
    blk.capture_radar_data(True)
    line = open(path).read()
    stamp, rest = line.split(", ", 1)
    assert len(stamp) == 12 and rest == golden


@pytest.mark.parametrize("name", ["C2", "C5"])
def test_scene_synthesis_on_device(jrc, name):
    """jrc_scene_synth (batched point-target RX symbols, the generator of the configs[4] sweep) against the NumPy float64
    model of the same formula (synth.rx_symbols); the noise it adds has the requested level."""
    import torch
    cfg = CFGS[name]
    n = 6 if name == "C5" else 40
    rng = np.random.default_rng(77)
    tx = synth.tx_symbols(cfg["T"], cfg["S"], cfg["N"])
    r, a, amp = (x.astype(np.float32).astype(np.float64) for x in synth.random_scene(rng, n, 3, cfg["N"], amp_db_span=12.0))
    ref = synth.rx_symbols(tx, cfg["R"], r, a, amp, chunk=8)          # the C ABI takes the scene as float32
    ch = jrc.Chain(cfg["N"], cfg["T"], cfg["R"], cfg["S"], 0, cfg["IR"], cfg["IA"])
    rx = torch.empty((n, cfg["R"], cfg["S"], cfg["N"]), dtype=torch.complex64, device="cuda")
    ch.scene_synth_ptr(tx, r, a, amp, rx.data_ptr())
    got = rx.cpu().numpy()
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 2e-5 * scale, np.abs(got - ref).max() / scale       # float32 range/azimuth inputs, float32 sums
    ch.scene_synth_ptr(tx, r, a, amp, rx.data_ptr(), noise_sigma=0.25, seed=5)
    nz = rx.cpu().numpy() - got
    assert abs(nz.real.std() - 0.25) < 0.01 and abs(nz.imag.std() - 0.25) < 0.01 and abs(nz.mean()) < 0.01
    ch.scene_synth_ptr(tx, r, a, amp, rx.data_ptr(), noise_sigma=0.25, seed=6)
    assert not np.array_equal(rx.cpu().numpy() - got, nz)                                  # another seed, another noise


def test_range_doppler_angle_cube(jrc, orc):
    """a burst of CPIs -> range-Doppler-angle cube (jrc_chain_run_burst).  No counterpart in the
    reference; parity against NumPy float64 on the oracle's complex maps, and a moving point target sits on the analytic
    Doppler bin."""
    import torch
    cfg = dict(T=4, R=2, S=4, N=64, IR=4, IA=4)
    nb, prf, fc = 32, 20e3, 24e9
    v = 15.0                                                        # m/s towards the radar
    fd = 2 * v * fc / 3e8                                           # 2.4 kHz
    tx = synth.tx_symbols(cfg["T"], cfg["S"], cfg["N"])
    t = np.arange(nb) / prf
    rx = synth.rx_symbols(tx, cfg["R"], np.full((nb, 1), 20.0), np.full((nb, 1), 15.0), np.ones((nb, 1)))
    rx = (rx * np.exp(2j * np.pi * fd * t)[:, None, None, None]).astype(np.complex64)
    rng = np.random.default_rng(3)
    rx += (0.01 * (rng.standard_normal(rx.shape) + 1j * rng.standard_normal(rx.shape))).astype(np.complex64)
    est = est_for(cfg)
    ch = gpu_chain(jrc, cfg, est)
    drx, dtx = torch.from_numpy(rx).cuda(), torch.from_numpy(tx).cuda()
    cube = torch.empty((ch.Nr, ch.Na, nb), dtype=torch.float32, device="cuda")
    per_ant = cfg["S"] * cfg["N"]
    torch.cuda.synchronize()
    ch.run_burst_ptr(drx.data_ptr(), cfg["R"] * per_ant, per_ant, dtx.data_ptr(), 0, per_ant, nb, cube.data_ptr())
    ch.sync()
    got = cube.cpu().numpy()
    _, cm, _ = oracle(orc, rx, tx, cfg, est, want_cmap=True)        # [nb][Nr][Na] complex64, oracle arithmetic
    ref = np.abs(np.fft.fftshift(np.fft.fft(cm.astype(np.complex128), axis=0), axes=0)) ** 2
    ref = np.moveaxis(ref, 0, 2)
    assert np.abs(got - ref).max() <= 2e-5 * ref.max(), np.abs(got - ref).max() / ref.max()
    n, i, d = np.unravel_index(np.argmax(got), got.shape)
    assert d == nb // 2 + int(round(fd / prf * nb)) and (n, i) == synth.expected_peak(20.0, 15.0, 64, 4, 8, 4)


@pytest.mark.parametrize("name,pre,n,cp", [("C2", 5, 300, 16), ("C1", 0, 64, 16), ("C3s", 2, 12, 64), ("C5", 1, 3, 512),
                                            ("C2", 1, 40, 15)])       # (odd prefix: rows not 16-byte aligned, generic kernel)
def test_chain_from_raw_time_samples(jrc, orc, name, pre, n, cp):
    """SURVEY.md 8(f) rank 1 for whole batches (jrc_chain_run_batch_time): cyclic-prefix removal + the RX OFDM FFT in
    front of the chain, on the device.  The demodulated symbols are bit-identical to the oracle's cp_remove + fft_vcc, so
    map and records equal, bit for bit, the ones of jrc_chain_run_batch on the oracle-demodulated symbols; and the
    detection lists equal the oracle chain's."""
    import torch
    cfg = CFGS[name]
    N, T, R, S = cfg["N"], cfg["T"], cfg["R"], cfg["S"]
    est = est_for(cfg)
    rx, tx, _ = scene(cfg, n, seed=17, n_targets=2, amp_db_span=6.0)
    rng = np.random.default_rng(5)
    sym_all = np.empty((n, R, pre + S, N), np.complex64)           # what the antennas' demodulators should deliver
    sym_all[:, :, :pre] = (rng.standard_normal((n, R, pre, N)) + 1j * rng.standard_normal((n, R, pre, N))).astype(np.complex64)
    sym_all[:, :, pre:] = rx
    td = np.fft.ifft(np.fft.ifftshift(sym_all.astype(np.complex128), axes=-1), axis=-1)
    td = np.concatenate([td[..., N - cp:], td], axis=-1).astype(np.complex64)      # [n][R][pre+S][cp+N]
    tx_all = np.concatenate([np.ones((T, pre, N), np.complex64), tx], axis=1)      # TX packets carry the preamble too
    # the oracle's front end, antenna row by antenna row
    dem = np.empty_like(sym_all)
    for c in range(n):
        for r in range(R):
            dem[c, r] = orc.fft_vcc(orc.cp_remove(td[c, r].ravel(), pre + S, N, cp), True, True)
    ch = jrc.Chain(N, T, R, S, pre, cfg["IR"], cfg["IA"])
    ch.set_estimator(**est)
    Nr, Na = ch.Nr, ch.Na
    dtd, dtx, ddem = torch.from_numpy(td).cuda(), torch.from_numpy(tx_all).cuda(), torch.from_numpy(dem).cuda()
    m1 = torch.empty((n, Nr, Na), dtype=torch.float32, device="cuda")
    m2 = torch.empty_like(m1)
    d1 = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    d2 = torch.zeros_like(d1)
    torch.cuda.synchronize()
    row_t, row_f = (pre + S) * (N + cp), (pre + S) * N
    ch.run_batch_time_ptr(dtd.data_ptr(), R * row_t, row_t, cp, dtx.data_ptr(), 0, row_f, n, 0, m1.data_ptr(), None, d1.data_ptr())
    path = ch.last_path
    ch.run_batch_ptr(ddem.data_ptr(), R * row_f, row_f, dtx.data_ptr(), 0, row_f, n, 0, m2.data_ptr(), None, d2.data_ptr())
    ch.sync()
    assert ch.last_path == path == (jrc.PATH_FUSED if name in ("C1", "C2") else jrc.PATH_TILED)
    assert torch.equal(m1, m2)
    assert torch.equal(d1, d2)
    mo, _, do = oracle(orc, np.ascontiguousarray(dem[:, :, pre:]), tx, cfg, est)
    err = np.abs(m1.cpu().numpy() - mo).reshape(n, -1).max(axis=1) / mo.reshape(n, -1).max(axis=1)
    assert err.max() <= 5e-6, err.max()
    check_detections(jrc, jrc.radar_chain.dets_to_numpy(d1), do, f"time samples {name}")
