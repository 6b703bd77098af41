"""BASELINE configs[0]: the shipped simulation flowgraph
(examples/simulation/radar/mimo_ofdm_jrc_radar_sim.grc) from the precoder's frequency-domain frame
through the TIME-DOMAIN channel to the detections:

  frame [sync x4 | SIG | MIMO-LTF x4] -> fft_vxx(IFFT 64, shift, window 1/8) -> cyclic prefix (16)
  -> zero_pad(tail = 3*80) -> target_simulator (one per TX, RX positions TXn_RXs) -> sum over TX
  -> ofdm_cyclic_prefix_remover -> fft_vxx(FFT 64, shift) -> mimo_ofdm_radar (N_pre = 5) -> ...

The channel is the oracle's restatement of target_simulator (pinned bit-for-bit against the reference's
own target_simulator_impl.cc in tests/test_oracle_vs_ref.py).  The CPU test checks the analytic peak
bins of SURVEY.md section 4 through the oracle chain; the GPU test runs the same frames through the
fused kernel and through the per-block mirrors."""
import numpy as np
import pytest

from mimo_ofdm_jrc import synth

FS, FC = 125_000_000, 24e9
LAM = 3e8 / FC
T, R, S, N, CP, PRE, IR, IA = 4, 2, 4, 64, 16, 5, 8, 16
# targets inside the cyclic prefix (16 samples = 19.2 m two-way): beyond it the previous OFDM symbol leaks into
# the FFT window of the real system as well, and the analytic bin no longer applies
KATS = [((10.0, 0.0), (67, 64)), ((10.0, -3.0), (67, 61)),
        ((15.0, 20.0), synth.expected_peak(15.0, 20.0, 64, 8, 8, 16)), ((5.0, -40.0), synth.expected_peak(5.0, -40.0, 64, 8, 8, 16))]


def sim_frames(orc, rng_m, az_deg):
    rng = np.random.default_rng(42)
    ltf = synth.tx_symbols(T, S, N)                                   # [T][S][N]
    qpsk = ((rng.integers(0, 2, (T, PRE, N)) * 2 - 1) + 1j * (rng.integers(0, 2, (T, PRE, N)) * 2 - 1)) / np.sqrt(2)
    qpsk[:, :, synth.LTF_64 == 0] = 0
    txf = np.concatenate([qpsk, ltf], axis=1).astype(np.complex64)    # [T][PRE+S][N] precoder output
    nsym = PRE + S
    rx_time = np.zeros((R, nsym * (N + CP) + 3 * (N + CP)), dtype=np.complex64)
    for t in range(T):
        td = orc.fft_vcc(txf[t] * np.float32(1 / 8), forward=False, shift=True)          # IFFT, window 1/sqrt(64)
        td = np.concatenate([td[:, -CP:], td], axis=1).reshape(-1)                       # cyclic prefixer
        pkt = orc.zero_pad(td, 0, 3 * (N + CP), seed=7 + t)
        pos = [(1 + 0.5 * t) * LAM, (3 + 0.5 * t) * LAM]                                 # TXn_RXs, ...radar_sim.grc:105-147
        rx_time += orc.target_simulator(pkt, [rng_m], [0.0], [10.0], [az_deg], pos, FS, FC)
    rxf = np.stack([orc.fft_vcc(orc.cp_remove(rx_time[r], nsym, N, CP), forward=True, shift=True) for r in range(R)])
    return txf, rxf.astype(np.complex64)                              # [T][9][64], [R][9][64]


@pytest.mark.parametrize("target,peak", KATS)
def test_time_domain_sim_hits_the_analytic_bins(orc, target, peak):
    txf, rxf = sim_frames(orc, *target)
    est = synth.default_estimator_params(N, T * R, IR, IA)
    m, _, d = orc.chain_batch(rxf[None], txf[None], N, T, R, S, IR, IA, est, n_pre=PRE)
    assert (d[0]["range_idx"], d[0]["angle_idx"]) == peak
    assert d[0]["flags"] == 1


@pytest.mark.gpu
@pytest.mark.parametrize("target,peak", KATS)
def test_time_domain_sim_on_gpu(jrc, orc, target, peak):
    txf, rxf = sim_frames(orc, *target)
    est = synth.default_estimator_params(N, T * R, IR, IA)
    mo, cmo, do = orc.chain_batch(rxf[None], txf[None], N, T, R, S, IR, IA, est, n_pre=PRE, want_cmap=True)
    # fused kernel on the LTF symbols
    ch = jrc.Chain(N, T, R, S, 0, IR, IA)
    ch.set_estimator(**est)
    m, d = ch.run_host(rxf[None, :, PRE:], txf[None, :, PRE:])
    assert (d[0]["range_idx"], d[0]["angle_idx"]) == peak
    assert np.abs(m[0] - mo[0]).max() <= 1e-4 * mo[0].max()
    np.testing.assert_allclose(d[0]["snr_db"], do[0]["snr_db"], atol=5e-3)
    # block by block, like the flowgraph (packets WITH the preamble, N_pre = 5): identical bits
    radar = jrc.mimo_ofdm_radar(N, T, R, S, PRE, False, False, 8, IR, False, "/tmp/jrc_sim_chan.csv")
    pad, tags, _ = radar.work(list(txf.reshape(T, -1)), list(rxf.reshape(R, -1)))
    y = jrc.fft_vcc(N * IR, False, shift=False).work(pad)
    cm = jrc.fft_vcc(T * R * IA, True, shift=True).work(jrc.matrix_transpose(N * IR, T * R, IA).work(y))
    assert np.array_equal(cm, cmo[0])
    blk = jrc.range_angle_estimator(T * R * IA, est["range_bins"], est["angle_bins"], est["noise_discard_range_m"],
                                    est["noise_discard_angle_deg"], 15.0, 0.0, "/tmp/jrc_sim_log.csv", False)
    det = blk.work(cm)
    assert (det["range_idx"], det["angle_idx"]) == peak and det["snr_db"] == do[0]["snr_db"]
    assert blk.messages and blk.messages[0][0][1][0] == est["range_bins"][peak[0]]


@pytest.mark.gpu
@pytest.mark.parametrize("target,peak", KATS)
def test_time_domain_sim_every_block_on_gpu(jrc, orc, target, peak):
    """The whole simulation flowgraph through the GPU blocks: TX IFFT, zero_pad, target_simulator (one per TX),
    ofdm_cyclic_prefix_remover fused with the RX FFT, fused radar chain.  With the oracle's pad noise the
    received symbols are bit-identical to the CPU flow; with the block's own noise the peak is unchanged."""
    rng_m, az_deg = target
    rng = np.random.default_rng(42)
    ltf = synth.tx_symbols(T, S, N)
    qpsk = ((rng.integers(0, 2, (T, PRE, N)) * 2 - 1) + 1j * (rng.integers(0, 2, (T, PRE, N)) * 2 - 1)) / np.sqrt(2)
    qpsk[:, :, synth.LTF_64 == 0] = 0
    txf = np.concatenate([qpsk, ltf], axis=1).astype(np.complex64)
    nsym = PRE + S
    ifft = jrc.fft_vcc(N, False, shift=True)
    zp = jrc.zero_pad(False, 0, 3 * (N + CP))
    cpr = jrc.ofdm_cyclic_prefix_remover(N, CP)
    est = synth.default_estimator_params(N, T * R, IR, IA)
    ch = jrc.Chain(N, T, R, S, 0, IR, IA)
    ch.set_estimator(**est)
    _, rxf_cpu = sim_frames(orc, rng_m, az_deg)
    for own_noise in (False, True):
        rx_time = np.zeros((R, nsym * (N + CP) + 3 * (N + CP)), dtype=np.complex64)
        for t in range(T):
            td = ifft.work(txf[t] * np.float32(1 / 8))
            td = np.concatenate([td[:, -CP:], td], axis=1).reshape(-1)
            pkt = zp.work(td) if own_noise else orc.zero_pad(td, 0, 3 * (N + CP), seed=7 + t)
            pos = [(1 + 0.5 * t) * LAM, (3 + 0.5 * t) * LAM]
            sim = jrc.target_simulator([rng_m], [0.0], [10.0], [az_deg], pos, FS, FC, 0.0)
            rx_time += sim.work(pkt)[0]
        rxf = np.stack([cpr.work(rx_time[r][:nsym * (N + CP)], demod=True) for r in range(R)]).astype(np.complex64)
        if not own_noise:
            assert np.array_equal(rxf, rxf_cpu)
        m, d = ch.run_host(rxf[None, :, PRE:], txf[None, :, PRE:])
        assert (d[0]["range_idx"], d[0]["angle_idx"]) == peak and d[0]["flags"] == 1
