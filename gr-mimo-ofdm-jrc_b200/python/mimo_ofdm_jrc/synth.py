"""Synthetic radar scenes for parity tests and bench.py.

TX symbols follow the shipped flowgraph's `ofdm_config` module
(examples/simulation/radar/mimo_ofdm_jrc_radar_sim.grc:1611 ff.): MIMO-LTF symbol s of TX
antenna t on subcarrier k is P_ltf[t, s] * ltf_64[k] (DC-centred subcarrier order, as
delivered by the shift=True OFDM FFTs).  RX symbols are the point-target response of
lib/target_simulator_impl.cc:177,188,296-303 evaluated directly in the frequency domain:
    Y_r[k, s] = sum_t X_t[k, s] * sum_j a_j exp(-j 2 pi tau_{j,t,r} (f_k + fc)),
    tau = (2 R_j - d_{t,r} sin(az_j)) / c,   d_{t,r} = lambda + (t + T r) lambda / 2
(the TXn_RXs antenna positions of ...radar_sim.grc:105-147).
Plain NumPy; nothing here is on the measured path.
"""
from __future__ import annotations

import numpy as np

C_LIGHT = 3e8

LTF_64 = np.array([0, 0, 0, 0, 1, 1, 1, 1, -1, -1, 1, 1, -1, 1, -1, 1, 1, 1, 1, 1, 1, -1, -1, 1, 1, -1, 1, -1,
                   1, 1, 1, 1, 0, 1, -1, -1, 1, 1, -1, 1, -1, 1, -1, -1, -1, -1, -1, 1, 1, -1, -1, 1, -1, 1,
                   -1, 1, 1, 1, 1, -1, -1, 0, 0, 0], dtype=np.float64)
P_LTF_4 = np.array([[1, -1, 1, 1], [1, 1, -1, 1], [1, 1, 1, -1], [-1, 1, 1, 1]], dtype=np.float64)


def ltf_sequence(nsc: int) -> np.ndarray:
    """+-1 training sequence with the ltf_64 null pattern (guards + DC) scaled to nsc."""
    if nsc == 64:
        return LTF_64.copy()
    rng = np.random.default_rng(nsc)
    seq = rng.integers(0, 2, nsc) * 2.0 - 1.0
    lo, hi = nsc // 16, nsc // 16 - 1
    seq[:lo] = 0
    if hi > 0:
        seq[-hi:] = 0
    seq[nsc // 2] = 0
    return seq


def p_matrix(T: int) -> np.ndarray:
    """T x T orthogonal +-1 cover matrix (ofdm_config.P_ltf for T = 4, Sylvester-Hadamard else)."""
    if T == 4:
        return P_LTF_4.copy()
    if T & (T - 1):
        raise ValueError("n_tx must be a power of two")
    H = np.array([[1.0]])
    while H.shape[0] < T:
        H = np.block([[H, H], [H, -H]])
    return H


def tx_symbols(T: int, S: int, nsc: int) -> np.ndarray:
    """[T][S][nsc] complex64 MIMO-LTF symbols."""
    P = p_matrix(T)
    ltf = ltf_sequence(nsc)
    X = P[:, np.arange(S) % T][:, :, None] * ltf[None, None, :]
    return np.ascontiguousarray(X.astype(np.complex64))


def range_bins(nsc: int, interp: int, samp_rate: float = 125e6) -> np.ndarray:
    """...radar_sim.grc:1400 -- np.linspace(0, 3e8*fft_len/(2*samp_rate), fft_len*interp)"""
    return np.linspace(0, 3e8 * nsc / (2 * samp_rate), nsc * interp).astype(np.float32)


def angle_bins(V: int, interp: int) -> np.ndarray:
    """...radar_sim.grc:155-167 (angle_axis)"""
    Na = V * interp
    return (np.arcsin(2 / Na * (np.arange(0, Na) - np.floor(Na / 2) + 0.5)) * 180 / np.pi).astype(np.float32)


def range_resolution(samp_rate: float = 125e6) -> float:
    return 3e8 / (2 * samp_rate)                       # R_res, ...radar_sim.grc:99


def angle_resolution(V: int) -> float:
    return float(np.rad2deg(np.arcsin(2 / V)))          # angle_res, ...radar_sim.grc:172


def default_estimator_params(nsc, V, ir, ia, samp_rate=125e6, snr_threshold=15.0, power_threshold=0.0):
    """The make() arguments the shipped flowgraph passes (...radar_sim.grc:1385-1411)."""
    return dict(range_bins=range_bins(nsc, ir, samp_rate), angle_bins=angle_bins(V, ia),
                noise_discard_range_m=np.float32(2 * range_resolution(samp_rate)),
                noise_discard_angle_deg=np.float32(2 * angle_resolution(V)),
                snr_threshold=np.float32(snr_threshold), power_threshold=np.float32(power_threshold))


def random_scene(rng, n_cpi, n_targets, nsc, samp_rate=125e6, amp_db_span=0.0):
    r_max = 3e8 * nsc / (2 * samp_rate)
    rng_m = rng.uniform(2.0, 0.9 * r_max, (n_cpi, n_targets))
    az = rng.uniform(-60.0, 60.0, (n_cpi, n_targets))
    amp = 10 ** (-rng.uniform(0, amp_db_span, (n_cpi, n_targets)) / 20.0) if amp_db_span > 0 else np.ones((n_cpi, n_targets))
    amp[:, 0] = 1.0
    return rng_m, az, amp


def rx_symbols(tx, R, ranges_m, az_deg, amps, samp_rate=125e6, center_freq=24e9, snr_db=None, rng=None,
               chunk=512):
    """tx [T][S][nsc] -> rx [n_cpi][R][S][nsc] complex64 for per-CPI point targets [n_cpi][n_targets]."""
    T, S, nsc = tx.shape
    ranges_m = np.atleast_2d(np.asarray(ranges_m, dtype=np.float64))
    az_deg = np.atleast_2d(np.asarray(az_deg, dtype=np.float64))
    amps = np.atleast_2d(np.asarray(amps, dtype=np.float64))
    n_cpi = ranges_m.shape[0]
    lam = C_LIGHT / center_freq
    t_idx, r_idx = np.arange(T), np.arange(R)
    d = lam + (t_idx[None, :] + T * r_idx[:, None]) * lam / 2                     # [R][T]
    fk = (np.arange(nsc) - nsc // 2) * samp_rate / nsc + center_freq             # [nsc]
    out = np.empty((n_cpi, R, S, nsc), dtype=np.complex64)
    X = tx.astype(np.complex128)
    for c0 in range(0, n_cpi, chunk):
        sl = slice(c0, min(n_cpi, c0 + chunk))
        tau = (2 * ranges_m[sl, :, None, None] - d[None, None] * np.sin(np.deg2rad(az_deg[sl]))[:, :, None, None]) / C_LIGHT
        ph = -2 * np.pi * np.mod(tau[..., None] * fk, 1.0)                       # [c][j][R][T][nsc]
        Hch = (amps[sl, :, None, None, None] * np.exp(1j * ph)).sum(axis=1)       # [c][R][T][nsc]
        Y = np.einsum("tsk,crtk->crsk", X, Hch)
        if snr_db is not None:
            g = rng if rng is not None else np.random.default_rng(0)
            sig = np.sqrt(np.mean(np.abs(Y) ** 2))
            sigma = sig * 10 ** (-snr_db / 20.0) / np.sqrt(2)
            Y = Y + sigma * (g.standard_normal(Y.shape) + 1j * g.standard_normal(Y.shape))
        out[sl] = Y.astype(np.complex64)
    return out


def expected_peak(range_m, az_deg, nsc, ir, V, ia, samp_rate=125e6):
    """Analytic peak bin of a single noise-free point target."""
    r_idx = int(np.rint(2 * range_m / C_LIGHT * samp_rate * ir)) % (nsc * ir)
    Na = V * ia
    a_idx = int(np.rint(Na * np.sin(np.deg2rad(az_deg)) / 2 + Na / 2)) % Na
    return r_idx, a_idx
