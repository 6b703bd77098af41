"""ctypes loader for the CPU oracle (oracle/libjrc_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libjrc_oracle.so")

DET_DTYPE = np.dtype([("range_idx", "<i4"), ("angle_idx", "<i4"), ("peak_power", "<f4"),
                      ("noise_power", "<f4"), ("snr_db", "<f4"), ("n_noise", "<i4"),
                      ("flags", "<u4"), ("cpi", "<i4")])
DBG_DTYPE = np.dtype([("angle_null_idx", "<i4"), ("discard_range_idx", "<i4"), ("discard_angle_idx", "<i4"),
                      ("start_range_idx", "<i4"), ("end_range_idx", "<i4"), ("start_angle_idx", "<i4"),
                      ("end_angle_idx", "<i4"), ("range_val", "<f4"), ("angle_val", "<f4")])


class ChainCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("fft_len", "n_tx", "n_rx", "n_sym", "n_pre", "interp_range",
                                        "interp_angle", "tx_interleave")] + \
               [("range_bins", C.c_void_p), ("angle_bins", C.c_void_p)] + \
               [(n, C.c_float) for n in ("noise_discard_range_m", "noise_discard_angle_deg",
                                          "snr_threshold", "power_threshold")]


class Peak1d(C.Structure):
    _fields_ = [("k", C.c_int32), ("freq", C.c_float), ("phase", C.c_float), ("mag", C.c_float)]


_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "libjrc_oracle.so"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        l = C.CDLL(LIB)
        vp, ci, cf, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
        l.orc_radar_create.restype = vp
        l.orc_radar_create.argtypes = [ci] * 10
        l.orc_radar_destroy.argtypes = [vp]
        l.orc_radar_set_background_record.argtypes = [vp, ci]
        l.orc_radar_work.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), sz, vp]
        l.orc_radar_work.restype = ci
        l.orc_fft_vcc_batch.argtypes = [vp, vp, ci, ci, ci, ci]
        l.orc_matrix_transpose.argtypes = [vp, ci, ci, ci, ci, vp]
        l.orc_mag_squared.argtypes = [vp, vp, sz]
        l.orc_range_angle_estimate.argtypes = [vp, ci, ci, vp, ci, vp, ci, cf, cf, cf, cf, vp, vp]
        l.orc_fft_peak_detect.argtypes = [vp, ci, ci, cf, cf, ci, C.POINTER(Peak1d)]
        l.orc_zero_pad.argtypes = [vp, ci, C.c_uint, C.c_uint, C.c_uint64, vp]
        l.orc_cp_remove.argtypes = [vp, ci, ci, ci, vp]
        l.orc_target_simulator.argtypes = [vp, ci, vp, vp, vp, vp, ci, vp, ci, ci, cf, ci, cf, ci, vp]
        l.orc_chain_batch.argtypes = [C.POINTER(ChainCfg), vp, vp, ci, ci, ci, vp, vp, vp]
        _lib = l
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def c64(a):
    return np.ascontiguousarray(a, dtype=np.complex64)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Radar:
    """orc_radar: the mimo_ofdm_radar block state machine (background ring included)."""

    def __init__(self, fft_len, n_tx, n_rx, n_sym, n_pre, background_removal, background_recording,
                 record_len, interp_factor, tx_interleave):
        self.args = (fft_len, n_tx, n_rx, n_sym, n_pre, int(background_removal), int(background_recording),
                     record_len, interp_factor, int(tx_interleave))
        self.h = lib().orc_radar_create(*self.args)
        self.V, self.Nr = n_tx * n_rx, fft_len * interp_factor

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_radar_destroy(self.h)
            self.h = None

    def set_background_record(self, on):
        lib().orc_radar_set_background_record(self.h, int(bool(on)))

    def work(self, tx_ports, rx_ports, tx_skip_items=0):
        txs = [c64(a) for a in tx_ports]
        rxs = [c64(a) for a in rx_ports]
        tp = (C.c_void_p * len(txs))(*[a.ctypes.data for a in txs])
        rp = (C.c_void_p * len(rxs))(*[a.ctypes.data for a in rxs])
        out = np.empty((self.V, self.Nr), dtype=np.complex64)
        lib().orc_radar_work(self.h, tp, rp, tx_skip_items, _p(out))
        return out


def fft_vcc(x, forward, shift):
    x = c64(x)
    n = x.shape[-1]
    out = np.empty_like(x)
    lib().orc_fft_vcc_batch(_p(x), _p(out), n, x.size // n, int(bool(forward)), int(bool(shift)))
    return out


def matrix_transpose(x, output_len, interp):
    x = c64(x)
    k, input_len = x.shape
    out = np.empty((input_len, output_len * interp), dtype=np.complex64)
    lib().orc_matrix_transpose(_p(x), k, input_len, output_len, interp, _p(out))
    return out


def mag_squared(x):
    x = c64(x)
    out = np.empty(x.shape, dtype=np.float32)
    lib().orc_mag_squared(_p(x), _p(out), x.size)
    return out


def range_angle_estimate(cmap, range_bins, angle_bins, noise_discard_range_m, noise_discard_angle_deg,
                         snr_threshold, power_threshold, want_dbg=False):
    cmap = c64(cmap)
    n_inputs, vlen = cmap.shape
    rb, ab = f32(range_bins), f32(angle_bins)
    det = np.zeros(1, dtype=DET_DTYPE)
    dbg = np.zeros(1, dtype=DBG_DTYPE)
    lib().orc_range_angle_estimate(_p(cmap), n_inputs, vlen, _p(rb), rb.size, _p(ab), ab.size,
                                   noise_discard_range_m, noise_discard_angle_deg, snr_threshold,
                                   power_threshold, _p(det), _p(dbg))
    return (det[0], dbg[0]) if want_dbg else det[0]


def fft_peak_detect(x, samp_rate, interp_factor, threshold_db, samp_protect):
    x = c64(x)
    out = Peak1d()
    lib().orc_fft_peak_detect(_p(x), x.size, samp_rate, interp_factor, threshold_db, samp_protect, C.byref(out))
    return out.k, out.freq, out.phase, out.mag


def zero_pad(x, pad_front, pad_tail, seed):
    x = c64(x)
    out = np.empty(x.size + pad_front + pad_tail, dtype=np.complex64)
    lib().orc_zero_pad(_p(x), x.size, pad_front, pad_tail, seed, _p(out))
    return out


def cp_remove(x, n_sym, fft_len, cp_len):
    x = c64(x)
    out = np.empty((n_sym, fft_len), dtype=np.complex64)
    lib().orc_cp_remove(_p(x), n_sym, fft_len, cp_len, _p(out))
    return out


def target_simulator(x, rng_m, velocity, rcs, azimuth, position_rx, samp_rate, center_freq,
                     self_coupling=False, self_coupling_db=0.0, accumulate=False):
    x = c64(x)
    r, v, s, a, p = f32(rng_m), f32(velocity), f32(rcs), f32(azimuth), f32(position_rx)
    out = np.empty((p.size, x.size), dtype=np.complex64)
    lib().orc_target_simulator(_p(x), x.size, _p(r), _p(v), _p(s), _p(a), r.size, _p(p), p.size,
                               int(samp_rate), center_freq, int(self_coupling), self_coupling_db,
                               int(accumulate), _p(out))
    return out


def chain_batch(rx, tx, fft_len, n_tx, n_rx, n_sym, interp_range, interp_angle, est, tx_interleave=False,
                n_pre=0, want_map=True, want_cmap=False, want_dets=True, cpi0=0):
    """rx [n_cpi][R][n_pre+S][N], tx [n_cpi or 1][T][n_pre+S][N] -> (map, cmap, dets)."""
    rx, tx = c64(rx), c64(tx)
    if tx.ndim == 3:
        tx = tx[None]
    n_cpi = rx.shape[0]
    tx_shared = tx.shape[0] == 1 and n_cpi > 1
    Nr, Na = fft_len * interp_range, n_tx * n_rx * interp_angle
    rb, ab = f32(est["range_bins"]), f32(est["angle_bins"])
    cfg = ChainCfg(fft_len, n_tx, n_rx, n_sym, n_pre, interp_range, interp_angle, int(tx_interleave),
                   rb.ctypes.data, ab.ctypes.data, est["noise_discard_range_m"], est["noise_discard_angle_deg"],
                   est["snr_threshold"], est["power_threshold"])
    m = np.empty((n_cpi, Nr, Na), dtype=np.float32) if want_map else None
    cm = np.empty((n_cpi, Nr, Na), dtype=np.complex64) if want_cmap else None
    d = np.zeros(n_cpi, dtype=DET_DTYPE) if want_dets else None
    lib().orc_chain_batch(C.byref(cfg), _p(rx), _p(tx), int(tx_shared), n_cpi, cpi0, _p(m), _p(cm), _p(d))
    return m, cm, d
