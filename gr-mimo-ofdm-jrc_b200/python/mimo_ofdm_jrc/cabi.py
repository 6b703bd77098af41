"""ctypes binding of libjrc_cuda.so (include/jrc_cuda.h).

This is the only place the Python side touches native code.  The library is the
product: if it is missing or cannot be loaded, importing a block raises -- there is
no Python/NumPy fallback for any computation.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.abspath(os.path.join(_HERE, "..", ".."))          # gr-mimo-ofdm-jrc_b200/
# JRC_CUDA_LIB selects an experiment build of the same library (kernel A/B runs); never a fallback
LIB_PATH = os.environ.get("JRC_CUDA_LIB") or os.path.join(PKG_ROOT, "libjrc_cuda.so")

JRC_OK, JRC_ERR_INVALID, JRC_ERR_CUDA, JRC_ERR_NO_DEVICE, JRC_ERR_STATE = 0, 1, 2, 3, 4
PATH_AUTO, PATH_FUSED, PATH_STAGED, PATH_TILED = 0, 1, 2, 3
DET_PASSED, DET_EXACT = 1, 2          # jrc_det.flags (include/jrc_cuda.h)


class JrcError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"jrc status {status}: {msg}")
        self.status = status


class ChainCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "fft_len", "n_tx", "n_rx", "n_sym", "n_pre", "interp_range", "interp_angle",
        "tx_interleave", "background_removal", "background_recording", "record_len", "device")]


class PortLayout(C.Structure):
    _fields_ = [("base", C.c_void_p), ("cpi_stride", C.c_int64), ("ant_stride", C.c_int64)]


class Peak1dOut(C.Structure):
    _fields_ = [("k", C.c_int32), ("freq", C.c_float), ("phase", C.c_float), ("mag", C.c_float)]


DET_DTYPE = np.dtype([("range_idx", "<i4"), ("angle_idx", "<i4"), ("peak_power", "<f4"),
                      ("noise_power", "<f4"), ("snr_db", "<f4"), ("n_noise", "<i4"),
                      ("flags", "<u4"), ("cpi", "<i4")])
assert DET_DTYPE.itemsize == 32

EXPORTS = [
    "jrc_chain_create", "jrc_chain_destroy", "jrc_last_error", "jrc_abi_version", "jrc_chain_stream",
    "jrc_chain_sync", "jrc_chain_set_estimator", "jrc_chain_set_thresholds",
    "jrc_chain_set_background_record", "jrc_chain_reset_background", "jrc_chain_run_batch",
    "jrc_chain_last_path", "jrc_chain_launch_count", "jrc_chain_run_host", "jrc_radar_estimate",
    "jrc_fft_vcc", "jrc_transpose_pad", "jrc_mag_squared", "jrc_target_sim", "jrc_nlog10", "jrc_estimate2d", "jrc_peak1d", "jrc_zero_pad",
    "jrc_cp_remove", "jrc_ofdm_demod", "jrc_chain_submit", "jrc_chain_poll", "jrc_chain_wait",
    "jrc_pinned_alloc", "jrc_pinned_free", "jrc_host_register", "jrc_host_unregister", "jrc_chain_exact_stats", "jrc_scene_synth", "jrc_chain_run_burst",
    "jrc_dev_alloc", "jrc_dev_free", "jrc_dev_copy", "jrc_ipc_export", "jrc_ipc_open", "jrc_ipc_close",
    "jrc_radar_estimate_fused", "jrc_fused_fetch_transposed", "jrc_fused_fetch_det", "jrc_chain_run_batch_time",
    "jrc_chain_copy_async",
]

_lib = None


def load():
    """Load libjrc_cuda.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built: run `make -C {PKG_ROOT}` (no CPU fallback exists)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, u32, u64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint32, C.c_uint64, C.c_size_t
    lib.jrc_last_error.restype = C.c_char_p
    lib.jrc_abi_version.restype = i32
    lib.jrc_chain_create.argtypes = [C.POINTER(ChainCfg), C.POINTER(vp)]
    lib.jrc_chain_destroy.argtypes = [vp]
    lib.jrc_chain_destroy.restype = None
    lib.jrc_chain_stream.argtypes = [vp]
    lib.jrc_chain_stream.restype = vp
    lib.jrc_chain_sync.argtypes = [vp]
    lib.jrc_chain_set_estimator.argtypes = [vp, vp, i32, vp, i32, f32, f32, f32, f32]
    lib.jrc_chain_set_thresholds.argtypes = [vp, f32, f32]
    lib.jrc_chain_set_background_record.argtypes = [vp, i32]
    lib.jrc_chain_reset_background.argtypes = [vp]
    lib.jrc_chain_run_batch.argtypes = [vp, PortLayout, PortLayout, i32, i32, vp, vp, vp, i32]
    lib.jrc_chain_last_path.argtypes = [vp]
    lib.jrc_chain_last_path.restype = i32
    lib.jrc_chain_launch_count.argtypes = [vp]
    lib.jrc_chain_launch_count.restype = i64
    lib.jrc_chain_run_host.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp]
    lib.jrc_radar_estimate.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), sz, vp, vp]
    lib.jrc_fft_vcc.argtypes = [vp, vp, vp, i32, i32, i32, i32]
    lib.jrc_transpose_pad.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.jrc_mag_squared.argtypes = [vp, vp, vp, sz]
    lib.jrc_target_sim.argtypes = [vp, vp, i32, vp, vp, vp, vp, i32, vp, i32, i32, C.c_float, i32, C.c_float, vp, i32, vp]
    lib.jrc_nlog10.argtypes = [vp, vp, vp, sz, C.c_float, C.c_float]
    lib.jrc_estimate2d.argtypes = [vp, vp, i32, i32, vp]
    lib.jrc_peak1d.argtypes = [vp, vp, i32, i32, f32, f32, i32, C.POINTER(Peak1dOut)]
    lib.jrc_zero_pad.argtypes = [vp, vp, i32, u32, u32, u64, vp]
    lib.jrc_cp_remove.argtypes = [vp, vp, i32, i32, i32, vp]
    lib.jrc_ofdm_demod.argtypes = [vp, vp, i32, i32, i32, vp]
    lib.jrc_chain_submit.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, C.POINTER(i64)]
    lib.jrc_chain_poll.argtypes = [vp, i64, C.POINTER(i32)]
    lib.jrc_chain_wait.argtypes = [vp, i64]
    lib.jrc_pinned_alloc.argtypes = [sz, C.POINTER(vp)]
    lib.jrc_pinned_free.argtypes = [vp]
    lib.jrc_host_register.argtypes = [vp, sz]
    lib.jrc_host_unregister.argtypes = [vp]
    lib.jrc_chain_exact_stats.argtypes = [vp, C.POINTER(i64)]
    lib.jrc_scene_synth.argtypes = [vp, vp, i32, i32, vp, vp, vp, C.c_double, C.c_double, f32, u64, vp]
    lib.jrc_chain_run_burst.argtypes = [vp, PortLayout, PortLayout, i32, vp]
    lib.jrc_dev_alloc.argtypes = [i32, sz, C.POINTER(vp)]
    lib.jrc_dev_free.argtypes = [vp]
    lib.jrc_dev_copy.argtypes = [vp, vp, sz]
    lib.jrc_ipc_export.argtypes = [vp, vp]
    lib.jrc_ipc_open.argtypes = [vp, i32, C.POINTER(vp)]
    lib.jrc_ipc_close.argtypes = [vp]
    lib.jrc_chain_run_batch_time.argtypes = [vp, PortLayout, i32, PortLayout, i32, i32, vp, vp, vp, i32]
    lib.jrc_chain_copy_async.argtypes = [vp, vp, vp, sz]
    lib.jrc_radar_estimate_fused.argtypes = [vp, vp, vp, sz, vp, vp, C.POINTER(i64)]
    lib.jrc_fused_fetch_transposed.argtypes = [vp, i64, vp]
    lib.jrc_fused_fetch_det.argtypes = [vp, i64, f32, f32, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("jrc_last_error", "jrc_abi_version", "jrc_chain_destroy", "jrc_chain_stream",
                        "jrc_chain_last_path", "jrc_chain_launch_count"):
            fn.restype = i32
    _lib = lib
    return lib


def check(status):
    if status != JRC_OK:
        raise JrcError(status, load().jrc_last_error().decode("utf-8", "replace"))


def np_ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


class Chain:
    """Owns one jrc_chain handle (one CUDA stream + scratch)."""

    def __init__(self, fft_len=64, n_tx=1, n_rx=1, n_sym=1, n_pre=0, interp_range=1, interp_angle=1,
                 tx_interleave=False, background_removal=False, background_recording=False,
                 record_len=0, device=0):
        lib = load()
        self.cfg = ChainCfg(fft_len, n_tx, n_rx, n_sym, n_pre, interp_range, interp_angle,
                            int(bool(tx_interleave)), int(bool(background_removal)),
                            int(bool(background_recording)), record_len, device)
        self._h = C.c_void_p()
        check(lib.jrc_chain_create(C.byref(self.cfg), C.byref(self._h)))
        self.V = n_tx * n_rx
        self.Nr = fft_len * interp_range
        self.Na = self.V * interp_angle

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            load().jrc_chain_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration ------------------------------------------------------
    def set_estimator(self, range_bins, angle_bins, noise_discard_range_m, noise_discard_angle_deg,
                      snr_threshold, power_threshold):
        rb = np.ascontiguousarray(range_bins, dtype=np.float32)
        ab = np.ascontiguousarray(angle_bins, dtype=np.float32)
        check(load().jrc_chain_set_estimator(self._h, np_ptr(rb), rb.size, np_ptr(ab), ab.size,
                                             noise_discard_range_m, noise_discard_angle_deg,
                                             snr_threshold, power_threshold))

    def set_thresholds(self, snr_threshold, power_threshold):
        check(load().jrc_chain_set_thresholds(self._h, snr_threshold, power_threshold))

    def set_background_record(self, on):
        check(load().jrc_chain_set_background_record(self._h, int(bool(on))))

    def reset_background(self):
        check(load().jrc_chain_reset_background(self._h))

    @property
    def stream(self):
        return load().jrc_chain_stream(self._h)

    def sync(self):
        check(load().jrc_chain_sync(self._h))

    @property
    def last_path(self):
        return load().jrc_chain_last_path(self._h)

    @property
    def launch_count(self):
        return load().jrc_chain_launch_count(self._h)

    # -- fused chain, device pointers ---------------------------------------
    def run_batch_ptr(self, rx_ptr, rx_cpi_stride, rx_ant_stride, tx_ptr, tx_cpi_stride, tx_ant_stride,
                      n_cpi, cpi0=0, map_ptr=None, cmap_ptr=None, dets_ptr=None, path=PATH_AUTO):
        rx = PortLayout(rx_ptr, rx_cpi_stride, rx_ant_stride)
        tx = PortLayout(tx_ptr, tx_cpi_stride, tx_ant_stride)
        check(load().jrc_chain_run_batch(self._h, rx, tx, n_cpi, cpi0, map_ptr, cmap_ptr, dets_ptr, path))

    def run_batch_time_ptr(self, rx_ptr, rx_cpi_stride, rx_ant_stride, cp_len, tx_ptr, tx_cpi_stride, tx_ant_stride,
                           n_cpi, cpi0=0, map_ptr=None, cmap_ptr=None, dets_ptr=None, path=PATH_AUTO):
        """The chain from the RX antennas' raw time samples (cyclic prefix + fft_len samples per symbol; device pointers)."""
        rx = PortLayout(rx_ptr, rx_cpi_stride, rx_ant_stride)
        tx = PortLayout(tx_ptr, tx_cpi_stride, tx_ant_stride)
        check(load().jrc_chain_run_batch_time(self._h, rx, cp_len, tx, n_cpi, cpi0, map_ptr, cmap_ptr, dets_ptr, path))

    def copy_async(self, dst_ptr, src_ptr, nbytes):
        """Device/peer copy queued on the handle's stream."""
        check(load().jrc_chain_copy_async(self._h, dst_ptr, src_ptr, nbytes))

    def run_burst_ptr(self, rx_ptr, rx_cpi_stride, rx_ant_stride, tx_ptr, tx_cpi_stride, tx_ant_stride, n_burst, cube_ptr):
        """Range-Doppler-angle cube [Nr][Na][n_burst] of a burst of CPIs (device pointers)."""
        rx = PortLayout(rx_ptr, rx_cpi_stride, rx_ant_stride)
        tx = PortLayout(tx_ptr, tx_cpi_stride, tx_ant_stride)
        check(load().jrc_chain_run_burst(self._h, rx, tx, n_burst, cube_ptr))

    # -- fused chain, host buffers ------------------------------------------
    def run_host_ptr(self, rx_ptr, tx_ptr, tx_shared, n_cpi, cpi0=0, map_ptr=None, dets_ptr=None):
        check(load().jrc_chain_run_host(self._h, rx_ptr, tx_ptr, int(bool(tx_shared)), n_cpi, cpi0,
                                        map_ptr, dets_ptr))

    def scene_synth_ptr(self, tx, ranges_m, az_deg, amps, rx_ptr, samp_rate=125e6, center_freq=24e9, noise_sigma=0.0, seed=0):
        """Batched point-target RX symbols on the device (mirror of synth.rx_symbols): tx [T][S][N] complex64 (NumPy), the
        per-CPI target parameters [n_cpi][J] (NumPy), rx_ptr = device pointer of [n_cpi][R][S][N] complex64."""
        tx = np.ascontiguousarray(tx, dtype=np.complex64)
        r = np.ascontiguousarray(np.atleast_2d(ranges_m), dtype=np.float32)
        a = np.ascontiguousarray(np.atleast_2d(az_deg), dtype=np.float32)
        m = np.ascontiguousarray(np.atleast_2d(amps), dtype=np.float32)
        assert r.shape == a.shape == m.shape
        check(load().jrc_scene_synth(self._h, np_ptr(tx), r.shape[0], r.shape[1], np_ptr(r), np_ptr(a), np_ptr(m),
                                     float(samp_rate), float(center_freq), float(noise_sigma), int(seed), rx_ptr))

    def exact_stats(self):
        """(marked, redone by k_est_exact, ties settled in the fused kernel) since the handle was created."""
        out = (C.c_int64 * 3)()
        check(load().jrc_chain_exact_stats(self._h, out))
        return {"marked": out[0], "redone": out[1], "ties_in_kernel": out[2]}

    # -- streaming form: up to 4 submissions in flight ------------------------
    def submit_ptr(self, rx_ptr, tx_ptr, tx_shared, n_cpi, cpi0=0, map_ptr=None, dets_ptr=None):
        t = C.c_int64()
        check(load().jrc_chain_submit(self._h, rx_ptr, tx_ptr, int(bool(tx_shared)), n_cpi, cpi0, map_ptr, dets_ptr,
                                      C.byref(t)))
        return t.value

    def poll(self, ticket):
        d = C.c_int32()
        check(load().jrc_chain_poll(self._h, ticket, C.byref(d)))
        return bool(d.value)

    def wait(self, ticket):
        check(load().jrc_chain_wait(self._h, ticket))

    def run_host(self, rx: np.ndarray, tx: np.ndarray, want_map=True, want_dets=True, cpi0=0):
        """rx [n_cpi][R][S][N] complex64, tx [n_cpi or 1][T][S][N] complex64 (NumPy, host)."""
        rx = np.ascontiguousarray(rx, dtype=np.complex64)
        tx = np.ascontiguousarray(tx, dtype=np.complex64)
        n_cpi = rx.shape[0]
        if tx.ndim == 3:
            tx = tx[None]
        tx_shared = tx.shape[0] == 1 and n_cpi > 1
        m = np.empty((n_cpi, self.Nr, self.Na), dtype=np.float32) if want_map else None
        d = np.zeros(n_cpi, dtype=DET_DTYPE) if want_dets else None
        self.run_host_ptr(np_ptr(rx), np_ptr(tx), tx_shared, n_cpi, cpi0,
                          np_ptr(m) if want_map else None, np_ptr(d) if want_dets else None)
        return m, d

    # -- per-block stage calls (NumPy host arrays) ---------------------------
    def radar_estimate(self, tx_ports, rx_ports, tx_skip_items=0, want_chan_est=False):
        T, R, N = self.cfg.n_tx, self.cfg.n_rx, self.cfg.fft_len
        txs = [np.ascontiguousarray(a, dtype=np.complex64) for a in tx_ports]
        rxs = [np.ascontiguousarray(a, dtype=np.complex64) for a in rx_ports]
        assert len(txs) == T and len(rxs) == R
        need = (self.cfg.n_pre + self.cfg.n_sym) * N
        for a in rxs:
            if a.size < need:
                raise ValueError("rx packet shorter than n_pre + n_sym symbols")
        for a in txs:
            if a.size < need + tx_skip_items * N:
                raise ValueError("tx packet shorter than skip + n_pre + n_sym symbols")
        tp = (C.c_void_p * T)(*[a.ctypes.data for a in txs])
        rp = (C.c_void_p * R)(*[a.ctypes.data for a in rxs])
        out = np.empty((self.V, self.Nr), dtype=np.complex64)
        ce = np.empty((self.V, N), dtype=np.complex64) if want_chan_est else None
        check(load().jrc_radar_estimate(self._h, tp, rp, tx_skip_items, np_ptr(out),
                                        np_ptr(ce) if want_chan_est else None))
        return (out, ce) if want_chan_est else out

    def fft_vcc(self, x, forward, shift):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        n = x.shape[-1]
        batch = x.size // n if n else 0
        out = np.empty_like(x)
        check(load().jrc_fft_vcc(self._h, np_ptr(x), np_ptr(out), n, batch, int(bool(forward)), int(bool(shift))))
        return out

    def transpose_pad(self, x, output_len, interp):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        k_items, input_len = x.shape
        out = np.empty((input_len, output_len * interp), dtype=np.complex64)
        check(load().jrc_transpose_pad(self._h, np_ptr(x), k_items, input_len, output_len, interp, np_ptr(out)))
        return out

    def mag_squared(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out = np.empty(x.shape, dtype=np.float32)
        check(load().jrc_mag_squared(self._h, np_ptr(x), np_ptr(out), x.size))
        return out

    def target_sim(self, x, range_m, velocity, rcs, azimuth, position_rx, samp_rate, center_freq,
                   self_coupling=False, self_coupling_db=0.0, target_phase=None, accumulate=False):
        """target_simulator::work for one packet x of time samples -> [n_rx][n]."""
        x = np.ascontiguousarray(x, dtype=np.complex64)
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)   # noqa: E731
        r, v, s, a, p = f32(range_m), f32(velocity), f32(rcs), f32(azimuth), f32(position_rx)
        assert r.size == v.size == s.size == a.size
        ph = None if target_phase is None else np.ascontiguousarray(target_phase, dtype=np.complex64)
        out = np.empty((p.size, x.size), dtype=np.complex64)
        check(load().jrc_target_sim(self._h, np_ptr(x), x.size, np_ptr(r), np_ptr(v), np_ptr(s), np_ptr(a), r.size,
                                    np_ptr(p), p.size, int(samp_rate), float(center_freq), int(bool(self_coupling)),
                                    float(self_coupling_db), None if ph is None else np_ptr(ph), int(bool(accumulate)),
                                    np_ptr(out)))
        return out

    def nlog10(self, x, n=10.0, k=0.0):
        """blocks_nlog10_ff in front of gui_heatmap_plot: n*log10(max(x, 1e-18)) + k."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty(x.shape, dtype=np.float32)
        check(load().jrc_nlog10(self._h, np_ptr(x), np_ptr(out), x.size, float(n), float(k)))
        return out

    def estimate2d(self, cmap):
        cmap = np.ascontiguousarray(cmap, dtype=np.complex64)
        n_inputs, vlen = cmap.shape
        det = np.zeros(1, dtype=DET_DTYPE)
        check(load().jrc_estimate2d(self._h, np_ptr(cmap), n_inputs, vlen, np_ptr(det)))
        return det[0]

    def peak1d(self, x, samp_rate, interp_factor, threshold_db, samp_protect):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out = Peak1dOut()
        check(load().jrc_peak1d(self._h, np_ptr(x), x.size, samp_rate, interp_factor, threshold_db,
                                samp_protect, C.byref(out)))
        return out.k, out.freq, out.phase, out.mag

    def cp_remove(self, x, n_sym, fft_len, cp_len):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        assert x.size >= n_sym * (fft_len + cp_len)
        out = np.empty((n_sym, fft_len), dtype=np.complex64)
        check(load().jrc_cp_remove(self._h, np_ptr(x), n_sym, fft_len, cp_len, np_ptr(out)))
        return out

    def ofdm_demod(self, x, n_sym, fft_len, cp_len):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        assert x.size >= n_sym * (fft_len + cp_len)
        out = np.empty((n_sym, fft_len), dtype=np.complex64)
        check(load().jrc_ofdm_demod(self._h, np_ptr(x), n_sym, fft_len, cp_len, np_ptr(out)))
        return out

    def zero_pad(self, x, pad_front, pad_tail, seed):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        out = np.empty(x.size + pad_front + pad_tail, dtype=np.complex64)
        check(load().jrc_zero_pad(self._h, np_ptr(x), x.size, pad_front, pad_tail, seed, np_ptr(out)))
        return out
