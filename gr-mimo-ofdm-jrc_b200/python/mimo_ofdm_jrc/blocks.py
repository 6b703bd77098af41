"""Host-side mirrors of the reference blocks on the radar path.

Same constructor arguments (the reference's make() signatures), same setter names, same
per-packet semantics and error behaviour; `work()` takes/returns NumPy arrays where the
GNU Radio block takes stream buffers.  All arithmetic happens in libjrc_cuda.so.
The C++ GNU Radio wrappers in ../lib/ are the production host side; these mirrors drive
the same C ABI from Python for tests, bench.py and notebooks.
"""
from __future__ import annotations

import datetime
import os

import numpy as np

from . import cabi


class mimo_ofdm_radar:
    """include/mimo_ofdm_jrc/mimo_ofdm_radar.h:48-60, lib/mimo_ofdm_radar_impl.cc"""

    def __init__(self, fft_len, N_tx, N_rx, N_sym, N_pre, background_removal, background_recording,
                 record_len, interp_factor, enable_tx_interleave, radar_chan_file,
                 len_tag_key="packet_len", debug=False, device=0):
        self.fft_len, self.N_tx, self.N_rx, self.N_sym, self.N_pre = fft_len, N_tx, N_rx, N_sym, N_pre
        self.interp_factor = interp_factor
        self.radar_chan_file = radar_chan_file
        self.len_tag_key = len_tag_key      # stored; the reference hard-codes "packet_len" (:167,176)
        self.debug = debug
        self._chain = cabi.Chain(fft_len, N_tx, N_rx, N_sym, N_pre, interp_factor, 1, enable_tx_interleave,
                                 background_removal, background_recording, record_len, device)
        self._chan_est = np.zeros((N_tx * N_rx, fft_len), dtype=np.complex64)
        self.nitems_written = 0
        self.alias = "mimo_ofdm_radar0"

    def set_background_record(self, background_recording):
        print(f"[MIMO OFDM RADAR] Background recording set to  {int(bool(background_recording))}")
        self._chain.set_background_record(background_recording)

    def capture_radar_data(self, capture_sig):
        """lib/mimo_ofdm_radar_impl.cc:348-387: append `time, N_tx, N_rx, fft_len:` + ';'-separated H."""
        if not capture_sig:
            return
        try:
            f = open(self.radar_chan_file, "a")
        except OSError:
            raise RuntimeError("[MIMO OFDM RADAR] Could not open file!!")
        with f:
            now = datetime.datetime.now()
            f.write(f"{now.strftime('%H:%M:%S')}.{now.microsecond // 1000:03d}, {self.N_tx}, {self.N_rx}, {self.fft_len}:")
            f.write(";".join(f"({z.real:.7g},{z.imag:.7g})" for z in self._chan_est.ravel()))
            f.write(";\n\n")
        print("[MIMO OFDM RADAR] Radar image captured!")

    def work(self, tx_ports, rx_ports, tx_tag_lens=None, rx_tag_lens=None):
        """One general_work() call with a whole frame in the buffers.

        tx_ports/rx_ports: per-port arrays of fft_len-vectors.  *_tag_lens: the values of the
        `packet_len` tags found on TX port 0 / RX port N_tx (lists); None = one tag covering
        the port.  Returns (out [V][fft_len*interp] or None, tags, consumed) like the block:
        no RX tag -> everything consumed, nothing produced (:216-231); more TX than RX tags
        -> the stale TX frames are skipped (:189-197).
        """
        N = self.fft_len
        rx_items = [np.asarray(a).size // N for a in rx_ports]
        tx_items = [np.asarray(a).size // N for a in tx_ports]
        if rx_tag_lens is None:
            rx_tag_lens = [rx_items[0]]
        if tx_tag_lens is None:
            tx_tag_lens = [tx_items[0]]
        if len(rx_tag_lens) == 0:
            return None, [], dict(tx=tx_items, rx=rx_items)
        if len(tx_tag_lens) == 0:
            raise ValueError("no packet_len tag on TX port 0 (undefined in the reference, :200)")
        off = max(0, len(tx_tag_lens) - len(rx_tag_lens))
        skip = int(sum(tx_tag_lens[:off]))
        rx_len, tx_len = int(rx_tag_lens[0]), int(tx_tag_lens[off])
        out, ce = self._chain.radar_estimate(tx_ports, rx_ports, skip, want_chan_est=True)
        self._chan_est = ce
        V = self.N_tx * self.N_rx
        tags = [dict(offset=self.nitems_written, key="packet_len", value=V, srcid=self.alias)]
        self.nitems_written += V
        return out, tags, dict(tx=[skip + tx_len] * self.N_tx, rx=[rx_len] * self.N_rx)


class matrix_transpose:
    """include/mimo_ofdm_jrc/matrix_transpose.h:48, lib/matrix_transpose_impl.cc:69-110"""

    def __init__(self, input_len, output_len, interp_factor, debug=False, len_key="packet_len", device=0):
        self.input_len, self.output_len, self.interp_factor = input_len, output_len, interp_factor
        self.len_key = len_key
        self._chain = cabi.Chain(device=device)

    def calculate_output_stream_length(self, ninput_items):
        return self.input_len

    def work(self, items, output_buffer_fullness=0.0):
        items = np.asarray(items, dtype=np.complex64).reshape(-1, self.input_len)
        k = items.shape[0]
        if k * float(self.input_len) / float(self.output_len) - k * self.input_len // self.output_len != 0:
            raise RuntimeError("[MATRIX TRANSPOSE] input_len and output_len do not match to packet length")
        if output_buffer_fullness > 0.001:      # :86-89 -- back-pressure drops the CPI
            return None
        return self._chain.transpose_pad(items, self.output_len, self.interp_factor)


class fft_vcc:
    """Stock gr::fft::fft_vcc as wired in the flowgraph (...radar_sim.grc:940-985), window = ones."""

    def __init__(self, fft_size, forward, window=None, shift=False, nthreads=1, device=0):
        self.fft_size, self.forward, self.shift = fft_size, bool(forward), bool(shift)
        if window is not None and len(window) and not np.allclose(window, 1.0):
            raise ValueError("only the rectangular window of the radar flowgraph is supported")
        self._chain = cabi.Chain(device=device)

    def work(self, items):
        items = np.asarray(items, dtype=np.complex64).reshape(-1, self.fft_size)
        return self._chain.fft_vcc(items, self.forward, self.shift)


class complex_to_mag_squared:
    def __init__(self, vlen=1, device=0):
        self.vlen = vlen
        self._chain = cabi.Chain(device=device)

    def work(self, items):
        return self._chain.mag_squared(items)


class target_simulator:
    """target_simulator (include/mimo_ofdm_jrc/target_simulator.h:48-71, lib/target_simulator_impl.cc): one input
    packet of time samples -> one packet per RX antenna.  `accumulate=True` sums the targets; the default keeps
    the reference's behaviour (its memcpy at :366 leaves only the last target in the output)."""

    def __init__(self, range, velocity, rcs, azimuth, position_rx, samp_rate, center_freq, self_coupling_db,
                 rndm_phaseshift=False, self_coupling=False, len_key="packet_len", debug=False, device=0,
                 accumulate=False):
        self.len_key, self.debug, self.accumulate = len_key, debug, accumulate
        self.chain = cabi.Chain(device=device)
        self.nitems_written = 0
        self.setup_targets(range, velocity, rcs, azimuth, position_rx, samp_rate, center_freq, self_coupling_db,
                           rndm_phaseshift, self_coupling)

    def setup_targets(self, range, velocity, rcs, azimuth, position_rx, samp_rate, center_freq, self_coupling_db,
                      rndm_phaseshift, self_coupling):
        self.range, self.velocity, self.rcs, self.azimuth = (np.asarray(v, dtype=np.float32) for v in
                                                             (range, velocity, rcs, azimuth))
        self.position_rx = np.asarray(position_rx, dtype=np.float32)
        self.samp_rate, self.center_freq = int(samp_rate), float(center_freq)
        self.self_coupling_db, self.rndm_phaseshift, self.self_coupling = float(self_coupling_db), bool(rndm_phaseshift), bool(self_coupling)

    def work(self, x):
        """-> (out [n_rx][n], tags): tags[l] = ("rx_time", (sec, frac), "stat_targ_sim") at the packet start (:333-336)."""
        phase = None
        if self.rndm_phaseshift:     # :311-320, one random phase per target and packet
            u = (np.random.randint(0, 1000, size=self.range.size) + 1) / 1000.0
            phase = np.exp(1j * 2 * np.pi * u.astype(np.float32)).astype(np.complex64)
        out = self.chain.target_sim(x, self.range, self.velocity, self.rcs, self.azimuth, self.position_rx, self.samp_rate,
                                    self.center_freq, self.self_coupling, self.self_coupling_db, phase, self.accumulate)
        sec = self.nitems_written // self.samp_rate
        frac = float(np.float32(self.nitems_written) / np.float32(self.samp_rate)) - sec
        tags = [("rx_time", (int(sec), frac), "stat_targ_sim")] * self.position_rx.size
        self.nitems_written += np.asarray(x).size
        return out, tags


class nlog10_ff:
    """blocks_nlog10_ff between complex_to_mag_squared and gui_heatmap_plot (...radar_sim.grc:725-745; bypassed
    in the simulation flowgraph, active in the USRP one): n*log10(max(x, 1e-18)) + k."""

    def __init__(self, n=10.0, vlen=1, k=0.0, device=0):
        self.n, self.k, self.vlen = float(n), float(k), int(vlen)
        self.chain = cabi.Chain(device=device)

    def work(self, x):
        return self.chain.nlog10(x, self.n, self.k)


def radar_log_read_last(path):
    """Consumer side of range_angle_estimator's CSV log: the last record as (time, power, snr, range, angle),
    or None when the file is missing / empty / ends with the "NEW RECORD" header
    (lib/mimo_precoder_impl.cc:903-950 reads the last line and takes the 5th comma-separated field)."""
    try:
        lines = [l for l in open(path).read().splitlines() if l.strip()]
    except OSError:
        return None
    if not lines:
        return None
    f = lines[-1].split(",")
    if len(f) != 5:
        return None
    try:
        return (f[0].strip(),) + tuple(float(v) for v in f[1:])
    except ValueError:
        return None


def radar_aided_steering_vector(angle_deg, n_tx):
    """a[i] = exp(j*pi*sin(angle)*i) for a half-wavelength TX array (lib/mimo_precoder_impl.cc:952-956)."""
    i = np.arange(n_tx, dtype=np.float64)
    return np.exp(1j * np.pi * np.sin(np.float32(angle_deg) / 180.0 * np.pi) * i).astype(np.complex64)


class range_angle_estimator:
    """include/mimo_ofdm_jrc/range_angle_estimator.h:48-58, lib/range_angle_estimator_impl.cc"""

    def __init__(self, vlen, range_bins, angle_bins, noise_discard_range_m, noise_discard_angle_deg,
                 snr_threshold, power_threshold, stats_path, stats_record, len_key="packet_len",
                 debug=False, device=0):
        self.vlen = vlen
        self.range_bins = np.asarray(range_bins, dtype=np.float32)
        self.angle_bins = np.asarray(angle_bins, dtype=np.float32)
        self.nd_range, self.nd_angle = np.float32(noise_discard_range_m), np.float32(noise_discard_angle_deg)
        self.snr_threshold, self.power_threshold = np.float32(snr_threshold), np.float32(power_threshold)
        self.stats_path, self.stats_record = stats_path, bool(stats_record)
        self._new_stat_started = False
        self.messages = []                      # what message_port_pub("params", ...) carried
        self._chain = cabi.Chain(device=device)
        self._push()
        if stats_path:
            try:
                open(stats_path, "a").close()
            except OSError:
                print(f"[RANGE-ANGLE ESTIMATOR] Could not open log file at {stats_path}")

    def _push(self):
        self._chain.set_estimator(self.range_bins, self.angle_bins, self.nd_range, self.nd_angle,
                                  self.snr_threshold, self.power_threshold)

    def set_snr_threshold(self, v):
        self.snr_threshold = np.float32(v)
        self._chain.set_thresholds(self.snr_threshold, self.power_threshold)

    def set_power_threshold(self, v):
        self.power_threshold = np.float32(v)
        self._chain.set_thresholds(self.snr_threshold, self.power_threshold)

    def set_stats_record(self, on):
        self.stats_record = bool(on)
        self._new_stat_started = False

    def work(self, cmap):
        """One packet = the whole complex map [n_inputs][vlen]; returns the detection record and
        publishes/logs exactly when the reference does (:234-279)."""
        cmap = np.asarray(cmap, dtype=np.complex64).reshape(-1, self.vlen)
        det = self._chain.estimate2d(cmap)
        if det["flags"] & 1:
            range_val = self.range_bins[det["range_idx"]]
            angle_val = self.angle_bins[det["angle_idx"]]
            msg = [("range", np.array([range_val], np.float32)), ("angle", np.array([angle_val], np.float32)),
                   ("power", np.array([det["peak_power"]], np.float32)), ("snr", np.array([det["snr_db"]], np.float32))]
            self.messages.append(msg)
            if self.stats_record:
                try:
                    f = open(self.stats_path, "a")
                except OSError:
                    raise RuntimeError("[STREAM DECODER] Could not open file!!")
                with f:
                    now = datetime.datetime.now()
                    if not self._new_stat_started:
                        f.write(f"\n NEW RECORD - {now.strftime('%m-%d-%Y %H:%M:%S')}\n")
                        self._new_stat_started = True
                    ts = f"{now.strftime('%H:%M:%S')}.{now.microsecond // 1000:03d}"
                    f.write(f"{ts}, \t{det['peak_power']:g}, \t{det['snr_db']:g}, \t{range_val:g}, \t{angle_val:g}\n")
        return det


class fft_peak_detect:
    """include/mimo_ofdm_jrc/fft_peak_detect.h:48, lib/fft_peak_detect_impl.cc:77-111"""

    def __init__(self, samp_rate, interp_factor, threshold, samp_protect, max_freq, cut_max_freq,
                 len_key="packet_len", device=0):
        self.samp_rate, self.interp_factor = int(samp_rate), np.float32(interp_factor)
        self.threshold, self.samp_protect = np.float32(threshold), int(samp_protect)
        self.max_freq, self.cut_max_freq = list(max_freq), bool(cut_max_freq)   # stored, unused (as in the reference)
        self._chain = cabi.Chain(device=device)

    def set_threshold(self, threshold):
        self.threshold = np.float32(threshold)

    def set_samp_protect(self, samp):
        self.samp_protect = int(samp)

    def set_max_freq(self, freq):
        self.max_freq = list(freq)

    def work(self, items):
        """Returns (k, freq, phase, mag); k == -1 when no sample passes the threshold (the
        reference then emits its one output item unwritten)."""
        return self._chain.peak1d(items, self.samp_rate, self.interp_factor, self.threshold, self.samp_protect)


class zero_pad:
    """include/mimo_ofdm_jrc/zero_pad.h:48, lib/zero_pad_impl.cc:61-94"""

    def __init__(self, debug=False, pad_front=0, pad_tail=0, device=0):
        self.pad_front, self.pad_tail = int(pad_front), int(pad_tail)
        self._chain = cabi.Chain(device=device)
        self._calls = 0

    def calculate_output_stream_length(self, ninput_items):
        return ninput_items + self.pad_front + self.pad_tail

    def work(self, items, seed=None):
        if seed is None:                       # the reference reseeds from std::random_device per call
            seed = int.from_bytes(os.urandom(8), "little")
        self._calls += 1
        return self._chain.zero_pad(items, self.pad_front, self.pad_tail, seed)


class ofdm_cyclic_prefix_remover:
    """include/mimo_ofdm_jrc/ofdm_cyclic_prefix_remover.h, lib/ofdm_cyclic_prefix_remover_impl.cc:62-99 -- the block
    in front of the radar path.  `demod=True` additionally applies the flowgraph's RX
    fft_vxx(fft_len, forward, shift) in the same kernel (jrc_ofdm_demod)."""

    def __init__(self, fft_len, cp_len, len_key="packet_len", device=0):
        self.fft_len, self.cp_len, self.len_key = int(fft_len), int(cp_len), len_key
        self._chain = cabi.Chain(device=device)

    def calculate_output_stream_length(self, ninput_items):
        return ninput_items // (self.fft_len + self.cp_len)

    def work(self, items, demod=False):
        items = np.asarray(items, dtype=np.complex64).reshape(-1)
        n_sym = items.size // (self.fft_len + self.cp_len)
        f = self._chain.ofdm_demod if demod else self._chain.cp_remove
        return f(items, n_sym, self.fft_len, self.cp_len)


class radar_chain:
    """The fused chain over a batch of CPIs resident in device memory (torch tensors are only
    used as device buffers).  rx [n_cpi][R][S][N] / tx [n_cpi or 1][T][S][N] complex64 CUDA."""

    def __init__(self, fft_len, N_tx, N_rx, N_sym, interp_range, interp_angle, enable_tx_interleave=False,
                 background_removal=False, background_recording=False, record_len=0, device=0,
                 estimator=None):
        self.chain = cabi.Chain(fft_len, N_tx, N_rx, N_sym, 0, interp_range, interp_angle, enable_tx_interleave,
                                background_removal, background_recording, record_len, device)
        self.device = device
        self.Nr, self.Na = self.chain.Nr, self.chain.Na
        self.per_ant = N_sym * fft_len
        self.N_tx, self.N_rx = N_tx, N_rx
        if estimator is not None:
            self.chain.set_estimator(**estimator)

    def run(self, rx, tx, want_map=True, want_dets=True, cpi0=0, path=cabi.PATH_AUTO, map_out=None, dets_out=None,
            sync_inputs=True, dets_ptr=None):
        """dets_ptr: raw device address for the records instead of a tensor (e.g. a slice of a shard.PeerRecordTable in
        another GPU's memory)."""
        import torch
        assert rx.is_cuda and tx.is_cuda and rx.dtype == torch.complex64 and tx.dtype == torch.complex64
        assert rx.is_contiguous() and tx.is_contiguous()
        n_cpi = rx.shape[0]
        tx_shared = tx.dim() == 3 or (tx.shape[0] == 1 and n_cpi > 1)
        if want_map and map_out is None:
            map_out = torch.empty((n_cpi, self.Nr, self.Na), dtype=torch.float32, device=rx.device)
        if want_dets and dets_out is None and dets_ptr is None:
            dets_out = torch.zeros((n_cpi, 32), dtype=torch.uint8, device=rx.device)
        if sync_inputs:
            torch.cuda.current_stream(rx.device).synchronize()     # inputs were produced on torch's stream
        self.chain.run_batch_ptr(rx.data_ptr(), self.N_rx * self.per_ant, self.per_ant,
                                 tx.data_ptr(), 0 if tx_shared else self.N_tx * self.per_ant, self.per_ant,
                                 n_cpi, cpi0,
                                 map_out.data_ptr() if want_map else None, None,
                                 (dets_ptr if dets_ptr is not None else dets_out.data_ptr()) if want_dets else None, path)
        return map_out if want_map else None, dets_out if want_dets else None

    def sync(self):
        self.chain.sync()

    @staticmethod
    def dets_to_numpy(dets):
        return dets.cpu().numpy().view(cabi.DET_DTYPE).reshape(-1)
