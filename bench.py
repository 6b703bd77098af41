#!/usr/bin/env python
"""bench.py -- range-angle CPIs/s of the radar hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the fused chain (mimo_ofdm_radar -> range IFFT -> transpose -> angle FFT
-> |.|^2 -> range_angle_estimator) over one batch of BASELINE configs[1]: 64 subcarriers,
4 TX x 2 RX = 8 virtual channels, 4 LTF symbols, range zero-pad 1024, angle zero-pad 64,
4096 CPIs per GPU.  CPIs are independent, so N GPUs each process their own 4096-CPI shard
(weak scaling); the 32-byte detection records of all K steps are gathered to rank 0 over NCCL once,
after the last step.

Prints ONE JSON line (rank 0).  Besides the contract's keys it carries
  configs   the other BASELINE configurations (configs[0] shipped flowgraph, configs[2], configs[4]) event-timed in
            the same process: CPI/s, kernels, fraction of the HBM roofline
  latency   configs[3]: one CPI per general_work() of the radar_chain block through the runtime stand-in
            (tests/cpp/latency_blocks.cc): p50/p99 per call, sustained CPI/s with the submit/wait pipeline
  sweep_c5  (--gpus N) the configs[4] sweep of 65 536 CPIs sharded over the N ranks
`--impl reference` times the reference's own CPU chain (oracle/_ref) on the host cores for the same metric,
configuration and batch.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

CFG = dict(T=4, R=2, S=4, N=64, IR=16, IA=8)          # BASELINE configs[1]
WORKLOAD = ("configs[1]: 64 subcarriers, 4TX x 2RX = 8 virtual channels, 4 LTF symbols, range zero-pad 1024, "
            "angle zero-pad 64, batch of 4096 CPIs per GPU per step, per-CPI TX symbols")
METRIC, UNIT = "range-angle CPIs/s", "CPI/s"
OTHER_CONFIGS = {
    "configs[0] shipped flowgraph 4x2, 64 sc, 512x128": dict(T=4, R=2, S=4, N=64, IR=8, IA=16, n=4096, targets=1,
                                                             kernels=["k_fused64x8<8,16>", "k_est_exact"]),
    "configs[2] 4x8, 256 sc, 4096x256, 5 targets": dict(T=4, R=8, S=4, N=256, IR=16, IA=8, n=888, targets=5,
                                                        kernels=["k_chan_est_tile<4,4>", "k_slice256", "k_map_finalize", "k_est_exact"]),
    "configs[4] 8x16, 2048 sc, 2048x128": dict(T=8, R=16, S=8, N=2048, IR=1, IA=1, n=222, targets=3,
                                               kernels=["k_wide_mac_angle<11,8>", "k_wide_range_mag<11>", "k_map_finalize", "k_est_exact"]),
}


def b_alg_per_cpi(c):
    """read (T+R)*S*Nsc complex once, write Nr*Na float32 once, 32 B record."""
    return (c["T"] + c["R"]) * c["S"] * c["N"] * 8 + (c["N"] * c["IR"]) * (c["T"] * c["R"] * c["IA"]) * 4 + 32


def config_dict(batch, world):
    """The same dict on both arms (the driver compares them)."""
    Nr, Na = CFG["N"] * CFG["IR"], CFG["T"] * CFG["R"] * CFG["IA"]
    return {"workload": WORKLOAD, "batch_per_gpu": batch, "map": [Nr, Na],
            "l2": "per-step working set (1.0 GiB map + 48 MiB symbols) exceeds the 126 MB L2; no flush needed",
            "parallelism": f"cpi-shard x{world}, detection records land on rank 0 over NVLink (one peer copy per rank at the drain into a CUDA-IPC mapped table; NCCL for the IPC handle and barriers)" if world > 1 else "single GPU"}


def make_inputs(batch, seed, cfg=CFG, targets=2, span=10.0):
    from mimo_ofdm_jrc import synth
    rng = np.random.default_rng(seed)
    tx = synth.tx_symbols(cfg["T"], cfg["S"], cfg["N"])
    r, a, amp = synth.random_scene(rng, batch, targets, cfg["N"], amp_db_span=span if targets > 1 else 0.0)
    rx = synth.rx_symbols(tx, cfg["R"], r, a, amp, snr_db=20.0, rng=rng, chunk=512 if cfg["N"] <= 64 else 16)
    txb = np.ascontiguousarray(np.broadcast_to(tx, (batch,) + tx.shape)) if cfg is CFG else tx
    est = synth.default_estimator_params(cfg["N"], cfg["T"] * cfg["R"], cfg["IR"], cfg["IA"])
    return rx, txb, est


def source_hash():
    """Hash of the sources of the dominant kernel (k_fused64x8 and what it includes): profiles/ncu_traffic.json is only
    quoted for the code it was measured on."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "csrc")
    for f in ("jrc_common.cuh", "jrc_exact.cuh", "jrc_fused.cuh"):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def bind_rank_to_cores(local, world):
    """One disjoint block of host cores per rank, BEFORE any pinned allocation (first touch decides the NUMA node of
    the staging buffers)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(1, world))
        mine = cores[local * per:(local + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:  # noqa: BLE001
        return None


# --------------------------------------------------------------------------------------
# clocks: NVML polled from a thread DURING the timed region
# --------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.power = [], set(), []
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
        self._stop = threading.Event()
        self._t = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.001)

    def start(self):
        if self.ok:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": max(self.power) if self.power else None}


# --------------------------------------------------------------------------------------
# CPU arm: the reference's own block sources (oracle/_ref) on the host cores
# --------------------------------------------------------------------------------------
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libjrc_ref.so")


def cpu_kind():
    """"reference": oracle/_ref = the reference's own block sources (built in the container by
    oracle/build_ref.sh; the two stock GNU Radio FFT blocks, absent from the reference tree, use the
    oracle's float32 FFT).  "port": the oracle restatement alone."""
    return "reference" if os.path.exists(REF_LIB) else "port"


def cpu_chain_rate(n_threads, per_thread, repeats=1, warm=0):
    """Runs the CPU chain on n_threads host threads (ctypes releases the GIL), each over its own
    per_thread CPIs.  Returns (CPIs/s, seconds per repeat list)."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    from oracle import orc
    orc.lib()
    rx, tx, est = make_inputs(n_threads * per_thread, seed=1234)
    ref = None
    if cpu_kind() == "reference":
        ref = C.CDLL(REF_LIB)
        ref.ref_chain_batch.argtypes = [C.POINTER(orc.ChainCfg), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
    Nr, Na = CFG["N"] * CFG["IR"], CFG["T"] * CFG["R"] * CFG["IA"]
    rb, ab = orc.f32(est["range_bins"]), orc.f32(est["angle_bins"])
    maps = [np.empty((per_thread, Nr, Na), np.float32) for _ in range(n_threads)]

    def work(i):
        sl = slice(i * per_thread, (i + 1) * per_thread)
        if ref is None:
            orc.chain_batch(rx[sl], tx[sl], CFG["N"], CFG["T"], CFG["R"], CFG["S"], CFG["IR"], CFG["IA"], est)
            return
        c = orc.ChainCfg(CFG["N"], CFG["T"], CFG["R"], CFG["S"], 0, CFG["IR"], CFG["IA"], 0, rb.ctypes.data, ab.ctypes.data,
                         est["noise_discard_range_m"], est["noise_discard_angle_deg"], est["snr_threshold"],
                         est["power_threshold"])
        a, b = np.ascontiguousarray(rx[sl]), np.ascontiguousarray(tx[sl])
        d = np.zeros(per_thread, orc.DET_DTYPE)
        ref.ref_chain_batch(C.byref(c), a.ctypes.data, b.ctypes.data, 0, per_thread, 0, maps[i].ctypes.data, None, d.ctypes.data)

    times = []
    with ThreadPoolExecutor(n_threads) as ex:
        for it in range(warm + repeats):
            t0 = time.perf_counter()
            list(ex.map(work, range(n_threads)))
            if it >= warm:
                times.append(time.perf_counter() - t0)
    return n_threads * per_thread * len(times) / sum(times), times


def fft_sanity(n_cpi=64):
    """BASELINE.md section 2's sanity point: the two FFT stages of the chain (V range IFFTs of Nr, Nr angle FFTs of Na per
    CPI) on ONE core with pocketfft complex64 (scipy.fft) and with the oracle's radix-2 FFT the CPU arm uses for the two
    stock GNU Radio FFT blocks -- how pessimistic that stand-in is against an optimised library."""
    try:
        import scipy.fft as sfft
        from oracle import orc
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)}
    V, Nr, Na = CFG["T"] * CFG["R"], CFG["N"] * CFG["IR"], CFG["T"] * CFG["R"] * CFG["IA"]
    rng = np.random.default_rng(0)
    a = (rng.standard_normal((n_cpi * V, Nr)) + 1j * rng.standard_normal((n_cpi * V, Nr))).astype(np.complex64)
    b = (rng.standard_normal((n_cpi * Nr, Na)) + 1j * rng.standard_normal((n_cpi * Nr, Na))).astype(np.complex64)
    res = {}
    for name, f in (("pocketfft_c64", lambda x, fwd: (sfft.fft if fwd else sfft.ifft)(x, axis=-1, workers=1)),
                    ("oracle_radix2", lambda x, fwd: orc.fft_vcc(x, fwd, False))):
        f(a[:V], False)
        t0 = time.perf_counter()
        f(a, False)
        f(b, True)
        res[name + "_cpi_per_s_one_core"] = n_cpi / (time.perf_counter() - t0)
    res["oracle_fft_slowdown_vs_pocketfft"] = res["pocketfft_c64_cpi_per_s_one_core"] / res["oracle_radix2_cpi_per_s_one_core"]
    res["what"] = "FFT stages only; the CPU arm's chain spends ~60 % of its time in the oracle FFT"
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    per_thread = max(1, args.batch // cores)                       # the same batch as the GPU arm, split over the cores
    batch = per_thread * cores
    rate, times = cpu_chain_rate(cores, per_thread, repeats=args.steps, warm=args.warmup)
    sample = f"{batch} CPIs per step ({per_thread} per thread x {cores} threads): the GPU arm's batch of the same workload"
    out = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config_dict(args.batch, world),
           "note": "reference chain on the host cores: the reference's own block sources (oracle/_ref) when built, else the "
                   "oracle restatement; no GNU Radio scheduler, float32 radix-2 FFT instead of gr-fft/FFTW (neither is "
                   "installable here; see cpu_baseline.fft_sanity)",
           "complex_msps": rate * CFG["R"] * CFG["S"] * CFG["N"] / 1e6,
           "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": cpu_kind(), "sample": sample,
                            "fft_sanity": fft_sanity()},
           "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def time_config(jrc, torch, name, c, dev, peak_gbs, reps):
    """One of the other BASELINE configurations: whole-chain CPI/s over a resident batch, event-timed."""
    cfg = {k: c[k] for k in ("T", "R", "S", "N", "IR", "IA")}
    n = c["n"]
    rx_h, tx_h, est = make_inputs(n, seed=5, cfg=cfg, targets=c["targets"], span=15.0)
    rc = jrc.radar_chain(cfg["N"], cfg["T"], cfg["R"], cfg["S"], cfg["IR"], cfg["IA"], device=dev.index, estimator=est)
    drx, dtx = torch.from_numpy(rx_h).to(dev), torch.from_numpy(tx_h).to(dev)
    dmap = torch.empty((n, rc.Nr, rc.Na), dtype=torch.float32, device=dev)
    ddet = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
    ext = torch.cuda.ExternalStream(rc.chain.stream, device=dev)
    torch.cuda.synchronize()
    l0 = None
    with torch.cuda.stream(ext):
        for _ in range(3):
            rc.run(drx, dtx, map_out=dmap, dets_out=ddet, sync_inputs=False)
        l0 = rc.chain.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(reps):
            rc.run(drx, dtx, map_out=dmap, dets_out=ddet, sync_inputs=False)
        e1.record(ext)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    launches = (rc.chain.launch_count - l0) // reps
    d = rc.dets_to_numpy(ddet)
    rate = n / (ms * 1e-3)
    path = {jrc.PATH_FUSED: "fused", jrc.PATH_TILED: "tiled", jrc.PATH_STAGED: "staged"}[rc.chain.last_path]
    alg = b_alg_per_cpi(cfg)
    out = {"path": path, "cpis_per_call": n, "ms_per_call": ms, "cpi_per_s": rate, "launches_per_call": int(launches),
           "complex_msps": rate * cfg["R"] * cfg["S"] * cfg["N"] / 1e6, "map": [rc.Nr, rc.Na],
           "algorithmic_bytes_per_cpi": alg, "achieved_gbs": rate * alg / 1e9, "roofline_frac": rate * alg / 1e9 / peak_gbs,
           "gate_pass_fraction": float((d["flags"] & 1).mean()), "exact_pass": rc.chain.exact_stats(),
           "kernels": c["kernels"] or (["k_chan_est(_tile)", "k_fft8_rows", "k_angle_mag", "k_map_finalize", "k_est_exact"]
                                       if path == "tiled" else None)}
    del drx, dtx, dmap, ddet, rc
    torch.cuda.empty_cache()
    return out


def time_raw_samples(jrc, torch, dev, peak_gbs, reps):
    """configs[1] fed with the RX antennas' raw time samples (jrc_chain_run_batch_time): cyclic-prefix removal and the RX OFDM
    FFT in front of the chain, on the device.  B_alg counts what this form reads: R*(n_sym)*(fft_len+cp) time samples (the
    preamble symbols are skipped, not read) + T*S*N TX symbols, and the same map + record out."""
    cfg, n, pre, cp = CFG, 4096, 0, 16
    N, T, R, S = cfg["N"], cfg["T"], cfg["R"], cfg["S"]
    rx_h, tx_h, est = make_inputs(n, seed=9, cfg=cfg, targets=1, span=15.0)
    td = np.fft.ifft(np.fft.ifftshift(rx_h.astype(np.complex128), axes=-1), axis=-1)
    td = np.concatenate([td[..., N - cp:], td], axis=-1).astype(np.complex64)          # [n][R][S][cp+N]
    ch = jrc.Chain(N, T, R, S, pre, cfg["IR"], cfg["IA"], device=dev.index)
    ch.set_estimator(**est)
    dtd, dtx = torch.from_numpy(td).to(dev), torch.from_numpy(tx_h).to(dev)
    dmap = torch.empty((n, ch.Nr, ch.Na), dtype=torch.float32, device=dev)
    ddet = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
    row_t, row_f = S * (N + cp), S * N
    tx_cpi = T * row_f if tx_h.ndim == 4 else 0
    ext = torch.cuda.ExternalStream(ch.stream, device=dev)
    torch.cuda.synchronize()

    def call():
        ch.run_batch_time_ptr(dtd.data_ptr(), R * row_t, row_t, cp, dtx.data_ptr(), tx_cpi, row_f, n, 0, dmap.data_ptr(), None,
                              ddet.data_ptr())
    with torch.cuda.stream(ext):
        for _ in range(3):
            call()
        l0 = ch.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(reps):
            call()
        e1.record(ext)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rate = n / (ms * 1e-3)
    alg = R * S * (N + cp) * 8 + T * S * N * 8 + ch.Nr * ch.Na * 4 + 32
    out = {"path": "fused", "cpis_per_call": n, "ms_per_call": ms, "cpi_per_s": rate, "launches_per_call": int((ch.launch_count - l0) // reps),
           "cp_len": cp, "algorithmic_bytes_per_cpi": alg, "achieved_gbs": rate * alg / 1e9, "roofline_frac": rate * alg / 1e9 / peak_gbs,
           "kernels": ["k_ofdm_demod64", "k_fused64x8<16,8>", "k_est_exact"]}
    del dtd, dtx, dmap, ddet, ch
    torch.cuda.empty_cache()
    return out


def latency_mode():
    """configs[3] through the C++ block harness (built by build()); its JSON is passed through."""
    exe = os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "build", "latency_blocks")
    if not os.path.exists(exe):
        return {"unavailable": "build/latency_blocks not built"}
    try:
        r = subprocess.run([exe, "4000"], capture_output=True, text=True, timeout=300)
        if r.returncode != 0:
            return {"unavailable": (r.stderr or r.stdout)[-300:]}
        return json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)}


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores_bound = bind_rank_to_cores(local, world) if world > 1 else None
    import torch
    import torch.distributed as dist
    import mimo_ofdm_jrc as jrc
    from mimo_ofdm_jrc import shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    B, K, W = args.batch, args.steps, max(3, args.warmup)
    rx_h, tx_h, est = make_inputs(B, seed=100 + rank)
    rc = jrc.radar_chain(CFG["N"], CFG["T"], CFG["R"], CFG["S"], CFG["IR"], CFG["IA"], device=local, estimator=est)
    Nr, Na = rc.Nr, rc.Na
    rx = torch.from_numpy(rx_h).to(dev)
    tx = torch.from_numpy(tx_h).to(dev)
    dmap = torch.empty((B, Nr, Na), dtype=torch.float32, device=dev)
    # Detection records of all K steps.  N > 1: the table lives in rank 0's memory and every rank's kernels store their
    # 32-byte records straight into it over NVLink (shard.PeerRecordTable, CUDA IPC): no collective on the timed path.
    # JRC_BENCH_GATHER=1 (or a failed IPC mapping) falls back to one NCCL gather after the last step.
    ddet_all = torch.zeros((K, B, 32), dtype=torch.uint8, device=dev)
    table = None
    if world > 1 and not os.environ.get("JRC_BENCH_GATHER"):
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            table = shard.PeerRecordTable(K * B, local)
        except Exception as e:  # noqa: BLE001
            print(f"rank {rank}: peer record table unavailable ({e}); NCCL gather instead", file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            table = None
    gath = torch.empty((world * K * B, 32), dtype=torch.uint8, device=dev) if (world > 1 and rank == 0 and table is None) else None
    ext = torch.cuda.ExternalStream(rc.chain.stream, device=dev)
    torch.cuda.synchronize()

    # JRC_BENCH_PEER_STORES=1: the kernels store their records into the table themselves (1.6 % slower steps on the
    # writing ranks: a kernel waits for its NVLink stores when it ends); default: local records, one peer copy at the drain
    peer_stores = table is not None and bool(os.environ.get("JRC_BENCH_PEER_STORES"))

    def step(k):
        if peer_stores:
            rc.run(rx, tx, map_out=dmap, dets_ptr=table.ptr((k % K) * B), path=jrc.PATH_FUSED, sync_inputs=False)
        else:
            rc.run(rx, tx, map_out=dmap, dets_out=ddet_all[k % K], path=jrc.PATH_FUSED, sync_inputs=False)

    def gather_all():
        if table is not None:
            if not peer_stores:
                table.push(rc.chain, ddet_all.view(K * B, 32))
        elif world > 1:
            shard.gather_detections(ddet_all.view(K * B, 32), dst=0, counts=[K * B] * world, out=gath)

    with torch.cuda.stream(ext):
        for k in range(W):
            step(k)
        gather_all()
    torch.cuda.synchronize()
    launches0 = rc.chain.launch_count
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    with torch.cuda.stream(ext):
        e0.record(ext)
        for k in range(K):
            ev[k][0].record(ext)
            step(k)
            ev[k][1].record(ext)
        gather_all()
        e1.record(ext)
    torch.cuda.synchronize()
    elapsed_local = e0.elapsed_time(e1)
    launches = rc.chain.launch_count - launches0
    # the timed region of a short run is a few milliseconds: keep the same kernel running until the clock record
    # holds enough samples (NOT part of the timed value)
    if sampler:
        t_end = time.perf_counter() + 0.25
        with torch.cuda.stream(ext):
            while time.perf_counter() < t_end:
                for k in range(8):
                    step(k)
                ext.synchronize()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["note"] = "sampled over the timed region and 0.25 s of the same step right after it"
    barrier()
    elapsed_ms = max_over_ranks(elapsed_local)
    step_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    rank_ms = [elapsed_local / K]
    if world > 1:        # every rank's own time per step: the job's value is set by the slowest GPU of the box
        t_all = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(t_all, torch.tensor([elapsed_local / K], dtype=torch.float64, device=dev))
        rank_ms = [float(t.item()) for t in t_all]
    value = world * B * K / (elapsed_ms * 1e-3)
    exact_stats = rc.chain.exact_stats()

    # ---- dominant kernel alone: k_fused64x8 without the (tiny) reference-order pass behind it ----
    with torch.cuda.stream(ext):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nk = max(5, min(K, 50))
        k0.record(ext)
        for _ in range(nk):
            rc.run(rx, tx, map_out=dmap, want_dets=False, path=jrc.PATH_FUSED, sync_inputs=False)
        k1.record(ext)
    torch.cuda.synchronize()
    map_only_ms = k0.elapsed_time(k1) / nk

    # ---- end to end through the host-buffer C-ABI call (pinned host memory) ----
    e2e = None
    if not args.no_e2e:
        Ke = max(3, min(K, 10))
        prx = torch.from_numpy(rx_h).pin_memory()
        ptx = torch.from_numpy(tx_h).pin_memory()
        pmap = torch.empty((B, Nr, Na), dtype=torch.float32).pin_memory()
        pdet = torch.zeros((B, 32), dtype=torch.uint8).pin_memory()

        def host_step(with_map):
            rc.chain.run_host_ptr(prx.data_ptr(), ptx.data_ptr(), False, B, 0,
                                  pmap.data_ptr() if with_map else None, pdet.data_ptr())

        res = {}
        for with_map in (True, False):
            host_step(with_map)
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(Ke):
                host_step(with_map)
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
            res[with_map] = world * B * Ke / dt
        # the ceiling of that path: a plain pinned D2H copy of the same 1 GiB, all ranks at once
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pmap.copy_(dmap, non_blocking=True)
        torch.cuda.synchronize()
        barrier()
        c0.record()
        for _ in range(3):
            pmap.copy_(dmap, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        d2h_gbs_rank = 3 * dmap.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        d2h_gbs_all = sum_over_ranks(d2h_gbs_rank)
        h2d = int(rx_h.nbytes + tx_h.nbytes)
        d2h = int(B * Nr * Na * 4 + B * 32)
        e2e = {"value": res[True], "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
               "what": "jrc_chain_run_host: pinned host symbols in, |.|^2 map + detection records out",
               "d2h_ceiling_gbs": d2h_gbs_all, "d2h_ceiling_what": "plain pinned cudaMemcpyAsync D2H of the same map, all ranks at once, sum over ranks",
               "pcie_frac": (res[True] / world) * (d2h / B) * world / 1e9 / d2h_gbs_all,
               "host_cores_per_rank": len(cores_bound) if cores_bound else None,
               "detections_only": {"value": res[False], "unit": UNIT, "h2d_bytes_per_step": h2d,
                                   "d2h_bytes_per_step": int(B * 32)}}
        del prx, ptx, pmap, pdet

    # ---- sanity: the timed output is the real thing --------------------------------
    if table is not None:
        table.complete()                 # every rank's stream is synchronised: the table in rank 0's memory is whole
        recs = table.records()
        if rank == 0:
            g = recs.cpu().numpy().view(jrc.DET_DTYPE).reshape(world, K, B)
            ddet_all.copy_(recs.view(world, K, B, 32)[0])
            for r_ in range(world):      # every rank's records arrived: CPI ids and the gate
                assert np.array_equal(g[r_, K - 1]["cpi"], np.arange(B)) and (g[r_, K - 1]["flags"] & 1).mean() > 0.9, r_
    d = rc.dets_to_numpy(ddet_all[(K - 1) % K])
    if rank == 0 or table is None:
        assert (d["flags"] & 1).mean() > 0.9 and d["range_idx"].max() < Nr and d["angle_idx"].max() < Na
    if gath is not None:
        g = gath.cpu().numpy().view(jrc.DET_DTYPE).reshape(world, K, B)
        assert np.array_equal(g[0, K - 1]["range_idx"], d["range_idx"])

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak_gbs, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks \
        else (6650.0, "fallback (B200_PROFILING.md)")

    # ---- configs[4] sweep: 65 536 CPIs sharded over the ranks (north_star) ----
    sweep = None
    if world > 1 or args.sweep:
        c5 = OTHER_CONFIGS["configs[4] 8x16, 2048 sc, 2048x128"]
        cfg5 = {k: c5[k] for k in ("T", "R", "S", "N", "IR", "IA")}
        total = 65536
        lo, hi = shard.shard_range(total, rank, world)
        nblk = 222                                  # one round of the two wide kernels: both grids filled exactly
        from mimo_ofdm_jrc import synth
        rx5_h, tx5_h, est5 = make_inputs(nblk, seed=7 + rank, cfg=cfg5, targets=3, span=15.0)
        rc5 = jrc.radar_chain(cfg5["N"], cfg5["T"], cfg5["R"], cfg5["S"], cfg5["IR"], cfg5["IA"], device=local, estimator=est5)
        rx5, tx5 = torch.from_numpy(rx5_h).to(dev), torch.from_numpy(tx5_h).to(dev)
        # every block of 222 CPIs is a NEW scene, synthesised on the device right before it is processed (jrc_scene_synth:
        # the 128 GiB of RX symbols of the sweep exist 444 MiB at a time); noise level as in make_inputs (20 dB)
        sigma5 = float(np.sqrt(np.mean(np.abs(rx5_h[:4]) ** 2)) * 10 ** (-20.0 / 20.0) / np.sqrt(2) / np.sqrt(1.01))
        rng5 = np.random.default_rng(1000 + rank)
        m5 = torch.empty((nblk, rc5.Nr, rc5.Na), dtype=torch.float32, device=dev)
        d5 = torch.zeros((hi - lo, 32), dtype=torch.uint8, device=dev)
        g5 = torch.empty((world * ((total + world - 1) // world), 32), dtype=torch.uint8, device=dev) if (world > 1 and rank == 0) else None
        ext5 = torch.cuda.ExternalStream(rc5.chain.stream, device=dev)
        counts5 = [shard.shard_range(total, r, world)[1] - shard.shard_range(total, r, world)[0] for r in range(world)]
        with torch.cuda.stream(ext5):
            rc5.run(rx5, tx5, map_out=m5, dets_out=d5[:nblk], sync_inputs=False)
            if world > 1:       # untimed: the first gather sets up NCCL's point-to-point connections (~1 s)
                shard.gather_detections(d5, dst=0, counts=counts5, out=g5)
        torch.cuda.synchronize()
        barrier()
        chain_ms, synth_s = 0.0, 0.0
        evs = []
        t_wall0 = time.perf_counter()
        with torch.cuda.stream(ext5):
            for c0 in range(0, hi - lo, nblk):
                nc = min(nblk, hi - lo - c0)
                r_, a_, amp_ = synth.random_scene(rng5, nc, 3, cfg5["N"], amp_db_span=15.0)
                ts = time.perf_counter()
                rc5.chain.scene_synth_ptr(tx5_h, r_, a_, amp_, rx5.data_ptr(), noise_sigma=sigma5, seed=(rank << 32) + c0)
                synth_s += time.perf_counter() - ts
                e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e_a.record(ext5)
                rc5.run(rx5[:nc], tx5, map_out=m5[:nc], dets_out=d5[c0:c0 + nc], cpi0=lo + c0, sync_inputs=False)
                e_b.record(ext5)
                evs.append((e_a, e_b))
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(ext5)
            if world > 1:
                shard.gather_detections(d5, dst=0, counts=counts5, out=g5)
            s1.record(ext5)
        torch.cuda.synchronize()
        wall5 = max_over_ranks(time.perf_counter() - t_wall0)
        chain_ms = sum(a.elapsed_time(b) for a, b in evs) + s0.elapsed_time(s1)
        d5n = rc5.dets_to_numpy(d5)
        assert (d5n["flags"] & 1).mean() > 0.9 and np.array_equal(d5n["cpi"], np.arange(lo, hi))
        ms5 = max_over_ranks(chain_ms)
        sweep = {"workload": "configs[4]: 2048 subcarriers, 8 x 16 virtual array, 65536 CPIs sharded over the ranks, detection records "
                             "gathered to rank 0 over NCCL; every block of 222 CPIs is a new 3-target scene synthesised on the device "
                             "(jrc_scene_synth) right before it is processed; ms = chain time (CUDA events around every chain call + "
                             "the gather, max over ranks), wall_s includes the scene synthesis",
                 "n_gpus": world, "cpis": total, "ms": ms5, "wall_s": wall5, "cpi_per_s": total / (ms5 * 1e-3),
                 "complex_gsps": total / (ms5 * 1e-3) * cfg5["R"] * cfg5["S"] * cfg5["N"] / 1e9,
                 "roofline_frac_per_gpu": total / (ms5 * 1e-3) * b_alg_per_cpi(cfg5) / 1e9 / peak_gbs / world}
        del rx5, tx5, m5, d5, rc5
        torch.cuda.empty_cache()

    if rank == 0:
        alg_bytes = b_alg_per_cpi(CFG) * B
        kern_ms = step_ms
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        traffic, traffic_note = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["k_fused64x8"]
            if tj.get("source_hash") == source_hash():
                traffic = tj["dram_bytes_per_launch"]
            else:
                traffic_note = "profiles/ncu_traffic.json was captured on other kernel sources (hash mismatch): not quoted"
        except Exception:  # noqa: BLE001
            pass
        cpu = None
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            r1, _ = cpu_chain_rate(1, 4)
            per_thread = max(1, min(256, int(r1 * 12.0)))           # ~12 s of work on every core
            rate, _ = cpu_chain_rate(cores, per_thread)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": cpu_kind(),
                   "sample": f"{cores * per_thread} CPIs of the same workload ({per_thread} per thread x {cores} threads); "
                             "kind reference = the reference's block sources (oracle/_ref) + float32 radix-2 FFT for the "
                             "two stock GNU Radio FFT blocks", "single_thread": r1, "fft_sanity": fft_sanity()}
        configs, latency = None, None
        if world == 1 and not args.no_configs:
            configs = {name: time_config(jrc, torch, name, c, dev, peak_gbs, reps=10) for name, c in OTHER_CONFIGS.items()}
            configs["configs[1] from raw RX time samples (CP removal + OFDM FFT on the device)"] = time_raw_samples(jrc, torch, dev, peak_gbs, reps=10)
            latency = latency_mode()
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic",
               "config": config_dict(B, world),
               "complex_msps": value * CFG["R"] * CFG["S"] * CFG["N"] / 1e6,
               "ms_per_step_by_rank": [round(x, 5) for x in rank_ms],
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                            "frac": achieved / peak_gbs, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                            "kernel": "k_fused64x8<16,8>", "kernel_ms": kern_ms, "map_only_kernel_ms": map_only_ms,
                            "kernel_ms_what": "mean CUDA-event time of one step on the kernel's stream = k_fused64x8 with its in-kernel "
                                              "estimator plus the (programmatically launched, normally empty) k_est_exact pass behind it; "
                                              "map_only_kernel_ms = the same kernel launched alone without the estimator",
                            "algorithmic_bytes_per_launch": alg_bytes},
               "exact_pass": exact_stats,
               "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
               "configs": configs, "latency": latency, "sweep_c5": sweep}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="run the configs[4] 65536-CPI sweep on one GPU as well")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:      # plain `python bench.py --gpus N`: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
