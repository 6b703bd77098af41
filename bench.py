#!/usr/bin/env python
"""bench.py -- range-angle CPIs/s of the radar hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the fused chain (mimo_ofdm_radar -> range IFFT -> transpose -> angle FFT
-> |.|^2 -> range_angle_estimator) over one batch of BASELINE configs[1]: 64 subcarriers,
4 TX x 2 RX = 8 virtual channels, 4 LTF symbols, range zero-pad 1024, angle zero-pad 64,
4096 CPIs per GPU.  CPIs are independent, so N GPUs each process their own 4096-CPI shard
(weak scaling) and only the 32-byte detection records are gathered to rank 0 over NCCL.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference
chain (oracle/) on the host cores for the same metric and configuration.
"""
import argparse
import json
import os
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


def b_alg_per_cpi(c):
    """read (T+R)*S*Nsc complex once, write Nr*Na float32 once, 32 B record."""
    return (c["T"] + c["R"]) * c["S"] * c["N"] * 8 + (c["N"] * c["IR"]) * (c["T"] * c["R"] * c["IA"]) * 4 + 32


def make_inputs(batch, seed):
    from mimo_ofdm_jrc import synth
    rng = np.random.default_rng(seed)
    tx = synth.tx_symbols(CFG["T"], CFG["S"], CFG["N"])
    r, a, amp = synth.random_scene(rng, batch, 2, CFG["N"], amp_db_span=10.0)
    rx = synth.rx_symbols(tx, CFG["R"], r, a, amp, snr_db=20.0, rng=rng)
    txb = np.ascontiguousarray(np.broadcast_to(tx, (batch,) + tx.shape))
    est = synth.default_estimator_params(CFG["N"], CFG["T"] * CFG["R"], CFG["IR"], CFG["IA"])
    return rx, txb, est


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
            time.sleep(0.002)

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
# CPU arm: the oracle restatement of the reference chain on the host cores
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

    def work(i):
        sl = slice(i * per_thread, (i + 1) * per_thread)
        if ref is None:
            orc.chain_batch(rx[sl], tx[sl], CFG["N"], CFG["T"], CFG["R"], CFG["S"], CFG["IR"], CFG["IA"], est)
            return
        c = orc.ChainCfg(CFG["N"], CFG["T"], CFG["R"], CFG["S"], 0, CFG["IR"], CFG["IA"], 0, rb.ctypes.data, ab.ctypes.data,
                         est["noise_discard_range_m"], est["noise_discard_angle_deg"], est["snr_threshold"],
                         est["power_threshold"])
        a, b = np.ascontiguousarray(rx[sl]), np.ascontiguousarray(tx[sl])
        m = np.empty((per_thread, Nr, Na), np.float32)
        d = np.zeros(per_thread, orc.DET_DTYPE)
        ref.ref_chain_batch(C.byref(c), a.ctypes.data, b.ctypes.data, 0, per_thread, 0, m.ctypes.data, None, d.ctypes.data)

    times = []
    with ThreadPoolExecutor(n_threads) as ex:
        for it in range(warm + repeats):
            t0 = time.perf_counter()
            list(ex.map(work, range(n_threads)))
            if it >= warm:
                times.append(time.perf_counter() - t0)
    return n_threads * per_thread * len(times) / sum(times), times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    r1, _ = cpu_chain_rate(1, 4)                                   # calibrate: CPIs/s on one thread
    per_thread = max(1, min(64, int(r1 * 1.0)))                     # ~1 s of work per thread per step
    rate, times = cpu_chain_rate(cores, per_thread, repeats=args.steps, warm=args.warmup)
    sample = f"{cores * per_thread} CPIs per step ({per_thread} per thread x {cores} threads) of the same workload"
    out = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "note": "reference chain on the host cores: the reference's own block sources "
                      "(oracle/_ref) when built, else the oracle restatement; no GNU Radio scheduler, float32 radix-2 FFT "
                      "instead of gr-fft/FFTW (neither is installable here)"},
           "complex_msps": rate * CFG["R"] * CFG["S"] * CFG["N"] / 1e6,
           "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": cpu_kind(), "sample": sample},
           "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import mimo_ofdm_jrc as jrc
    from mimo_ofdm_jrc import shard

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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

    B, K, W = args.batch, args.steps, max(3, args.warmup)
    rx_h, tx_h, est = make_inputs(B, seed=100 + rank)
    rc = jrc.radar_chain(CFG["N"], CFG["T"], CFG["R"], CFG["S"], CFG["IR"], CFG["IA"], device=local, estimator=est)
    Nr, Na = rc.Nr, rc.Na
    rx = torch.from_numpy(rx_h).to(dev)
    tx = torch.from_numpy(tx_h).to(dev)
    dmap = torch.empty((B, Nr, Na), dtype=torch.float32, device=dev)
    ddet = torch.zeros((B, 32), dtype=torch.uint8, device=dev)
    ext = torch.cuda.ExternalStream(rc.chain.stream, device=dev)
    torch.cuda.synchronize()

    counts = [B] * world                       # contiguous equal shards: no size exchange needed
    # two detection buffers per rank: the gather of step k overlaps the kernel of step k+1
    ddets = [ddet, torch.zeros_like(ddet)] if world > 1 else [ddet]
    gbufs = [torch.empty((world * B, 32), dtype=torch.uint8, device=dev) for _ in range(2)] if (world > 1 and rank == 0) else [None, None]
    pending = [None, None]

    def step(k=0):
        slot = k & 1 if world > 1 else 0
        if pending[slot] is not None:
            pending[slot].wait()               # the gather that last read this buffer
            pending[slot] = None
        rc.run(rx, tx, map_out=dmap, dets_out=ddets[slot], path=jrc.PATH_FUSED, sync_inputs=False)
        if world > 1:
            pending[slot], _ = shard.gather_detections(ddets[slot], dst=0, counts=counts, out=gbufs[slot], async_op=True)

    def drain():
        for i in range(2):
            if pending[i] is not None:
                pending[i].wait()
                pending[i] = None

    with torch.cuda.stream(ext):
        for k in range(W):
            step(k)
        drain()
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
        drain()
        e1.record(ext)
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    barrier()
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    launches = rc.chain.launch_count - launches0
    value = world * B * K / (elapsed_ms * 1e-3)

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
        h2d = int(rx_h.nbytes + tx_h.nbytes)
        e2e = {"value": res[True], "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(B * Nr * Na * 4 + B * 32), "steps": Ke,
               "what": "jrc_chain_run_host: pinned host symbols in, |.|^2 map + detection records out",
               "detections_only": {"value": res[False], "unit": UNIT, "h2d_bytes_per_step": h2d,
                                   "d2h_bytes_per_step": int(B * 32)}}
        del prx, ptx, pmap, pdet

    # ---- sanity: the timed output is the real thing --------------------------------
    d = rc.dets_to_numpy(ddet)
    assert (d["flags"] & 1).mean() > 0.9 and d["range_idx"].max() < Nr and d["angle_idx"].max() < Na

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak_gbs, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks \
            else (6650.0, "fallback (B200_PROFILING.md)")
        alg_bytes = b_alg_per_cpi(CFG) * B
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["k_fused64x8"]["dram_bytes_per_launch"]
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
                             "two stock GNU Radio FFT blocks", "single_thread": r1}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic",
               "config": {"workload": WORKLOAD, "batch_per_gpu": B, "map": [Nr, Na],
                          "l2": "per-step working set (1.0 GiB map + 48 MiB symbols) exceeds the 126 MB L2; no flush needed",
                          "parallelism": f"cpi-shard x{world}, detections gathered to rank 0" if world > 1 else "single GPU"},
               "complex_msps": value * CFG["R"] * CFG["S"] * CFG["N"] / 1e6,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                            "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                            "kernel": "k_fused64x8<16,8>", "kernel_ms": kern_ms,
                            "algorithmic_bytes_per_launch": alg_bytes},
               "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
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
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:      # plain `python bench.py --gpus N`: relaunch under torchrun
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
