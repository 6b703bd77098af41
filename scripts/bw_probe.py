"""GPU probe: HBM write-only / read-only / copy bandwidth with torch ops, and the fused kernel
with and without the map store (how much of the step is the store stream)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "python"))
import numpy as np, torch
import mimo_ofdm_jrc as jrc
import bench

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3

N = 1 << 28   # 1 GiB of float32
FUSED_ONLY = '--fused-only' in sys.argv
if not FUSED_ONLY:
  a = torch.empty(N, dtype=torch.float32, device="cuda")
  b = torch.empty(N, dtype=torch.float32, device="cuda")
  t = timeit(lambda: a.fill_(1.0)); print(f"fill_ (write only): {N*4/t/1e9:.0f} GB/s")
  t = timeit(lambda: a.zero_()); print(f"zero_ (memset): {N*4/t/1e9:.0f} GB/s")
  t = timeit(lambda: b.copy_(a)); print(f"copy_ (read+write): {2*N*4/t/1e9:.0f} GB/s")
  t = timeit(lambda: a.sum()); print(f"sum (read only): {N*4/t/1e9:.0f} GB/s")

B = 4096
rx_h, tx_h, est = bench.make_inputs(B, seed=1)
C = bench.CFG
rc = jrc.radar_chain(C["N"], C["T"], C["R"], C["S"], C["IR"], C["IA"], estimator=est)
rx, tx = torch.from_numpy(rx_h).cuda(), torch.from_numpy(tx_h).cuda()
dmap = torch.empty((B, rc.Nr, rc.Na), dtype=torch.float32, device="cuda")
ddet = torch.zeros((B, 32), dtype=torch.uint8, device="cuda")
ext = torch.cuda.ExternalStream(rc.chain.stream)
torch.cuda.synchronize()
for name, kw in (("map+dets", dict(want_map=True, want_dets=True)), ("dets only", dict(want_map=False, want_dets=True)),
                 ("map only", dict(want_map=True, want_dets=False))):
    def f():
        rc.run(rx, tx, map_out=dmap if kw["want_map"] else None, dets_out=ddet if kw["want_dets"] else None, sync_inputs=False, **kw)
    with torch.cuda.stream(ext):
        for _ in range(5): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(50): f()
        e1.record(ext)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"fused {name}: {ms*1e3:.1f} us per 4096 CPIs -> {B/ms/1e3:.2f} M CPI/s")
