"""k_fused_tc (experiment build) against k_fused64x8 on the bench scene: map error, time per 4096 CPIs.
    make -C gr-mimo-ofdm-jrc_b200 libjrc_cuda_tc.so && JRC_CUDA_LIB=$PWD/gr-mimo-ofdm-jrc_b200/libjrc_cuda_tc.so python scripts/tc_check.py
JRC_TC_DBG=1/2/3: without the range passes / the map stores / both (elimination runs)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "python")]
import numpy as np, torch
import mimo_ofdm_jrc as jrc
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
want_dets = len(sys.argv) > 2 and sys.argv[2] == "dets"
C = bench.CFG
rx_h, tx_h, est = bench.make_inputs(B, seed=100)
rx, tx = torch.from_numpy(rx_h).cuda(), torch.from_numpy(tx_h).cuda()
out = {}
res = {}
for mode in ("0", "1"):
    os.environ["JRC_TC"] = mode
    rc = jrc.radar_chain(C["N"], C["T"], C["R"], C["S"], C["IR"], C["IA"], estimator=est)
    m = torch.empty((B, rc.Nr, rc.Na), dtype=torch.float32, device="cuda")
    d = torch.zeros((B, 32), dtype=torch.uint8, device="cuda") if want_dets else None
    ext = torch.cuda.ExternalStream(rc.chain.stream)
    torch.cuda.synchronize()
    def step():
        rc.run(rx, tx, map_out=m, dets_out=d, want_dets=want_dets, path=jrc.PATH_FUSED, sync_inputs=False)
    with torch.cuda.stream(ext):
        for _ in range(3): step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(20): step()
        e1.record(ext)
    torch.cuda.synchronize()
    out["ms_tc" if mode == "1" else "ms_simt"] = e0.elapsed_time(e1) / 20
    res[mode] = (m.clone(), rc.dets_to_numpy(d) if want_dets else None)
    del rc
m0, m1 = res["0"][0], res["1"][0]
pk = m0.reshape(B, -1).max(dim=1).values
out["map_err_of_peak"] = float(((m1 - m0).abs().reshape(B, -1).max(dim=1).values / pk).max())
if want_dets:
    d0, d1 = res["0"][1], res["1"][1]
    out["idx_equal"] = bool(np.array_equal(d0["range_idx"], d1["range_idx"]) and np.array_equal(d0["angle_idx"], d1["angle_idx"]))
    out["flags_equal"] = bool(np.array_equal(d0["flags"] & 1, d1["flags"] & 1))
print(json.dumps(out))
