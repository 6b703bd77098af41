"""How many records of the bench scene are redone in the reference's order (csrc/jrc_exact.cuh), and why."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "python")]
import numpy as np, torch
import mimo_ofdm_jrc as jrc
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rx_h, tx_h, est = bench.make_inputs(B, seed=100)
C = bench.CFG
rc = jrc.radar_chain(C["N"], C["T"], C["R"], C["S"], C["IR"], C["IA"], estimator=est)
rx, tx = torch.from_numpy(rx_h).cuda(), torch.from_numpy(tx_h).cuda()
m, d = rc.run(rx, tx, path=jrc.PATH_FUSED)
rc.sync()
d = rc.dets_to_numpy(d)
ex = (d["flags"] & jrc.DET_EXACT) != 0
print(json.dumps({"batch": B, "stats": rc.chain.exact_stats(), "exact": int(ex.sum()), "pending_left": int(((d["flags"] >> 31) & 1).sum()),
                  "snr_min": float(np.nanmin(d["snr_db"])), "snr_max": float(np.nanmax(d["snr_db"])),
                  "passed": int((d["flags"] & 1).sum())}))
