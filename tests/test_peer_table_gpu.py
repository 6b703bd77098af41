"""GPU, two processes on ONE device over gloo: the multi-GPU detection table (shard.PeerRecordTable) -- rank 0's table mapped
into the other process by CUDA IPC, filled (a) by one peer copy per rank behind its last batch and (b) by the chain kernels'
own stores -- against each rank's locally produced records.  (NCCL refuses two ranks on one GPU; the table only needs
torch.distributed for the 64-byte handle and the barrier, and gloo carries both.)"""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, os.path.join({root!r}, "gr-mimo-ofdm-jrc_b200", "python"))
    import numpy as np, torch, torch.distributed as dist
    import mimo_ofdm_jrc as jrc
    from mimo_ofdm_jrc import shard, synth
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(0)
    cfg = dict(T=4, R=2, S=4, N=64, IR=16, IA=8)
    n = 96
    rng = np.random.default_rng(100 + rank)
    tx = synth.tx_symbols(cfg["T"], cfg["S"], cfg["N"])
    r, a, amp = synth.random_scene(rng, n, 2, cfg["N"], amp_db_span=6.0)
    rx = synth.rx_symbols(tx, cfg["R"], r, a, amp, snr_db=20.0, rng=rng)
    est = synth.default_estimator_params(cfg["N"], cfg["T"] * cfg["R"], cfg["IR"], cfg["IA"])
    rc = jrc.radar_chain(cfg["N"], cfg["T"], cfg["R"], cfg["S"], cfg["IR"], cfg["IA"], device=0, estimator=est)
    drx, dtx = torch.from_numpy(rx).cuda(), torch.from_numpy(tx).cuda()
    local = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    rc.run(drx, dtx, want_map=False, dets_out=local, cpi0=1000 * rank, sync_inputs=False)
    rc.sync()
    mine = rc.dets_to_numpy(local)
    assert np.array_equal(mine["cpi"], 1000 * rank + np.arange(n)) and (mine["range_idx"] >= 0).all(), mine[:4]
    for mode in ("push", "stores"):
        table = shard.PeerRecordTable(n, 0)
        if mode == "push":
            table.push(rc.chain, local)                      # one peer copy behind the batch
        else:
            rc.run(drx, dtx, want_map=False, dets_ptr=table.ptr(0), cpi0=1000 * rank, sync_inputs=False)   # the kernels store
        rc.sync()
        table.complete()
        recs = table.records()
        if rank == 0:
            g = recs.cpu().numpy().view(jrc.DET_DTYPE).reshape(world, n)
            assert np.array_equal(g[0], mine), mode
            for r_ in range(world):
                assert np.array_equal(g[r_]["cpi"], 1000 * r_ + np.arange(n)) and (g[r_]["range_idx"] >= 0).all(), (mode, r_)
        # every rank checks its own slice through a gather of the raw bytes over gloo
        mine_t = torch.from_numpy(mine.view(np.uint8).reshape(n, 32).copy())
        allm = shard.gather_detections(mine_t, dst=0, counts=[n] * world)
        if rank == 0:
            assert np.array_equal(allm.numpy().view(jrc.DET_DTYPE).reshape(world, n), g), mode
        try:
            table.push(rc.chain, local, i=1)
            raise SystemExit("a slice overflow was accepted")
        except ValueError:
            pass
        table.close()
    if rank == 0:
        print("PEER_TABLE_OK")
    dist.barrier(); dist.destroy_process_group()
""")


def test_peer_record_table_two_processes_one_gpu(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "PEER_TABLE_OK" in r.stdout
