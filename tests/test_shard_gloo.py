"""CPU-only, world_size 2 over gloo: the multi-GPU host logic (contiguous CPI shards + the
detection gather to rank 0) without a GPU."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from mimo_ofdm_jrc import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_batch_contiguously():
    for n in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            edges = [shard.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, os.path.join({root!r}, "gr-mimo-ofdm-jrc_b200", "python"))
    import numpy as np, torch, torch.distributed as dist
    from mimo_ofdm_jrc import shard
    from mimo_ofdm_jrc.cabi import DET_DTYPE
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 37                                   # ragged on purpose
    lo, hi = shard.shard_range(n, rank, world)
    d = np.zeros(hi - lo, dtype=DET_DTYPE)
    d["cpi"] = np.arange(lo, hi); d["range_idx"] = 3 * np.arange(lo, hi); d["snr_db"] = rank + 0.5
    t = torch.from_numpy(d.view(np.uint8).reshape(-1, 32).copy())
    out = shard.gather_detections(t, dst=0)
    if rank == 0:
        g = out.numpy().view(DET_DTYPE).reshape(-1)
        assert g.size == n and np.array_equal(g["cpi"], np.arange(n)) and np.array_equal(g["range_idx"], 3 * np.arange(n))
        assert set(np.unique(g["snr_db"])) == {{0.5, 1.5}}
        print("GATHER_OK")
    else:
        assert out is None
    # equal shards with known counts (n_cpi % world == 0, e.g. 4096 CPIs on 2/4/8 GPUs): no size exchange, and the SAME
    # return type as the ragged case -- one [world * n][32] tensor in rank order
    e = torch.full((5, 32), rank, dtype=torch.uint8)
    out = shard.gather_detections(e, dst=0, counts=[5] * world)
    if rank == 0:
        assert isinstance(out, torch.Tensor) and tuple(out.shape) == (5 * world, 32)
        assert all(int(out[5 * r, 0]) == r for r in range(world))
        assert out.numpy().view(DET_DTYPE).reshape(-1).size == 5 * world
    else:
        assert out is None
    # preallocated receive tensor, asynchronous form
    recv = torch.empty((5 * world, 32), dtype=torch.uint8) if rank == 0 else None
    work, out = shard.gather_detections(e, dst=0, counts=[5] * world, out=recv, async_op=True)
    work.wait()
    if rank == 0:
        assert out is recv and all(int(recv[5 * r, 0]) == r for r in range(world))
    # an asynchronous ragged gather is refused on every rank BEFORE any collective is issued
    try:
        shard.gather_detections(e, dst=0, counts=[5, 4], async_op=True)
        raise SystemExit("ragged async gather was accepted")
    except ValueError:
        pass
    dist.barrier()                      # still in step: nobody is stuck in a half-issued collective
    dist.barrier(); dist.destroy_process_group()
""")


def test_detection_gather_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GATHER_OK" in r.stdout
