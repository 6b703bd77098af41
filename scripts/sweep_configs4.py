"""BASELINE configs[4]: 2048 subcarriers, 8 x 16 virtual array (128 channels), 65536 CPIs sharded over the
GPUs of one node, detection records gathered to rank 0 over NCCL.

  python scripts/sweep_configs4.py                                  # 1 GPU
  python -m torchrun --nproc-per-node 8 --master-addr 127.0.0.1 ... scripts/sweep_configs4.py --gpus 8

Every rank owns a contiguous shard of 65536 / world CPIs (mimo_ofdm_jrc.shard) and walks it in batches of
--batch CPIs; the symbols of one batch (3 MiB per CPI) are resident in HBM and re-used for every batch of
the shard (65536 distinct CPIs would be 192 GiB of input), the maps stay on the producing GPU, the 32-byte
detection records go to rank 0 (asynchronous, double buffered).  Prints one JSON line on rank 0."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-mimo-ofdm-jrc_b200", "python"))
import numpy as np

CFG = dict(T=8, R=16, S=8, N=2048, IR=1, IA=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--cpis", type=int, default=65536)
    ap.add_argument("--batch", type=int, default=128)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import mimo_ofdm_jrc as jrc
    from mimo_ofdm_jrc import shard, synth

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T, R, S, N, IR, IA = (CFG[k] for k in ("T", "R", "S", "N", "IR", "IA"))
    B = args.batch
    lo, hi = shard.shard_range(args.cpis, rank, world)
    n_batches = (hi - lo + B - 1) // B
    rng = np.random.default_rng(50 + rank)
    tx = synth.tx_symbols(T, S, N)
    r, a, amp = synth.random_scene(rng, B, 3, N, amp_db_span=12.0)
    rx = synth.rx_symbols(tx, R, r, a, amp, snr_db=20.0, rng=rng, chunk=16)
    est = synth.default_estimator_params(N, T * R, IR, IA)
    rc = jrc.radar_chain(N, T, R, S, IR, IA, device=local, estimator=est)
    drx, dtx = torch.from_numpy(rx).to(dev), torch.from_numpy(tx).to(dev)
    dmap = torch.empty((B, rc.Nr, rc.Na), dtype=torch.float32, device=dev)
    ddets = [torch.zeros((B, 32), dtype=torch.uint8, device=dev) for _ in range(2)]
    gbufs = [[torch.empty_like(ddets[0]) for _ in range(world)] for _ in range(2)] if (world > 1 and rank == 0) else [None, None]
    pending = [None, None]
    ext = torch.cuda.ExternalStream(rc.chain.stream, device=dev)
    counts = [B] * world

    def step(k):
        slot = k & 1
        if pending[slot] is not None:
            pending[slot].wait(); pending[slot] = None
        rc.run(drx, dtx, map_out=dmap, dets_out=ddets[slot], cpi0=lo + k * B, sync_inputs=False)
        if world > 1:
            pending[slot], _ = shard.gather_detections(ddets[slot], dst=0, counts=counts, bufs=gbufs[slot], async_op=True)

    def drain():
        for i in range(2):
            if pending[i] is not None:
                pending[i].wait(); pending[i] = None

    torch.cuda.synchronize()
    with torch.cuda.stream(ext):
        for k in range(3):
            step(k)
        drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ext):
        e0.record(ext)
        for k in range(n_batches):
            step(k)
        drain()
        e1.record(ext)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    d = rc.dets_to_numpy(ddets[(n_batches - 1) & 1])
    assert (d["flags"] & 1).mean() > 0.9 and int(d["cpi"][0]) == lo + (n_batches - 1) * B
    if rank == 0:
        done = world * n_batches * B
        rate = done / (ms * 1e-3)
        b_alg = (T + R) * S * N * 8 + rc.Nr * rc.Na * 4 + 32
        print(json.dumps({"workload": "configs[4]: 2048 subcarriers, 8x16 array, CPI-sharded sweep", "n_gpus": world,
                          "cpis": done, "batch_per_gpu": B, "ms": ms, "cpis_per_s": rate,
                          "complex_msps": rate * R * S * N / 1e6, "alg_gbs": rate * b_alg / 1e9,
                          "path": {jrc.PATH_FUSED: "fused", jrc.PATH_TILED: "tiled", jrc.PATH_STAGED: "staged"}[rc.chain.last_path],
                          "scaling": "strong (65536 CPIs in total)", "input": "one resident batch per GPU re-used for the whole shard"}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
