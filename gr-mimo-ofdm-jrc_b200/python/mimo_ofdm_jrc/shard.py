"""CPI sharding across GPUs.

Every CPI is independent (lib/mimo_ofdm_radar_impl.cc:243-244 clears all per-frame state), so a
batch is split into contiguous blocks, one per rank, with no collective on the data path.  The
only exchange is the gather of the 32-byte detection records to the host rank (rank 0).
torch.distributed is plumbing here: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations


def shard_range(n_cpi: int, rank: int, world: int):
    """Contiguous block [lo, hi) of rank; the first n_cpi % world ranks get one extra CPI.
    Contiguity keeps the output order and the background sliding window of
    lib/mimo_ofdm_radar_impl.cc:276-300 meaningful within a shard."""
    if world < 1 or not (0 <= rank < world) or n_cpi < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_cpi, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_detections(dets, dst: int = 0, group=None, counts=None, bufs=None, async_op=False):
    """dets: uint8 tensor [n_local][32] (device tensor under NCCL, CPU tensor under gloo).
    Returns on rank dst the concatenation over ranks in rank order (ragged shards allowed),
    None elsewhere.

    counts: per-rank record counts when the caller already knows them (contiguous shards of a known
    batch: shard_range) -- skips the size exchange and its host synchronisation, so the gather is
    a single asynchronous NCCL call on the current stream.  bufs: optional preallocated receive
    buffers on rank dst (list of world tensors [max(counts)][32]) to keep the step allocation-free.
    async_op=True (needs counts): returns (work, result) without making the current stream wait for
    the collective, so the next step's kernel overlaps the gather; call work.wait() before the
    records (or the send buffer) are touched again."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return dets
    if counts is None:
        n_local = torch.tensor([dets.shape[0]], dtype=torch.int64, device=dets.device)
        cts = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(cts, n_local, group=group)
        counts = [int(c.item()) for c in cts]
    n_max = max(counts)
    padded = dets
    if dets.shape[0] < n_max:
        padded = torch.zeros((n_max, dets.shape[1]), dtype=dets.dtype, device=dets.device)
        padded[: dets.shape[0]] = dets
    if rank == dst and bufs is None:
        bufs = [torch.empty_like(padded) for _ in range(world)]
    work = dist.gather(padded.contiguous(), bufs if rank == dst else None, dst=dst, group=group, async_op=async_op)
    if rank != dst:
        res = None
    elif all(c == n_max for c in counts):
        res = bufs             # equal shards: the per-rank blocks, in rank order, no copy
    else:
        if async_op:
            raise ValueError("async gather needs equal shard sizes")
        res = torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
    return (work, res) if async_op else res
