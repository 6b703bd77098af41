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


def gather_detections(dets, dst: int = 0, group=None, counts=None, out=None, async_op=False):
    """dets: uint8 tensor [n_local][32] (device tensor under NCCL, CPU tensor under gloo).
    Returns on rank dst ONE tensor [sum(counts)][32]: the records of all ranks in rank order (ragged shards
    allowed); None on the other ranks.

    counts: per-rank record counts when the caller already knows them (contiguous shards of a known batch:
    shard_range) -- skips the size exchange and its host synchronisation, so the gather is a single asynchronous
    NCCL call on the current stream.  out: optional preallocated receive tensor [world * max(counts)][32] on
    rank dst (keeps the step allocation-free; with equal shards the result IS this tensor).
    async_op=True (needs counts and equal shard sizes, checked before anything is launched): returns
    (work, result) without making the current stream wait for the collective; call work.wait() before the records
    (or the send buffer) are touched again."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return (None, dets) if async_op else dets
    if async_op:
        if counts is None:
            raise ValueError("async gather needs the per-rank counts")
        if len(set(counts)) != 1:
            raise ValueError("async gather needs equal shard sizes")      # raised on EVERY rank, before the collective
    if counts is None:
        n_local = torch.tensor([dets.shape[0]], dtype=torch.int64, device=dets.device)
        cts = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(cts, n_local, group=group)
        counts = [int(c.item()) for c in cts]
    if len(counts) != world:
        raise ValueError("counts must have one entry per rank")
    n_max = max(counts)
    padded = dets
    if dets.shape[0] < n_max:
        padded = torch.zeros((n_max, dets.shape[1]), dtype=dets.dtype, device=dets.device)
        padded[: dets.shape[0]] = dets
    bufs = None
    if rank == dst:
        if out is None:
            out = torch.empty((world * n_max, dets.shape[1]), dtype=dets.dtype, device=dets.device)
        if out.shape[0] != world * n_max:
            raise ValueError("out must be [world * max(counts)][32]")
        bufs = [out[r * n_max:(r + 1) * n_max] for r in range(world)]      # views: the gather fills `out` in place
    work = dist.gather(padded.contiguous(), bufs, dst=dst, group=group, async_op=async_op)
    res = None
    if rank == dst:
        res = out if all(c == n_max for c in counts) else torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
    return (work, res) if async_op else res


class PeerRecordTable:
    """Detection table of a multi-GPU job in the memory of rank `dst`, written by every rank's kernels directly (CUDA IPC
    mapping, NVLink peer stores) instead of being gathered: `ptr(i)` is the address this rank passes as `dets` for its i-th
    record; after `complete()` (stream synchronised by the caller, then a barrier) rank `dst` reads `records()`.
    torch.distributed is only used to hand the 64-byte IPC handle around and for the barrier."""

    def __init__(self, n_per_rank, device, dst=0, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import cabi
        self._cabi, self._dist, self._group = cabi, dist, group
        self.world, self.rank, self.dst = dist.get_world_size(group), dist.get_rank(group), dst
        self.n, self.device = int(n_per_rank), device
        lib = cabi.load()
        self._base = C.c_void_p()
        self._owner = self.rank == dst
        h = torch.zeros(64, dtype=torch.uint8, device=torch.device("cuda", device))
        if self._owner:
            cabi.check(lib.jrc_dev_alloc(device, self.world * self.n * 32, C.byref(self._base)))
            raw = (C.c_ubyte * 64)()
            cabi.check(lib.jrc_ipc_export(self._base, raw))
            h.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
        dist.broadcast(h, src=dst, group=group)
        if not self._owner:
            raw = (C.c_ubyte * 64).from_buffer_copy(bytes(h.cpu().numpy().tobytes()))
            cabi.check(lib.jrc_ipc_open(raw, device, C.byref(self._base)))

    def ptr(self, i=0):
        return self._base.value + (self.rank * self.n + i) * 32

    def push(self, chain, dets, i=0):
        """Queues ONE peer copy of this rank's records `dets` ([n][32] uint8, device) to slot i.. of its slice on the chain's
        stream: the drain-time alternative to handing `ptr()` to the kernels."""
        n = int(dets.shape[0])
        if i < 0 or i + n > self.n:
            raise ValueError("records do not fit this rank's slice")
        chain.copy_async(self.ptr(i), dets.data_ptr(), n * 32)

    def complete(self):
        self._dist.barrier(group=self._group)

    def records(self):
        """rank dst: uint8 tensor [world * n][32] (a copy); None elsewhere"""
        if not self._owner:
            return None
        import torch
        out = torch.empty((self.world * self.n, 32), dtype=torch.uint8, device=torch.device("cuda", self.device))
        torch.cuda.synchronize(self.device)
        self._cabi.check(self._cabi.load().jrc_dev_copy(out.data_ptr(), self._base, out.numel()))
        return out

    def close(self):
        lib = self._cabi.load()
        if self._base.value:
            if self._owner:
                self._dist.barrier(group=self._group)       # nobody maps it any more
                lib.jrc_dev_free(self._base)
            else:
                lib.jrc_ipc_close(self._base)
                self._dist.barrier(group=self._group)
            self._base.value = None
