"""Multi-GPU sharding of the deflate / inflate path.

One process per GPU (``torch.distributed``, backend ``nccl`` on GPUs, ``gloo`` on
CPU for the host-logic tests).  Segments are independent streams, so the data
path has no collective: rank r owns the contiguous segment range
``shard_range(nseg, world, r)`` and compresses / decompresses it locally.  The
one exchange step is the frame assembly:

1. every rank all-gathers its per-segment compressed sizes (int64, ``nseg_r``
   entries, ragged shards padded for the collective and compacted afterwards, so
   the header holds the true ``nseg`` entries in segment order; NCCL all-gather
   over NVLink),
2. from the gathered sizes every rank derives its payload offset in the frame
   (exclusive scan), and
3. the payload of every rank lands in GPU 0's frame buffer at that offset.
   On GPUs this is ``PeerFrame``: the frame buffer is exported with CUDA IPC,
   every rank maps it and puts its payload there with an asynchronous peer
   copy (copy engines over NVLink, no SM time), so the transfer runs beside the
   rank's next kernels; ``assemble_frame`` is the same exchange with NCCL /
   gloo send/recv (used by the CPU tests and when IPC is unavailable).

The decompress side reads the frame back: ``PeerFrame.get`` (C ABI ``fb200_mg_get``:
a kernel reads the sizes from the header, one peer copy moves the range) or
``scatter_frame`` (send/recv) hand every rank the compressed streams of its
segment range, which it inflates locally.

The frame is an addition -- the reference defines no container:
``magic "FB2\\0" u32 | seg_size u32 | nseg u64 | comp_size u32[nseg] | streams``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

FRAME_MAGIC = 0x00324246


def shard_range(nseg: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous segment range [first, last) owned by `rank`."""
    base, rem = divmod(nseg, world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def frame_header_bytes(nseg: int) -> int:
    return 16 + 4 * nseg


def gather_sizes(local_sizes: torch.Tensor, world: int, nseg_total: Optional[int] = None):
    """All-gather of per-segment compressed sizes.  local_sizes: int64[nseg_r], the sizes of the segments
    ``shard_range(nseg_total, world, rank)`` owns (nseg_total defaults to world * nseg_r: equal shards).  Ragged
    shards are padded to the longest one for the collective and compacted afterwards.  Returns
    (all_sizes int64[nseg_total] in segment order, totals int64[world] = payload bytes per rank)."""
    n_loc = int(local_sizes.numel())
    if nseg_total is None:
        nseg_total = n_loc * world
    n_max = (nseg_total + world - 1) // world if world else 0
    assert n_loc <= n_max, "shard longer than shard_range allows"
    padded = torch.zeros(n_max, dtype=torch.int64, device=local_sizes.device)
    padded[:n_loc] = local_sizes.to(torch.int64)
    out = torch.empty(world * n_max, dtype=torch.int64, device=local_sizes.device)
    if world == 1:
        out.copy_(padded)
    else:
        dist.all_gather_into_tensor(out, padded)
    out = out.view(world, n_max)
    totals = out.sum(dim=1)
    if nseg_total == world * n_max:
        return out.reshape(-1), totals
    parts = []
    for r in range(world):
        first, last = shard_range(nseg_total, world, r)
        parts.append(out[r, : last - first])
    return torch.cat(parts), totals


def payload_offsets(totals: torch.Tensor) -> torch.Tensor:
    """Exclusive offsets of the ranks' payloads behind the frame header (int64[world])."""
    return torch.cumsum(totals, 0) - totals


def _frame_header(all_sizes: torch.Tensor, seg_size: int, device) -> torch.Tensor:
    """magic | seg_size | nseg (u64) | comp_size u32[nseg] as a uint8 tensor on `device`."""
    nseg = int(all_sizes.numel())
    assert nseg == 0 or int(all_sizes.max()) < 2 ** 31, "a compressed segment does not fit the u32 size field"
    head = torch.tensor([FRAME_MAGIC, seg_size, nseg & 0xFFFFFFFF, nseg >> 32], dtype=torch.int64).to(torch.int32)
    return torch.cat([head.view(torch.uint8).to(device), all_sizes.to(torch.int32).contiguous().view(torch.uint8).to(device)])


def assemble_frame(payload: torch.Tensor, local_sizes: torch.Tensor, seg_size: int, rank: int, world: int,
                   frame: Optional[torch.Tensor], nseg_total: Optional[int] = None) -> int:
    """Build the framed output on rank 0 (send/recv transport: gloo on CPU, NCCL when CUDA IPC is unavailable).

    payload: uint8[>= sum(local_sizes)] this rank's compacted streams (device on GPU runs);
    frame:   rank 0 only, uint8 buffer large enough for header + all payloads.
    Returns the total frame length (valid on every rank)."""
    all_sizes, totals = gather_sizes(local_sizes, world, nseg_total)
    totals_h = totals.cpu().tolist()
    offs_h = payload_offsets(totals).cpu().tolist()
    hdr = frame_header_bytes(int(all_sizes.numel()))
    total = hdr + int(sum(totals_h))
    my_len = int(totals_h[rank])
    if rank == 0:
        assert frame is not None and frame.numel() >= total
        frame[:hdr].copy_(_frame_header(all_sizes, seg_size, frame.device))
        frame[hdr + offs_h[0]: hdr + offs_h[0] + my_len].copy_(payload[:my_len])
        if world > 1:
            ops = []
            for r in range(1, world):
                if totals_h[r]:
                    ops.append(dist.P2POp(dist.irecv, frame[hdr + offs_h[r]: hdr + offs_h[r] + int(totals_h[r])], r))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
    elif my_len:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, payload[:my_len], 0)]):
            w.wait()
    return total


def scatter_frame(frame: Optional[torch.Tensor], rank: int, world: int, device=None):
    """Decompress side with the send/recv transport: rank 0 holds the frame, every rank receives the compressed
    streams of the segments ``shard_range(nseg, world, rank)`` owns.  Returns (seg_size, nseg, first, sizes int64
    (cpu), payload uint8) -- sizes and payload of this rank's range."""
    meta = [None]
    if rank == 0:
        seg_size, nseg, sizes, hdr = parse_frame(frame)
        meta[0] = (seg_size, nseg, sizes.tolist(), hdr)
    if world > 1:
        dist.broadcast_object_list(meta, src=0)
    seg_size, nseg, sizes_l, hdr = meta[0]
    sizes = torch.tensor(sizes_l, dtype=torch.int64)
    offs = torch.cumsum(sizes, 0) - sizes
    first, last = shard_range(nseg, world, rank)
    mine = int(sizes[first:last].sum())
    dev = device if device is not None else (frame.device if frame is not None else torch.device("cpu"))
    payload = torch.empty(mine, dtype=torch.uint8, device=dev)
    if rank == 0:
        payload.copy_(frame[hdr + int(offs[first]) if last > first else hdr: (hdr + int(offs[first]) if last > first else hdr) + mine])
        ops = []
        for r in range(1, world):
            f, l = shard_range(nseg, world, r)
            n = int(sizes[f:l].sum())
            if n:
                ops.append(dist.P2POp(dist.isend, frame[hdr + int(offs[f]): hdr + int(offs[f]) + n].contiguous(), r))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
    elif mine:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.irecv, payload, 0)]):
            w.wait()
    return seg_size, nseg, first, sizes[first:last], payload


class _DevPtr:
    """Raw device memory as a __cuda_array_interface__ object (for torch.as_tensor)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


class PeerFrame:
    """The frame buffer on rank 0's GPU, mapped into every rank of the box through CUDA IPC
    (fb200_mg_frame_alloc / _open).  ``put`` = all-gather of the segment sizes (NCCL) + header (rank 0) +
    an asynchronous peer copy of this rank's payload to its offset; ``wait`` = the copies of every rank
    have landed.  Nothing here uses SMs besides the all-gather, so the transfer overlaps the kernels the
    caller launches between ``put`` and ``wait``."""

    def __init__(self, ctx, rank: int, world: int, capacity: int, device: torch.device):
        self.ctx, self.rank, self.world, self.capacity, self.device = ctx, rank, world, int(capacity), device
        self.ptr = 0
        ok = 1
        handle = [None]
        if rank == 0:
            try:
                self.ptr, handle[0] = ctx.mg_frame_alloc(self.capacity)
            except Exception as e:  # noqa: BLE001 -- reported to every rank below
                handle[0] = None
                self.error = str(e)
        if world > 1:
            dist.broadcast_object_list(handle, src=0)
        if handle[0] is None:
            ok = 0
        elif rank != 0:
            try:
                self.ptr = ctx.mg_frame_open(handle[0])
            except Exception as e:  # noqa: BLE001
                ok = 0
                self.error = str(e)
        if not hasattr(self, "error"):
            self.error = None
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.available = bool(int(flag.item()))
        self._holder = None
        self._hdr_flag = None
        self.length = 0
        self.view: Optional[torch.Tensor] = None
        if not self.available:
            self.close()
        elif rank == 0:
            self._holder = _DevPtr(self.ptr, self.capacity)
            self.view = torch.as_tensor(self._holder, device=device)

    def put(self, payload: torch.Tensor, local_sizes: torch.Tensor, seg_size: int, nseg_total: Optional[int] = None) -> int:
        """payload: this rank's compacted streams (device).  Returns the total frame length."""
        all_sizes, totals = gather_sizes(local_sizes, self.world, nseg_total)
        totals_h = totals.cpu().tolist()
        offs_h = payload_offsets(totals).cpu().tolist()
        hdr = frame_header_bytes(int(all_sizes.numel()))
        total = hdr + int(sum(totals_h))
        if total > self.capacity:
            raise ValueError(f"frame capacity {self.capacity} < {total}")
        if self.rank == 0:
            self.view[:hdr].copy_(_frame_header(all_sizes, seg_size, self.device))
        self.ctx.mg_put(self.ptr, hdr + int(offs_h[self.rank]), payload.data_ptr(), int(totals_h[self.rank]))
        self.length = total
        return total

    def get(self, first: int, count: int, d_comp: torch.Tensor, d_comp_off: torch.Tensor):
        """Decompress side: the compressed streams of segments [first, first + count) travel from the frame on
        rank 0's GPU into d_comp (one peer copy over NVLink), their offsets into d_comp_off (count + 1 int64,
        starting at 0): ready for fb200_inflate_batch_dev.  Returns (seg_size, nseg_total, bytes)."""
        return self.ctx.mg_get(self.ptr, self.length, first, count, d_comp.data_ptr(), int(d_comp.numel()),
                               d_comp_off.data_ptr())

    def get_begin(self, first: int, count: int, d_comp: torch.Tensor, d_comp_off: torch.Tensor):
        """``get`` without the final wait: the offsets are in d_comp_off and the payload is on its way when this
        returns; ``ctx.mg_wait()`` blocks until it has arrived.  Whatever the caller runs in between (the deflate of
        its next batch) overlaps the transfer."""
        return self.ctx.mg_get_async(self.ptr, self.length, first, count, d_comp.data_ptr(), int(d_comp.numel()),
                                     d_comp_off.data_ptr())

    def wait(self):
        """This rank's payload has landed in the frame and the header (written by rank 0) is complete: everything
        this rank reads back with ``get`` is there.  No barrier: a rank does not wait for the payloads of the others,
        so while late ranks still put (into GPU 0) early ranks already get (out of GPU 0) -- NVLink is full duplex."""
        self.ctx.mg_wait()
        if self.world > 1:
            if self._hdr_flag is None:
                self._hdr_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
            dist.broadcast(self._hdr_flag, src=0)  # ordered behind rank 0's header write on its stream
            torch.cuda.current_stream(self.device).synchronize()

    def wait_all(self):
        """Every rank's payload is in the frame (the assembled frame is complete)."""
        self.ctx.mg_wait()
        if self.world > 1:
            dist.barrier()

    def close(self):
        if self.ptr:
            self.view = None
            self._holder = None
            try:
                self.ctx.mg_frame_close(self.ptr, self.rank == 0)
            finally:
                self.ptr = 0


def parse_frame(frame: torch.Tensor):
    """-> (seg_size, nseg, sizes int64[nseg] (cpu), payload offset)"""
    head = frame[:16].cpu().view(torch.int32).to(torch.int64)
    assert int(head[0]) == FRAME_MAGIC
    seg_size = int(head[1]) & 0xFFFFFFFF
    nseg = (int(head[2]) & 0xFFFFFFFF) | (int(head[3]) << 32)
    hdr = frame_header_bytes(nseg)
    sizes = frame[16:hdr].cpu().view(torch.int32).to(torch.int64)
    return seg_size, nseg, sizes, hdr
