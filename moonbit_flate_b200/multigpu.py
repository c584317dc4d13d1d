"""Multi-GPU sharding of the deflate / inflate path.

One process per GPU (``torch.distributed``, backend ``nccl`` on GPUs, ``gloo`` on
CPU for the host-logic tests).  Segments are independent streams, so the data
path has no collective: rank r owns the contiguous segment range
``shard_range(nseg, world, r)`` and compresses / decompresses it locally.  The
one exchange step is the frame assembly:

1. every rank all-gathers its per-segment compressed sizes (int64, ``nseg_r``
   entries; NCCL all-gather over NVLink),
2. from the gathered sizes every rank derives its payload offset in the frame
   (exclusive scan), and
3. the payload of every rank lands in GPU 0's frame buffer at that offset.
   On GPUs this is ``PeerFrame``: the frame buffer is exported with CUDA IPC,
   every rank maps it and puts its payload there with an asynchronous peer
   copy (copy engines over NVLink, no SM time), so the transfer runs beside the
   rank's next kernels; ``assemble_frame`` is the same exchange with NCCL /
   gloo send/recv (used by the CPU tests and when IPC is unavailable).

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


def gather_sizes(local_sizes: torch.Tensor, world: int) -> torch.Tensor:
    """All-gather of per-segment compressed sizes.  local_sizes: int64[nseg_local]
    (equal length on every rank; pad with zeros if the split is ragged).  Returns
    int64[world, nseg_local]."""
    out = torch.empty(world * local_sizes.numel(), dtype=local_sizes.dtype, device=local_sizes.device)
    if world == 1:
        out.copy_(local_sizes)
    else:
        dist.all_gather_into_tensor(out, local_sizes.contiguous())
    return out.view(world, -1)


def payload_offsets(all_sizes: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """From the gathered sizes: per-rank payload totals and exclusive offsets (int64[world] each)."""
    totals = all_sizes.sum(dim=1)
    offs = torch.cumsum(totals, 0) - totals
    return totals, offs


def assemble_frame(payload: torch.Tensor, local_sizes: torch.Tensor, seg_size: int, rank: int, world: int,
                   frame: Optional[torch.Tensor]) -> int:
    """Build the framed output on rank 0.

    payload: uint8[>= sum(local_sizes)] this rank's compacted streams (device on GPU runs);
    frame:   rank 0 only, uint8 buffer large enough for header + all payloads.
    Returns the total frame length (valid on every rank)."""
    all_sizes = gather_sizes(local_sizes, world)
    totals, offs = payload_offsets(all_sizes)
    totals_h = totals.cpu().tolist()
    offs_h = offs.cpu().tolist()
    nseg = all_sizes.numel()
    hdr = frame_header_bytes(nseg)
    total = hdr + int(sum(totals_h))
    my_len = int(totals_h[rank])
    if rank == 0:
        assert frame is not None and frame.numel() >= total
        head = torch.tensor([FRAME_MAGIC, seg_size, nseg & 0xFFFFFFFF, nseg >> 32], dtype=torch.int64).to(torch.int32)
        frame[:16].copy_(head.view(torch.uint8).to(frame.device))
        frame[16:hdr].copy_(all_sizes.reshape(-1).to(torch.int32).view(torch.uint8))
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
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.available = bool(int(flag.item()))
        self._holder = None
        self.view: Optional[torch.Tensor] = None
        if not self.available:
            self.close()
        elif rank == 0:
            self._holder = _DevPtr(self.ptr, self.capacity)
            self.view = torch.as_tensor(self._holder, device=device)

    def put(self, payload: torch.Tensor, local_sizes: torch.Tensor, seg_size: int) -> int:
        """payload: this rank's compacted streams (device).  Returns the total frame length."""
        all_sizes = gather_sizes(local_sizes, self.world)
        totals, offs = payload_offsets(all_sizes)
        totals_h = totals.cpu().tolist()
        offs_h = offs.cpu().tolist()
        nseg = all_sizes.numel()
        hdr = frame_header_bytes(nseg)
        total = hdr + int(sum(totals_h))
        if total > self.capacity:
            raise ValueError(f"frame capacity {self.capacity} < {total}")
        if self.rank == 0:
            head = torch.tensor([FRAME_MAGIC, seg_size, nseg & 0xFFFFFFFF, nseg >> 32], dtype=torch.int64).to(torch.int32)
            self.view[:16].copy_(head.view(torch.uint8).to(self.device))
            self.view[16:hdr].copy_(all_sizes.reshape(-1).to(torch.int32).view(torch.uint8))
        self.ctx.mg_put(self.ptr, hdr + int(offs_h[self.rank]), payload.data_ptr(), int(totals_h[self.rank]))
        return total

    def wait(self):
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
