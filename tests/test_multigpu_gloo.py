"""CPU suite, part 4: the N>1 host logic (segment sharding, all-gather of per-segment sizes,
payload offsets, frame assembly on rank 0) with world_size 2 and 3 over gloo.  The per-rank
"compressed payload" here is the ORACLE's output (test infrastructure standing in for the GPU
kernels, which need a device); the collective / framing code under test is the product's
moonbit_flate_b200.multigpu, unchanged."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, Corpus, Oracle

SEG = 65536


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nseg_total, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from moonbit_flate_b200 import multigpu as mg

    orc, corpus = Oracle(), Corpus()
    first, last = mg.shard_range(nseg_total, world, rank)
    streams = []
    for i in range(first, last):
        streams.append(orc.deflate(corpus.unit(SEG, seed=1, index=i, klass=-1)))
    sizes = torch.tensor([len(s) for s in streams], dtype=torch.int64)
    payload = torch.from_numpy(np.frombuffer(b"".join(streams) or b"\0", dtype=np.uint8).copy())
    frame = None
    if rank == 0:
        frame = torch.zeros(mg.frame_header_bytes(nseg_total) + nseg_total * (SEG + SEG // 8 + 1024), dtype=torch.uint8)
    total = mg.assemble_frame(payload, sizes, SEG, rank, world, frame, nseg_total=nseg_total)
    if rank == 0:
        np.save(os.path.join(outdir, "frame.npy"), frame[:total].numpy())
    # decompress side: every rank gets the streams of its segment range back out of the frame and inflates them
    seg_size, nseg, f2, my_sizes, my_payload = mg.scatter_frame(frame[:total] if rank == 0 else None, rank, world)
    ok = seg_size == SEG and nseg == nseg_total and f2 == first and my_sizes.tolist() == sizes.tolist()
    pos = 0
    for k, sz in enumerate(my_sizes.tolist()):
        st, out, _, cons = orc.inflate(my_payload[pos: pos + sz].numpy().tobytes(), SEG + 1)
        ok = ok and st == 0 and cons == sz and out == corpus.unit(SEG, seed=1, index=first + k, klass=-1)
        pos += sz
    ok = ok and pos == my_payload.numel()
    open(os.path.join(outdir, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nseg", [(2, 8), (3, 7), (3, 2)])
def test_frame_assembly_and_scatter_gloo(tmp_path, world, nseg):
    """Frame assembly on rank 0 from ragged shards (7 segments over 3 ranks; 2 over 3: one rank owns nothing):
    the header carries the true segment count and the sizes in segment order (the documented format), the payload
    is the concatenation of the streams; then the way back: every rank receives its range and inflates it."""
    from moonbit_flate_b200 import multigpu as mg

    port = _free_port()
    mp.spawn(_worker, args=(world, port, nseg, str(tmp_path)), nprocs=world, join=True)
    frame = torch.from_numpy(np.load(os.path.join(tmp_path, "frame.npy")))
    seg_size, n, sizes, hdr = mg.parse_frame(frame)
    assert seg_size == SEG and n == nseg and sizes.numel() == nseg and hdr == 16 + 4 * nseg
    orc, corpus = Oracle(), Corpus()
    pos = hdr
    for i in range(nseg):
        want = orc.deflate(corpus.unit(SEG, seed=1, index=i, klass=-1))
        assert int(sizes[i]) == len(want)
        assert frame[pos: pos + len(want)].numpy().tobytes() == want, i
        pos += len(want)
    assert pos == frame.numel()
    for r in range(world):
        assert open(os.path.join(tmp_path, f"ok{r}")).read() == "1", f"rank {r}: scatter / inflate from the frame"


def test_shard_range_partitions():
    from moonbit_flate_b200 import multigpu as mg

    for nseg in (0, 1, 7, 16384, 131072):
        for world in (1, 2, 3, 4, 8):
            r = [mg.shard_range(nseg, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == nseg
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
