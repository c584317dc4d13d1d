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
    nloc_max = (nseg_total + world - 1) // world
    streams = []
    for i in range(first, last):
        streams.append(orc.deflate(corpus.unit(SEG, seed=1, index=i, klass=-1)))
    sizes = torch.zeros(nloc_max, dtype=torch.int64)
    sizes[: len(streams)] = torch.tensor([len(s) for s in streams], dtype=torch.int64)
    payload = torch.from_numpy(np.frombuffer(b"".join(streams) or b"\0", dtype=np.uint8).copy())
    frame = None
    if rank == 0:
        frame = torch.zeros(mg.frame_header_bytes(nloc_max * world) + world * nloc_max * (SEG + SEG // 8 + 1024),
                            dtype=torch.uint8)
    total = mg.assemble_frame(payload, sizes, SEG, rank, world, frame)
    if rank == 0:
        np.save(os.path.join(outdir, "frame.npy"), frame[:total].numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nseg", [(2, 8), (3, 7)])
def test_frame_assembly_gloo(tmp_path, world, nseg):
    from moonbit_flate_b200 import multigpu as mg

    port = _free_port()
    mp.spawn(_worker, args=(world, port, nseg, str(tmp_path)), nprocs=world, join=True)
    frame = torch.from_numpy(np.load(os.path.join(tmp_path, "frame.npy")))
    seg_size, n_slots, sizes, hdr = mg.parse_frame(frame)
    assert seg_size == SEG
    orc, corpus = Oracle(), Corpus()
    # slots: world x ceil(nseg/world); ragged shards leave zero-size slots at the end of a rank's range
    nloc_max = (nseg + world - 1) // world
    assert n_slots == nloc_max * world
    pos = hdr
    seen = 0
    for r in range(world):
        first, last = mg.shard_range(nseg, world, r)
        for k in range(nloc_max):
            sz = int(sizes[r * nloc_max + k])
            if first + k < last:
                want = orc.deflate(corpus.unit(SEG, seed=1, index=first + k, klass=-1))
                assert sz == len(want)
                assert frame[pos: pos + sz].numpy().tobytes() == want, (r, k)
                seen += 1
            else:
                assert sz == 0
            pos += sz
    assert seen == nseg and pos == frame.numel()


def test_shard_range_partitions():
    from moonbit_flate_b200 import multigpu as mg

    for nseg in (0, 1, 7, 16384, 131072):
        for world in (1, 2, 3, 4, 8):
            r = [mg.shard_range(nseg, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == nseg
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
