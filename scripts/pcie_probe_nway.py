"""N-way concurrent pinned H2D / D2H bandwidth: one process per GPU (torchrun), every rank copies at the same time
between its own pinned host buffers and its GPU; rank 0 prints the per-rank and aggregate GB/s per direction for
H2D alone, D2H alone and both together.  The aggregate `both` figure bounds the end-to-end (host buffer) codec
throughput of the box: a step moves N + C bytes in each direction per GPU.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/pcie_probe_nway.py > profiles/rXX_pcie_8way.json
"""
import json, os, sys, time
import torch, torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
try:
    import pynvml
    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(lr))
except Exception:
    pass
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.zero_()
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2.zero_()
d = torch.empty(n, dtype=torch.uint8, device=dev); d2 = torch.zeros(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for name in ("h2d", "d2h", "both"):
    for _ in range(2):  # warm-up
        d.copy_(h, non_blocking=True); h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter(); k = 0
    while k < 12:
        if name in ("h2d", "both"):
            with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
        if name in ("d2h", "both"):
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize(); k += 1
    dt = time.perf_counter() - t0
    mine = torch.tensor([k * n / dt / 1e9], dtype=torch.float64, device=dev)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(allv, mine)
    else:
        allv = [mine]
    res[name] = {"per_rank_gbs_per_direction": [round(float(v.item()), 1) for v in allv],
                 "aggregate_gbs_per_direction": round(sum(float(v.item()) for v in allv), 1)}
if rank == 0:
    # codec bound: per step and GPU N + C bytes travel in EACH direction (deflate: N in, C out; inflate: C in, N out)
    c_over_n = 0.4723
    agg = res["both"]["aggregate_gbs_per_direction"]
    bound = 2.0 / (1.0 + c_over_n) * agg
    print(json.dumps({"what": f"{world} ranks, 1 GiB pinned copies, all ranks at once", "world": world, **res,
                      "e2e_bound_gbs": round(bound, 1),
                      "e2e_bound_note": "2 N uncompressed bytes per step against (N + C) bytes per direction, C/N = 0.4723, both directions busy"},
                     indent=1))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
