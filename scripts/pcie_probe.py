"""Concurrent pinned H2D / D2H bandwidth of this process' GPU (argv[1]); run one per GPU at the same time to see
what the platform gives several GPUs at once.  usage: python scripts/pcie_probe.py <gpu> [seconds]"""
import sys, time, torch
g = int(sys.argv[1]); secs = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
torch.cuda.set_device(g)
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.zero_()
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name in ("h2d", "d2h", "both"):
    torch.cuda.synchronize(); t0 = time.perf_counter(); k = 0
    while time.perf_counter() - t0 < secs:
        if name in ("h2d", "both"):
            with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
        if name in ("d2h", "both"):
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize(); k += 1
    dt = time.perf_counter() - t0
    print(f"gpu{g} {name}: {k * n / dt / 1e9:.1f} GB/s per direction", flush=True)
