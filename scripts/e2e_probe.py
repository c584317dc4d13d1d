"""Host-buffer (e2e) path probe: PCIe copy bandwidth and fb200_deflate_segments / fb200_inflate_batch wall time.
usage: python scripts/e2e_probe.py [segments]   (FB200_CHUNK_MB selects the pipeline chunk size)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import moonbit_flate_b200 as fb
from helpers import Corpus
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
SEG = 65536
n = nseg * SEG
ctx = fb.Context(0)
h_src = torch.empty(n, dtype=torch.uint8, pin_memory=True)
Corpus().L.fb_corpus_fill(h_src.numpy().ctypes.data, 0, nseg, SEG, 1, -1)
cap = n + n // 8 + nseg * 1024
h_dst = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
h_off = torch.zeros(nseg + 1, dtype=torch.int64, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_ooff = (torch.arange(nseg + 1, dtype=torch.int64) * SEG).pin_memory()
h_olen = torch.zeros(nseg, dtype=torch.int64, pin_memory=True)
h_st = torch.zeros(nseg, dtype=torch.int32, pin_memory=True)
h_eo = torch.zeros(nseg, dtype=torch.int64, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, a, b in (("H2D", d, h_src), ("D2H", h_out, d)):
    a.copy_(b); torch.cuda.synchronize()
    t0 = time.perf_counter(); a.copy_(b, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name}: {n/dt/1e9:.1f} GB/s")
for rep in range(3):
    t0 = time.perf_counter()
    cl = ctx.deflate_segments_ptr(h_src.data_ptr(), n, SEG, h_dst.data_ptr(), cap, h_off.data_ptr())
    t1 = time.perf_counter()
    ctx.inflate_batch_ptr(h_dst.data_ptr(), h_off.data_ptr(), nseg, h_out.data_ptr(), h_ooff.data_ptr(), h_olen.data_ptr(), h_st.data_ptr(), h_eo.data_ptr())
    t2 = time.perf_counter()
    print(f"rep {rep}: deflate {1e3*(t1-t0):.1f} ms  inflate {1e3*(t2-t1):.1f} ms  C={cl}")
assert torch.equal(h_out, h_src) and int(h_st.abs().sum()) == 0
print("ok")
