"""Small-stream inflate run (device resident): python scripts/prof_small_inflate.py [streams] ; FB200_TRACE=1 prints decode rounds"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import moonbit_flate_b200 as fb
from helpers import Corpus
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
ctx = fb.Context(0)
src, off = Corpus().fill_var(ns, seed=1)
n = int(off[-1])
d_src = torch.from_numpy(src).cuda(); d_off = torch.from_numpy(off.astype(np.int64)).cuda()
cap = n + n // 8 + ns * 64
d_dst = torch.empty(cap, dtype=torch.uint8, device="cuda"); d_doff = torch.zeros(ns + 1, dtype=torch.int64, device="cuda")
c = ctx.deflate_streams_dev(d_src.data_ptr(), d_off.data_ptr(), ns, n, d_dst.data_ptr(), cap, d_doff.data_ptr())
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
d_olen = torch.zeros(ns, dtype=torch.int64, device="cuda"); d_st = torch.zeros(ns, dtype=torch.int32, device="cuda"); d_eo = torch.zeros(ns, dtype=torch.int64, device="cuda")
for r in range(3):
    ctx.inflate_batch_dev(d_dst.data_ptr(), d_doff.data_ptr(), ns, d_out.data_ptr(), d_off.data_ptr(), d_olen.data_ptr(), d_st.data_ptr(), d_eo.data_ptr())
    ms = ctx.last_stage_ms()["inflate"]
    print(f"rep {r}: {ns} streams, {n/1e6:.0f} MB: inflate {ms:.3f} ms = {n/ms/1e6:.1f} GB/s, fallbacks {int(ctx.last_stats().inflate_fallbacks)}")
assert torch.equal(d_out, d_src)
