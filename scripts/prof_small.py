"""Small-stream deflate run for ncu captures: python scripts/prof_small.py [streams]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import moonbit_flate_b200 as fb
from helpers import Corpus
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ctx = fb.Context(0)
src, off = Corpus().fill_var(ns, seed=1)
n = int(off[-1])
d_src = torch.from_numpy(src).cuda(); d_off = torch.from_numpy(off.astype(np.int64)).cuda()
cap = n + n // 8 + ns * 64
d_dst = torch.empty(cap, dtype=torch.uint8, device="cuda"); d_doff = torch.zeros(ns + 1, dtype=torch.int64, device="cuda")
for r in range(2):
    c = ctx.deflate_streams_dev(d_src.data_ptr(), d_off.data_ptr(), ns, n, d_dst.data_ptr(), cap, d_doff.data_ptr())
    print({k: round(v, 3) for k, v in ctx.last_stage_ms().items()})
