"""Small device-resident deflate+inflate run for ncu captures: python scripts/prof_run.py [segments] [reps] [klass]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import moonbit_flate_b200 as fb
from helpers import Corpus
nseg = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
klass = int(sys.argv[3]) if len(sys.argv) > 3 else -1
SEG = 65536
ctx = fb.Context(0)
src = torch.from_numpy(Corpus().fill(nseg, SEG, seed=1, klass=klass)).cuda()
n = nseg * SEG
cap = n + n // 8 + nseg * 1024
dst = torch.zeros(cap, dtype=torch.uint8, device="cuda")
off = torch.zeros(nseg + 1, dtype=torch.int64, device="cuda")
out = torch.zeros(n, dtype=torch.uint8, device="cuda")
ooff = torch.arange(nseg + 1, dtype=torch.int64, device="cuda") * SEG
olen = torch.zeros(nseg, dtype=torch.int64, device="cuda")
st = torch.zeros(nseg, dtype=torch.int32, device="cuda")
eo = torch.zeros(nseg, dtype=torch.int64, device="cuda")
for r in range(reps):
    c = ctx.deflate_segments_dev(src.data_ptr(), n, SEG, dst.data_ptr(), cap, off.data_ptr())
    ms_d = ctx.last_stage_ms()
    ctx.inflate_batch_dev(dst.data_ptr(), off.data_ptr(), nseg, out.data_ptr(), ooff.data_ptr(), olen.data_ptr(), st.data_ptr(), eo.data_ptr())
    ms_i = ctx.last_stage_ms(); fbk = int(ctx.last_stats().inflate_fallbacks)
    print(f"rep {r}: C/N={c/n:.4f} " + " ".join(f"{k}={v:.3f}" for k, v in ms_d.items() if k != "inflate") + f" inflate={ms_i['inflate']:.3f} fallbacks={fbk}")
assert torch.equal(out, src) and int(st.abs().sum()) == 0
print("ok")
