"""Phase cycle counts of the CTA-per-stream inflate kernel (variant build with -DFB_CTA_PROF=1)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import moonbit_flate_b200 as fb
from helpers import Corpus, Oracle
ctx = fb.Context(0)
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 16
data = Corpus().fill((mib << 20) // 65536, 65536, seed=1, klass=0).tobytes()
comp = Oracle().deflate(data)
out = np.zeros(8, np.uint64)
f = fb._lib.fb200_debug_cta_prof
f.argtypes = [C.c_void_p, C.c_int]
for rep in range(3):
    f(out.ctypes.data, 1)
    t0 = time.perf_counter()
    got, st, eo, cons = ctx.inflate(comp, len(data) + 1)
    t1 = time.perf_counter()
    ms = ctx.last_stage_ms()["inflate"]
    f(out.ctypes.data, 0)
    tot = out.sum()
    names = ["header", "rounds", "scan+write pass", "replay", "copy-out", "tail"]
    print(f"rep {rep}: wall {1e3*(t1-t0):.1f} ms, kernel stage {ms:.2f} ms, cycles {tot/1e6:.1f} M: " +
          ", ".join(f"{n} {100*out[i]/tot:.1f}%" for i, n in enumerate(names)), "ok" if got == data and st == 0 else "MISMATCH")
