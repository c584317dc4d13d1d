"""One long stream through the host-buffer API (the Writer path): wall time, GB/s, parity with the oracle.
usage: python scripts/long_stream_probe.py [MiB] [klass]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import moonbit_flate_b200 as fb
from helpers import Corpus, Oracle
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
klass = int(sys.argv[2]) if len(sys.argv) > 2 else -1
n = mib << 20
src = Corpus().fill(n // 65536, 65536, seed=3, klass=klass)
off = np.array([0, n], dtype=np.uint64)
ctx = fb.Context(0)
t0 = time.perf_counter(); want = np.frombuffer(Oracle().deflate(src), dtype=np.uint8); t1 = time.perf_counter()
print(f"oracle (1 thread): {1e3*(t1-t0):.0f} ms")
for rep in range(3):
    t0 = time.perf_counter(); comp, doff = ctx.deflate_streams(src, off); t1 = time.perf_counter()
    same = comp.size == want.size and np.array_equal(comp, want)
    first = -1 if same else int(np.argmax(comp[:min(comp.size, want.size)] != want[:min(comp.size, want.size)]))
    print(f"rep {rep}: {mib} MiB single stream: deflate {1e3*(t1-t0):.1f} ms ({n/(t1-t0)/1e9:.2f} GB/s), C/N {comp.size/n:.4f}, identical to the oracle: {same} (sizes {comp.size} / {want.size}, first difference at {first})")
if os.environ.get("PROBE_INFLATE", "1") == "1":
    t1 = time.perf_counter(); out, olen, st, eo, cons = ctx.inflate_batch(comp, doff, off); t2 = time.perf_counter()
    print("fallbacks", int(ctx.last_stats().inflate_fallbacks), "consumed", int(cons[0]), "of", comp.size)
    print(f"inflate {1e3*(t2-t1):.1f} ms ({n/(t2-t1)/1e9:.2f} GB/s), status {int(st[0])}, len {int(olen[0])}, round trip {np.array_equal(out, src)}")
