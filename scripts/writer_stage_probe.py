"""Stage times of ONE stream through one deflate call (block-parallel parse): python scripts/writer_stage_probe.py [MiB ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import moonbit_flate_b200 as fb
from helpers import Corpus

ctx = fb.Context(0)
corpus = Corpus()
for mib in [int(x) for x in sys.argv[1:]] or [1, 16]:
    n = mib << 20
    data = corpus.fill(n // 65536, 65536, seed=1, klass=Corpus.TEXT).tobytes()
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter(); c = ctx.deflate(data); t1 = time.perf_counter()
        best = min(best, t1 - t0)
    st = ctx.last_stats()
    print(mib, "MiB: call", round(best * 1e3, 2), "ms; stages", {k: round(v, 3) for k, v in ctx.last_stage_ms().items()},
          {f: getattr(st, f) for f, _ in st._fields_})
