"""Reader on ONE incompressible stream (literal-only blocks with a flat code): python scripts/reader_random_probe.py [MiB ...]"""
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
    data = corpus.fill(n // 65536, 65536, seed=3, klass=2).tobytes()
    comp = ctx.deflate(data)
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter(); got, err = fb.Reader.new(comp, ctx).read_all(); t1 = time.perf_counter()
        assert err is None and got == data
        best = min(best, t1 - t0)
    print(mib, "MiB incompressible: reader", round(best * 1e3, 2), "ms", round(n / best / 1e6, 1), "MB/s")
