"""The drop-in objects on ONE stream (BASELINE configs[0] and longer): Writer::new/write/close and
Reader::new/read through the Python mirror of the reference API, beside the oracle on one host thread.
usage: python scripts/single_stream_probe.py [MiB ...]   -> one JSON object on stdout"""
import io, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import moonbit_flate_b200 as fb
from helpers import Corpus, Oracle

sizes = [int(x) for x in sys.argv[1:]] or [1, 16, 128]
ctx = fb.Context(0)
corpus, oracle = Corpus(), Oracle()
res = []
for mib in sizes:
    n = mib << 20
    data = corpus.fill(n // 65536, 65536, seed=1, klass=Corpus.TEXT).tobytes()
    t0 = time.perf_counter(); want = oracle.deflate(data); t1 = time.perf_counter()
    st, back, _, _ = oracle.inflate(want, n + 1); t2 = time.perf_counter()
    assert back == data
    best_w = best_r = best_w1 = 1e9
    for _ in range(4):
        buf = io.BytesIO()
        t3 = time.perf_counter()
        w = fb.Writer.new(buf, ctx); w.write(data); w.close()
        t4 = time.perf_counter()
        got, err = fb.Reader.new(buf.getvalue(), ctx).read_all()
        t5 = time.perf_counter()
        assert buf.getvalue() == want and err is None and got == data
        best_w, best_r = min(best_w, t4 - t3), min(best_r, t5 - t4)
        t6 = time.perf_counter(); c = ctx.deflate(data); t7 = time.perf_counter()  # the whole stream in one call
        assert c == want
        best_w1 = min(best_w1, t7 - t6)
    res.append({"mib": mib, "compressed_bytes": len(want),
                "gpu_writer_ms": round(best_w * 1e3, 2), "gpu_deflate_one_call_ms": round(best_w1 * 1e3, 2),
                "gpu_reader_ms": round(best_r * 1e3, 2),
                "gpu_reader_mb_s": round(n / best_r / 1e6, 1), "gpu_writer_mb_s": round(n / best_w / 1e6, 1),
                "oracle_1_thread_deflate_ms": round((t1 - t0) * 1e3, 2), "oracle_1_thread_inflate_ms": round((t2 - t1) * 1e3, 2)})
    print(res[-1], file=sys.stderr)
print(json.dumps({"what": "one text stream through Writer::new/write/close and Reader::new/read (host buffers, Python mirror), "
                          "best of 4; byte-identical to the oracle, round trip exact", "results": res}, indent=1))
