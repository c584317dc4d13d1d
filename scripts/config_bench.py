"""Device-resident throughput of the BASELINE configs that are not the headline bench line:
   configs[2] 1 M independent small streams (1..16 KiB uncompressed each), configs[4] 1 GiB incompressible
   random bytes and 1 GiB long-run data (64 KiB segments).  For each: deflate and inflate ms / GB/s (CUDA events
   of the library's own stage timers), compression ratio, round-trip check, and a sample of streams compared
   with the oracle (checker only).

   python scripts/config_bench.py [--streams 1000000] [--segments 16384] [--reps 3] > profiles/rXX_configs.json
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import moonbit_flate_b200 as fb
from helpers import Corpus, Oracle

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=1000000)
ap.add_argument("--segments", type=int, default=16384)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
ctx = fb.Context(0)
corpus, oracle = Corpus(), Oracle()
dev = torch.device("cuda", 0)


def run(name, src_np, off_np, sample):
    ns = off_np.size - 1
    n = int(off_np[-1])
    d_src = torch.from_numpy(src_np).to(dev)
    d_off = torch.from_numpy(off_np.astype(np.int64)).to(dev)
    cap = n + n // 8 + ns * 64 + 4096
    d_dst = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_doff = torch.zeros(ns + 1, dtype=torch.int64, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    d_olen = torch.zeros(ns, dtype=torch.int64, device=dev)
    d_st = torch.zeros(ns, dtype=torch.int32, device=dev)
    d_eo = torch.zeros(ns, dtype=torch.int64, device=dev)
    best_d, best_i, stages = 1e30, 1e30, None
    for _ in range(args.reps + 1):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        c = ctx.deflate_streams_dev(d_src.data_ptr(), d_off.data_ptr(), ns, n, d_dst.data_ptr(), cap, d_doff.data_ptr())
        t1 = time.perf_counter()
        ms_d = ctx.last_stage_ms()
        t2 = time.perf_counter()
        ctx.inflate_batch_dev(d_dst.data_ptr(), d_doff.data_ptr(), ns, d_out.data_ptr(), d_off.data_ptr(),
                              d_olen.data_ptr(), d_st.data_ptr(), d_eo.data_ptr())
        t3 = time.perf_counter()
        ms_i = ctx.last_stage_ms()["inflate"]
        fbk = int(ctx.last_stats().inflate_fallbacks)
        dk = sum(v for k, v in ms_d.items() if k != "inflate")
        if dk < best_d:
            best_d, stages = dk, {k: round(v, 3) for k, v in ms_d.items() if k != "inflate"}
        best_i = min(best_i, ms_i)
        wall = (round((t1 - t0) * 1e3, 2), round((t3 - t2) * 1e3, 2))
    assert torch.equal(d_out, d_src) and int(d_st.abs().sum()) == 0 and bool((d_olen == (d_off[1:] - d_off[:-1])).all())
    doff = d_doff.cpu().numpy()
    for i in sample:
        got = d_dst[int(doff[i]): int(doff[i + 1])].cpu().numpy().tobytes()
        assert got == oracle.deflate(src_np[int(off_np[i]): int(off_np[i + 1])].tobytes()), (name, i)
    res = {"config": name, "streams": ns, "uncompressed_bytes": n, "compressed_bytes": int(c), "ratio": round(c / n, 5),
           "deflate_ms": round(best_d, 3), "deflate_gbs": round(n / best_d / 1e6, 2), "deflate_stage_ms": stages,
           "inflate_ms": round(best_i, 3), "inflate_gbs": round(n / best_i / 1e6, 2), "inflate_fallbacks": fbk,
           "call_wall_ms_last": {"deflate": wall[0], "inflate": wall[1]},
           "hbm_frac_deflate": round((n + c) / best_d / 1e6 / 6537.0, 5), "hbm_frac_inflate": round((n + c) / best_i / 1e6 / 6537.0, 5),
           "checked": f"round trip bit-exact on all streams; {len(sample)} streams byte-identical to the oracle"}
    print(json.dumps(res), file=sys.stderr)
    del d_src, d_dst, d_out
    torch.cuda.empty_cache()
    return res


def config0():
    """configs[0]: deflate-fast compress + inflate round trip of a 1 MiB synthetic text buffer as ONE stream through the
    Writer / Reader mirror of the reference API (the reference's own CPU-runnable case): wall time on the GPU path
    beside the oracle on one host thread.  A single stream is latency bound here (17 blocks, block-parallel parse)."""
    import io
    data = corpus.fill(16, 65536, seed=1, klass=Corpus.TEXT).tobytes()
    best_w, best_r = 1e9, 1e9
    for _ in range(args.reps + 1):
        buf = io.BytesIO()
        t0 = time.perf_counter()
        w = fb.Writer.new(buf, ctx); w.write(data); w.close()
        t1 = time.perf_counter()
        got, err = fb.Reader.new(buf.getvalue(), ctx).read_all()
        t2 = time.perf_counter()
        best_w, best_r = min(best_w, t1 - t0), min(best_r, t2 - t1)
    assert err is None and got == data
    t0 = time.perf_counter(); want = oracle.deflate(data); t1 = time.perf_counter()
    st, back, _, _ = oracle.inflate(want, len(data) + 1); t2 = time.perf_counter()
    assert buf.getvalue() == want and back == data
    res = {"config": "configs[0] 1 MiB text, one stream, Writer::new/write/close + Reader::new/read (host buffers)",
           "uncompressed_bytes": len(data), "compressed_bytes": len(want),
           "gpu_writer_ms": round(best_w * 1e3, 2), "gpu_reader_ms": round(best_r * 1e3, 2),
           "oracle_1_thread_deflate_ms": round((t1 - t0) * 1e3, 2), "oracle_1_thread_inflate_ms": round((t2 - t1) * 1e3, 2),
           "checked": "GPU stream byte-identical to the oracle's, round trip exact"}
    print(json.dumps(res), file=sys.stderr)
    return res


out = []
out.append(config0())
SEG = 65536
nseg = args.segments
seg_off = (np.arange(nseg + 1, dtype=np.uint64) * SEG)
samp = list(range(0, nseg, max(1, nseg // 12)))[:12]
out.append(run("configs[1] mixed corpus, 64 KiB segments", corpus.fill(nseg, SEG, seed=1), seg_off, samp))
out.append(run("configs[4]a incompressible random bytes, 64 KiB segments", corpus.fill(nseg, SEG, seed=1, klass=Corpus.RANDOM), seg_off, samp))
out.append(run("configs[4]b long runs (byte repeated 1..4096 times), 64 KiB segments", corpus.fill(nseg, SEG, seed=1, klass=Corpus.RUNS), seg_off, samp))
out.append(run("configs[4]b single repeated byte per segment (maximum-length matches)", corpus.fill(nseg, SEG, seed=1, klass=Corpus.CONST), seg_off, samp))
out.append(run("configs[4]b period-7 pattern", corpus.fill(nseg, SEG, seed=1, klass=Corpus.PERIOD7), seg_off, samp))
src, off = corpus.fill_var(args.streams, seed=1)
out.append(run(f"configs[2] {args.streams} independent small streams (1..16 KiB uncompressed each, mixed classes)", src, off,
               list(range(0, args.streams, max(1, args.streams // 24)))[:24]))
print(json.dumps({"gpu": torch.cuda.get_device_name(0), "hbm_peak_gbs": 6537.0, "timing": "library stage timers (CUDA events on the context stream), best of %d" % args.reps,
                  "results": out}, indent=1))
