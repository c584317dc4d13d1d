#!/usr/bin/env python
"""bench.py -- deflate/inflate GB/s (uncompressed) of the B200 path, with the
CPU reference arm beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload segments|small|random|runs] [--scaling weak|strong]

Default workload (BASELINE.json configs[1]): a synthetic
mixed-entropy corpus of 64 KiB independent segments -- 16384 segments = 1 GiB
per GPU (at N GPUs rank r owns segments [r*16384, (r+1)*16384): weak scaling,
the 8-GPU run is the 8 GiB corpus of configs[3]; --scaling strong splits a
fixed 8 GiB corpus over the ranks instead).  One step = one deflate pass over
the rank's shard, at N > 1 the all-gather of segment sizes and the frame
assembly on GPU 0 over NVLink followed by the fetch of the rank's range back
out of the frame, and one inflate pass over those streams (at N > 1 the K
steps are software-pipelined: every batch still takes all of these stages, but
its way back out of the frame travels beside the deflate of the next batch and
the next batch's way in beside its inflate -- copy engines next to kernels).  value =
uncompressed bytes through the codec per second = 2 * N_uncompressed / t_step,
inputs resident in HBM.  The other BASELINE configs are selected with
--workload (small = configs[2], random / runs = configs[4]); the default N=1
line also carries a reduced-size measurement of each under "configs".

Only this file's cpu_baseline / --impl reference legs and the parity checks
touch oracle/ (the CPU restatement of the reference); the timed GPU path is
libflate_b200.so through its C ABI.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEG = 65536
METRIC = "deflate+inflate GB/s (uncompressed bytes through the codec, per step: 1 deflate pass + 1 inflate pass)"
KLASS = {"segments": -1, "small": -1, "random": 2, "runs": 3}
WORKLOAD_TEXT = {
    "segments": "64 KiB independent segments, mixed corpus 50% text / 25% records / 15% random / 10% runs (BASELINE configs[1]; configs[3] over 8 GPUs)",
    "small": "independent small streams, 1..16 KiB uncompressed each, mixed corpus (BASELINE configs[2])",
    "random": "64 KiB segments of incompressible random bytes: literal-only Huffman blocks (BASELINE configs[4]a)",
    "runs": "64 KiB segments of long runs, a byte repeated 1..4096 times: maximum-length matches (BASELINE configs[4]b)",
}


def _ensure_helpers():
    for rel, d in (("oracle/libflate_oracle.so", "oracle"), ("tools/libfb_corpus.so", "tools")):
        if not os.path.exists(os.path.join(ROOT, rel)):
            subprocess.check_call(["make", "-s", "-C", d], cwd=ROOT)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_source_hash():
    """sha256 over the kernel sources: a DRAM-traffic figure captured under ncu is only quoted for the very
    sources it was captured from (profiles/traffic.json records the hash of its capture)."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "moonbit_flate_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region.  The sampler is started a little early
    (nvidia-smi needs ~100 ms to come up); every sample carries a timestamp and only those that fall inside
    [mark_begin, mark_end] are used -- if the region was shorter than one sampling period, the samples of the
    identical warm-up load right before it are reported and flagged."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.p = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        import datetime
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]),
                             [nm for nm, v in zip(names, f[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= self.t1]
        note = "samples inside the timed region"
        if not inside:
            inside = [r for r in rows if self.t0 is None or r[0] <= self.t1][-4:]
            note = "timed region shorter than the sampling period: last samples of the identical warm-up load"
        sm = [r[1] for r in inside]
        reasons = sorted({x for r in inside for x in r[3]})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max((r[2] for r in inside), default=None),
                "samples": len(sm), "reasons": reasons, "window": note}


# ---------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of the reference) on the host cores.

def _load_oracle():
    """The oracle for the timing legs, rebuilt on this very host with -O3 -march=native into
    oracle/libflate_oracle_native.so; the portable build the tests use if that fails."""
    from helpers import Oracle
    o = Oracle()
    native = os.path.join(ROOT, "oracle", "libflate_oracle_native.so")
    try:
        subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(ROOT, "oracle"), "native"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        L = C.CDLL(native)
        return L, "-O3 -march=native, built on this host"
    except Exception:  # noqa: BLE001
        return o.L, "portable build (-O3 -march=x86-64-v2)"


def _cpu_codec_pass(oracle_lib, src: np.ndarray, off: np.ndarray, threads: int):
    """deflate + inflate of the streams src[off[i]:off[i+1]] with `threads` host threads (thread t takes streams
    t, t + threads, ...; the loop over a thread's streams runs in C).  Returns seconds."""
    ns = off.size - 1
    off = np.ascontiguousarray(off, dtype=np.uint64)
    ok = [0] * threads
    fn = oracle_lib.orc_codec_pass
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]

    def work(t):
        ok[t] = fn(src.ctypes.data, off.ctypes.data, ns, t, threads)

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    assert all(ok), "oracle round trip failed"
    return dt


def cpu_baseline(src: np.ndarray, off: np.ndarray, budget_s: float = 12.0):
    """Bounded sample of the same workload on the box's host cores (rank 0, N=1)."""
    L, flags = _load_oracle()
    cores = os.cpu_count() or 1
    ns = off.size - 1
    # calibrate on 64 streams single-threaded, then size the all-core sample for ~budget_s of wall time
    k = min(64, ns)
    t1 = _cpu_codec_pass(L, src, off[: k + 1], 1)
    per_byte = t1 / max(1, int(off[k] - off[0]))
    n_all = ns
    while n_all > cores * 8 and per_byte * int(off[n_all] - off[0]) / cores > budget_s:
        n_all //= 2
    tN = _cpu_codec_pass(L, src, off[: n_all + 1], cores)
    nb = int(off[n_all] - off[0])
    return {
        "value": round(2 * nb / tN / 1e9, 4), "unit": "GB/s", "cores": cores, "kind": "port",
        "sample": f"first {n_all} of {ns} streams of the workload ({nb / 2**20:.0f} MiB), deflate+inflate, {cores} threads; "
                  f"C restatement of the MoonBit reference (moon toolchain absent), {flags}",
        "one_thread_gbs": round(2 / per_byte / 1e9, 4),
    }


# ---------------------------------------------------------------------------

_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner, nvidia-smi helpers, ...) may write to fd 1; the contract is ONE JSON
    line on stdout.  Everything but that line is sent to stderr."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def _emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def make_workload(corpus, kind: str, first: int, nunits: int, seed: int, pinned: bool):
    """-> (host uint8 torch tensor, offsets uint64[nunits+1], seg_size or 0)"""
    import torch
    klass = KLASS[kind]
    if kind == "small":
        off = np.zeros(nunits + 1, np.uint64)
        total = corpus.L.fb_corpus_fill_var(None, off.ctypes.data, first, nunits, seed, klass)
        t = torch.empty(max(int(total), 1), dtype=torch.uint8, pin_memory=pinned)[: int(total)]
        corpus.L.fb_corpus_fill_var(t.numpy().ctypes.data, off.ctypes.data, first, nunits, seed, klass)
        return t, off, 0
    t = torch.empty(nunits * SEG, dtype=torch.uint8, pin_memory=pinned)
    corpus.L.fb_corpus_fill(t.numpy().ctypes.data, first, nunits, SEG, seed, klass)
    return t, np.arange(nunits + 1, dtype=np.uint64) * SEG, SEG


class DeviceCodec:
    """Device-resident deflate + inflate of one workload through the C ABI (fb200_*_dev)."""

    def __init__(self, torch, ctx, dev, h_src, off: np.ndarray, seg: int):
        self.torch, self.ctx, self.seg = torch, ctx, seg
        self.ns = off.size - 1
        self.n = int(off[-1])
        self.d_src = h_src.to(dev)
        self.d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
        self.cap = self.n + self.n // 8 + self.ns * 1024 + 4096
        self.d_dst = torch.zeros(self.cap, dtype=torch.uint8, device=dev)
        self.d_doff = torch.zeros(self.ns + 1, dtype=torch.int64, device=dev)
        self.d_out = torch.zeros(self.n, dtype=torch.uint8, device=dev)
        self.d_olen = torch.zeros(self.ns, dtype=torch.int64, device=dev)
        self.d_st = torch.zeros(self.ns, dtype=torch.int32, device=dev)
        self.d_eo = torch.zeros(self.ns, dtype=torch.int64, device=dev)
        self.clen = 0

    def deflate(self):
        if self.seg:
            self.clen = self.ctx.deflate_segments_dev(self.d_src.data_ptr(), self.n, self.seg, self.d_dst.data_ptr(),
                                                      self.cap, self.d_doff.data_ptr())
        else:
            self.clen = self.ctx.deflate_streams_dev(self.d_src.data_ptr(), self.d_off.data_ptr(), self.ns, self.n,
                                                     self.d_dst.data_ptr(), self.cap, self.d_doff.data_ptr())
        return self.clen

    def inflate(self, d_comp=None, d_coff=None):
        self.ctx.inflate_batch_dev((d_comp if d_comp is not None else self.d_dst).data_ptr(),
                                   (d_coff if d_coff is not None else self.d_doff).data_ptr(), self.ns,
                                   self.d_out.data_ptr(), self.d_off.data_ptr(), self.d_olen.data_ptr(),
                                   self.d_st.data_ptr(), self.d_eo.data_ptr())

    def check(self, h_np, off, oracle, nsample=24):
        """Round trip bit-exact on the device + a sample of streams byte-identical to the oracle's."""
        t = self.torch
        assert bool((self.d_st == 0).all()) and bool((self.d_olen == (self.d_off[1:] - self.d_off[:-1])).all()), "inflate status"
        assert t.equal(self.d_out, self.d_src), "round trip mismatch"
        doff = self.d_doff.cpu().numpy()
        for i in list(range(0, self.ns, max(1, self.ns // nsample)))[:nsample]:
            got = self.d_dst[int(doff[i]): int(doff[i + 1])].cpu().numpy().tobytes()
            assert got == oracle.deflate(h_np[int(off[i]): int(off[i + 1])].tobytes()), f"stream {i} differs from the oracle"
        return nsample


def side_config(torch, ctx, dev, corpus, oracle, kind, nunits, seed, reps=3):
    """Reduced-size device-resident measurement of another BASELINE config (default N=1 line, "configs")."""
    h, off, seg = make_workload(corpus, kind, 0, nunits, seed, pinned=False)
    dc = DeviceCodec(torch, ctx, dev, h, off, seg)
    best_d, best_i, stages = 1e30, 1e30, None
    for _ in range(reps + 1):
        dc.deflate()
        ms = ctx.last_stage_ms()
        d = sum(v for k, v in ms.items() if k != "inflate")
        dc.inflate()
        i = ctx.last_stage_ms()["inflate"]
        if d < best_d:
            best_d, stages = d, {k: round(v, 3) for k, v in ms.items() if k != "inflate"}
        best_i = min(best_i, i)
    ns = dc.check(h.numpy(), off, oracle, 12)
    peak, _ = _peaks()
    res = {"workload": f"{nunits} x " + WORKLOAD_TEXT[kind], "uncompressed_bytes": dc.n, "compressed_bytes": int(dc.clen),
           "ratio": round(dc.clen / dc.n, 5), "deflate_ms": round(best_d, 3), "inflate_ms": round(best_i, 3),
           "deflate_gbs": round(dc.n / best_d / 1e6, 2), "inflate_gbs": round(dc.n / best_i / 1e6, 2),
           "value_gbs": round(2 * dc.n / (best_d + best_i) / 1e6, 2), "deflate_stage_ms": stages,
           "hbm_frac_deflate": round((dc.n + dc.clen) / best_d / 1e6 / peak, 5),
           "hbm_frac_inflate": round((dc.n + dc.clen) / best_i / 1e6 / peak, 5),
           "timing": f"library stage timers (CUDA events on the context stream), best of {reps}",
           "checked": f"round trip bit-exact on all streams; {ns} streams byte-identical to the oracle"}
    del dc
    torch.cuda.empty_cache()
    return res


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="segments", choices=list(KLASS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--segments", type=int, default=16384, help="64 KiB segments per GPU (16384 = 1 GiB); weak scaling")
    ap.add_argument("--total-segments", type=int, default=131072, help="segments of the whole corpus (8 GiB); strong scaling")
    ap.add_argument("--streams", type=int, default=1000000, help="small streams per GPU (--workload small)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the reduced-size measurements of the other BASELINE configs")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    _ensure_helpers()
    from helpers import Corpus, Oracle
    from moonbit_flate_b200.multigpu import shard_range

    kind = args.workload
    strong = args.scaling == "strong"
    if kind == "small":
        units_total = args.streams * (1 if strong else world)
    else:
        units_total = args.total_segments if strong else args.segments * world
    first, last = shard_range(units_total, world, rank)
    nunits = last - first
    config = {"workload": f"{units_total} x " + WORKLOAD_TEXT[kind] + f"; seed {args.seed}",
              "units_total": units_total, "units_per_gpu": nunits,
              "sharding": f"contiguous unit ranges over {world} GPU(s), {'fixed total (strong)' if strong else 'fixed per GPU (weak)'}",
              "l2": "inputs larger than the 126 MB L2; no flush needed"}
    corpus = Corpus()

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        L, flags = _load_oracle()
        cores = os.cpu_count() or 1
        # the same units rank 0's GPU arm compresses (its whole shard, or a stated prefix of it when the whole
        # shard would take more than ~20 s per step on this host)
        h, off, _ = make_workload(corpus, kind, first, nunits, args.seed, pinned=False)
        src = h.numpy()
        k = min(32, nunits)
        per_byte = _cpu_codec_pass(L, src, off[: k + 1], 1) / max(1, int(off[k]))
        n_s = nunits
        while n_s > cores * 8 and per_byte * int(off[n_s]) / cores > 20.0:
            n_s //= 2
        offs = off[: n_s + 1]
        for _ in range(min(args.warmup, 1)):
            _cpu_codec_pass(L, src, offs, cores)
        ts = [_cpu_codec_pass(L, src, offs, cores) for _ in range(args.steps)]
        t = sum(ts) / len(ts)
        nb = int(offs[-1])
        v = 2 * nb / t / 1e9
        sample = (f"{n_s} of the {nunits} units of rank 0's shard per step ({nb / 2**20:.0f} MiB), deflate+inflate on {cores} "
                  f"host threads; C restatement of the MoonBit reference (moon toolchain absent, the reference cannot be "
                  f"compiled here), {flags}")
        line = {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": "GB/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t * 1e3, 3),
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": round(v, 4), "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": round(v, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        _emit(line)
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist

    import moonbit_flate_b200 as fb
    from moonbit_flate_b200 import multigpu as mg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # host buffers next to the GPU: bind this rank to the CPUs NVML reports as local to its GPU before any pinned
    # allocation (first touch decides the NUMA node); without it the e2e copies of several ranks cross sockets
    affinity = "unchanged"
    if os.environ.get("FB200_BENCH_AFFINITY", "1") != "0":
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            affinity = f"nvml ideal cpus ({len(os.sched_getaffinity(0))} cpus)"
        except Exception as e:  # noqa: BLE001 -- restricted cpusets: keep the default placement
            affinity = f"unchanged ({type(e).__name__})"
    config["host_affinity"] = affinity

    ctx = fb.Context(local_rank)
    lib_stream = torch.cuda.ExternalStream(ctx.cuda_stream(), device=dev)
    oracle = Oracle()

    # synthetic shard of this rank, generated into pinned host memory
    h_src, off, seg = make_workload(corpus, kind, first, nunits, args.seed, pinned=True)
    h_np = h_src.numpy()
    nbytes = int(off[-1])
    dc = DeviceCodec(torch, ctx, dev, h_src, off, seg)

    # frame on GPU 0 (N > 1): CUDA IPC + copy-engine peer copies; NCCL send/recv if IPC is unavailable on this box.
    # The inflate pass of every rank reads its streams back OUT OF THE FRAME (fb200_mg_get), not its local buffer.
    frame = None
    peer = None
    d_fcomp = d_fcoff = None
    if world > 1:
        frame_cap = mg.frame_header_bytes(units_total) + world * dc.cap
        if os.environ.get("FB200_FRAME_TRANSPORT", "ipc") == "ipc":
            peer = mg.PeerFrame(ctx, rank, world, frame_cap, dev)
            if not peer.available:
                peer = None
        if peer is None and rank == 0:
            frame = torch.empty(frame_cap, dtype=torch.uint8, device=dev)
        d_fcomp = torch.zeros(dc.cap, dtype=torch.uint8, device=dev)
        d_fcoff = torch.zeros(nunits + 1, dtype=torch.int64, device=dev)
        config["frame"] = ("assembled on GPU 0 with cuda-ipc peer copies (copy engines over NVLink) and read back by every rank "
                           "with fb200_mg_get_async before its inflate pass; steps software-pipelined: the way back of batch i "
                           "beside the deflate of batch i + 1, the way in of batch i + 1 beside the inflate of batch i" if peer else
                           "assembled on GPU 0 and scattered back with nccl send/recv before the inflate pass")

    stage_acc = {}
    launches = [0]
    frame_len = [0]

    def step(record=True):
        dc.deflate()
        if record:
            launches[0] += int(ctx.last_stats().kernel_launches)
            for k, v in ctx.last_stage_ms().items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v
        if world > 1:
            sizes = dc.d_doff[1:] - dc.d_doff[:-1]
            if peer:
                frame_len[0] = peer.put(dc.d_dst, sizes, seg or 0, nseg_total=units_total)
                peer.wait()  # this rank's payload and the header are in GPU 0's frame
                peer.get(first, nunits, d_fcomp, d_fcoff)
                dc.inflate(d_fcomp, d_fcoff)
            else:
                frame_len[0] = mg.assemble_frame(dc.d_dst, sizes, seg or 0, rank, world, frame, nseg_total=units_total)
                _, _, _, my_sizes, my_payload = mg.scatter_frame(frame[: frame_len[0]] if rank == 0 else None, rank, world, dev)
                d_fcoff.copy_(torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(my_sizes, 0)]).to(dev))
                dc.inflate(my_payload, d_fcoff)
        else:
            dc.inflate()
        if record:
            launches[0] += int(ctx.last_stats().kernel_launches)
            stage_acc["inflate"] = stage_acc.get("inflate", 0.0) + ctx.last_stage_ms()["inflate"]

    def run_steps(k, record):
        """k steps.  With the frame on GPU 0 (N > 1, peer copies) the steps are software-pipelined: every batch is
        still deflated, put into the frame, read back out of it and inflated, but the way back of batch i travels
        beside the deflate of batch i + 1 and the way in of batch i + 1 beside the inflate of batch i (copy engines
        over NVLink next to the kernels), instead of all ranks pushing and then all ranks pulling through GPU 0's
        ports while the SMs wait.  (Before anyone puts batch i + 1, every rank has its batch i back: the size
        all-gather inside put comes after the rank's mg_wait.)"""
        if k <= 0:
            return
        if not (world > 1 and peer):
            for _ in range(k):
                step(record)
            return

        def deflate_rec():
            dc.deflate()
            if record:
                launches[0] += int(ctx.last_stats().kernel_launches)
                for kk, v in ctx.last_stage_ms().items():
                    stage_acc[kk] = stage_acc.get(kk, 0.0) + v

        def put():
            frame_len[0] = peer.put(dc.d_dst, dc.d_doff[1:] - dc.d_doff[:-1], seg or 0, nseg_total=units_total)

        deflate_rec()
        put()
        for i in range(k):
            peer.wait()  # this rank's payload and the header of batch i are in GPU 0's frame
            peer.get_begin(first, nunits, d_fcomp, d_fcoff)
            if i + 1 < k:
                deflate_rec()  # batch i + 1, beside the way back of batch i
            ctx.mg_wait()  # batch i has arrived
            if i + 1 < k:
                put()  # batch i + 1 on its way in, beside the inflate below
            dc.inflate(d_fcomp, d_fcoff)
            if record:
                launches[0] += int(ctx.last_stats().kernel_launches)
                stage_acc["inflate"] = stage_acc.get("inflate", 0.0) + ctx.last_stage_ms()["inflate"]

    # warm-up (untimed) + parity check of the resident result
    sampler = ClockSampler(local_rank)
    sampler.start()
    run_steps(args.warmup, record=False)
    torch.cuda.synchronize()
    dc.d_out.zero_()
    run_steps(1, record=False)  # same load right before the timed region (also keeps the clocks up for the sampler)
    torch.cuda.synchronize()
    if peer:
        peer.wait_all()
    n_checked = dc.check(h_np, off, oracle)  # inflate(what came out of the frame) == input; sample == oracle
    clen = dc.clen
    if world > 1 and rank == 0:  # the assembled frame: sample streams of every rank inflate to that rank's units
        import zlib
        fr = peer.view if peer else frame
        seg_size, nseg_f, sizes_f, hdr_f = mg.parse_frame(fr)
        assert seg_size == (seg or 0) and nseg_f == units_total and hdr_f + int(sizes_f.sum()) == frame_len[0]
        offs_f = np.concatenate([[0], np.cumsum(sizes_f.numpy())]) + hdr_f
        for r in range(world):
            f_r, l_r = shard_range(units_total, world, r)
            for i in (f_r, (f_r + l_r) // 2, l_r - 1):
                comp_i = fr[int(offs_f[i]): int(offs_f[i + 1])].cpu().numpy().tobytes()
                hh, _, _ = make_workload(corpus, kind, i, 1, args.seed, pinned=False)
                assert zlib.decompress(comp_i, -15) == hh.numpy().tobytes(), f"frame stream {i} (rank {r}) does not inflate to its unit"

    # timed region: events on the stream the kernels are launched on
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record(lib_stream)
    run_steps(args.steps, record=True)
    torch.cuda.synchronize()
    e1.record(lib_stream)
    e1.synchronize()
    sampler.mark_end()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(clen), float(nbytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_step = float(t_ms.item()) / args.steps
    total_unc = float(sums[1].item())
    value = 2 * total_unc / (ms_step * 1e-3) / 1e9

    # ------------------------------------------------------------------ e2e through the host-buffer C ABI
    e2e = None
    if not args.no_e2e and nbytes <= (3 << 30):
        # Four contexts: deflate calls alternate between two, inflate calls between two more, all through the
        # asynchronous form of the host-buffer C ABI (fb200_*_async + fb200_wait).  Two deflate calls are kept in
        # flight -- the input of the second travels while the kernels of the first run -- and inflate call i starts as
        # soon as deflate call i has delivered its streams to (pinned) host memory; the H2D copy of one call travels
        # beside the D2H copy of another (PCIe is full duplex) and the library orders copies and kernel phases of
        # the calls FIFO on the GPU.  The timed region holds K deflate and K inflate calls, fill and drain included.
        ctxD = [ctx, fb.Context(local_rank)]
        ctxI = [fb.Context(local_rank), fb.Context(local_rank)]
        ns = nunits
        cap = dc.cap
        NB = 4  # compressed host buffers: two being written by deflate calls, two being read by inflate calls
        h_off = torch.from_numpy(off.astype(np.int64)).pin_memory()
        h_dst = [torch.empty(cap, dtype=torch.uint8, pin_memory=True) for _ in range(NB)]
        h_doff = [torch.zeros(ns + 1, dtype=torch.int64, pin_memory=True) for _ in range(NB)]
        h_out = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        h_out_len = [torch.zeros(ns, dtype=torch.int64, pin_memory=True) for _ in range(2)]
        h_status = [torch.zeros(ns, dtype=torch.int32, pin_memory=True) for _ in range(2)]
        h_err_off = [torch.zeros(ns, dtype=torch.int64, pin_memory=True) for _ in range(2)]

        def deflate_async(i, c=None):
            b = i % NB
            c = c or ctxD[i & 1]
            if seg:
                c.deflate_segments_async_ptr(h_src.data_ptr(), nbytes, seg, h_dst[b].data_ptr(), cap, h_doff[b].data_ptr())
            else:
                c.deflate_streams_async_ptr(h_src.data_ptr(), h_off.data_ptr(), ns, h_dst[b].data_ptr(), cap, h_doff[b].data_ptr())

        def inflate_async(i, c=None):
            b, o = i % NB, i & 1
            (c or ctxI[o]).inflate_batch_async_ptr(h_dst[b].data_ptr(), h_doff[b].data_ptr(), ns, h_out[o].data_ptr(),
                                                   h_off.data_ptr(), h_out_len[o].data_ptr(), h_status[o].data_ptr(),
                                                   h_err_off[o].data_ptr())

        def e2e_run(k):
            """k deflate + k inflate calls: two deflate calls are kept in flight, inflate call i is issued as soon as
            deflate call i has returned (its streams are in host memory) and inflate call i - 2 has freed its context."""
            deflate_async(0)
            if k > 1:
                deflate_async(1)
            cl = 0
            for i in range(k):
                cl = ctxD[i & 1].wait()
                if i + 2 < k:
                    deflate_async(i + 2)
                if i >= 2:
                    ctxI[i & 1].wait()
                inflate_async(i)
            for c in ctxI:
                c.wait()
            return cl

        def e2e_ok():
            return all(int(h_status[o].abs().sum()) == 0 and torch.equal(h_out[o], h_src) for o in range(2))

        def e2e_serial_step():
            deflate_async(0, ctx)
            cl = ctx.wait()
            inflate_async(0, ctx)
            ctx.wait()
            return cl

        cl = e2e_run(4)  # warm-up of all contexts + check
        assert e2e_ok(), "e2e round trip mismatch"
        for o in range(2):
            h_out[o].zero_()
        ksteps = max(4, min(args.steps, 16))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        cl = e2e_run(ksteps)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / ksteps
        assert e2e_ok(), "e2e round trip mismatch (timed run)"
        # the same two calls one after the other on one context (round-1 definition), for comparison
        e2e_serial_step()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, ksteps // 2)):
            e2e_serial_step()
        dts = (time.perf_counter() - t0) / max(1, ksteps // 2)
        t_e = torch.tensor([dt, dts], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        meta = (ns + 1) * 8
        e2e = {"value": round(2 * total_unc / float(t_e[0].item()) / 1e9, 3), "unit": "GB/s",
               "h2d_bytes_per_step": int((nbytes + cl + 3 * meta) * world),
               "d2h_bytes_per_step": int((cl + nbytes + meta + ns * 20) * world),
               "steps": ksteps, "ms_per_step": round(float(t_e[0].item()) * 1e3, 3),
               "path": "fb200_deflate_*_async on two contexts + fb200_inflate_batch_async on two more, pinned host buffers; "
                       "inflate call i consumes what deflate call i delivered to the host; two deflate calls in flight; "
                       "pipeline fill and drain inside the timed region",
               "serial": {"value": round(2 * total_unc / float(t_e[1].item()) / 1e9, 3),
                          "ms_per_step": round(float(t_e[1].item()) * 1e3, 3),
                          "path": "one deflate call, then one inflate call, each waited for, one context"}}
        pb = os.path.join(ROOT, "profiles", "pcie_bound.json")
        if os.path.exists(pb):
            try:
                e2e["platform_bound_gbs"] = json.load(open(pb)).get(str(world))
            except Exception:
                pass
        for c in ctxI:
            c.close()
        ctxD[1].close()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ roofline of the dominant kernel + report
    peak, peak_src = _peaks()
    k = args.steps
    c_local = float(clen)
    st_ms = {kk: v / k for kk, v in stage_acc.items()}
    deflate_ms = sum(v for kk, v in st_ms.items() if kk != "inflate")
    inflate_ms = st_ms.get("inflate", 0.0)
    dom = max(st_ms, key=lambda kk: st_ms[kk])
    kernel_name = {"parse": "k_parse (K1 lz77 parse)", "inflate": "k_inflate_par (K6 inflate)"}.get(dom, dom)
    alg_bytes = nbytes + c_local  # per launch: N + C (deflate) == C + N (inflate), SURVEY 8d
    achieved = alg_bytes / (st_ms[dom] * 1e-3) / 1e9
    traffic, traffic_note = None, "no ncu capture on record for this workload"
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and kind == "segments" and nunits == 16384:
        try:
            tj = json.load(open(tpath))
            cur = kernel_source_hash()
            if tj.get("kernel_source_sha256_16") == cur:
                traffic = tj.get(dom)
                traffic_note = f"{tj.get('capture')} (kernel sources {cur})"
            else:
                traffic_note = (f"withheld: the capture on record ({tj.get('capture')}) is of kernel sources "
                                f"{tj.get('kernel_source_sha256_16')}, this run is {cur}")
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(alg_bytes),
                "kernel_ms_per_launch": round(st_ms[dom], 4),
                "stage_ms_per_step": {kk: round(v, 4) for kk, v in st_ms.items()},
                "deflate": {"gbs": round(nbytes / (deflate_ms * 1e-3) / 1e9, 2),
                            "alg_gbs": round(alg_bytes / (deflate_ms * 1e-3) / 1e9, 2),
                            "frac": round(alg_bytes / (deflate_ms * 1e-3) / 1e9 / peak, 5)},
                "inflate": {"gbs": round(nbytes / (inflate_ms * 1e-3) / 1e9, 2),
                            "alg_gbs": round(alg_bytes / (inflate_ms * 1e-3) / 1e9, 2),
                            "frac": round(alg_bytes / (inflate_ms * 1e-3) / 1e9 / peak, 5)}}

    line = {"metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
            "compression_ratio": round(float(sums[0].item()) / total_unc, 5),
            "parity": f"round trip bit-exact on every stream; {n_checked} streams byte-identical to the oracle",
            "roofline": roofline, "clocks": clocks, "gpu_launches": launches[0]}
    if world > 1:
        # what a step spends outside the kernels: size all-gather, peer copies into GPU 0's frame, header, peer copies
        # back out of it (all ranks' streams pass through one GPU's NVLink ports, once in each direction)
        line["frame_exchange_ms_per_step"] = round(ms_step - sum(st_ms.values()), 3)
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(h_np, off)
    if world == 1 and kind == "segments" and not args.no_configs and not strong:
        del dc
        torch.cuda.empty_cache()
        cfgs = {}
        for name, kd, nu in (("configs[2]", "small", 262144), ("configs[4]a", "random", 16384), ("configs[4]b", "runs", 16384)):
            try:
                cfgs[name] = side_config(torch, ctx, dev, corpus, oracle, kd, nu, args.seed)
            except Exception as e:  # noqa: BLE001 -- the headline line must survive a side measurement
                cfgs[name] = {"error": f"{type(e).__name__}: {e}"}
        line["configs"] = cfgs
    _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
