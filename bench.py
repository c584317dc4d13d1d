#!/usr/bin/env python
"""bench.py -- deflate/inflate GB/s (uncompressed) of the B200 path, with the
CPU reference arm beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): a synthetic mixed-entropy
corpus of 64 KiB independent segments -- 16384 segments = 1 GiB per GPU (at N
GPUs rank r owns segments [r*16384, (r+1)*16384): weak scaling, the 8-GPU run
is the 8 GiB corpus of configs[3]).  One step = one deflate pass over the
rank's shard (+ at N>1 the all-gather of segment sizes and the frame assembly
on GPU 0 over NVLink) followed by one inflate pass over the streams it
produced.  value = uncompressed bytes through the codec per second =
2 * N_uncompressed * world / t_step, inputs resident in HBM.

Only this file's cpu_baseline / --impl reference legs and the warm-up parity
check touch oracle/ (the CPU restatement of the reference); the timed GPU path
is libflate_b200.so through its C ABI.
"""
from __future__ import annotations

import argparse
import ctypes as C
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

def _cpu_codec_pass(oracle_lib, src: np.ndarray, nseg: int, threads: int):
    """deflate + inflate of nseg 64 KiB segments with `threads` host threads.  Returns seconds."""
    bound = oracle_lib.orc_deflate_bound(SEG)
    ok = [True] * threads

    def work(t):
        dst = np.empty(bound, np.uint8)
        out = np.empty(SEG, np.uint8)
        ol = C.c_size_t()
        eo = C.c_int64()
        cons = C.c_int64()
        for i in range(t, nseg, threads):
            p = src.ctypes.data + i * SEG
            n = oracle_lib.orc_deflate(p, SEG, dst.ctypes.data, bound)
            st = oracle_lib.orc_inflate(dst.ctypes.data, n, out.ctypes.data, SEG, C.byref(ol), C.byref(eo), C.byref(cons))
            if n < 0 or st != 0 or ol.value != SEG:
                ok[t] = False

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    assert all(ok), "oracle round trip failed"
    return dt


def _load_oracle():
    from helpers import Oracle
    return Oracle().L


def cpu_baseline(src: np.ndarray, nseg_total: int, budget_s: float = 12.0):
    """Bounded sample of the same workload on the box's host cores (rank 0, N=1)."""
    L = _load_oracle()
    cores = os.cpu_count() or 1
    # calibrate on 64 segments single-threaded, then size the all-core sample for ~budget_s of wall time
    t1 = _cpu_codec_pass(L, src, min(64, nseg_total), 1)
    per_seg = t1 / min(64, nseg_total)
    one_thread_gbs = 2 * SEG / per_seg / 1e9
    n_all = int(min(nseg_total, max(cores * 8, budget_s * cores / per_seg * 0.8)))
    tN = _cpu_codec_pass(L, src, n_all, cores)
    return {
        "value": round(2 * SEG * n_all / tN / 1e9, 4), "unit": "GB/s", "cores": cores, "kind": "port",
        "sample": f"first {n_all} of {nseg_total} segments (64 KiB each), deflate+inflate, {cores} threads; "
                  f"C restatement of the MoonBit reference (moon toolchain absent)",
        "one_thread_gbs": round(one_thread_gbs, 4),
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


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--segments", type=int, default=16384, help="64 KiB segments per GPU (16384 = 1 GiB)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    _ensure_helpers()
    from helpers import Corpus

    nseg = args.segments
    nbytes = nseg * SEG
    config = {"workload": f"{nseg} x 64 KiB independent segments per GPU "
                          f"({nbytes / 2**30:.3g} GiB/GPU, mixed corpus 50% text / 25% records / 15% random / 10% runs, "
                          f"seed {args.seed}); BASELINE configs[1] per GPU, configs[3] at 8 GPUs",
              "segments_per_gpu": nseg, "segment_bytes": SEG, "sharding": f"segments x{world} (weak)",
              "l2": "inputs (1 GiB/GPU) larger than the 126 MB L2; no flush needed"}

    corpus = Corpus()

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        L = _load_oracle()
        cores = os.cpu_count() or 1
        src = corpus.fill(min(nseg, 4096), SEG, seed=args.seed, first=0)
        t1 = _cpu_codec_pass(L, src, 32, 1) / 32
        # each step: a bounded sample sized for ~4 s of wall time on all cores
        n_s = int(min(src.size // SEG, max(cores * 4, 4.0 * cores / t1 * 0.8)))
        for _ in range(args.warmup):
            _cpu_codec_pass(L, src, min(n_s, cores * 4), cores)
        ts = [_cpu_codec_pass(L, src, n_s, cores) for _ in range(args.steps)]
        t = sum(ts) / len(ts)
        v = 2 * SEG * n_s / t / 1e9
        line = {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": "GB/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t * 1e3, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": round(v, 4), "unit": "GB/s", "cores": cores, "kind": "port",
                                 "sample": f"{n_s} segments of 64 KiB per step, deflate+inflate on {cores} host threads; "
                                           "C restatement of the MoonBit reference (moon toolchain absent, "
                                           "reference cannot be compiled here)"},
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

    # synthetic shard of this rank, generated into pinned host memory
    h_src = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_np = h_src.numpy()
    corpus.L.fb_corpus_fill(h_np.ctypes.data, rank * nseg, nseg, SEG, args.seed, -1)
    d_src = h_src.to(dev, non_blocking=False)
    dst_cap = nbytes + nbytes // 8 + nseg * 1024
    d_dst = torch.zeros(dst_cap, dtype=torch.uint8, device=dev)
    d_seg_off = torch.zeros(nseg + 1, dtype=torch.int64, device=dev)
    d_out = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    d_out_off = (torch.arange(nseg + 1, dtype=torch.int64, device=dev) * SEG)
    d_out_len = torch.zeros(nseg, dtype=torch.int64, device=dev)
    d_status = torch.zeros(nseg, dtype=torch.int32, device=dev)
    d_err_off = torch.zeros(nseg, dtype=torch.int64, device=dev)
    # frame assembly on GPU 0 (N > 1): CUDA IPC + copy-engine peer copies that run beside the inflate pass;
    # NCCL send/recv if IPC is unavailable on this box
    frame = None
    peer = None
    frame_cap = mg.frame_header_bytes(nseg * world) + world * dst_cap
    if world > 1:
        if os.environ.get("FB200_FRAME_TRANSPORT", "ipc") == "ipc":
            peer = mg.PeerFrame(ctx, rank, world, frame_cap, dev)
            if not peer.available:
                peer = None
        if peer is None and rank == 0:
            frame = torch.empty(frame_cap, dtype=torch.uint8, device=dev)
        config["frame_transport"] = ("cuda-ipc peer copies (copy engines over NVLink), overlapped with the inflate pass"
                                     if peer else "nccl send/recv")

    stage_acc = {}
    launches = [0]
    clen_box = [0]
    frame_len = [0]

    def step(record=True):
        clen = ctx.deflate_segments_dev(d_src.data_ptr(), nbytes, SEG, d_dst.data_ptr(), dst_cap, d_seg_off.data_ptr())
        clen_box[0] = clen
        if record:
            launches[0] += int(ctx.last_stats().kernel_launches)
            for k, v in ctx.last_stage_ms().items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v
        if world > 1:
            sizes = d_seg_off[1:] - d_seg_off[:-1]
            if peer:  # asynchronous: the copies run while this rank inflates (no device-wide sync here)
                frame_len[0] = peer.put(d_dst, sizes, SEG)
            else:
                frame_len[0] = mg.assemble_frame(d_dst, sizes, SEG, rank, world, frame)
                torch.cuda.synchronize()
        ctx.inflate_batch_dev(d_dst.data_ptr(), d_seg_off.data_ptr(), nseg, d_out.data_ptr(), d_out_off.data_ptr(),
                              d_out_len.data_ptr(), d_status.data_ptr(), d_err_off.data_ptr())
        if peer:
            peer.wait()  # every rank's payload is in GPU 0's frame before the step ends
        if record:
            launches[0] += int(ctx.last_stats().kernel_launches)
            stage_acc["inflate"] = stage_acc.get("inflate", 0.0) + ctx.last_stage_ms()["inflate"]

    # warm-up (untimed) + parity check of the resident result
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step(record=False)
    torch.cuda.synchronize()
    assert bool((d_status == 0).all()) and bool((d_out_len == SEG).all()), "inflate status"
    assert torch.equal(d_out, d_src), "round trip mismatch"
    clen = clen_box[0]
    if rank == 0:  # byte parity of a sample of GPU streams against the oracle (checker)
        from helpers import Oracle
        orc = Oracle()
        offs = d_seg_off.cpu().numpy()
        for i in list(range(0, nseg, max(1, nseg // 24)))[:24]:
            got = d_dst[int(offs[i]): int(offs[i + 1])].cpu().numpy().tobytes()
            assert got == orc.deflate(h_np[i * SEG:(i + 1) * SEG].tobytes()), f"segment {i} differs from the oracle"

    if world > 1 and rank == 0:  # the assembled frame: sample streams of every rank inflate to that rank's segments
        import zlib
        fr = peer.view if peer else frame
        seg_size, nseg_f, sizes_f, hdr_f = mg.parse_frame(fr)
        assert seg_size == SEG and nseg_f == nseg * world and hdr_f + int(sizes_f.sum()) == frame_len[0]
        offs_f = np.concatenate([[0], np.cumsum(sizes_f.numpy())]) + hdr_f
        for r in range(world):
            for i in (r * nseg, r * nseg + nseg // 2, (r + 1) * nseg - 1):
                comp_i = fr[int(offs_f[i]): int(offs_f[i + 1])].cpu().numpy().tobytes()
                want = corpus.fill(1, SEG, seed=args.seed, first=i).tobytes()
                assert zlib.decompress(comp_i, -15) == want, f"frame stream {i} (rank {r}) does not inflate to its segment"

    # timed region: events on the stream the kernels are launched on
    step(record=False)  # same load right before the timed region (also keeps the clocks up for the sampler)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record(lib_stream)
    for _ in range(args.steps):
        step(record=True)
    torch.cuda.synchronize()
    e1.record(lib_stream)
    e1.synchronize()
    sampler.mark_end()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    c_sum = torch.tensor([float(clen)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(c_sum, op=dist.ReduceOp.SUM)
    ms_step = float(t_ms.item()) / args.steps
    total_unc = nbytes * world
    value = 2 * total_unc / (ms_step * 1e-3) / 1e9

    # ------------------------------------------------------------------ e2e through the host-buffer C ABI
    e2e = None
    if not args.no_e2e:
        h_dst = torch.empty(dst_cap, dtype=torch.uint8, pin_memory=True)
        h_seg_off = torch.zeros(nseg + 1, dtype=torch.int64, pin_memory=True)
        h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        h_out_off = (torch.arange(nseg + 1, dtype=torch.int64) * SEG).pin_memory()
        h_out_len = torch.zeros(nseg, dtype=torch.int64, pin_memory=True)
        h_status = torch.zeros(nseg, dtype=torch.int32, pin_memory=True)
        h_err_off = torch.zeros(nseg, dtype=torch.int64, pin_memory=True)

        def e2e_step():
            cl = ctx.deflate_segments_ptr(h_src.data_ptr(), nbytes, SEG, h_dst.data_ptr(), dst_cap, h_seg_off.data_ptr())
            ctx.inflate_batch_ptr(h_dst.data_ptr(), h_seg_off.data_ptr(), nseg, h_out.data_ptr(), h_out_off.data_ptr(),
                                  h_out_len.data_ptr(), h_status.data_ptr(), h_err_off.data_ptr())
            return cl

        cl = e2e_step()
        assert int(h_status.abs().sum()) == 0 and torch.equal(h_out, h_src), "e2e round trip mismatch"
        ksteps = max(1, min(args.steps, 10))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        per_step = []
        t0 = time.perf_counter()
        for _ in range(ksteps):
            ts = time.perf_counter()
            cl = e2e_step()
            per_step.append(time.perf_counter() - ts)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / ksteps
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        meta = (nseg + 1) * 8
        e2e = {"value": round(2 * total_unc / float(t_e.item()) / 1e9, 3), "unit": "GB/s",
               "h2d_bytes_per_step": int((nbytes + cl + 3 * meta) * world),
               "d2h_bytes_per_step": int((cl + nbytes + meta + nseg * 20) * world),
               "steps": ksteps, "ms_per_step": round(float(t_e.item()) * 1e3, 3),
               "ms_per_step_min_max_rank0": [round(min(per_step) * 1e3, 3), round(max(per_step) * 1e3, 3)],
               "path": "fb200_deflate_segments + fb200_inflate_batch with pinned host buffers"}

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
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": traffic, "peak_source": peak_src,
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
            "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
            "compression_ratio": round(float(c_sum.item()) / total_unc, 5),
            "roofline": roofline, "clocks": clocks, "gpu_launches": launches[0]}
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(h_np, nseg)
    _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
