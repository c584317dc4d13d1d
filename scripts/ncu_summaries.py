"""Turns the raw ncu outputs of a round into the summaries kept under profiles/.

  python scripts/ncu_summaries.py launches gpurun_out/rXX_launches.csv profiles/rXX_launch_list_summary.txt "<command line>"
  python scripts/ncu_summaries.py full gpurun_out/rXX_full.ncu-rep profiles/rXX_ncu_full_summary.json [profiles/traffic.json]

`launches`: the CSV log of `ncu --metrics gpu__time_duration.sum --clock-control none --csv`.
`full`: a `--set full` report; read with `ncu -i ... --page raw --csv` (needs ncu on PATH, no GPU).
"""
import csv, io, json, subprocess, sys
from collections import OrderedDict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def kernel_source_hash():
    """the same hash bench.py computes: sha256 over moonbit_flate_b200/csrc/*.{cu,cuh,h} of the tree the capture was made from"""
    import hashlib, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "moonbit_flate_b200", "csrc")
    h = hashlib.sha256()
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if r]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    acc = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        ms = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6) * v
        a = acc.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in acc.values())
    with open(dst, "w") as f:
        f.write(f"ncu launch list (gpu__time_duration.sum, --clock-control none) of: {cmd}\n")
        f.write("per-launch times are serialised and cold-cache; the SHARE is what must agree with bench.py's stage_ms\n")
        f.write(f"{'kernel':70} {'launches':>8} {'total ms':>10} {'ms/launch':>10} {'share':>7}\n")
        for k, (n, ms) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:70]:70} {n:8d} {ms:10.3f} {ms / n:10.4f} {100 * ms / tot:6.1f}%\n")
    print(open(dst).read())


def full(src, dst, traffic=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = OrderedDict()
        d["Kernel Name"] = r[hdr.index("Kernel Name")]
        us = {}
        for k in KEEP:
            if k in hdr:
                d[k] = r[hdr.index(k)]
                us[k] = units[hdr.index(k)]
        d["units"] = {k: us[k] for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum") if k in us}
        res.append(d)
    json.dump(res, open(dst, "w"), indent=1)
    print(json.dumps(res, indent=1))
    if traffic:
        def gb(d, k):
            v = float(d[k].replace(",", ""))
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[d["units"][k]]
        t = {}
        for d in res:
            name = "parse" if "k_parse<0>" in d["Kernel Name"] or "k_parse<(bool)0>" in d["Kernel Name"] else (
                "inflate" if "k_inflate_par" in d["Kernel Name"] else None)
            if name:
                t[name] = int(gb(d, "dram__bytes_read.sum") + gb(d, "dram__bytes_write.sum"))
        t["capture"] = f"ncu --set full ({dst}): dram__bytes_read.sum + dram__bytes_write.sum per launch, 16384 x 64 KiB mixed segments"
        t["kernel_source_sha256_16"] = kernel_source_hash()
        json.dump(t, open(traffic, "w"), indent=1)
        print(t)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
