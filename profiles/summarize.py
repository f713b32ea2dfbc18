"""Turn the ncu CSVs of profiles/capture.sh into the small files that are committed:
    python profiles/summarize.py r01 [gpurun_out]
  launches_<R>.csv      -> launch_shares_<R>.txt   (time per kernel over one short bench run)
  <kernel>_<R>_raw.csv  -> ncu_summary_<R>.json    (the handful of metrics DESIGN.md / bench.py quote)
"""
import collections
import csv
import json
import os
import re
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
SRC = sys.argv[2] if len(sys.argv) > 2 else os.path.dirname(os.path.abspath(__file__))
OUT = os.path.dirname(os.path.abspath(__file__))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
STALLS = "smsp__average_warps_issue_stalled_"


def read_rows(path):
    with open(path, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    return list(csv.reader(lines))


def launch_shares():
    p = os.path.join(SRC, f"launches_{R}.csv")
    if not os.path.exists(p):
        return
    rows = read_rows(p)
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    t = collections.defaultdict(float)
    n = collections.Counter()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        name = re.sub(r"\(.*", "", r[ik]).replace("ska::", "").strip()
        t[name] += v
        n[name] += 1
    tot = sum(t.values())
    with open(os.path.join(OUT, f"launch_shares_{R}.txt"), "w") as fh:
        fh.write(f"launch list (bench.py --steps 2 --warmup 1 --no-cpu-baseline under ncu): total kernel time {tot:.1f} ms\n")
        for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
            fh.write(f"  {k:44s} x{n[k]:3d} {v:9.3f} ms {100 * v / tot:5.1f}%\n")


def kernel_summaries():
    out = {}
    for f in sorted(os.listdir(SRC)):
        m = re.match(rf"(.+)_{R}_raw\.csv$", f)
        if not m or m.group(1) == "kernels":
            continue
        rows = read_rows(os.path.join(SRC, f))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP or h == "Kernel Name" or (h.startswith(STALLS) and h.endswith("_per_issue_active.ratio")):
                d[h] = [v, u]
        out[m.group(1)] = d
    # round 2: one raw CSV with every kernel of a build + align (capture.sh with ONE_PROCESS=1), one row per launch
    multi = os.path.join(SRC, f"kernels_{R}_raw.csv")
    if os.path.exists(multi):
        rows = read_rows(multi)
        hdr, units = rows[0], rows[1]
        ik = hdr.index("Kernel Name")
        seen = collections.Counter()
        for vals in rows[2:]:
            if len(vals) <= ik:
                continue
            base = re.sub(r"^void\s+", "", vals[ik]).split("(")[0].replace("ska::", "")
            short = re.sub(r"_kernel.*", "", base)
            tmpl = re.search(r"_kernel(<.*>)", base)
            seen[base] += 1
            name = short + (tmpl.group(1) if tmpl else "") + (f"#{seen[base]}" if seen[base] > 1 else "")
            d = {}
            for h, u, v in zip(hdr, units, vals):
                if h in KEEP or h == "Kernel Name" or (h.startswith(STALLS) and h.endswith("_per_issue_active.ratio")):
                    d[h] = [v, u]
            out[name] = d
    if out:
        with open(os.path.join(OUT, f"ncu_summary_{R}.json"), "w") as fh:
            json.dump(out, fh, indent=1, sort_keys=True)


launch_shares()
kernel_summaries()
