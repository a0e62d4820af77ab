#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the tables of profiles/README.md and profiles/traffic.json.

    python profiles/summarize.py launches gpurun_out/launches_r1g.csv            # launch list -> markdown table
    python profiles/summarize.py full gpurun_out/prof_r1g.ncu-rep [--traffic c2 "FluxStage"]   # --set full capture

`full` calls `ncu -i <rep> --page raw --csv` (works without a GPU).  With --traffic WORKLOAD KERNEL it also records
dram__bytes_read.sum + dram__bytes_write.sum of the first launch whose name contains KERNEL into
profiles/traffic.json, which bench.py reports as roofline.traffic.
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def short(name):
    return name.replace("void kernel_entry<", "").replace(">(Params)", "").replace("astrea::", "")


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v, u = float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
        ms = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
        k = short(row["Kernel Name"])
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + ms)
    total = sum(t for _, t in agg.values())
    out = ["| kernel | launches | total ms | avg us | share |", "|---|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {n} | {t:.3f} | {t / n * 1e3:.1f} | {100 * t / total:.1f}% |")
    out.append(f"\ntotal device time in the capture: {total:.2f} ms")
    return "\n".join(out)


def full(rep, traffic=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    names = [short(d[idx["Kernel Name"]]) for d in data]
    out = ["| metric | unit | " + " | ".join(names) + " |", "|---|---|" + "---|" * len(names)]
    for m in METRICS:
        if m in idx:
            label = m.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", "")
            out.append(f"| {label} | {units[idx[m]]} | " + " | ".join(d[idx[m]] for d in data) + " |")
    if traffic:
        workload, kernel = traffic
        for d, n in zip(data, names):
            if kernel in n:
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd = float(d[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
                wr = float(d[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
                path = os.path.join(HERE, "traffic.json")
                store = json.load(open(path)) if os.path.exists(path) else {}
                store.setdefault(workload, {})
                store[workload].setdefault("kernels", {})[n] = {"dram_bytes_read": rd, "dram_bytes_write": wr}
                store[workload]["source"] = f"ncu --set full, {os.path.basename(rep)}"
                json.dump(store, open(path, "w"), indent=1, sort_keys=True)
                break
    return "\n".join(out)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        print(launches(sys.argv[2]))
    else:
        traffic = None
        if "--traffic" in sys.argv:
            k = sys.argv.index("--traffic")
            traffic = (sys.argv[k + 1], sys.argv[k + 2])
        print(full(sys.argv[2], traffic))
