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


def work(path, workload, cells, steps):
    """Counters of every launch inside the profiled range of profiles/step_capture.py -> per cell-update figures."""
    lines = [l for l in open(path) if not l.startswith("==")]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "inst": 1, "ns": 1e-6, "us": 1e-3, "ms": 1.0}
    tot = collections.defaultdict(float)
    per_kernel = collections.OrderedDict()
    ids = set()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", "")) * scale.get(row["Metric Unit"], 1)
        tot[row["Metric Name"]] += v
        k = short(row["Kernel Name"])
        per_kernel.setdefault(k, collections.defaultdict(float))[row["Metric Name"]] += v
        ids.add(row["ID"])
    n = float(cells) * float(steps)
    src = f"ncu counters of {steps} step(s) at {cells} cells, {os.path.basename(path)} ({len(ids)} launches)"
    for name, key, value in (("traffic.json", "bytes_per_cell_update", (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / n),
                             ("fp64_work.json", "fp64_warp_inst_per_cell_update", tot["smsp__inst_executed_pipe_fp64.sum"] / n)):
        fpath = os.path.join(HERE, name)
        store = json.load(open(fpath)) if os.path.exists(fpath) else {}
        store.setdefault(workload, {})
        store[workload][key] = value
        store[workload]["source"] = src
        if name == "fp64_work.json":
            store[workload]["warp_inst_per_cell_update"] = tot["smsp__inst_executed.sum"] / n
        json.dump(store, open(fpath, "w"), indent=1, sort_keys=True)
    out = ["| kernel | ms | DRAM MB | fp64 warp inst (M) | warp inst (M) |", "|---|---|---|---|---|"]
    for k, m in sorted(per_kernel.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        out.append(f"| {k} | {m['gpu__time_duration.sum']:.3f} | {(m['dram__bytes_read.sum'] + m['dram__bytes_write.sum']) / 1e6:.1f} | "
                   f"{m['smsp__inst_executed_pipe_fp64.sum'] / 1e6:.2f} | {m['smsp__inst_executed.sum'] / 1e6:.2f} |")
    out.append(f"\nper cell-update: DRAM {(tot['dram__bytes_read.sum'] + tot['dram__bytes_write.sum']) / n:.1f} B, fp64 warp instructions "
               f"{tot['smsp__inst_executed_pipe_fp64.sum'] / n:.3f}, all warp instructions {tot['smsp__inst_executed.sum'] / n:.3f}; "
               f"device time {tot['gpu__time_duration.sum'] / float(steps):.3f} ms/step under ncu ({src})")
    return "\n".join(out)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        print(launches(sys.argv[2]))
    elif mode == "work":
        print(work(sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])))
    else:
        traffic = None
        if "--traffic" in sys.argv:
            k = sys.argv.index("--traffic")
            traffic = (sys.argv[k + 1], sys.argv[k + 2])
        print(full(sys.argv[2], traffic))
