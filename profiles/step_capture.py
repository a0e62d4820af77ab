#!/usr/bin/env python
"""Exactly `--steps` Runge-Kutta steps of one workload inside a cudaProfilerStart/Stop range, for ncu:

    ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/<tag>_work_c5.csv \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum \
        python profiles/step_capture.py --workload c5 --steps 1

`python profiles/summarize.py work <csv> <workload> <cells> <steps>` then sums the counters over every launch of the
range and divides by cells x steps: DRAM bytes and fp64 warp instructions per cell-update, which bench.py reads from
profiles/traffic.json and profiles/fp64_work.json (`roofline.traffic`, `roofline.fp64.pipe_busy`).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from bench import WORKLOADS
    from astrea_b200.simulation import Simulation
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--cells", type=int, default=0)
    ap.add_argument("--steps", type=int, default=1)
    a = ap.parse_args()
    config, cells, dim, subgrid, solver, timestep, _ = WORKLOADS[a.workload]
    cells = a.cells or cells
    torch.cuda.set_device(0)
    sim = Simulation(config, cells, dim, subgrid, solver, timestep, device=0, cells_x=cells if dim == 2 else None)
    sim.set_time(0.0)
    sim.ctx.save_state()
    sim.step_async()                       # warm-up: kernel attributes, caches
    sim.restore_state()
    sim.set_time(0.0)
    sim.sync()
    torch.cuda.profiler.start()
    for _ in range(a.steps):
        sim.step_async()
    sim.sync()
    torch.cuda.profiler.stop()
    print("captured", a.steps, "steps of", a.workload, "cells", cells, "launches/step", None)
    sim.time()
    sim.close()


if __name__ == "__main__":
    main()
