"""Where the end-to-end step of the drop-in seam spends its time (host wall clock around synchronised calls).
python profiles/e2e_breakdown.py c2|c5  ->  one JSON line."""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from astrea_b200 import _native as N
from astrea_b200.simulation import Simulation

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
config, cells, dim, subgrid, solver, timestep, _ = bench.WORKLOADS[wl]
sim = Simulation(config, cells, dim, subgrid, solver, timestep, device=0, cells_x=cells)
ctx = sim.ctx
pool = N.PinnedPool(ctx.lib, 0)
pageable = ctx.download()
pinned = pool.empty(ctx.shape); pinned[...] = pageable
out = pool.empty(ctx.shape)
nbytes = pageable.nbytes

def timed(f, n=5):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))

import os
res = {"workload": wl, "bytes": nbytes, "host_cpus": os.cpu_count()}
rng = np.random.default_rng(1)
probe = rng.standard_normal(ctx.shape)
ctx.upload(probe)
back = ctx.download()
res["pageable_round_trip_identical"] = bool(np.array_equal(probe, back))
ctx.upload(pinned)
res["upload_pinned_ms"] = timed(lambda: ctx.upload(pinned))
res["upload_pageable_ms"] = timed(lambda: ctx.upload(pageable))
res["download_pinned_ms"] = timed(lambda: ctx.download(out=out))
res["download_pageable_ms"] = timed(lambda: ctx.download(out=pageable))
eig = [None]
def space(): ctx.upload(pinned); eig[0] = ctx.evolve_space(0)
t_us = timed(space)
res["evolve_space_ms"] = t_us - res["upload_pinned_ms"]
def both():
    ctx.upload(pinned); e = ctx.evolve_space(0); ctx.evolve_time(sim.cfl * min(sim.dx / x for x in e)); ctx.download(out=out)
res["seam_step_ms"] = timed(both)
res["evolve_time_ms"] = res["seam_step_ms"] - t_us - res["download_pinned_ms"]
# raw DMA rates of the box, one direction at a time
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
res["raw_h2d_GBps"] = nbytes / timed(lambda: d.copy_(h, non_blocking=True)) / 1e6
res["raw_d2h_GBps"] = nbytes / timed(lambda: h.copy_(d, non_blocking=True)) / 1e6
res["pool_empty_ms"] = timed(lambda: pool.empty(ctx.shape))
print(json.dumps(res))
