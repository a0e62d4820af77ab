#!/usr/bin/env python
"""Benchmark of the per-timestep finite-volume update (BASELINE.json metric: cell-updates/s, fp64, per RK step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4|c5]

A "step" is one full Runge-Kutta step (all stages) of every cell of the workload.  Default workload, for every N, is
the configuration the north star quotes its target on, BASELINE config 5: Lax-Liu 6, 8192^2 PER GPU, PPM + HLLC,
SSPRK(3,3), periodic (the largest single-GPU configuration: 47 GB of the 180 GB HBM).  With N > 1 (torchrun, one rank
per GPU) every rank holds a slab of that size (weak scaling) of a global (N*8192) x 8192 periodic grid; ranks exchange
ghost rows over NVLink before every spatial operator and reduce the wave speeds once per step.  The other BASELINE
configurations are behind --workload.

Before anything is timed the run checks itself (``parity_check`` in the JSON line): the N ranks advance the 64^2
golden case of config 5 (tests/golden/, output of the unmodified reference) as N slabs with the halo exchange of the
timed path, and a 256^2 Lax-Liu 3 grid whose assembled result must equal, bit for bit, the same grid advanced by one
GPU.  A failed check ends the run with a non-zero exit code.

The reference's own run of this configuration stops with LinAlgError after a few steps (the horizon
is re-measured here at run time), so the timed loop returns to the initial state every `horizon` steps with a
device-to-device copy that is inside the timed region.

`--impl reference` times the CPU implementation of the same path (the numpy oracle, a restatement pinned bit for
bit to the reference; the reference itself does not exist on the GPU box) on all host cores as independent
single-threaded replicas, on a bounded sample (the reference is single-threaded and cannot hold 2048^2 in memory).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (config, cells per GPU side, dimension, subgrid, solver, timestep, description)
    "c1": ("sod", 1024, 1, "plm", "lf", "ssprk(2,2)", "1D Sod 1024 PLM+minmod+LLF SSPRK(2,2) edge"),
    "c2": ("ll3", 2048, 2, "ppm", "hllc", "ssprk(3,3)", "2D Lax-Liu 3 2048^2 PPM+HLLC SSPRK(3,3) periodic"),
    "c3": ("khi", 4096, 2, "weno5", "hllc", "ssprk(3,3)", "2D Kelvin-Helmholtz 4096^2 WENO5+HLLC SSPRK(3,3) periodic"),
    "c4": ("orszag-tang", 4096, 2, "plm", "hlld", "ssprk(3,3)", "2D MHD Orszag-Tang 4096^2 PLM+HLLD + constrained transport SSPRK(3,3) periodic"),
    "c5": ("ll6", 8192, 2, "ppm", "hllc", "ssprk(3,3)", "2D Lax-Liu 6 8192^2 per GPU PPM+HLLC SSPRK(3,3) periodic"),
}
MAX_HORIZON = 8


def alg_bytes_per_cell_update(stages):
    """64 B of state per cell; stage 1 reads u and writes k, later stages read u, k and write k'."""
    return 64 * (2 + 3 * (stages - 1))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def _oracle_sample(args):
    """One bounded sample on one core: `steps` full RK steps of the workload's scheme on a cells^2 (or cells) grid."""
    config, cells, dim, subgrid, solver, timestep, steps, eigen = args
    from astrea_b200.initial import initial_state, problem
    from oracle import OracleConfig, advance
    prob = problem(config, cells, 1.4)
    from astrea_b200.selectors import MAGNETIC_2D
    cfg = OracleConfig(config=config, cells=cells, dimension=dim, subgrid=subgrid, solver=solver, timestep=timestep,
                       boundary=prob["boundary"], dx=prob["dx"], eigen=eigen, magnetic_2d=config in MAGNETIC_2D)
    g0 = initial_state(config, cells, dim, 1.4, cfg.high_order)
    t0 = time.perf_counter()
    g, dts = advance(g0, cfg, steps)
    return time.perf_counter() - t0, bool(np.isfinite(g).all())


def cpu_sample_spec(workload, cells):
    config, _, dim, subgrid, solver, timestep, _ = WORKLOADS[workload]
    steps = 2 if dim == 2 else 200
    return (config, cells, dim, subgrid, solver, timestep, steps, "lapack")


def run_reference_arm(a):
    """The CPU implementation of the path on all host cores (independent single-threaded replicas)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    config, _, dim, subgrid, solver, timestep, desc = WORKLOADS[a.workload]
    cells = 128 if dim == 2 else 1024
    spec = cpu_sample_spec(a.workload, cells)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    cells_per_sample = cells ** dim * spec[6]
    with mp.get_context("spawn").Pool(cores) as pool:
        for _ in range(a.warmup):
            pool.map(_oracle_sample, [spec] * cores)
        times = []
        for _ in range(a.steps):
            t0 = time.perf_counter()
            pool.map(_oracle_sample, [spec] * cores)
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = cores * cells_per_sample * a.steps / total
    sample = (f"{cores} replicas x ({config} {cells}{'^2' if dim == 2 else ''} {subgrid}+{solver} {timestep}, {spec[6]} RK steps) per bench step; "
              "numpy oracle pinned bit-for-bit to the reference (np.linalg.eigvals wave speeds)")
    line = {"impl": "reference", "metric": "cell-updates/sec (fp64, per RK step)", "value": value, "unit": "cell-updates/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "cpu_sample_cells": cells, "note": "reference numpy path is single-threaded; all cores used as replicas"},
            "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    OUT.emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------ self-check
def _rel_l1(a, b):
    axes = tuple(range(a.ndim - 1))
    den = np.abs(b).sum(axis=axes)
    return float(np.max(np.abs(a - b).sum(axis=axes) / np.where(den > 0, den, 1)))


def parity_check(world, rank, local):
    """Correctness evidence inside the benchmark run, through the same Simulation / halo-exchange path that is timed.

    golden_c5_64:  BASELINE config 5 at 64^2 (tests/golden/c5_ll6_ppm_hllc_ssprk33.npz: initial grid, dt sequence and
                   final grid of the unmodified reference), advanced as `world` slabs with the reference's dt sequence;
                   tolerance 1e-10 relative L1 per variable after its 4 steps (north star).
    slabs_256:     Lax-Liu 3 at 256^2, PPM + HLLC, SSPRK(3,3), 2 steps with the device-side dt, as `world` slabs against
                   the same grid on one GPU (rank 0); must be bit-identical, dt sequence included.
    With world == 1 the second check compares the asynchronous device-clock path with the synchronous one."""
    import torch
    import torch.distributed as dist
    from astrea_b200.initial import initial_state
    from astrea_b200.simulation import Simulation
    dev = torch.device("cuda", local)

    def gather(slab):
        if world == 1:
            return slab
        mine = torch.from_numpy(np.ascontiguousarray(slab)).to(dev)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        return np.concatenate([p.cpu().numpy() for p in parts], axis=0)

    out = {}
    data = np.load(os.path.join(ROOT, "tests", "golden", "c5_ll6_ppm_hllc_ssprk33.npz"))
    g0, want, dts = data["g0"], data["g"], [float(d) for d in data["dts"]]
    cells = g0.shape[0]
    if cells % world == 0 and cells // world >= 8:
        rows = cells // world
        sim = Simulation("ll6", cells, 2, "ppm", "hllc", "ssprk(3,3)", device=local, rank=rank, world=world, cells_x=rows,
                         grid=g0[rank * rows:(rank + 1) * rows])
        for dt in dts:
            sim.step(dt=dt)
        got = gather(sim.state())
        sim.close()
        err = _rel_l1(got, want)
        out["golden_c5_64"] = {"max_rel_l1": err, "bit_identical": bool(np.array_equal(got, want)), "steps": len(dts),
                               "tolerance": 1e-10, "ok": bool(err <= 1e-10)}
    cells, steps = 256, 2
    g0 = initial_state("ll3", cells, 2, 1.4, True)
    rows = cells // world
    sim = Simulation("ll3", cells, 2, "ppm", "hllc", "ssprk(3,3)", device=local, rank=rank, world=world, cells_x=rows,
                     grid=g0[rank * rows:(rank + 1) * rows])
    sim.set_time(0.0)
    for _ in range(steps):
        sim.step_async()
    sim.time()
    got, got_dts = gather(sim.state()), sim.ctx.dt_history(steps)
    sim.close()
    if rank == 0:
        one = Simulation("ll3", cells, 2, "ppm", "hllc", "ssprk(3,3)", device=local, grid=g0)
        want_dts = one.run(steps)
        want = one.state()
        one.close()
        same = bool(np.array_equal(got, want)) and got_dts == want_dts
        out["slabs_256"] = {"max_rel_l1": _rel_l1(got, want), "bit_identical": same, "steps": steps, "ranks": world, "ok": same}
    ok = all(v["ok"] for v in out.values())
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.broadcast(flag, src=0)
        ok = bool(flag.item())
    out["ok"] = ok
    return out


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    from astrea_b200 import _native as N
    from astrea_b200 import evolvers
    from astrea_b200.simulation import Simulation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torchrun (one rank per GPU), see the module docstring")
    # host threads and page-locked arrays of this rank on the socket its GPU hangs off (before anything is allocated)
    from astrea_b200.hostbind import bind_host_to_gpu
    host_binding = {"bound": False, "why": "--no-host-bind"} if a.no_host_bind else bind_host_to_gpu(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N.device_library()       # fail loudly if the CUDA library is missing
    check = None if a.no_parity_check else parity_check(world, rank, local)
    if check is not None and not check["ok"]:
        if rank == 0:
            OUT.emit(json.dumps({"metric": "cell-updates/sec (fp64, per RK step)", "value": None, "n_gpus": world,
                                 "parity_check": check, "error": "parity self-check failed: nothing was timed"}))
        raise SystemExit(1)

    config, cells, dim, subgrid, solver, timestep, desc = WORKLOADS[a.workload]
    if a.cells:
        cells = a.cells
    from astrea_b200.selectors import integrator_enum, stages_of
    stages = stages_of(integrator_enum(timestep))
    geometry = {}
    if a.threads_2d:
        geometry["threads_2d"] = a.threads_2d
    if a.segment_2d:
        geometry["segment_2d"] = a.segment_2d
    if a.no_recon_bulk:
        geometry["recon_bulk"] = False
    if a.flux_block_tile is not None:
        geometry["flux_block_tile"] = bool(a.flux_block_tile)
    if a.stage_speeds:
        geometry["stage_speeds"] = True
    sim = Simulation(config, cells, dim, subgrid, solver, timestep, device=local, rank=rank, world=world,
                     cells_x=cells if dim == 2 else None, overlap=a.overlap, **geometry)
    ctx = sim.ctx
    stream = torch.cuda.ExternalStream(ctx.stream_handle, device=local)
    cells_per_rank = cells ** dim
    ctx.save_state()

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # finite horizon of this configuration (the reference raises LinAlgError beyond it)
    horizon = 0
    try:
        for _ in range(MAX_HORIZON):
            sim.step()
            sim.check_finite()       # a non-finite wave speed in any stage ends the reference's run (fv.py:158)
            horizon += 1
    except np.linalg.LinAlgError:
        pass
    if world > 1:
        h = torch.tensor([horizon], device=f"cuda:{local}")
        dist.all_reduce(h, op=dist.ReduceOp.MIN)
        horizon = int(h.item())
    if horizon == 0:
        raise SystemExit("the workload produces non-finite wave speeds in its first step")
    sim.restore_state()

    # a configuration that never went non-finite (the 1D Sod tube) needs no return to the initial state: batches of 256
    batch = 256 if (horizon == MAX_HORIZON and dim == 1) else horizon

    def run_steps(n):
        # steps are enqueued back to back: dt = cfl*min(dx/eigmax) is evaluated on the device (astrea_step_async;
        # small grids replay each step as one CUDA graph)
        done = 0
        while done < n:
            sim.restore_state()
            sim.set_time(0.0)
            k = min(batch, n - done)
            sim.run_steps(k)
            done += k

    run_steps(a.warmup)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run_steps(a.steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    sim.time()                   # raises if any timed step saw a non-finite wave speed
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count - launches0
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * cells_per_rank * a.steps / (ms * 1e-3)

    # per-kernel-class device time of `horizon` steps (CUDA events around every launch, on the launching stream)
    sim.restore_state()
    ctx.profile(True)
    sim.set_time(0.0)
    for _ in range(horizon):
        sim.step_async()
    prof = ctx.profile_read()
    ctx.profile(False)
    total_ms = sum(v[0] for v in prof.values())
    peak, peak_src = load_peaks()
    alg = alg_bytes_per_cell_update(stages)
    sweeps_per_step = stages * dim
    ms_step = ms / a.steps
    # Whole step against the HBM roofline: SURVEY 8(d) algorithmic bytes of one cell-update x cells of this GPU / the
    # measured step time (max over ranks).  This is `frac`.  The per-class split below comes from CUDA events around
    # every launch of `horizon` further steps on the launching stream (astrea_profile).
    step_achieved = cells_per_rank * alg / (ms_step * 1e-3) / 1e9
    sweep_ms = prof["flux"][0] + prof["prim"][0] + prof["recon"][0]
    sweep_n = max(1, prof["flux"][1])
    bytes_per_sweep = cells_per_rank * alg / sweeps_per_step
    flux_avg_ms = prof["flux"][0] / max(1, prof["flux"][1])
    roofline = {"bound": "hbm", "achieved": step_achieved, "peak": peak, "unit": "GB/s", "frac": step_achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "kernel": "whole Runge-Kutta step (all kernels of `stages` operator evaluations and register updates)",
                "alg_bytes_per_cell_update": alg, "alg_bytes_per_launch": cells_per_rank * alg, "kernel_avg_ms": ms_step,
                "class_ms_per_step": {k: v[0] / horizon for k, v in prof.items()},
                "class_launches_per_step": {k: v[1] / horizon for k, v in prof.items()},
                "dominant_kernel": {"name": "FluxStage" if dim == 2 else "Sweep1D", "avg_ms": flux_avg_ms,
                                    "share_of_step": prof["flux"][0] / total_ms if total_ms else None},
                # the sweep-only figure of round 1 (all algorithmic bytes of a sweep over the time of its three stages,
                # register update excluded) is kept for comparison only
                "sweep_only": {"kernels": "PrimBothStage/2 + ReconStage + FluxStage" if dim == 2 else "Sweep1D",
                               "alg_bytes": bytes_per_sweep, "avg_ms": sweep_ms / sweep_n,
                               "frac": bytes_per_sweep / (sweep_ms / sweep_n * 1e-3) / 1e9 / peak if sweep_ms else None},
                "note": "the path is bound by the fp64 pipe, not by HBM (see the fp64 object and DESIGN.md)"}
    # The bound that really applies: the fp64 pipe.  FMA rate of this GPU measured now (astrea_fp64_probe), and the
    # fp64 instructions a step executes (ncu count per cell-update of these kernels, profiles/fp64_work.json, produced
    # by profiles/summarize.py from the committed capture) -> how busy the pipe is at the measured step time.
    sim.restore_state()
    tflops = ctx.fp64_probe()
    fp64 = {"peak_tflops_measured": tflops}
    work_file = os.path.join(ROOT, "profiles", "fp64_work.json")
    if os.path.exists(work_file) and tflops > 0:
        with open(work_file) as fh:
            work = json.load(fh).get(a.workload)
        if work:
            warp_inst = work["fp64_warp_inst_per_cell_update"] * cells_per_rank      # per step
            rate = tflops * 1e12 / 2 / 32                                              # warp-level fp64 instructions per second
            floor_ms = warp_inst / rate * 1e3
            fp64.update({"fp64_warp_inst_per_cell_update": work["fp64_warp_inst_per_cell_update"],
                         "pipe_floor_ms_per_step": floor_ms, "pipe_busy": floor_ms / ms_step, "count_source": work.get("source")})
    roofline["fp64"] = fp64
    # DRAM bytes a step really moves, per launch like `achieved`: ncu --set full capture of these kernels
    # (profiles/traffic.json, written by profiles/summarize.py), launches per step counted live
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as fh:
            tr = json.load(fh).get(a.workload)
        if tr and not a.cells and tr.get("bytes_per_cell_update"):
            roofline["traffic"] = tr["bytes_per_cell_update"] * cells_per_rank
            roofline["traffic_per_cell_update"] = tr["bytes_per_cell_update"]
            roofline["traffic_source"] = tr.get("source")

    # end to end through the reference-facing calls, host buffers in and out every step
    e2e = None if a.no_e2e else measure_e2e(a, sim, horizon, world, rank, local, stream, barrier)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        spec = cpu_sample_spec(a.workload, 256 if dim == 2 else 1024)
        secs, finite = _oracle_sample(spec)
        n = spec[1] ** dim * spec[6]
        cpu = {"value": n / secs, "unit": "cell-updates/s", "cores": 1, "kind": "port",
               "sample": f"{spec[0]} {spec[1]}{'^2' if dim == 2 else ''} {spec[3]}+{spec[4]} {spec[5]}, {spec[6]} RK steps, {secs:.1f} s, "
                         "numpy oracle (bit-identical to the reference, single-threaded like it)"}

    if rank == 0:
        line = {"metric": "cell-updates/sec (fp64, per RK step)", "value": value, "unit": "cell-updates/s", "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc if not a.cells else desc + f" (cells overridden: {cells})", "cells_per_gpu": cells_per_rank,
                           "global_cells": world * cells_per_rank, "stages_per_step": stages, "decomposition": f"x-slabs x{world}",
                           "halo_exchange": "none (one GPU)" if world == 1 else ("NCCL send/recv on a second stream, overlapped with the interior rows of the register update" if a.overlap else "NCCL send/recv in order on the compute stream"),
                           "finite_horizon_steps": horizon, "dt": "computed on the device every step (cfl*min(dx/eigmax)), no host round trip",
                           "l2": "state per register (%.0f MB) exceeds the 126 MB L2" % (cells_per_rank * 64 / 1e6)
                                 if cells_per_rank * 64 > 126e6 else "working set fits L2 (small workload)",
                           "restore": f"device-to-device return to the initial state every {batch} steps, inside the timed region",
                           "host_binding": host_binding},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "parity_check": check}
        if cpu:
            line["cpu_baseline"] = cpu
        OUT.emit(json.dumps(line))
    sim.close()
    evolvers.release()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_e2e(a, sim, horizon, world, rank, local, stream, barrier):
    """Same metric with the grid crossing the host/device boundary every step, the way the reference's loop would drive
    the drop-in: ``grid`` starts as an ordinary (pageable) numpy array, as constructor.initialise returns it
    (astrea.py:35), every later ``grid`` is the array the previous evolve_time returned (astrea.py:81), which the
    drop-in allocates page-locked (evolvers.py, _native.PinnedPool).  Every `horizon` steps the loop starts again from
    the pageable initial array.  N = 1: the pair evolvers.evolve_space / evolve_time (astrea.py:67,81).  N > 1: upload /
    step / download of each rank's slab with the same memory kinds."""
    import torch
    import torch.distributed as dist
    from collections import namedtuple
    from astrea_b200 import _native as N
    from astrea_b200 import evolvers
    ctx = sim.ctx
    shape = tuple(ctx.shape)
    sim.restore_state()
    ic_np = ctx.download()                     # np.empty: pageable
    nbytes = ic_np.nbytes
    steps = horizon
    if world == 1:
        SV = namedtuple("simulation_variables", "dimension cells boundary gamma dx cfl subgrid solver solver_category timestep magnetic_2d permutations")
        config, cells, dim, subgrid, solver, timestep, _ = WORKLOADS[a.workload]
        cells = a.cells or cells
        perms = {0: (0, 1)} if dim == 1 else {0: (0, 1, 2), 1: (1, 0, 2)}
        sv0 = SV(dim, cells, sim.boundary, sim.gamma, sim.dx, sim.cfl, subgrid, solver, "hll" if solver.startswith("hll") else "lax",
                 timestep, sim.magnetic_2d, perms)

        def one_pass(n_steps):
            sv, grid = sv0, ic_np
            for n in range(n_steps):
                fluxes = evolvers.evolve_space(grid, sv, device=local)
                dt = sv.cfl * min(sv.dx / f["eigmax"] for f in fluxes.values())
                grid = evolvers.evolve_time(grid, fluxes, dt, sv, device=local)
                sv = sv._replace(permutations=dict(reversed(list(sv.permutations.items()))))
            return grid
        one_pass(min(2, steps))         # warm-up: creates the context and the first pinned arrays
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        one_pass(steps)
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
    else:
        pool = N.PinnedPool(ctx.lib, local)

        def one_pass(n_steps):
            grid = ic_np
            ctx.parity = 0
            for n in range(n_steps):
                sim._halo_ready = False            # a fresh upload: ghost rows are stale
                ctx.upload_ptr(grid.ctypes.data)
                sim.step()
                grid = ctx.download(out=pool.empty(shape))
            return grid
        one_pass(min(2, steps))
        barrier()
        t0 = time.perf_counter()
        one_pass(steps)
        barrier()
        secs = time.perf_counter() - t0
        pool.close()
        t = torch.tensor([secs], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    cells_per_rank = int(np.prod(shape[:-1]))
    return {"value": world * cells_per_rank * steps / secs, "unit": "cell-updates/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
            "steps": steps, "ms_per_step": 1e3 * secs / steps,
            "host_memory": "step 1 of every pass uploads a pageable numpy array; later steps upload the page-locked array the "
                           "previous evolve_time returned; every download lands in a fresh page-locked array",
            "api": "evolvers.evolve_space + evolvers.evolve_time (numpy in / numpy out)" if world == 1
            else "Context.upload + Simulation.step + Context.download per rank", "timer": "host wall clock around synchronised calls"}


class StdoutToStderr:
    """Libraries (NCCL's version banner) write to fd 1; the contract is ONE JSON line on stdout.  Everything but
    the final line goes to stderr."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


OUT = None


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c5")
    ap.add_argument("--cells", type=int, default=0, help="override the cells per side (per GPU)")
    ap.add_argument("--threads-2d", type=int, default=0)
    ap.add_argument("--segment-2d", type=int, default=0)
    ap.add_argument("--no-recon-bulk", action="store_true", help="A/B: reconstruction march with register prefetch instead of cp.async.bulk")
    ap.add_argument("--stage-speeds", action="store_true", help="A/B: evaluate the interface wave speeds in every operator of a step")
    ap.add_argument("--flux-block-tile", type=int, default=None, help="A/B: 0 = warp-wide, 1 = block-wide rows in the flux stage (default: by grid width)")
    ap.add_argument("--no-host-bind", action="store_true", help="A/B: leave the rank's CPU affinity alone (default: the CPUs local to its GPU)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the self-check before the timed region (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--overlap", action="store_true", help="multi-GPU: exchange ghost rows behind the register update (measured slower, see Simulation)")
    a = ap.parse_args()
    with StdoutToStderr() as OUT:
        if a.impl == "reference":
            run_reference_arm(a)
        else:
            run_ours(a)


if __name__ == "__main__":
    main()
