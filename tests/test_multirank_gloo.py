"""Slab decomposition over two processes (gloo, CPU): the host logic of the multi-GPU path.

Each rank owns half of the rows, exchanges ghost rows before every spatial-operator evaluation and all-reduces
the wave speeds; the assembled result must equal the single-process result bit for bit.  The native side is the
host-simulated build of the kernels (test infrastructure) so that this runs without GPUs; on the B200 box the
same driver runs over NCCL (tests/test_gpu_parity.py, bench.py --gpus N).
"""
import os
import tempfile

import numpy as np
import pytest


def _cells(spec, world):
    return spec[1] // world * world


def _run_spec(rank, world, spec, outdir, lib):
    from astrea_b200.simulation import Simulation
    config, _, subgrid, solver, timestep, bc, steps = spec[:7]
    cells = _cells(spec, world)
    author = spec[7] if len(spec) > 7 else "mc"
    full = np.load(os.path.join(outdir, "g0.npy"))
    rows = cells // world
    sim = Simulation(config, cells, 2, subgrid, solver, timestep, boundary=bc, rank=rank, world=world, cells_x=rows,
                     grid=full[rank * rows:(rank + 1) * rows], overlap=(world == 3), _lib=lib, threads_2d=32, segment_2d=9,
                     ppm_author=author)
    dts = sim.run(steps)
    np.save(os.path.join(outdir, f"slab{rank}.npy"), sim.state())
    np.save(os.path.join(outdir, f"dts{rank}.npy"), np.array(dts))
    # the same steps again through the asynchronous path (device-side dt after an in-place all-reduce)
    sim.upload(full[rank * rows:(rank + 1) * rows])
    sim.ctx.parity = 0
    sim.set_time(0.0)
    for _ in range(steps):
        sim.step_async()
    t, n, last = sim.time()
    assert n == steps and sim.ctx.dt_history(steps) == dts, (sim.ctx.dt_history(steps), dts)
    np.save(os.path.join(outdir, f"aslab{rank}.npy"), sim.state())
    np.save(os.path.join(outdir, f"snap{rank}.npy"), sim.snapshot())         # astrea.py:47 of this rank's rows: (ny, rows, 8)
    sim.close()


def _worker(rank, world, rendezvous, specs, root):
    """One process of a group that runs every spec in turn (a process start costs more than the steps of a spec)."""
    import traceback
    import torch.distributed as dist
    from astrea_b200 import _native, build
    dist.init_process_group("gloo", init_method="file://" + rendezvous, rank=rank, world_size=world)
    lib = _native.bind(build.HOSTSIM_LIB)
    for k, spec in enumerate(specs):
        outdir = os.path.join(root, f"spec{k}")
        failed = 0
        try:
            _run_spec(rank, world, spec, outdir, lib)
        except Exception:
            failed = 1
            with open(os.path.join(outdir, f"error{rank}.txt"), "w") as f:
                f.write(traceback.format_exc())
        # a rank that failed may have left its peers inside a collective of that spec: the group is only usable for the
        # next spec if every rank got through
        import torch
        flag = torch.tensor([failed])
        dist.all_reduce(flag)
        if int(flag.item()):
            for later in range(k + 1, len(specs)):
                with open(os.path.join(root, f"spec{later}", f"error{rank}.txt"), "w") as f:
                    f.write(f"not run: spec {k} failed on some rank")
            break
    dist.barrier()
    dist.destroy_process_group()


SPECS = [("ll3", 32, "ppm", "hllc", "ssprk(3,3)", "wrap", 2),
         ("ll4", 32, "ppm", "lf", "ssprk(3,3)", "edge", 2),
         ("khi", 32, "weno5", "hllc", "ssprk(2,2)", "wrap", 3),
         ("ll12", 36, "plm", "lf", "rk4", "edge", 2),
         ("ll3", 32, "weno7", "hllc", "ssprk(5,4)", "wrap", 1),
         # constrained transport on slabs: face states recomputed in the ghost rows, refine_grid behind its own exchange
         ("orszag-tang", 32, "plm", "hlld", "ssprk(3,3)", "wrap", 3),
         ("orszag-tang", 36, "ppm", "hlld", "ssprk(2,2)", "edge", 2),
         ("mhd rotor", 32, "weno5", "hllc", "ssprk(3,3)", "wrap", 2),
         ("orszag-tang", 32, "pcm", "hlld", "ssprk(10,4)", "wrap", 1),
         # PPM authors 'c' / 'ph': the grid-wide any() switches are OR-ed across the slabs (astrea_set_flag_reducer)
         ("ll3", 32, "ppm", "hllc", "ssprk(2,2)", "wrap", 2, "c"), ("khi", 36, "ppm", "lf", "ssprk(3,3)", "wrap", 2, "ph"),
         ("ll4", 32, "ppm", "lf", "euler", "edge", 2, "ph"), ("sod", 32, "ppm", "hllc", "ssprk(2,2)", "edge", 2, "c"),
         # Lax-Wendroff: the column pick over the whole padded array is the minimum of the slabs' search keys
         # (astrea_set_key_reducer)
         ("ll3", 32, "plm", "lw", "ssprk(2,2)", "wrap", 2), ("ll4", 32, "ppm", "lw", "ssprk(3,3)", "edge", 2),
         ("sod", 32, "pcm", "lw", "euler", "edge", 3), ("khi", 36, "weno3", "lw", "ssprk(2,2)", "wrap", 2)]


def _initial(spec, world):
    from astrea_b200.initial import initial_state
    config, _, subgrid, _, _, bc = spec[:6]
    high = subgrid.startswith("w") or subgrid == "ppm"
    return initial_state(config, _cells(spec, world), 2, 1.4, high, boundary=bc)


@pytest.fixture(scope="module", params=[2, 3])
def group_run(request):
    """All specs on one group of ``world`` gloo processes; yields (world, directory with one sub-directory per spec)."""
    import torch.multiprocessing as mp
    world = request.param
    with tempfile.TemporaryDirectory() as root:
        for k, spec in enumerate(SPECS):
            os.makedirs(os.path.join(root, f"spec{k}"))
            np.save(os.path.join(root, f"spec{k}", "g0.npy"), _initial(spec, world))
        mp.spawn(_worker, args=(world, os.path.join(root, "rdv"), SPECS, root), nprocs=world, join=True)
        yield world, root


@pytest.mark.parametrize("k", range(len(SPECS)), ids=["-".join(map(str, s[:6])) for s in SPECS])
def test_two_ranks_equal_one(hostsim_lib, group_run, k):
    from astrea_b200.simulation import Simulation
    world, root = group_run
    spec = SPECS[k]
    tmp = os.path.join(root, f"spec{k}")
    errors = [open(os.path.join(tmp, f)).read() for f in sorted(os.listdir(tmp)) if f.startswith("error")]
    assert not errors, "\n".join(errors)
    config, _, subgrid, solver, timestep, bc, steps = spec[:7]
    cells = _cells(spec, world)
    author = spec[7] if len(spec) > 7 else "mc"
    g0 = np.load(os.path.join(tmp, "g0.npy"))
    single = Simulation(config, cells, 2, subgrid, solver, timestep, boundary=bc, grid=g0, _lib=hostsim_lib, ppm_author=author)
    want_dts = single.run(steps)
    want = single.state()
    want_snapshot = single.snapshot()
    single.close()
    got = np.concatenate([np.load(os.path.join(tmp, f"slab{r}.npy")) for r in range(world)], axis=0)
    again = np.concatenate([np.load(os.path.join(tmp, f"aslab{r}.npy")) for r in range(world)], axis=0)
    snapshot = np.concatenate([np.load(os.path.join(tmp, f"snap{r}.npy")) for r in range(world)], axis=1)
    assert np.array_equal(snapshot, want_snapshot, equal_nan=True)
    for r in range(world):
        assert list(np.load(os.path.join(tmp, f"dts{r}.npy"))) == want_dts
    assert np.array_equal(again, want, equal_nan=True)
    assert np.array_equal(got, want, equal_nan=True)
