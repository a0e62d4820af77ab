"""Slab decomposition over two B200s (NCCL): the assembled result equals the single-GPU result bit for bit.
Needs two visible GPUs (skipped otherwise); the same host logic runs over gloo in tests/test_multirank_gloo.py."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, spec, outdir, overlap):
    import torch
    import torch.distributed as dist
    from astrea_b200.simulation import Simulation
    torch.cuda.set_device(rank)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    config, cells, subgrid, solver, timestep, bc, steps = spec[:7]
    author = spec[7] if len(spec) > 7 else "mc"
    full = np.load(os.path.join(outdir, "g0.npy"))
    rows = cells // world
    sim = Simulation(config, cells, 2, subgrid, solver, timestep, boundary=bc, device=rank, rank=rank, world=world,
                     cells_x=rows, grid=full[rank * rows:(rank + 1) * rows], overlap=overlap, ppm_author=author)
    sim.set_time(0.0)
    for _ in range(steps):
        sim.step_async()
    t, n, last = sim.time()
    np.save(os.path.join(outdir, f"slab{rank}.npy"), sim.state())
    np.save(os.path.join(outdir, f"dts{rank}.npy"), np.array(sim.ctx.dt_history(steps)))
    sim.close()
    dist.barrier()
    dist.destroy_process_group()


SPECS = [("ll3", 256, "ppm", "hllc", "ssprk(3,3)", "wrap", 2), ("ll4", 256, "weno5", "lf", "ssprk(2,2)", "edge", 3),
         ("khi", 192, "plm", "hllc", "rk4", "wrap", 2), ("orszag-tang", 192, "plm", "hlld", "ssprk(3,3)", "wrap", 3),
         ("orszag-tang", 160, "ppm", "hlld", "ssprk(2,2)", "edge", 2),
         # PPM authors 'c' / 'ph': grid-wide any() switches OR-ed across the slabs with one NCCL all-reduce per flag pass
         ("ll3", 256, "ppm", "hllc", "ssprk(2,2)", "wrap", 2, "c"), ("khi", 192, "ppm", "lf", "ssprk(3,3)", "wrap", 2, "ph"),
         # Lax-Wendroff: minimum of the slabs' column-search keys (one NCCL all-reduce per sweep)
         ("ll3", 256, "plm", "lw", "ssprk(2,2)", "wrap", 2), ("ll4", 192, "ppm", "lw", "ssprk(3,3)", "edge", 2)]


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("overlap", [True, False], ids=["overlap", "in-order"])
@pytest.mark.parametrize("spec", SPECS, ids=["-".join(map(str, s[:6])) for s in SPECS])
def test_two_gpus_equal_one(spec, overlap):
    import torch.multiprocessing as mp
    from astrea_b200.initial import initial_state
    from astrea_b200.simulation import Simulation
    config, cells, subgrid, solver, timestep, bc, steps = spec[:7]
    author = spec[7] if len(spec) > 7 else "mc"
    high = subgrid.startswith("w") or subgrid == "ppm"
    g0 = initial_state(config, cells, 2, 1.4, high, boundary=bc)
    single = Simulation(config, cells, 2, subgrid, solver, timestep, boundary=bc, grid=g0, ppm_author=author)
    want_dts = single.run(steps)
    want = single.state()
    single.close()
    with tempfile.TemporaryDirectory() as tmp:
        np.save(os.path.join(tmp, "g0.npy"), g0)
        mp.spawn(_worker, args=(2, 29700 + (hash(spec) % 200), spec, tmp, overlap), nprocs=2, join=True)
        got = np.concatenate([np.load(os.path.join(tmp, f"slab{r}.npy")) for r in range(2)], axis=0)
        for r in range(2):
            assert list(np.load(os.path.join(tmp, f"dts{r}.npy"))) == want_dts
    assert np.array_equal(got, want, equal_nan=True)


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_two_devices_driven_by_one_host_thread():
    """Contexts on different GPUs driven alternately from one thread (every entry point selects its context's device and
    restores the caller's; kernel attributes are kept per device): both give the single-device result."""
    import torch
    from astrea_b200 import _native as N
    from astrea_b200.initial import initial_state, problem
    from astrea_b200.selectors import make_cfg
    cells = 96
    prob = problem("ll6", cells, 1.4)
    g0 = initial_state("ll6", cells, 2, 1.4, True)
    ctxs = []
    for dev in (0, 1):
        cfg = make_cfg(dimension=2, cells=cells, boundary="wrap", gamma=1.4, dx=prob["dx"], cfl=.5, subgrid="ppm", solver="hllc",
                       timestep="ssprk(3,3)", device=dev)
        ctxs.append(N.Context(cfg))
    before = torch.cuda.current_device()
    for c in ctxs:
        c.upload(g0)
    dts = [[], []]
    for _ in range(3):
        for k, c in enumerate(ctxs):          # interleaved: device 0, device 1, device 0, ...
            dts[k].append(c.step())
    a, b = ctxs[0].download(), ctxs[1].download()
    assert torch.cuda.current_device() == before
    for c in ctxs:
        c.close()
    assert dts[0] == dts[1] and np.array_equal(a, b)
