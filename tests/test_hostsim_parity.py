"""Kernel logic against the oracle without a GPU.

The kernel sources build a second way (g++, -DASTREA_HOSTSIM): the same block/phase code, executed thread by
thread on the CPU behind the same C ABI.  That library is test infrastructure (the astrea_b200 package never
loads it); it lets every index map, halo rule and limiter branch of the CUDA kernels be checked against the oracle
in the build container.  The `-m gpu` tests repeat these comparisons on the device library.
"""
import numpy as np
import pytest

from conftest import golden_case, golden_index, rel_l1
from cases import run_native, run_oracle
from astrea_b200.initial import initial_state, problem

HYDRO = sorted(golden_index())      # every golden case, constrained-transport MHD included


@pytest.mark.parametrize("cid", HYDRO)
def test_golden_cases_bit_exact_vs_oracle(hostsim_lib, cid):
    meta, data = golden_case(cid)
    want, dts = run_oracle(meta, data["g0"], meta["steps"])
    got, used, _ = run_native(hostsim_lib, meta, data["g0"], meta["steps"])
    assert used == dts
    assert np.array_equal(got, want, equal_nan=True)


@pytest.mark.parametrize("cid", HYDRO)
def test_golden_cases_vs_reference(hostsim_lib, cid):
    """Against the reference's own output, fed the reference's dt sequence through evolve_space / evolve_time."""
    meta, data = golden_case(cid)
    got, _, eigs = run_native(hostsim_lib, meta, data["g0"], meta["steps"], dts=list(data["dts"]))
    tol = 1e-10 if "weno7" in cid else 1e-12 * meta["steps"]
    assert np.all(rel_l1(got, data["g"]) <= tol)
    # eigmax[a] is per sweep axis a; the reference lists them in iteration order (reversed on odd steps)
    for n, e in enumerate(eigs):
        ref = list(data["eigmax"][n])
        ref = ref[::-1] if (n % 2 and meta["dimension"] == 2) else ref
        assert np.allclose(e, ref, rtol=1e-11, atol=0)


def _meta(config, cells, dim, subgrid, solver, timestep, bc, mhd=False):
    prob = problem(config, cells, 1.4)
    return dict(config=config, cells=cells, dimension=dim, subgrid=subgrid, solver=solver, timestep=timestep,
                boundary=bc or prob["boundary"], dx=prob["dx"], gamma=1.4, cfl=.5, magnetic_2d=mhd)


# every scheme x solver, both boundary modes, odd sizes, with the grid cut into many blocks
MATRIX = [(cfg, sub, sol, bc, dim)
          for dim, cfg in ((1, "sod"), (2, "ll4"))
          for sub in ("pcm", "plm", "ppm", "weno3", "weno5", "weno7")
          for sol in ("lf", "hllc")
          for bc in ("edge", "wrap")]


@pytest.mark.parametrize("config,subgrid,solver,bc,dim", MATRIX, ids=["-".join(map(str, m)) for m in MATRIX])
def test_scheme_solver_matrix_multiblock(hostsim_lib, config, subgrid, solver, bc, dim):
    cells = 61 if dim == 1 else 38
    meta = _meta(config, cells, dim, subgrid, solver, "ssprk(3,3)", bc)
    high = subgrid.startswith("w") or subgrid == "ppm"
    g0 = initial_state(config, cells, dim, 1.4, high, boundary=bc)
    want, dts = run_oracle(meta, g0, 2)
    got, used, _ = run_native(hostsim_lib, meta, g0, 2, threads_2d=32, segment_2d=11, tile_1d=13)
    assert used == dts
    assert np.array_equal(got, want, equal_nan=True)
    if dim == 2:
        # the same through the general 8-variable kernels (a grid without v_z / B normally takes the hydro variants)
        got, used, _ = run_native(hostsim_lib, meta, g0, 2, threads_2d=32, segment_2d=11, general_path=True)
        assert used == dts
        assert np.array_equal(got, want, equal_nan=True)


# constrained transport (magnetic_2d): every scheme, both HLL solvers, both boundary modes, every integrator family
MHD = [("orszag-tang", "plm", "hlld", "ssprk(3,3)", "wrap"), ("orszag-tang", "ppm", "hlld", "ssprk(3,3)", "wrap"),
       ("orszag-tang", "weno5", "hllc", "ssprk(2,2)", "wrap"), ("orszag-tang", "pcm", "hlld", "euler", "wrap"),
       ("mhd rotor", "plm", "hlld", "ssprk(3,3)", "wrap"), ("orszag-tang", "plm", "hlld", "ssprk(10,4)", "wrap"),
       ("orszag-tang", "weno3", "hlld", "ssprk(5,3)", "edge"), ("orszag-tang", "weno7", "hlld", "ssprk(5,4)", "wrap"),
       ("orszag-tang", "plm", "hllc", "rk4", "edge"), ("orszag-tang", "ppm", "hlld", "ssprk(4,3)", "edge")]


@pytest.mark.parametrize("config,subgrid,solver,timestep,bc", MHD, ids=["-".join(m) for m in MHD])
def test_constrained_transport(hostsim_lib, config, subgrid, solver, timestep, bc):
    """mag_field.py: transverse PPM to the corners, upwinded corner E_z, induction update, face-average overwrite
    (Q14) and inverse reconstruction after every register update — two steps, so that the odd-step role swap of
    SURVEY Q1b is exercised."""
    cells = 26
    meta = _meta(config, cells, 2, subgrid, solver, timestep, bc, mhd=True)
    high = subgrid.startswith("w") or subgrid == "ppm"
    g0 = initial_state(config, cells, 2, 1.4, high, boundary=bc)
    want, dts = run_oracle(meta, g0, 2)
    got, used, _ = run_native(hostsim_lib, meta, g0, 2, segment_2d=9)
    assert np.isfinite(want).all()
    assert used == dts
    assert np.array_equal(got, want, equal_nan=True)


def test_mhd_blast_raises_like_the_reference(hostsim_lib):
    """The reference's MHD blast run dies of LinAlgError in its first step (negative pressure -> NaN wave speed)."""
    from astrea_b200 import _native as N
    from cases import native_cfg, oracle_cfg
    from oracle import advance
    meta = _meta("mhd blast", 24, 2, "plm", "hlld", "ssprk(2,2)", None, mhd=True)
    g0 = initial_state("mhd blast", 24, 2, 1.4, False)
    with pytest.raises(np.linalg.LinAlgError):
        advance(np.copy(g0), oracle_cfg(meta), 2)
    ctx = N.Context(native_cfg(meta), lib=hostsim_lib)
    ctx.upload(g0)
    with pytest.raises(np.linalg.LinAlgError):
        ctx.step()
        ctx.step()
        ctx.read_eigmax()
    ctx.close()


def test_face_field_download(hostsim_lib):
    """evolve_time overwrites Bx, By of the caller's grid with the face averages of the stage-1 operator (Q14)."""
    import ctypes
    from astrea_b200 import _native as N
    from cases import native_cfg, oracle_cfg
    from oracle import space_operator
    meta = _meta("orszag-tang", 20, 2, "plm", "hlld", "ssprk(3,3)", None, mhd=True)
    g0 = initial_state("orszag-tang", 20, 2, 1.4, False)
    fl = space_operator(np.copy(g0), oracle_cfg(meta))
    ctx = N.Context(native_cfg(meta), lib=hostsim_lib)
    ctx.upload(g0)
    ctx.evolve_space(0)
    out = np.empty((20, 20, 2))
    ctx._check(ctx.lib.astrea_download_face_field(ctx._h, out.ctypes.data))
    ctx.close()
    assert np.array_equal(out[..., 0], fl[0]["face_avg"][..., 5])
    assert np.array_equal(out[..., 1], fl[1]["face_avg"].transpose(1, 0, 2)[..., 6])


INTEGRATORS = ["euler", "rk4", "ssprk(2,2)", "ssprk(3,3)", "ssprk(4,3)", "ssprk(5,3)", "ssprk(5,4)", "ssprk(10,4)"]


@pytest.mark.parametrize("timestep", INTEGRATORS)
@pytest.mark.parametrize("dim", [1, 2])
def test_every_integrator(hostsim_lib, timestep, dim):
    cells = 64 if dim == 1 else 24
    config = "sod" if dim == 1 else "ll12"
    meta = _meta(config, cells, dim, "plm", "hllc", timestep, None)
    g0 = initial_state(config, cells, dim, 1.4, False)
    want, dts = run_oracle(meta, g0, 3)
    got, used, _ = run_native(hostsim_lib, meta, g0, 3)
    assert used == dts
    assert np.array_equal(got, want, equal_nan=True)


@pytest.mark.parametrize("limiter", ["minmod", "vanleer", "ospre", "vanalbada", "koren", "superbee"])
def test_slope_limiters(hostsim_lib, limiter):
    """limiters.py:10-49: only minmod is wired into plm.py:27; the others are selectable through the C ABI."""
    meta = _meta("sod", 96, 1, "plm", "lf", "ssprk(2,2)", None)
    g0 = initial_state("sod", 96, 1, 1.4, False)
    from cases import oracle_cfg
    from oracle import advance
    cfg = oracle_cfg(meta)
    cfg.slope_limiter = limiter
    want, dts = advance(np.copy(g0), cfg, 3)
    got, used, _ = run_native(hostsim_lib, meta, g0, 3, limiter=limiter)
    assert used == dts
    assert np.array_equal(got, want, equal_nan=True)


def test_geometry_does_not_change_results(hostsim_lib):
    """Block shape is an implementation detail: any tiling gives bit-identical states."""
    meta = _meta("ll3", 40, 2, "ppm", "hllc", "ssprk(3,3)", None)
    g0 = initial_state("ll3", 40, 2, 1.4, True)
    base, _, _ = run_native(hostsim_lib, meta, g0, 2)
    for threads, seg in ((32, 7), (48, 40), (64, 16)):
        got, _, _ = run_native(hostsim_lib, meta, g0, 2, threads_2d=threads, segment_2d=seg)
        assert np.array_equal(got, base, equal_nan=True)
    # flux stage with block-wide / warp-wide rows of transverse points, reconstruction without the bulk-copy ring
    for kw in (dict(flux_block_tile=True), dict(flux_block_tile=True, threads_2d=96), dict(flux_block_tile=False),
               dict(recon_bulk=False), dict(flux_block_tile=True, recon_bulk=False, segment_2d=11)):
        got, _, _ = run_native(hostsim_lib, meta, g0, 2, **kw)
        assert np.array_equal(got, base, equal_nan=True), kw
    for bc in ("edge", "wrap"):
        meta2 = _meta("ll4", 37, 2, "weno5", "lf", "ssprk(2,2)", bc)
        g2 = initial_state("ll4", 37, 2, 1.4, True, boundary=bc)
        a, _, _ = run_native(hostsim_lib, meta2, g2, 2)
        b, _, _ = run_native(hostsim_lib, meta2, g2, 2, flux_block_tile=True, recon_bulk=False)
        assert np.array_equal(a, b, equal_nan=True), bc


def test_primitive_download(hostsim_lib):
    """astrea.py:47 snapshots sim_variables.convert_conservative(grid): 4th-order for PPM/WENO, pointwise otherwise."""
    from astrea_b200 import _native as N
    from cases import native_cfg, oracle_cfg
    from oracle.gridops import prim_avg_of_cons_avg
    for dim, sub, bc in ((1, "plm", "edge"), (1, "ppm", "edge"), (2, "weno5", "wrap"), (2, "ppm", "edge"), (2, "pcm", "wrap")):
        cells = 50 if dim == 1 else 20
        config = "sod" if dim == 1 else "ll3"
        meta = _meta(config, cells, dim, sub, "lf", "euler", bc)
        g0 = initial_state(config, cells, dim, 1.4, sub != "plm" and sub != "pcm", boundary=bc)
        ctx = N.Context(native_cfg(meta), lib=hostsim_lib)
        ctx.upload(g0)
        assert np.array_equal(ctx.download(), g0)
        assert np.array_equal(ctx.download(primitive=True), prim_avg_of_cons_avg(g0, oracle_cfg(meta)))
        ctx.close()


def test_nonfinite_wave_speed_raises(hostsim_lib):
    """Where np.linalg.eigvals raises LinAlgError in the reference (fv.py:158) the ABI returns
    ASTREA_E_NONFINITE and the binding raises a LinAlgError subclass."""
    from astrea_b200 import _native as N
    from cases import native_cfg
    meta = _meta("sod", 64, 1, "plm", "lf", "ssprk(2,2)", None)
    g0 = initial_state("sod", 64, 1, 1.4, False)
    g0[10, 4] = np.nan
    ctx = N.Context(native_cfg(meta), lib=hostsim_lib)
    ctx.upload(g0)
    with pytest.raises(np.linalg.LinAlgError):
        ctx.step()
    ctx.close()


def test_call_order_is_enforced(hostsim_lib):
    from astrea_b200 import _native as N
    from cases import native_cfg
    meta = _meta("sod", 64, 1, "plm", "lf", "ssprk(2,2)", None)
    ctx = N.Context(native_cfg(meta), lib=hostsim_lib)
    ctx.upload(initial_state("sod", 64, 1, 1.4, False))
    with pytest.raises(N.AstreaError) as err:
        ctx.evolve_time(1e-3)
    assert err.value.code == N.E_STATE
    ctx.close()


def test_unsupported_selectors_fail_loudly(hostsim_lib):
    """Lax-Wendroff is reproduced for spectra without v_z / B only: anything else is refused, not guessed."""
    from astrea_b200 import _native as N
    from cases import native_cfg
    # constrained transport + Lax-Wendroff: refused at creation
    meta = _meta("orszag-tang", 16, 2, "plm", "lw", "ssprk(2,2)", None, mhd=True)
    with pytest.raises(N.AstreaError):
        N.Context(native_cfg(meta), lib=hostsim_lib)
    # a 1D MHD state (B != 0) + Lax-Wendroff: refused when the operator runs
    meta = _meta("brio-wu", 64, 1, "plm", "lw", "ssprk(2,2)", None)
    ctx = N.Context(native_cfg(meta), lib=hostsim_lib)
    ctx.upload(initial_state("brio-wu", 64, 1, 1.4, False))
    with pytest.raises(N.AstreaError) as err:
        ctx.step()
    assert err.value.code == N.E_ARG and "Lax-Wendroff" in str(err.value)
    ctx.close()
    # a slab needs the grid-wide hooks: Lax-Wendroff's column search (astrea_set_key_reducer) and the any() switches of
    # the PPM authors 'c' / 'ph' (astrea_set_flag_reducer); without them the operator stops instead of using local values
    from astrea_b200.selectors import make_cfg
    for solver, author, hook in (("lw", "mc", "astrea_set_key_reducer"), ("hllc", "c", "astrea_set_flag_reducer")):
        cfg = make_cfg(dimension=2, nx=12, ny=24, boundary="wrap", gamma=1.4, dx=1.0 / 24, cfl=.5, subgrid="ppm", solver=solver,
                       timestep="ssprk(2,2)", nx_global=24, x_offset=0, ppm_author=author)
        ctx = N.Context(cfg, lib=hostsim_lib)
        ctx.upload(initial_state("ll6", 24, 2, 1.4, True)[:12])
        with pytest.raises(N.AstreaError) as err:
            ctx.run_instr(0, external_rows=True)
        assert err.value.code == N.E_STATE and hook in str(err.value)
        ctx.close()
    # unknown enums
    cfg = native_cfg(_meta("sod", 64, 1, "plm", "lf", "ssprk(2,2)", None))
    cfg.solver = 9
    with pytest.raises(N.AstreaError):
        N.Context(cfg, lib=hostsim_lib)


@pytest.mark.parametrize("spec", [("sod", 1, "plm", "lf", "ssprk(2,2)", 96), ("ll6", 2, "ppm", "hllc", "ssprk(3,3)", 32),
                                  ("orszag-tang", 2, "plm", "hlld", "ssprk(3,3)", 24)], ids=["sod1d", "ll6", "ot"])
def test_async_stepping_with_device_clock(hostsim_lib, spec):
    """astrea_step_async: dt = cfl*min(dx/eigmax) and the t_stop clip (astrea.py:70-78) evaluated on the device."""
    from astrea_b200 import _native as N
    from astrea_b200.selectors import MAGNETIC_2D
    from cases import native_cfg, oracle_cfg
    from oracle import advance
    config, dim, subgrid, solver, timestep, cells = spec
    meta = _meta(config, cells, dim, subgrid, solver, timestep, None, mhd=config in MAGNETIC_2D)
    g0 = initial_state(config, cells, dim, 1.4, subgrid == "ppm")
    ctx = N.Context(native_cfg(meta), lib=hostsim_lib)
    ctx.upload(g0)
    # free-running steps
    want, dts = run_oracle(meta, g0, 4)
    ctx.set_time(0.0, 0.0)
    for _ in range(4):
        ctx.step_async()
    t, steps, last = ctx.get_time()
    assert steps == 4 and last == dts[-1] and ctx.dt_history(4) == dts
    tt = 0.0
    for d in dts:
        tt += d
    assert t == tt
    assert np.array_equal(ctx.download(), want, equal_nan=True)
    # clipped at t_stop: the last step lands exactly on it (astrea.py:74-75)
    t_stop = dts[0] + 0.4 * dts[1]
    cfg = oracle_cfg(meta)
    want, used = advance(np.copy(g0), cfg, 2, t=0.0, t_end=t_stop)
    ctx.upload(g0)
    ctx.parity = 0
    ctx.set_time(0.0, t_stop)
    ctx.step_async()
    ctx.step_async()
    t, steps, last = ctx.get_time()
    assert steps == 2 and ctx.dt_history(2) == used and t == used[0] + used[1]
    assert np.array_equal(ctx.download(), want, equal_nan=True)
    ctx.close()


def test_drop_in_behind_the_reference_time_loop(hostsim_lib):
    """The reference's own loop body (astrea.py:67-85, as driven by tests/golden/refharness.py) with its `evolvers`
    module swapped for astrea_b200.evolvers: same sim_variables namedtuple, same calls, same results.  Needs the
    reference checkout (build container only)."""
    import sys
    sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__file__), "golden"))
    import refharness as rh
    if not rh.available():
        pytest.skip("reference checkout not present")
    from astrea_b200 import evolvers as ours
    for config, cells, dim, subgrid, solver, timestep, steps in (("ll6", 32, 2, "ppm", "hllc", "ssprk(3,3)", 3),
                                                                 ("sod", 128, 1, "plm", "lf", "ssprk(2,2)", 4),
                                                                 ("orszag-tang", 24, 2, "plm", "hlld", "ssprk(3,3)", 3)):
        sv = rh.make_sim_variables(config, cells, dim, subgrid, solver, timestep)
        g0 = rh.initial_grid(sv)
        want, dts, eigs, _ = rh.run_steps(sv, steps, grid=np.copy(g0))
        grid, used = np.copy(g0), []
        with np.errstate(all="ignore"):
            for n in range(steps):
                fluxes = ours.evolve_space(grid, sv, _lib=hostsim_lib)
                dt = sv.cfl * min(sv.dx / f["eigmax"] for f in fluxes.values())
                grid = ours.evolve_time(grid, fluxes, dt, sv, _lib=hostsim_lib)
                sv = sv._replace(permutations=dict(reversed(list(sv.permutations.items()))))
                used.append(dt)
        ours.release()
        assert np.allclose(used, dts, rtol=1e-12, atol=0)
        assert np.all(rel_l1(grid, want[-1]) <= 1e-10), (config, rel_l1(grid, want[-1]))


@pytest.mark.parametrize("dim,subgrid", [(1, "plm"), (2, "ppm"), (2, "plm")])
def test_device_diagnostics(hostsim_lib, dim, subgrid):
    """Conservation totals and total variation (functions/analytic.py:48-77) reduced on the device."""
    from astrea_b200 import _native as N
    from cases import native_cfg, oracle_cfg
    from oracle.gridops import prim_avg_of_cons_avg
    cells = 300 if dim == 1 else 70
    config = "sod" if dim == 1 else "khi"
    meta = _meta(config, cells, dim, subgrid, "hllc", "ssprk(2,2)", None)
    g0 = initial_state(config, cells, dim, 1.4, subgrid == "ppm")
    ctx = N.Context(native_cfg(meta), lib=hostsim_lib)
    ctx.upload(g0)
    ctx.step()
    q = ctx.download()
    tot, tv = ctx.diagnostics()
    ctx.close()
    w = prim_avg_of_cons_avg(q, oracle_cfg(meta))
    d = w
    for ax in range(dim):
        d = np.diff(d, axis=ax)
    axes = tuple(range(dim))
    scale = np.abs(q).sum(axis=axes)          # summation order differs: compare against the size of the terms
    assert np.all(np.abs(tot - q.sum(axis=axes)) <= 1e-12 * np.where(scale > 0, scale, 1))
    assert np.allclose(tv, np.abs(d).sum(axis=axes), rtol=1e-12, atol=1e-13)


def test_snapshot_layout(hostsim_lib):
    """astrea.py:47: the stored snapshot is the primitive grid transposed by ortho_axis."""
    from astrea_b200.simulation import Simulation
    from cases import oracle_cfg
    from oracle.gridops import prim_avg_of_cons_avg
    sim = Simulation("ll3", 24, 2, "ppm", "hllc", "ssprk(3,3)", _lib=hostsim_lib)
    meta = _meta("ll3", 24, 2, "ppm", "hllc", "ssprk(3,3)", None)
    want = prim_avg_of_cons_avg(sim.state(), oracle_cfg(meta)).transpose(1, 0, 2)
    assert np.array_equal(sim.snapshot(), want)
    sim.close()


def test_tiny_grids(hostsim_lib):
    """4 and 6 cells per side: every stencil wraps or clamps across the whole grid (ghost depth > grid size)."""
    from astrea_b200.selectors import MAGNETIC_2D
    for dim, config in ((1, "sod"), (2, "ll3"), (2, "orszag-tang")):
        for cells in (4, 6):
            for subgrid in ("pcm", "plm", "ppm", "weno5", "weno7"):
                for solver, bc in (("lf", "wrap"), ("hllc", "edge"), ("hlld", "wrap")):
                    mhd = config in MAGNETIC_2D
                    meta = _meta(config, cells, dim, subgrid, solver, "ssprk(2,2)", bc, mhd=mhd)
                    g0 = initial_state(config, cells, dim, 1.4, subgrid in ("ppm", "weno5", "weno7"), boundary=bc)
                    try:
                        want, dts = run_oracle(meta, g0, 2)
                    except np.linalg.LinAlgError:
                        continue
                    got, used, _ = run_native(hostsim_lib, meta, g0, 2)
                    assert used == dts and np.array_equal(got, want, equal_nan=True), (dim, config, cells, subgrid, solver, bc)


def test_step_program_contract(hostsim_lib):
    """Call-order rules of the instruction-level interface (include/astrea_b200.h): violations are reported, not ignored."""
    from astrea_b200 import _native as N
    from cases import native_cfg
    meta = _meta("ll3", 24, 2, "plm", "hllc", "ssprk(3,3)", None)
    ctx = N.Context(native_cfg(meta), lib=hostsim_lib)
    ctx.upload(initial_state("ll3", 24, 2, 1.4, False))
    prog, upd, readers = ctx.program(), ctx.updates(), ctx.halo_readers()
    assert prog == [True, False, True, False, True, False] and upd == [not p for p in prog] and readers == prog
    with pytest.raises(N.AstreaError) as err:
        ctx.run_instr(2)                           # instruction 0 has not run
    assert err.value.code == N.E_STATE
    with pytest.raises(N.AstreaError) as err:
        ctx.halo_ptrs(1)                           # a register update reads no ghost rows
    assert err.value.code == N.E_ARG
    ctx.run_instr(0)
    for call in (lambda: ctx.download(), lambda: ctx.diagnostics(), lambda: ctx.save_state(), lambda: ctx.finish_step()):
        with pytest.raises(N.AstreaError) as err:
            call()                                 # a step is in flight
        assert err.value.code == N.E_STATE
    with pytest.raises(N.AstreaError):
        ctx.run_update_part(0, 0)                  # not a register update
    ctx.set_dt(1e-4)
    ctx.run_update_part(1, 0)
    ctx.run_update_part(1, 1)
    for i in range(2, len(prog)):
        ctx.run_instr(i)
    ctx.finish_step()
    assert ctx.parity == 1
    with pytest.raises(N.AstreaError):
        ctx.dt_history(5)                          # no asynchronous step has been taken
    ctx.close()


def test_split_update_equals_whole_update(hostsim_lib):
    """astrea_run_update_part (edge rows, then interior rows) gives the same register as the undivided update."""
    from astrea_b200 import _native as N
    from cases import native_cfg
    meta = _meta("ll6", 80, 2, "ppm", "hllc", "ssprk(3,3)", None)
    g0 = initial_state("ll6", 80, 2, 1.4, True)
    results = []
    for split in (False, True):
        ctx = N.Context(native_cfg(meta), lib=hostsim_lib)
        ctx.upload(g0)
        ctx.run_instr(0)
        ctx.set_dt(2e-3)
        for i, is_update in enumerate(ctx.updates()):
            if i == 0:
                continue
            if is_update and split:
                ctx.run_update_part(i, 0)
                ctx.run_update_part(i, 1)
            else:
                ctx.run_instr(i)
        ctx.finish_step()
        results.append(ctx.download())
        ctx.close()
    assert np.array_equal(results[0], results[1])


def test_hllc_low_mach_switch(hostsim_lib):
    """solvers.py:118-122: the low-Mach rescaling of the HLLC wave speeds (off by default; no caller in the reference
    turns it on).  It goes through sin(), whose last bit may differ between libraries: north-star tolerance."""
    from cases import oracle_cfg
    from oracle import advance
    for dim, config, cells in ((1, "sod", 128), (2, "khi", 32)):
        meta = _meta(config, cells, dim, "plm", "hllc", "ssprk(2,2)", None)
        g0 = initial_state(config, cells, dim, 1.4, False)
        cfg = oracle_cfg(meta)
        cfg.low_mach = True
        want, dts = advance(np.copy(g0), cfg, 2)
        got, used, _ = run_native(hostsim_lib, meta, g0, 2, low_mach=True)
        assert np.allclose(used, dts, rtol=1e-13, atol=0)
        assert np.all(rel_l1(got, want) <= 2e-12)
        plain, _, _ = run_native(hostsim_lib, meta, g0, 2)
        assert not np.array_equal(plain, got)      # the switch does something


PPM_AUTHORS = [(cfg, dim, sol, ts, bc, author)
               for cfg, dim, sol, ts, bc in (("sod", 1, "hllc", "ssprk(3,3)", None), ("sin", 1, "lf", "rk4", None),
                                             ("ll3", 2, "hllc", "ssprk(3,3)", None), ("ll4", 2, "lf", "ssprk(2,2)", "edge"),
                                             ("orszag-tang", 2, "hlld", "ssprk(2,2)", None))
               for author in ("c", "ph")]


@pytest.mark.parametrize("config,dim,solver,timestep,bc,author", PPM_AUTHORS, ids=["-".join(map(str, m)) for m in PPM_AUTHORS])
def test_ppm_authors_colella_and_peterson_hammett(hostsim_lib, config, dim, solver, timestep, bc, author):
    """ppm.run(author='c' | 'ph'): interface limiter + extrapolant limiter with their grid-wide any() switches."""
    from astrea_b200.selectors import MAGNETIC_2D
    cells = 96 if dim == 1 else 30
    meta = _meta(config, cells, dim, "ppm", solver, timestep, bc, mhd=config in MAGNETIC_2D)
    meta["ppm_author"] = author
    g0 = initial_state(config, cells, dim, 1.4, True, boundary=meta["boundary"])
    want, dts = run_oracle(meta, g0, 2)
    for general in (False, True):
        got, used, _ = run_native(hostsim_lib, meta, g0, 2, segment_2d=11, tile_1d=19, general_path=general)
        assert used == dts
        assert np.array_equal(got, want, equal_nan=True)


@pytest.mark.parametrize("author", ["c", "ph"])
def test_ppm_author_switches_off(hostsim_lib, author):
    """A strictly monotone profile with outflow boundaries has no face extremum: `local_extrema.any()` is False and the
    interface limiter returns the faces untouched for every cell (limiters.py:58,76-78)."""
    from oracle.reconstruct import ppm_face_value
    from oracle.gridops import prim_avg_of_cons_avg
    from cases import oracle_cfg
    cells = 80
    x = (np.arange(cells) + .5) / cells
    w = np.zeros((cells, 8))
    w[:, 0], w[:, 1], w[:, 4] = 1 + x, .2 + .3 * x, 1 + 2 * x
    w[:, 2], w[:, 3], w[:, 5], w[:, 6], w[:, 7] = .1 + .1 * x, .3 + x, .2 + .1 * x, .4 + .2 * x, .1 + .3 * x
    from oracle.gridops import cons_of_prim
    g0 = cons_of_prim(w, 1.4)
    meta = dict(config="sod", cells=cells, dimension=1, subgrid="ppm", solver="lf", timestep="euler", boundary="edge", dx=1 / cells,
                gamma=1.4, cfl=.5, magnetic_2d=False, ppm_author=author)
    wS = prim_avg_of_cons_avg(g0, oracle_cfg(meta))
    face = ppm_face_value(wS, "edge")
    assert not ((face - wS) * (np.roll(wS, -1, axis=0) - face) < 0)[:-1].any()      # the premise of the test
    want, dts = run_oracle(meta, g0, 1)
    got, used, _ = run_native(hostsim_lib, meta, g0, 1, tile_1d=23)
    assert used == dts
    assert np.array_equal(got, want, equal_nan=True)


def test_clean_grid_after_nonfinite_error(hostsim_lib):
    """The non-finite flag is sticky until it has been reported, not beyond: a context (the drop-in keeps them in a
    module-level cache) must accept a clean grid after a bad one."""
    from astrea_b200 import _native as N
    from cases import native_cfg
    meta = _meta("sod", 64, 1, "plm", "lf", "ssprk(2,2)", None)
    good = initial_state("sod", 64, 1, 1.4, False)
    bad = np.copy(good)
    bad[10, 4] = np.nan
    want, dts = run_oracle(meta, good, 2)
    ctx = N.Context(native_cfg(meta), lib=hostsim_lib)
    ctx.upload(bad)
    with pytest.raises(np.linalg.LinAlgError):
        ctx.step()
    # the same context, a clean grid: synchronous and asynchronous stepping both work again
    ctx.upload(good)
    ctx.parity = 0
    assert [ctx.step(), ctx.step()] == dts
    assert np.array_equal(ctx.download(), want)
    ctx.upload(bad)
    ctx.set_time(0.0, 0.0)
    ctx.step_async()
    with pytest.raises(np.linalg.LinAlgError):
        ctx.get_time()
    ctx.get_time()                              # reported once, then cleared
    ctx.upload(good)
    ctx.parity = 0
    ctx.set_time(0.0, 0.0)
    ctx.step_async()
    ctx.step_async()
    assert ctx.get_time()[1] == 2 and np.array_equal(ctx.download(), want)
    ctx.close()


def test_download_validates_the_output_array(hostsim_lib):
    from astrea_b200 import _native as N
    from cases import native_cfg
    meta = _meta("ll3", 12, 2, "plm", "lf", "euler", None)
    ctx = N.Context(native_cfg(meta), lib=hostsim_lib)
    g0 = initial_state("ll3", 12, 2, 1.4, False)
    ctx.upload(g0)
    for wrong in (np.empty((12, 11, 8)), np.empty((12, 12, 8), dtype=np.float32), np.empty((12, 12, 16))[..., ::2],
                  np.empty((12, 8, 12)).transpose(0, 2, 1)):
        with pytest.raises(ValueError):
            ctx.download(out=wrong)
    ro = np.empty((12, 12, 8))
    ro.flags.writeable = False
    with pytest.raises(ValueError):
        ctx.download(out=ro)
    out = np.empty((12, 12, 8))
    assert ctx.download(out=out) is out and np.array_equal(out, g0)
    ctx.close()


@pytest.mark.parametrize("rows", [8, 9, 11, 15, 16, 17])
def test_split_update_on_short_slabs(hostsim_lib, rows):
    """astrea_run_update_part on slabs shorter than two edge blocks: after part 0 every row that travels to a
    neighbour (the first and last GHOST rows) holds its final value, and parts 0 + 1 update every row exactly once."""
    import ctypes
    from astrea_b200 import _native as N
    from astrea_b200.selectors import make_cfg
    cells = 24
    g0 = initial_state("ll6", cells, 2, 1.4, True)[:rows]
    results, sent = [], []
    for split in (False, True):
        cfg = make_cfg(dimension=2, nx=rows, ny=cells, boundary="wrap", gamma=1.4, dx=1.0 / cells, cfl=.5, subgrid="ppm",
                       solver="hllc", timestep="ssprk(3,3)", nx_global=2 * rows, x_offset=0)
        ctx = N.Context(cfg, lib=hostsim_lib)
        on_host = ctx.lib.astrea_is_device_build() == 0
        ctx.upload(g0)
        ctx.run_instr(0)
        ctx.set_dt(1e-3)
        if split:
            ctx.run_update_part(1, 0)
        else:
            ctx.run_instr(1)
        if on_host:       # host simulation: the send blocks of the operator that follows are plain memory
            _, count = ctx.halo_info()
            sent.append([np.array((ctypes.c_double * count).from_address(p)) for p in ctx.halo_ptrs(2)[:2]])
        if split:
            ctx.run_update_part(1, 1)
        for i in range(2, len(ctx.program())):
            ctx.run_instr(i)
        ctx.finish_step()
        results.append(ctx.download())
        ctx.close()
    assert np.array_equal(results[0], results[1], equal_nan=True)
    if sent:
        # ghost columns are filled later (halo_prepare); compare the interior columns of the rows that travel
        g = 8
        width = cells + 2 * g
        for whole, part in zip(sent[0], sent[1]):
            a, b = whole.reshape(-1, 8, width)[..., g:-g], part.reshape(-1, 8, width)[..., g:-g]
            assert np.array_equal(a, b, equal_nan=True)


def test_async_snapshots_while_stepping(hostsim_lib):
    """astrea.py:47-50 off the critical path: a snapshot started before further steps are enqueued holds the state of
    its own step (primitive, transposed by ortho_axis), for ragged sizes, 1D, and more tickets than event slots."""
    from astrea_b200.simulation import Simulation
    from cases import oracle_cfg
    from oracle.gridops import prim_avg_of_cons_avg
    for config, cells, dim, subgrid in (("ll6", 45, 2, "ppm"), ("khi", 33, 2, "plm"), ("sod", 77, 1, "weno5")):
        sim = Simulation(config, cells, dim, subgrid, "hllc", "ssprk(3,3)", _lib=hostsim_lib)
        meta = _meta(config, cells, dim, subgrid, "hllc", "ssprk(3,3)", None)
        shots, states = [], []
        sim.set_time(0.0)
        for n in range(6):
            states.append(None)
            shots.append(sim.snapshot_async())
            states[-1] = sim.state()                   # the conservative grid the snapshot was taken of
            sim.step_async()
        for (out, ticket), q in zip(shots, states):
            sim.ctx.snapshot_wait(ticket)
            want = prim_avg_of_cons_avg(q, oracle_cfg(meta))
            want = want.transpose(1, 0, 2) if dim == 2 else want
            assert np.array_equal(out, want, equal_nan=True), (config, ticket)
        with pytest.raises(ValueError):
            sim.ctx.snapshot_begin(np.empty((3, 3, 8)))
        sim.close()


@pytest.mark.parametrize("spec", [("sod", "plm", "lf", "ssprk(2,2)", 1024, None), ("sod", "ppm", "hllc", "ssprk(3,3)", 300, None),
                                  ("shu-osher", "weno5", "lf", "rk4", 257, None), ("brio-wu", "plm", "hlld", "ssprk(5,4)", 200, None),
                                  ("sod", "pcm", "lw", "euler", 64, None), ("square", "weno7", "lf", "ssprk(10,4)", 90, None),
                                  ("sod", "weno3", "hllc", "ssprk(5,3)", 130, "wrap"), ("ryu-jones", "ppm", "hlld", "ssprk(4,3)", 128, None)],
                         ids=lambda s: "-".join(map(str, s[:5])))
def test_run_steps_equals_single_steps(hostsim_lib, spec):
    """astrea_run_steps on 1D grids (on the device: CUDA-graph replay of each step) gives the same grid, clock and dt
    history as plainly launched single steps, for every integrator family, across calls and after a new upload."""
    from astrea_b200 import _native as N
    from cases import native_cfg
    config, subgrid, solver, timestep, cells, bc = spec
    meta = _meta(config, cells, 1, subgrid, solver, timestep, bc)
    g0 = initial_state(config, cells, 1, 1.4, subgrid in ("ppm", "weno3", "weno5", "weno7"), boundary=meta["boundary"])
    ref = N.Context(native_cfg(meta, step_graph=False), lib=hostsim_lib)          # plain launches
    ref.upload(g0)
    ref.set_time(0.0, 0.0)
    for _ in range(7):
        ref.step_async()
    want, want_clock, want_dts = ref.download(), ref.get_time(), ref.dt_history(7)
    ctx = N.Context(native_cfg(meta), lib=hostsim_lib)
    ctx.upload(g0)
    ctx.set_time(0.0, 0.0)
    ctx.run_steps(3)
    ctx.run_steps(0)
    ctx.run_steps(4)
    assert ctx.get_time() == want_clock and ctx.dt_history(7) == want_dts
    assert np.array_equal(ctx.download(), want, equal_nan=True)
    # t_stop clip inside a replayed batch, after a fresh upload (astrea.py:74-75)
    t_stop = want_dts[0] + want_dts[1] + 0.3 * want_dts[2]
    for c in (ref, ctx):
        c.upload(g0)
        c.parity = 0
        c.set_time(0.0, t_stop)
    for _ in range(3):
        ref.step_async()
    ctx.run_steps(3)
    assert ctx.get_time() == ref.get_time() and ctx.get_time()[0] == t_stop
    assert np.array_equal(ctx.download(), ref.download(), equal_nan=True)
    ref.close()
    ctx.close()


def test_run_steps_falls_back_in_2d(hostsim_lib):
    from astrea_b200 import _native as N
    from cases import native_cfg
    meta = _meta("ll6", 24, 2, "ppm", "hllc", "ssprk(3,3)", None)
    g0 = initial_state("ll6", 24, 2, 1.4, True)
    want, dts = run_oracle(meta, g0, 3)
    ctx = N.Context(native_cfg(meta), lib=hostsim_lib)
    ctx.upload(g0)
    ctx.set_time(0.0, 0.0)
    ctx.run_steps(3)
    assert ctx.get_time()[1] == 3 and ctx.dt_history(3) == dts
    assert np.array_equal(ctx.download(), want, equal_nan=True)
    ctx.close()


@pytest.mark.parametrize("case", [("ll3", 64, "ppm", "hllc", "ssprk(3,3)"), ("ll6", 32, "ppm", "hllc", "ssprk(3,3)"),
                                  ("ll3", 32, "weno7", "hllc", "rk4"), ("ll3", 32, "weno5", "hllc", "ssprk(3,3)"),
                                  ("orszag-tang", 32, "ppm", "hlld", "ssprk(3,3)"), ("ll12", 32, "ppm", "hllc", "ssprk(3,3)")],
                         ids=lambda c: "-".join(map(str, c)))
def test_later_operators_still_raise_where_the_reference_does(hostsim_lib, case):
    """The operators after the first of a step evaluate the interface wave speeds only where they could be non-finite
    (FluxStage, `need_speed`): the reference computes them in every evolve_space call but uses them only to raise on
    NaN / Inf (fv.py:158).  The step and the half of the seam (evolve_space / evolve_time) that raises, and every grid
    before it, must not depend on that: same as with the speeds evaluated everywhere (flags bit 6)."""
    from astrea_b200 import _native as N
    from astrea_b200.simulation import Simulation
    config, cells, subgrid, solver, timestep = case
    runs = []
    for speeds in (False, True):
        sim = Simulation(config, cells, 2, subgrid, solver, timestep, _lib=hostsim_lib, stage_speeds=speeds)
        ctx, where, grids = sim.ctx, None, []
        for n in range(16):
            try:
                eig = ctx.evolve_space(n & 1)
            except N.NonFiniteError:
                where = (n, "evolve_space")
                break
            try:
                ctx.evolve_time(sim.cfl * min(sim.dx / e for e in eig))
            except N.NonFiniteError:
                where = (n, "evolve_time")
                break
            grids.append(ctx.download())
        runs.append((where, grids))
        sim.close()
    assert runs[0][0] == runs[1][0] and runs[0][0] is not None
    assert len(runs[0][1]) == len(runs[1][1])
    for a, b in zip(runs[0][1], runs[1][1]):
        assert np.array_equal(a, b)
