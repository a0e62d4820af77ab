"""Parity of the sm_100a library with the oracle and the reference's golden vectors, through the C ABI, on a B200.

Tolerances (north star): 1e-12 relative L1 per variable after one step, 1e-10 after N steps.  The device arithmetic
is compiled without FMA contraction and with IEEE division / square root, so the comparison with the closed-form
oracle is expected to be bit-exact; the assertion is the north-star tolerance, the bit-equality is reported.
"""
import numpy as np
import pytest

from conftest import golden_case, golden_index, rel_l1
from cases import run_native, run_oracle
from astrea_b200.initial import initial_state, problem

pytestmark = pytest.mark.gpu
HYDRO = sorted(golden_index())      # every golden case, constrained-transport MHD included


@pytest.fixture(scope="module")
def lib():
    from astrea_b200 import _native
    return _native.device_library()      # raises if the CUDA library is missing: there is no fallback


@pytest.mark.parametrize("cid", HYDRO)
def test_golden_vs_reference_output(lib, cid):
    meta, data = golden_case(cid)
    got, _, eigs = run_native(lib, meta, data["g0"], meta["steps"], dts=list(data["dts"]))
    tol = 1e-10 if (meta["steps"] > 1 or "weno7" in cid) else 1e-12
    assert np.all(rel_l1(got, data["g"]) <= tol), rel_l1(got, data["g"])
    for n, e in enumerate(eigs):
        ref = list(data["eigmax"][n])
        ref = ref[::-1] if (n % 2 and meta["dimension"] == 2) else ref
        assert np.allclose(e, ref, rtol=1e-11, atol=0)


@pytest.mark.parametrize("cid", HYDRO)
def test_golden_single_step_vs_oracle(lib, cid):
    meta, data = golden_case(cid)
    want, dts = run_oracle(meta, data["g0"], 1)
    got, used, _ = run_native(lib, meta, data["g0"], 1)
    assert np.allclose(used, dts, rtol=1e-14, atol=0)
    assert np.all(rel_l1(got, want) <= 1e-12), rel_l1(got, want)


@pytest.mark.parametrize("cid", HYDRO)
def test_golden_all_steps_bit_exact_vs_oracle(lib, cid):
    meta, data = golden_case(cid)
    want, dts = run_oracle(meta, data["g0"], meta["steps"])
    got, used, _ = run_native(lib, meta, data["g0"], meta["steps"])
    assert np.all(rel_l1(got, want) <= 1e-10), rel_l1(got, want)
    assert used == dts and np.array_equal(got, want, equal_nan=True), "within tolerance but not bit-identical"


def _meta(config, cells, dim, subgrid, solver, timestep, bc, mhd=False):
    prob = problem(config, cells, 1.4)
    return dict(config=config, cells=cells, dimension=dim, subgrid=subgrid, solver=solver, timestep=timestep,
                boundary=bc or prob["boundary"], dx=prob["dx"], gamma=1.4, cfl=.5, magnetic_2d=mhd)


MATRIX = [(cfg, sub, sol, bc, dim)
          for dim, cfg in ((1, "sod"), (2, "ll4"))
          for sub in ("pcm", "plm", "ppm", "weno3", "weno5", "weno7")
          for sol in ("lf", "hllc")
          for bc in ("edge", "wrap")]


@pytest.mark.parametrize("config,subgrid,solver,bc,dim", MATRIX, ids=["-".join(map(str, m)) for m in MATRIX])
def test_scheme_solver_matrix(lib, config, subgrid, solver, bc, dim):
    """Ragged sizes and many blocks: 1D 1531 cells, 2D 203 x 203 (not multiples of any tile)."""
    cells = 1531 if dim == 1 else 203
    meta = _meta(config, cells, dim, subgrid, solver, "ssprk(3,3)", bc)
    high = subgrid.startswith("w") or subgrid == "ppm"
    g0 = initial_state(config, cells, dim, 1.4, high, boundary=bc)
    want, dts = run_oracle(meta, g0, 2)
    got, used, _ = run_native(lib, meta, g0, 2, segment_2d=37)
    assert np.all(rel_l1(got, want) <= 1e-10), rel_l1(got, want)
    assert np.allclose(used, dts, rtol=1e-13, atol=0)
    if dim == 2:
        # the same through the general 8-variable kernels (a grid without v_z / B normally takes the hydro variants)
        gen, used, _ = run_native(lib, meta, g0, 2, segment_2d=37, general_path=True)
        assert np.array_equal(gen, got, equal_nan=True)


MHD = [("orszag-tang", "plm", "hlld", "ssprk(3,3)", "wrap"), ("orszag-tang", "ppm", "hlld", "ssprk(3,3)", "wrap"),
       ("orszag-tang", "weno5", "hllc", "ssprk(2,2)", "wrap"), ("orszag-tang", "pcm", "hlld", "euler", "wrap"),
       ("mhd rotor", "plm", "hlld", "ssprk(3,3)", "wrap"), ("orszag-tang", "plm", "hlld", "ssprk(10,4)", "wrap"),
       ("orszag-tang", "weno3", "hlld", "ssprk(5,3)", "edge"), ("orszag-tang", "weno7", "hlld", "ssprk(5,4)", "wrap"),
       ("orszag-tang", "plm", "hllc", "rk4", "edge"), ("orszag-tang", "ppm", "hlld", "ssprk(4,3)", "edge")]


@pytest.mark.parametrize("config,subgrid,solver,timestep,bc", MHD, ids=["-".join(m) for m in MHD])
def test_constrained_transport(lib, config, subgrid, solver, timestep, bc):
    """BASELINE config 4 family (magnetic_2d): 150^2 (ragged against every tile), 3 steps (odd-step role swap, Q1b)."""
    cells = 150
    meta = _meta(config, cells, 2, subgrid, solver, timestep, bc, mhd=True)
    high = subgrid.startswith("w") or subgrid == "ppm"
    g0 = initial_state(config, cells, 2, 1.4, high, boundary=bc)
    want, dts = run_oracle(meta, g0, 3)
    got, used, _ = run_native(lib, meta, g0, 3)
    assert np.isfinite(want).all()
    assert np.all(rel_l1(got, want) <= 1e-10), rel_l1(got, want)
    assert np.allclose(used, dts, rtol=1e-13, atol=0)


@pytest.mark.parametrize("cells", [131, 77])
@pytest.mark.parametrize("subgrid,solver,bc", [("plm", "hlld", "wrap"), ("ppm", "hlld", "edge"), ("weno5", "hllc", "wrap")])
def test_constrained_transport_odd_sizes(lib, cells, subgrid, solver, bc):
    """Odd grid sizes: the column pitch is rounded to an even count for the bulk copies, warps of the staged flush are
    partly beyond the range, tiles of the refinement are ragged."""
    meta = _meta("orszag-tang", cells, 2, subgrid, solver, "ssprk(2,2)", bc, mhd=True)
    g0 = initial_state("orszag-tang", cells, 2, 1.4, subgrid != "plm", boundary=bc)
    want, dts = run_oracle(meta, g0, 2)
    got, used, _ = run_native(lib, meta, g0, 2)
    assert np.all(rel_l1(got, want) <= 1e-10), rel_l1(got, want)
    assert used == dts and np.array_equal(got, want, equal_nan=True)


def test_orszag_tang_1024_properties(lib):
    """Orszag-Tang at 1024^2 (the 4096^2 run of BASELINE config 4 is in test_gpu_fullsize.py): the result is independent
    of the launch geometry, and mass, momentum and energy totals are conserved to round-off."""
    cells = 1024
    meta = _meta("orszag-tang", cells, 2, "plm", "hlld", "ssprk(3,3)", None, mhd=True)
    g0 = initial_state("orszag-tang", cells, 2, 1.4, False)
    a, dts, _ = run_native(lib, meta, g0, 2)
    b, _, _ = run_native(lib, meta, g0, 2, segment_2d=100, threads_2d=64)
    assert np.array_equal(a, b)
    scale = np.abs(g0).sum(axis=(0, 1))
    drift = np.abs(a.sum(axis=(0, 1)) - g0.sum(axis=(0, 1)))
    assert np.all(drift[[0, 1, 2, 4]] <= 1e-9 * np.where(scale > 0, scale, 1)[[0, 1, 2, 4]])   # 1e6 cells of round-off


@pytest.mark.parametrize("timestep", ["euler", "rk4", "ssprk(2,2)", "ssprk(3,3)", "ssprk(4,3)", "ssprk(5,3)", "ssprk(5,4)", "ssprk(10,4)"])
def test_every_integrator(lib, timestep):
    meta = _meta("ll12", 96, 2, "plm", "hllc", timestep, None)
    g0 = initial_state("ll12", 96, 2, 1.4, False)
    want, dts = run_oracle(meta, g0, 3)
    got, used, _ = run_native(lib, meta, g0, 3)
    assert np.all(rel_l1(got, want) <= 1e-10)
    assert np.allclose(used, dts, rtol=1e-13, atol=0)


def test_sod_full_run(lib):
    """BASELINE config 1 to t_end = 0.2: 896 steps, parity 1e-10 with the oracle at the end."""
    from astrea_b200 import _native as N
    from cases import native_cfg, oracle_cfg
    from oracle import advance
    meta = _meta("sod", 1024, 1, "plm", "lf", "ssprk(2,2)", None)
    g0 = initial_state("sod", 1024, 1, 1.4, False)
    ctx = N.Context(native_cfg(meta), lib=lib)
    ctx.upload(g0)
    t, steps, dts = 0.0, 0, []
    while t < 0.2:
        dt = ctx.step(t, 0.2)
        dts.append(dt)
        t += dt
        steps += 1
    got = ctx.download()
    ctx.close()
    assert steps == 896
    cfg = oracle_cfg(meta)
    want = np.copy(g0)
    for dt in dts:
        want, _ = advance(want, cfg, 1, dts=[dt])
        cfg.step_parity = 0
    assert np.all(rel_l1(got, want) <= 1e-10)


def test_nonfinite_raises_linalgerror(lib):
    from astrea_b200 import _native as N
    from cases import native_cfg
    meta = _meta("ll3", 64, 2, "ppm", "hllc", "ssprk(3,3)", None)
    g0 = initial_state("ll3", 64, 2, 1.4, True)
    g0[5, 7, 0] = np.nan
    ctx = N.Context(native_cfg(meta), lib=lib)
    ctx.upload(g0)
    with pytest.raises(np.linalg.LinAlgError):
        ctx.step()
    ctx.close()


def test_reference_horizon_is_reproduced(lib):
    """The reference's LL3 PPM+HLLC run dies of LinAlgError in its third step: so does the device run."""
    from astrea_b200 import _native as N
    from cases import native_cfg
    meta = _meta("ll3", 128, 2, "ppm", "hllc", "ssprk(3,3)", None)
    ctx = N.Context(native_cfg(meta), lib=lib)
    ctx.upload(initial_state("ll3", 128, 2, 1.4, True))
    ctx.step()
    ctx.step()
    with pytest.raises(np.linalg.LinAlgError):
        ctx.step()
        ctx.step()
    ctx.close()


def test_drop_in_evolvers_functions(lib):
    """evolve_space / evolve_time with the reference's call signature and a sim_variables-like namedtuple."""
    from collections import namedtuple
    from astrea_b200 import evolvers
    from cases import oracle_cfg
    from oracle import advance
    meta = _meta("ll6", 64, 2, "ppm", "hllc", "ssprk(3,3)", None)
    SV = namedtuple("simulation_variables", "dimension cells boundary gamma dx cfl subgrid solver solver_category timestep magnetic_2d permutations")
    perms = {0: (0, 1, 2), 1: (1, 0, 2)}
    sv = SV(2, 64, meta["boundary"], 1.4, meta["dx"], .5, "ppm", "hllc", "hll", "ssprk(3,3)", False, perms)
    grid = initial_state("ll6", 64, 2, 1.4, True)
    want, dts = advance(np.copy(grid), oracle_cfg(meta), 3)
    for n in range(3):
        fluxes = evolvers.evolve_space(grid, sv)
        dt = sv.cfl * min(sv.dx / f["eigmax"] for f in fluxes.values())
        grid = evolvers.evolve_time(grid, fluxes, dt, sv)
        sv = sv._replace(permutations=dict(reversed(list(sv.permutations.items()))))     # astrea.py:85
    evolvers.release()
    assert np.all(rel_l1(grid, want) <= 1e-10)


# ---------------------------------------------------------------------------------------------- full-size properties
def _run_big(lib, cells, subgrid, solver, steps, g0, **geometry):
    meta = _meta("ll3", cells, 2, subgrid, solver, "ssprk(3,3)", "wrap")
    return run_native(lib, meta, g0, steps, **geometry)


def test_full_size_shift_equivariance_and_conservation(lib):
    """BASELINE config 2 size (2048^2, PPM+HLLC, SSPRK3, periodic): the oracle cannot run this size, so use
    properties.  (a) shifting the periodic initial data by (s_x, s_y) cells shifts the result, bit for bit — this
    exercises every block seam and the halo fill; (b) the conserved totals change only by round-off."""
    cells = 2048
    g0 = initial_state("ll3", cells, 2, 1.4, True)
    a, dts_a, _ = _run_big(lib, cells, "ppm", "hllc", 1, g0)
    shift = (301, 77)
    b, dts_b, _ = _run_big(lib, cells, "ppm", "hllc", 1, np.ascontiguousarray(np.roll(g0, shift, axis=(0, 1))))
    assert dts_a == dts_b
    assert np.array_equal(np.roll(a, shift, axis=(0, 1)), b)
    tot0, tot1 = g0.sum(axis=(0, 1)), a.sum(axis=(0, 1))
    scale = np.abs(g0).sum(axis=(0, 1))
    assert np.all(np.abs(tot1 - tot0) <= 1e-9 * np.where(scale > 0, scale, 1))       # 4e6 cells of round-off


def test_full_size_tiling_independence(lib):
    cells = 1024
    g0 = initial_state("ll3", cells, 2, 1.4, True)
    a, _, _ = _run_big(lib, cells, "ppm", "hllc", 2, g0)
    b, _, _ = _run_big(lib, cells, "ppm", "hllc", 2, g0, threads_2d=96, segment_2d=100)
    assert np.array_equal(a, b, equal_nan=True)


def test_medium_size_vs_oracle(lib):
    """256^2 PPM+HLLC SSPRK3, the largest size the oracle finishes in seconds: both finite steps."""
    meta = _meta("ll3", 256, 2, "ppm", "hllc", "ssprk(3,3)", None)
    g0 = initial_state("ll3", 256, 2, 1.4, True)
    want, dts = run_oracle(meta, g0, 2)
    got, used, _ = run_native(lib, meta, g0, 2)
    assert np.all(rel_l1(got, want) <= 1e-10)
    assert np.allclose(used, dts, rtol=1e-13, atol=0)


@pytest.mark.parametrize("spec", [("sod", 1, "plm", "lf", "ssprk(2,2)", 1024), ("ll6", 2, "ppm", "hllc", "ssprk(3,3)", 96),
                                  ("orszag-tang", 2, "plm", "hlld", "ssprk(3,3)", 64)], ids=["sod1d", "ll6", "ot"])
def test_async_stepping_with_device_clock(lib, spec):
    """astrea_step_async: dt and the t_stop clip (astrea.py:70-78) evaluated on the device, steps enqueued back to back."""
    from astrea_b200 import _native as N
    from astrea_b200.selectors import MAGNETIC_2D
    from cases import native_cfg, oracle_cfg
    from oracle import advance
    config, dim, subgrid, solver, timestep, cells = spec
    meta = _meta(config, cells, dim, subgrid, solver, timestep, None, mhd=config in MAGNETIC_2D)
    g0 = initial_state(config, cells, dim, 1.4, subgrid == "ppm")
    ctx = N.Context(native_cfg(meta), lib=lib)
    ctx.upload(g0)
    want, dts = run_oracle(meta, g0, 4)
    ctx.set_time(0.0, 0.0)
    for _ in range(4):
        ctx.step_async()
    t, steps, last = ctx.get_time()
    assert steps == 4 and np.allclose(ctx.dt_history(4), dts, rtol=1e-13, atol=0)
    assert np.all(rel_l1(ctx.download(), want) <= 1e-10)
    t_stop = dts[0] + 0.4 * dts[1]
    want, used = advance(np.copy(g0), oracle_cfg(meta), 2, t=0.0, t_end=t_stop)
    ctx.upload(g0)
    ctx.parity = 0
    ctx.set_time(0.0, t_stop)
    ctx.step_async()
    ctx.step_async()
    t, steps, last = ctx.get_time()
    assert steps == 2 and abs(t - t_stop) <= 1e-15 and np.allclose(ctx.dt_history(2), used, rtol=1e-12, atol=0)
    assert np.all(rel_l1(ctx.download(), want) <= 1e-10)
    ctx.close()


def test_device_diagnostics(lib):
    """Conservation totals and total variation (functions/analytic.py:48-77) reduced on the device, 2D and 1D."""
    from astrea_b200 import _native as N
    from cases import native_cfg, oracle_cfg
    from oracle.gridops import prim_avg_of_cons_avg
    for dim, config, cells, subgrid in ((2, "khi", 300, "ppm"), (1, "sod", 1000, "plm")):
        meta = _meta(config, cells, dim, subgrid, "hllc", "ssprk(2,2)", None)
        g0 = initial_state(config, cells, dim, 1.4, subgrid == "ppm")
        ctx = N.Context(native_cfg(meta), lib=lib)
        ctx.upload(g0)
        ctx.step()
        q = ctx.download()
        tot, tv = ctx.diagnostics()
        ctx.close()
        d = prim_avg_of_cons_avg(q, oracle_cfg(meta))
        for ax in range(dim):
            d = np.diff(d, axis=ax)
        axes = tuple(range(dim))
        scale = np.abs(q).sum(axis=axes)
        assert np.all(np.abs(tot - q.sum(axis=axes)) <= 1e-12 * np.where(scale > 0, scale, 1))   # summation order differs
        assert np.allclose(tv, np.abs(d).sum(axis=axes), rtol=1e-12, atol=1e-12)


def test_hllc_low_mach_switch(lib):
    """solvers.py:118-122 (off by default): goes through sin(), so the north-star tolerance applies, not bit equality."""
    from cases import oracle_cfg
    from oracle import advance
    for dim, config, cells in ((1, "sod", 512), (2, "khi", 128)):
        meta = _meta(config, cells, dim, "plm", "hllc", "ssprk(2,2)", None)
        g0 = initial_state(config, cells, dim, 1.4, False)
        cfg = oracle_cfg(meta)
        cfg.low_mach = True
        want, dts = advance(np.copy(g0), cfg, 2)
        got, used, _ = run_native(lib, meta, g0, 2, low_mach=True)
        assert np.allclose(used, dts, rtol=1e-13, atol=0)
        assert np.all(rel_l1(got, want) <= 2e-12)


@pytest.mark.parametrize("author", ["c", "ph"])
@pytest.mark.parametrize("config,dim,solver,bc", [("sod", 1, "hllc", None), ("ll3", 2, "hllc", None), ("ll4", 2, "lf", "edge"),
                                                  ("orszag-tang", 2, "hlld", None)])
def test_ppm_authors_colella_and_peterson_hammett(lib, config, dim, solver, bc, author):
    """ppm.run(author='c' | 'ph') (limiters.py:53-78,144-201): grid-wide any() switches via flag passes."""
    from astrea_b200.selectors import MAGNETIC_2D
    cells = 700 if dim == 1 else 130
    meta = _meta(config, cells, dim, "ppm", solver, "ssprk(3,3)", bc, mhd=config in MAGNETIC_2D)
    meta["ppm_author"] = author
    g0 = initial_state(config, cells, dim, 1.4, True, boundary=meta["boundary"])
    want, dts = run_oracle(meta, g0, 2)
    got, used, _ = run_native(lib, meta, g0, 2)
    assert np.all(rel_l1(got, want) <= 1e-10), rel_l1(got, want)
    assert np.allclose(used, dts, rtol=1e-13, atol=0)


def test_fast_division_and_square_root_are_ieee(lib):
    """The branch-free division / square root of the kernels (csrc/common.cuh, Fast) against the compiler's IEEE
    routines on 2^26 generated operand pairs: every operation Fast accepts must give the IEEE bits; the ordinary
    operand classes (half of the samples) must be accepted."""
    from astrea_b200 import _native as N
    from astrea_b200.selectors import make_cfg
    ctx = N.Context(make_cfg(dimension=1, cells=64, boundary="edge", gamma=1.4, dx=1 / 64, cfl=.5, subgrid="plm", solver="lf",
                             timestep="ssprk(2,2)"), lib=lib)
    try:
        accepted = wrong = declined = 0
        for seed in (1, 2, 3, 4):
            a, w, d = ctx.arith_check(1 << 24, seed)
            accepted += a; wrong += w; declined += d
    finally:
        ctx.close()
    assert wrong == 0, (accepted, wrong, declined)
    assert accepted > 0.45 * (accepted + declined), (accepted, declined)
