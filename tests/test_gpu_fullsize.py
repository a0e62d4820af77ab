"""BASELINE.json's full sizes on a B200, beyond what the oracle can run: size-independent properties.

  * periodic shift equivariance, bit for bit: advancing the initial data rolled by (s_x, s_y) cells gives the rolled
    result and the same dt.  Every block seam, tile edge and the halo fill sit somewhere else in the two runs.
  * conservation: the totals of the conserved variables change by summation round-off only (periodic box).
  * independence of the launch geometry, bit for bit.
Config 2 (2048^2) is in test_gpu_parity.py; here: config 3 (KHI 4096^2 WENO5 + HLLC), config 4 (Orszag-Tang 4096^2
PLM + HLLD + constrained transport) and config 5, the headline (Lax-Liu 6 8192^2 PPM + HLLC), all SSPRK(3,3).
"""
import numpy as np
import pytest

from astrea_b200.initial import initial_state, problem

pytestmark = pytest.mark.gpu


def _context(config, cells, subgrid, solver, mhd=False, **geometry):
    from astrea_b200 import _native as N
    from astrea_b200.selectors import make_cfg
    prob = problem(config, cells, 1.4)
    cfg = make_cfg(dimension=2, cells=cells, boundary="wrap", gamma=1.4, dx=prob["dx"], cfl=.5, subgrid=subgrid, solver=solver,
                   timestep="ssprk(3,3)", magnetic_2d=mhd, **geometry)
    return N.Context(cfg, lib=N.device_library())


def _conserved(ctx):
    tot, _ = ctx.diagnostics()
    return tot


def _shift_and_conservation(ctx, g0, shift, steps=1):
    ctx.upload(g0)
    tot0 = _conserved(ctx)
    dts_a = [ctx.step() for _ in range(steps)]
    tot1 = _conserved(ctx)
    a = ctx.download()
    rolled = np.ascontiguousarray(np.roll(g0, shift, axis=(0, 1)))
    ctx.upload(rolled)
    del rolled
    ctx.parity = 0
    dts_b = [ctx.step() for _ in range(steps)]
    b = ctx.download()
    assert dts_a == dts_b
    a = np.roll(a, shift, axis=(0, 1))
    assert np.array_equal(a, b)
    # sums of `cells` terms against the size of the terms (a total may cancel to ~0): round-off ~ sqrt(cells) * 2^-53
    scale = np.abs(g0).sum(axis=(0, 1))
    assert np.all(np.abs(tot1 - tot0) <= 1e-10 * np.where(scale > 0, scale, 1.0)), (tot0, tot1, scale)


def test_config3_khi_4096_weno5_hllc():
    cells = 4096
    ctx = _context("khi", cells, "weno5", "hllc")
    try:
        g0 = initial_state("khi", cells, 2, 1.4, True)
        _shift_and_conservation(ctx, g0, (1301, 77))
    finally:
        ctx.close()


def test_config5_ll6_8192_ppm_hllc():
    """The headline configuration at its full size (47 GB of HBM): device-side initial conditions, one step."""
    from astrea_b200.initial import piecewise_spec
    cells = 8192
    ctx = _context("ll6", cells, "ppm", "hllc")
    try:
        ctx.init_piecewise(piecewise_spec("ll6", cells, 1.4))
        g0 = ctx.download()
        # the device-side initial grid equals the reference's constructor output on a strip the host can afford
        want = initial_state("ll6", cells, 2, 1.4, True)
        assert np.array_equal(g0, want)
        del want
        _shift_and_conservation(ctx, g0, (4099, 513))
    finally:
        ctx.close()


def test_config4_orszag_tang_4096_plm_hlld_ct():
    """(a) conservation of mass, momentum and energy totals; (b) independence of the launch geometry, bit for bit.
    (div B of the face field is not a property of the reference: it re-derives the face field from the cell averages
    by reconstruction every step, mag_field.py:191-211, and the re-derived field is not divergence-free.)"""
    cells = 4096
    g0 = initial_state("orszag-tang", cells, 2, 1.4, False)
    results = []
    for geometry in ({}, {"segment_2d": 100, "threads_2d": 64}):
        ctx = _context("orszag-tang", cells, "plm", "hlld", mhd=True, **geometry)
        try:
            ctx.upload(g0)
            tot0 = _conserved(ctx)
            dts = [ctx.step(), ctx.step()]
            tot1 = _conserved(ctx)
            results.append((ctx.download(), dts))
            scale = np.abs(g0).sum(axis=(0, 1))
            assert np.all(np.abs(tot1 - tot0)[[0, 1, 2, 4]] <= 1e-10 * np.where(scale > 0, scale, 1)[[0, 1, 2, 4]])
        finally:
            ctx.close()
    assert results[0][1] == results[1][1]
    assert np.array_equal(results[0][0], results[1][0])
