"""The two-pass structure of the guarded kernels (csrc/common.cuh: Fast first, Exact for the flagged warps / blocks),
exercised on the CPU: the audit build of the host simulation runs the same control flow with a guard whose results are
IEEE and whose flag follows Fast's rules.  A repeated pass must reproduce the single-pass result bit for bit — this is
the property an in-place kernel would break — on cases where flagged operations really occur (negative pressures
and zero denominators of the Lax-Liu problems near the end of their finite horizon)."""
import numpy as np
import pytest

from cases import run_native
from astrea_b200.initial import initial_state, problem


@pytest.fixture(scope="module")
def audit_lib():
    from astrea_b200 import _native, build
    return _native.bind(build.build(hostsim=True, variant="audit", defines=("ASTREA_AUDIT=1",)))


@pytest.mark.parametrize("config,cells,subgrid,solver,steps,bc", [
    ("ll3", 48, "ppm", "hllc", 2, None), ("ll4", 53, "weno5", "hllc", 2, "edge"), ("ll4", 53, "ppm", "lf", 2, "wrap"),
    ("ll6", 40, "plm", "hllc", 3, None)])
def test_repeated_pass_reproduces_single_pass(hostsim_lib, audit_lib, config, cells, subgrid, solver, steps, bc):
    prob = problem(config, cells, 1.4)
    meta = dict(config=config, cells=cells, dimension=2, subgrid=subgrid, solver=solver, timestep="ssprk(3,3)",
                boundary=bc or prob["boundary"], dx=prob["dx"], gamma=1.4, cfl=.5, magnetic_2d=False)
    g0 = initial_state(config, cells, 2, 1.4, subgrid != "plm", boundary=meta["boundary"])
    two, dts_two, _ = run_native(audit_lib, meta, g0, steps, segment_2d=19)
    one, dts_one, _ = run_native(hostsim_lib, meta, g0, steps, segment_2d=19)
    assert dts_two == dts_one
    assert np.array_equal(two, one, equal_nan=True)
