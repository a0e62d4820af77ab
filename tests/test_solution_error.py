"""functions/analytic.py:24-44 calculate_solution_error as a device reduction, against vectors the unmodified
reference produced (tests/golden/f_solution_error.npz, made by tests/golden/make_function_golden.py).  The per-cell
differences are bit-identical to the reference's; only the summation order differs (block partials vs numpy's
pairwise sum), hence the 1e-12 relative tolerance.  Host simulation here, the sm_100a library under ``-m gpu``."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

VECTORS = np.load(os.path.join(GOLDEN, "f_solution_error.npz"))
META = json.load(open(os.path.join(GOLDEN, "f_solution_error.json")))


def _check(lib, key):
    from astrea_b200.simulation import Simulation
    m = META[key]
    sim = Simulation(m["config"], m["cells"], m["dimension"], m["subgrid"], m["solver"], m["timestep"], grid=VECTORS[key + "|g"], _lib=lib)
    try:
        for norm in (0, 1, 2, 3, 11):
            want = VECTORS[key + f"|err{norm}"]
            got = sim.solution_error(norm)
            assert got.shape == (10,)
            assert np.allclose(got, want, rtol=1e-12, atol=1e-300), (key, norm, got, want)
    finally:
        sim.close()


@pytest.mark.parametrize("key", sorted(META))
def test_hostsim_solution_error(hostsim_lib, key):
    _check(hostsim_lib, key)


def test_theoretical_state_equals_reference_initialise():
    """initial.theoretical_primitives == constructor.initialise(sim_variables) (build container only)."""
    import sys
    sys.path.insert(0, GOLDEN)
    import refharness as rh
    if not rh.available():
        pytest.skip("reference checkout not present")
    from astrea_b200.initial import theoretical_primitives
    constructor = rh._import_ref()[0]
    for key, m in META.items():
        sv = rh.make_sim_variables(m["config"], m["cells"], m["dimension"], m["subgrid"], m["solver"], m["timestep"])
        want = constructor.initialise(sv)
        got = theoretical_primitives(m["config"], m["cells"], m["dimension"], m["gamma"])
        assert np.array_equal(got, want), key


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(META))
def test_gpu_solution_error(key):
    from astrea_b200 import _native
    _check(_native.device_library(), key)
