"""The numpy oracle against the golden vectors produced by the unmodified reference (tests/golden/make_golden.py)
and against the known answers recorded in BASELINE.md §3."""
import numpy as np
import pytest

from conftest import golden_case, golden_index, rel_l1
from cases import run_oracle
from astrea_b200.initial import initial_state
from oracle import OracleConfig, advance

CASES = sorted(golden_index())


@pytest.mark.parametrize("cid", CASES)
def test_oracle_reproduces_reference(cid):
    meta, data = golden_case(cid)
    g, used = run_oracle(meta, data["g0"], meta["steps"], eigen="lapack")
    # bit-identical in the build container; LAPACK's eigenvalue kernels may differ in the last bits elsewhere
    assert np.all(rel_l1(g, data["g"]) <= 1e-13)
    assert np.allclose(used, data["dts"], rtol=1e-13, atol=0)


@pytest.mark.parametrize("cid", CASES)
def test_closed_form_wave_speed_within_tolerance(cid):
    """|v_n| + c_fast in place of np.linalg.eigvals (what the device computes): well inside 1e-12 per step."""
    meta, data = golden_case(cid)
    g, used = run_oracle(meta, data["g0"], meta["steps"], eigen="closed")
    tol = 1e-10 if "weno7" in cid else 1e-12 * meta["steps"]
    assert np.all(rel_l1(g, data["g"]) <= tol)
    assert np.allclose(used, data["dts"], rtol=1e-11, atol=0)


@pytest.mark.parametrize("cid", CASES)
def test_initial_conditions_match_reference(cid):
    """astrea_b200.initial (host-side input preparation) against constructor.initialise of the reference."""
    meta, data = golden_case(cid)
    high = meta["subgrid"].startswith("w") or meta["subgrid"] in ("ppm", "parabolic", "p")
    g0 = initial_state(meta["config"], meta["cells"], meta["dimension"], meta["gamma"], high)
    assert np.array_equal(g0, data["g0"])


# BASELINE.md §3 — Σ|q| per variable after N steps and the dt sequence, generated from the reference
KNOWN = [
    ("sod", 1024, 1, "plm", "lf", "ssprk22", [4.126729759416585e-04, 3.106894820148085e-04, 2.782649959369863e-04, 2.648711179254350e-04],
     [5.760000000000000e+02, 1.167205083788287e+00, 0, 0, 1.408000000000000e+03, 0, 0, 0]),
    ("ll3", 64, 2, "ppm", "hllc", "ssprk33", [3.730167773024258e-03, 3.355688651911434e-03],
     [2.767462399999970e+03, 8.431282767772111e+02, 8.569207489018229e+02, 0, 6.448459795479216e+03, 0, 0, 0]),
    ("khi", 64, 2, "weno5", "hllc", "ssprk33", [9.282306462417374e-03, 9.251613532992821e-03, 8.743117630514439e-03, 8.534747437164800e-03],
     [6.143999999999997e+03, 2.972315159769582e+03, 1.949773245710208e+03, 0, 1.140156405993541e+04, 0, 0, 0]),
    ("orszag-tang", 64, 2, "plm", "hlld", "ssprk33", [3.632973295678505e-03, 3.602804527740569e-03, 3.548939560030396e-03, 3.501517168760009e-03],
     [9.054147873672281e+02, 5.761635496661055e+02, 5.763949374637637e+02, 0, 1.973804236460551e+03, 7.358803264760007e+02, 7.324133150726813e+02, 0]),
    ("ll6", 64, 2, "ppm", "hllc", "ssprk33", [3.955728315410522e-03, 3.949776294963158e-03, 3.929906665944331e-03, 3.939807856733347e-03],
     [7.168000000000009e+03, 5.253762699583806e+03, 3.512605743431352e+03, 0, 1.318359722222227e+04, 0, 0, 0]),
]


@pytest.mark.parametrize("config,cells,dim,subgrid,solver,timestep,dts,sums", KNOWN, ids=[k[0] for k in KNOWN])
def test_baseline_known_answers(config, cells, dim, subgrid, solver, timestep, dts, sums):
    from astrea_b200.initial import problem
    from astrea_b200.selectors import MAGNETIC_2D
    prob = problem(config, cells, 1.4)
    cfg = OracleConfig(config=config, cells=cells, dimension=dim, subgrid=subgrid, solver=solver, timestep=timestep,
                       boundary=prob["boundary"], dx=prob["dx"], magnetic_2d=config in MAGNETIC_2D)
    g0 = initial_state(config, cells, dim, 1.4, cfg.high_order)
    g, used = advance(g0, cfg, len(dts))
    assert np.allclose(used, dts, rtol=1e-12, atol=0)
    got = np.sum(np.abs(g), axis=tuple(range(dim)))
    assert np.allclose(got, sums, rtol=1e-12, atol=1e-300)


def test_sod_full_run_known_answer():
    """BASELINE.md §3 last row: 896 steps to t_end = 0.2, Σ|w| of the primitive solution."""
    from astrea_b200.initial import problem
    from oracle.gridops import prim_avg_of_cons_avg
    prob = problem("sod", 1024, 1.4)
    cfg = OracleConfig(config="sod", cells=1024, dimension=1, subgrid="plm", solver="lf", timestep="ssprk22",
                       boundary="edge", dx=prob["dx"])
    g = initial_state("sod", 1024, 1, 1.4, False)
    t, steps = 0.0, 0
    while t < prob["t_end"]:
        g, used = advance(g, cfg, 1, t=t, t_end=prob["t_end"])
        cfg.step_parity = 0          # 1D has a single permutation
        t += used[0]
        steps += 1
    assert steps == 896
    w = prim_avg_of_cons_avg(g, cfg)
    got = np.sum(np.abs(w), axis=0)
    want = [5.760000000000003e+02, 4.518987009026167e+02, 0, 0, 5.335174584852718e+02, 0, 0, 0]
    assert np.allclose(got, want, rtol=1e-12, atol=1e-300)
