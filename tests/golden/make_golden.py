"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

For every case the reference's own ``constructor.initialise`` builds the initial grid and the reference's
``evolve_space`` / ``evolve_time`` are driven exactly as astrea.py:67-85 does (tests/golden/refharness.py).
Stored per case: the initial grid ``g0``, the grid after the last step ``g``, every ``dt`` and the per-axis
``eigmax`` of every step.  The script also re-checks that the numpy oracle reproduces each case bit for bit
(that is how the oracle is pinned; SURVEY.md §8c says the reference has no golden vectors of its own).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refharness as rh  # noqa: E402
from oracle import OracleConfig, advance  # noqa: E402

# (id, config, cells, dimension, subgrid, solver, timestep, steps, boundary override)
CASES = [
    ("c1_sod_plm_llf_ssprk22", "sod", 1024, 1, "plm", "lf", "ssprk(2,2)", 4, None),
    ("c2_ll3_ppm_hllc_ssprk33", "ll3", 64, 2, "ppm", "hllc", "ssprk(3,3)", 2, None),
    ("c3_khi_weno5_hllc_ssprk33", "khi", 64, 2, "weno5", "hllc", "ssprk(3,3)", 4, None),
    ("c4_ot_plm_hlld_ssprk33", "orszag-tang", 64, 2, "plm", "hlld", "ssprk(3,3)", 4, None),
    ("c5_ll6_ppm_hllc_ssprk33", "ll6", 64, 2, "ppm", "hllc", "ssprk(3,3)", 4, None),
    ("x_ll3_pcm_llf_euler", "ll3", 32, 2, "pcm", "lf", "euler", 3, None),
    ("x_sod_ppm_hllc_ssprk54", "sod", 128, 1, "ppm", "hllc", "ssprk(5,4)", 3, None),
    ("x_ll3_weno7_llf_rk4", "ll3", 32, 2, "weno7", "lf", "rk4", 2, None),
    ("x_khi_weno3_hllc_ssprk104", "khi", 32, 2, "weno3", "hllc", "ssprk(10,4)", 2, None),
    ("x_sod_weno5_llf_ssprk53", "sod", 64, 1, "weno", "lf", "ssprk(5,3)", 3, None),
    ("x_sod_pcm_hllc_ssprk43", "sod", 64, 1, "pcm", "hllc", "ssprk(4,3)", 3, None),
    ("x_sod2d_ppm_hllc_ssprk33_edge", "sod", 32, 2, "ppm", "hllc", "ssprk(3,3)", 3, None),
    ("x_sod2d_plm_llf_ssprk22_edge", "sod", 32, 2, "plm", "lf", "ssprk(2,2)", 3, None),
    ("x_sedov2d_ppm_llf_ssprk33", "sedov", 32, 2, "ppm", "lf", "ssprk(3,3)", 3, None),
    ("x_shu_weno5_hllc_ssprk33_edge", "shu-osher", 128, 1, "weno5", "hllc", "ssprk(3,3)", 3, None),
    ("x_bw_plm_hlld_ssprk22", "brio-wu", 128, 1, "plm", "hlld", "ssprk(2,2)", 3, None),
    ("x_rj_ppm_hlld_ssprk33", "ryu-jones", 128, 1, "ppm", "hlld", "ssprk(3,3)", 3, None),
    ("x_ot_plm_llf_ssprk22_ct", "orszag-tang", 32, 2, "plm", "lf", "ssprk(2,2)", 3, None),
    ("x_ot_ppm_hllc_ssprk33_ct", "orszag-tang", 32, 2, "ppm", "hllc", "ssprk(3,3)", 2, None),
    # Lax-Wendroff (solvers.py:79-88), states whose averaged pressures stay positive (real spectrum)
    ("x_sod_plm_lw_ssprk22", "sod", 128, 1, "plm", "lw", "ssprk(2,2)", 3, None),
    ("x_khi_plm_lw_ssprk22", "khi", 32, 2, "plm", "lw", "ssprk(2,2)", 3, None),
    ("x_toro2_pcm_lw_euler", "toro2", 64, 1, "pcm", "lw", "euler", 3, None),
    ("x_sq_weno3_lw_ssprk22", "square", 48, 1, "weno3", "lw", "ssprk(2,2)", 2, None),
]


# PPM authors 'c' / 'ph' (limiters.py:53-78,144-201) are reachable only through ppm.run(author=...): evolve_space
# hard-codes 'mc' (evolvers.py:17).  One forward-Euler step from ppm.run(author) -> calculate_Riemann_flux -> evolve_time.
AUTHOR_CASES = [
    ("a_sod_ppm_c_hllc", "sod", 128, 1, "hllc", "c"),
    ("a_ll3_ppm_ph_hllc", "ll3", 32, 2, "hllc", "ph"),
    ("a_khi_ppm_c_llf", "khi", 32, 2, "lf", "c"),
    ("a_shu_ppm_ph_llf", "shu-osher", 128, 1, "lf", "ph"),
]


def author_cases(index):
    rh._import_ref()
    from schemes import ppm
    from num_methods import evolvers, solvers
    from oracle import space_operator, time_update
    for cid, config, cells, dim, solver, author in AUTHOR_CASES:
        sv = rh.make_sim_variables(config, cells, dim, "ppm", solver, "euler")
        g0 = rh.initial_grid(sv)
        with np.errstate(all="ignore"):
            fluxes = solvers.calculate_Riemann_flux(ppm.run(np.copy(g0), sv, author=author), sv)
            eig = [float(v["eigmax"]) for v in fluxes.values()]
            dt = sv.cfl * min(sv.dx / e for e in eig)
            g1 = evolvers.evolve_time(np.copy(g0), fluxes, dt, sv)
        cfg = OracleConfig(config=config, cells=cells, dimension=dim, subgrid="ppm", solver=solver, timestep="euler",
                           boundary=sv.boundary, dx=sv.dx, ppm_author=author)
        with np.errstate(all="ignore"):
            go = time_update(np.copy(g0), space_operator(np.copy(g0), cfg), dt, cfg)
        pinned = bool(np.array_equal(go, g1, equal_nan=True))
        assert pinned, f"oracle differs from the reference on {cid}"
        np.savez_compressed(os.path.join(HERE, cid + ".npz"), g0=g0, g=g1, dts=np.array([dt]), eigmax=np.array([eig]))
        index[cid] = dict(config=config, cells=cells, dimension=dim, subgrid="ppm", solver=solver, timestep="euler", steps=1,
                          boundary=sv.boundary, dx=sv.dx, gamma=sv.gamma, cfl=sv.cfl, magnetic_2d=False,
                          finite=bool(np.isfinite(g1).all()), oracle_bit_equal=pinned, ppm_author=author)
        print(cid, "oracle bit-equal:", pinned)


def main():
    index = {}
    author_cases(index)
    for cid, config, cells, dim, subgrid, solver, timestep, steps, bc in CASES:
        sv = rh.make_sim_variables(config, cells, dim, subgrid, solver, timestep, boundary=bc)
        g0 = rh.initial_grid(sv)
        grids, dts, eigs, _ = rh.run_steps(sv, steps, grid=np.copy(g0))
        cfg = OracleConfig(config=config, cells=cells, dimension=dim, subgrid=subgrid, solver=solver, timestep=timestep,
                           boundary=sv.boundary, dx=sv.dx, magnetic_2d=sv.magnetic_2d)
        go, used = advance(np.copy(g0), cfg, steps)
        pinned = bool(np.array_equal(go, grids[-1], equal_nan=True) and used == dts)
        assert pinned, f"oracle differs from the reference on {cid}"
        np.savez_compressed(os.path.join(HERE, cid + ".npz"), g0=g0, g=grids[-1], dts=np.array(dts), eigmax=np.array(eigs))
        index[cid] = dict(config=config, cells=cells, dimension=dim, subgrid=subgrid, solver=solver, timestep=timestep,
                          steps=steps, boundary=sv.boundary, dx=sv.dx, gamma=sv.gamma, cfl=sv.cfl,
                          magnetic_2d=bool(sv.magnetic_2d), finite=bool(np.isfinite(grids[-1]).all()),
                          oracle_bit_equal=pinned)
        print(cid, "finite" if index[cid]["finite"] else "NON-FINITE", "oracle bit-equal:", pinned)
    with open(os.path.join(HERE, "index.json"), "w") as fh:
        json.dump(index, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
