"""Function-level golden vectors from the UNMODIFIED reference (build container only).

    python tests/golden/make_function_golden.py

f_ppm_flattener.npz   schemes/ppm.py:111-134 ``apply_flattener(wS, axis, boundary)``: for a list of evolved states
                      (the final grids ``g`` of step-level golden cases, so shocks and contacts are smeared over a few
                      cells and the coefficient takes values in (0, 1)) and for every sweep axis: the primitive input
                      ``wS = convert_conservative(g.transpose(axes))`` exactly as ppm.run builds it (ppm.py:23-26) and
                      the coefficient ``chi = eta[..., 0]`` (the reference repeats it over the 8 variables).
                      Also the same states with strengthened pressure jumps (``sharp``), which reach the
                      z > z1 and the compressive-and-weak branches.
f_solution_error.npz  functions/analytic.py:24-44 ``calculate_solution_error(w, sim_variables, norm)`` on the primitive
                      snapshot (astrea.py:47) of a few smooth problems after a few steps of the reference's own loop, for
                      norm = 0, 1, 2, 3 and 11 (maximum): the conservative grid the snapshot was made of and the 10 errors.
``apply_artificial_viscosity`` (ppm.py:138-170) cannot produce vectors: it raises a broadcast error for every 1D grid
with N != 8 cells and for every 2D grid (ppm.py:164 / :154-156); recorded here as ``viscosity_raises``.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refharness as rh  # noqa: E402

SOURCES = ["x_sod_ppm_hllc_ssprk54", "x_rj_ppm_hlld_ssprk33", "x_shu_weno5_hllc_ssprk33_edge", "c5_ll6_ppm_hllc_ssprk33",
           "c2_ll3_ppm_hllc_ssprk33", "x_sod2d_ppm_hllc_ssprk33_edge", "x_sedov2d_ppm_llf_ssprk33", "c3_khi_weno5_hllc_ssprk33"]


def main():
    rh._import_ref()
    from schemes import ppm
    from oracle.reconstruct import ppm_flattener
    index = json.load(open(os.path.join(HERE, "index.json")))
    out, raises = {}, {}
    for cid in SOURCES:
        meta = index[cid]
        sv = rh.make_sim_variables(meta["config"], meta["cells"], meta["dimension"], "ppm", "hllc", "euler")
        g = np.load(os.path.join(HERE, cid + ".npz"))["g"]
        for axis, axes in sv.permutations.items():
            with np.errstate(all="ignore"):
                wS = sv.convert_conservative(g.transpose(axes), sv)
                for tag in ("", "sharp"):
                    w = np.copy(wS)
                    if tag == "sharp":
                        w[..., 4] = w[..., 4] ** 3          # stronger pressure jumps: other branches of zeta
                    eta = ppm.apply_flattener(np.copy(w), axis, sv.boundary)
                    assert all(np.array_equal(eta[..., 0], eta[..., k], equal_nan=True) for k in range(8))
                    chi = np.ascontiguousarray(eta[..., 0])
                    mine = ppm_flattener(np.copy(w), axis, sv.boundary)[..., 0]
                    assert np.array_equal(mine, chi, equal_nan=True), f"oracle flattener differs on {cid} axis {axis} {tag}"
                    key = f"{cid}|{axis}|{tag}"
                    out[key + "|w"] = np.ascontiguousarray(w)
                    out[key + "|chi"] = chi
                    print(key, w.shape, "chi in [%g, %g], %d cells strictly inside (0, 1)" % (np.nanmin(chi), np.nanmax(chi),
                                                                                         int(((chi > 0) & (chi < 1)).sum())))
                try:
                    ppm.apply_artificial_viscosity(np.copy(wS), axis, sv)
                    raises[f"{cid}|{axis}"] = None
                except ValueError as err:
                    raises[f"{cid}|{axis}"] = str(err)
    np.savez_compressed(os.path.join(HERE, "f_ppm_flattener.npz"), **out)
    meta = {"sources": SOURCES, "boundary": {cid: index[cid]["boundary"] for cid in SOURCES},
            "dimension": {cid: index[cid]["dimension"] for cid in SOURCES}, "viscosity_raises": raises}
    with open(os.path.join(HERE, "f_ppm_flattener.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)


ERROR_CASES = [("sin", 64, 1, "ppm", "hllc", "ssprk(3,3)", 5), ("sin", 50, 1, "plm", "lf", "ssprk(2,2)", 4),
               ("gauss", 32, 2, "weno5", "lf", "ssprk(3,3)", 3), ("gauss", 96, 1, "weno7", "hllc", "rk4", 3),
               ("ivc", 24, 2, "ppm", "hllc", "ssprk(3,3)", 2)]


def solution_error():
    rh._import_ref()
    from functions import analytic
    out, meta = {}, {}
    for config, cells, dim, subgrid, solver, timestep, steps in ERROR_CASES:
        sv = rh.make_sim_variables(config, cells, dim, subgrid, solver, timestep)
        g0 = rh.initial_grid(sv)
        grids, dts, _, _ = rh.run_steps(sv, steps, grid=np.copy(g0))
        g = grids[-1]
        with np.errstate(all="ignore"):
            w = sv.convert_conservative(np.copy(g), sv)
            key = f"{config}|{cells}|{dim}|{subgrid}"
            out[key + "|g"] = g
            for norm in (0, 1, 2, 3, 11):
                out[key + f"|err{norm}"] = analytic.calculate_solution_error(np.copy(w), sv, norm)
        meta[key] = dict(config=config, cells=cells, dimension=dim, subgrid=subgrid, solver=solver, timestep=timestep, steps=steps,
                         boundary=sv.boundary, gamma=sv.gamma)
        print(key, out[key + "|err1"][:5])
    np.savez_compressed(os.path.join(HERE, "f_solution_error.npz"), **out)
    with open(os.path.join(HERE, "f_solution_error.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
    solution_error()
