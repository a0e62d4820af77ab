"""Function-level golden vectors from the UNMODIFIED reference (build container only).

    python tests/golden/make_function_golden.py

f_ppm_flattener.npz   schemes/ppm.py:111-134 ``apply_flattener(wS, axis, boundary)``: for a list of evolved states
                      (the final grids ``g`` of step-level golden cases, so shocks and contacts are smeared over a few
                      cells and the coefficient takes values in (0, 1)) and for every sweep axis: the primitive input
                      ``wS = convert_conservative(g.transpose(axes))`` exactly as ppm.run builds it (ppm.py:23-26) and
                      the coefficient ``chi = eta[..., 0]`` (the reference repeats it over the 8 variables).
                      Also the same states with strengthened pressure jumps (``sharp``), which reach the
                      z > z1 and the compressive-and-weak branches.
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


if __name__ == "__main__":
    main()
