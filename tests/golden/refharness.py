"""Drive the UNMODIFIED reference (mervyzr/astrea, read-only at /root/reference) from Python.

Test infrastructure only.  This file is used in the build container to (a) pin the numpy
oracle in ``oracle/`` bit-for-bit against the reference and (b) generate the golden vectors
committed under ``tests/golden/``.  ``/root/reference`` does not exist on the GPU box, so
nothing here is imported by ``-m gpu`` tests, ``bench.py`` or ``__graft_entry__.smoke()``.

The reference's ``astrea.py`` / ``functions/generic.py`` cannot be imported (h5py, tinydb are
absent), so ``sim_variables`` is assembled by hand following ``functions/generic.py:159-286``
and ``astrea.py:119-133``; the time loop below is ``astrea.py:67,70-71,81,84-85``.
"""
import itertools
import os
import sys
from collections import namedtuple

import numpy as np

REF_ROOT = os.environ.get("ASTREA_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "num_methods"))


def _import_ref():
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from functions import constructor, fv  # noqa
    from num_methods import evolvers  # noqa
    from static import tests as ictable  # noqa
    return constructor, fv, evolvers, ictable


_SOLVER_CATEGORY = {  # static/.db.json 'solver' rows
    "lax": ["lf", "friedrich", "lax-friedrich", "llf", "local lax-friedrich", "lw", "lax-wendroff", "wendroff"],
    "hll": ["hllc", "c", "hlld", "d"],
    "complete": ["os", "osher", "solomon", "osher-solomon", "osher solomon", "es", "entropy", "entropy-stable"],
}
_MAG2D = ["orszag-tang", "orszag", "tang", "ot", "mhd rotor", "mhd-rotor", "rotor", "mhd blast", "mhd-blast",
          "mhd blast wave", "mhd-blast-wave"]


def make_sim_variables(config, cells, dimension, subgrid, solver, timestep, cfl=0.5, gamma=1.4, boundary=None):
    constructor, fv, evolvers, ictable = _import_ref()
    config, subgrid, solver, timestep = config.lower(), subgrid.lower(), solver.lower(), timestep.lower()
    d = dict(config=config, cells=int(cells), cfl=cfl, gamma=gamma, dimension=dimension, precision="float64",
             subgrid=subgrid, timestep=timestep, solver=solver, run_type="single", checkpoints=1,
             live_plot=False, take_snaps=False, save_plots=False, save_video=False, save_file=False,
             quiet=True, seed=0)
    d["solver_category"] = [k for k, v in _SOLVER_CATEGORY.items() if solver in v][0]
    d["magnetic_2d"] = config in _MAG2D
    if subgrid.startswith("w") or subgrid in ["ppm", "parabolic", "p"]:
        d["convert_primitive"] = fv.high_order_convert_primitive
        d["convert_conservative"] = fv.high_order_convert_conservative
    else:
        d["convert_primitive"] = fv.point_convert_primitive
        d["convert_conservative"] = fv.point_convert_conservative
    perms = [a for a in itertools.permutations(range(dimension + 1)) if a[-1] == dimension]
    d["permutations"] = {i: a for i, a in enumerate(perms)}
    d["ortho_axis"] = perms[-1]
    d.update(ictable.generate_test_conditions(config, cells, gamma))
    if boundary is not None:
        d["boundary"] = boundary
    return namedtuple("simulation_variables", d)(**d)


def initial_grid(sv):
    constructor, _, _, _ = _import_ref()
    return constructor.initialise(sv, convert=True)


def run_steps(sv, nsteps, grid=None, dts=None, record=True):
    """astrea.py:45-85 without HDF5/printing.  Returns (grids after each step, dts, eigmaxes)."""
    _, _, evolvers, _ = _import_ref()
    if grid is None:
        grid = initial_grid(sv)
    out, used_dt, eigs = [], [], []
    with np.errstate(all="ignore"):
        for n in range(nsteps):
            fluxes = evolvers.evolve_space(grid, sv)
            eig = [v["eigmax"] for v in fluxes.values()]
            dt = sv.cfl * min(sv.dx / e for e in eig) if dts is None else dts[n]
            grid = evolvers.evolve_time(grid, fluxes, dt, sv)
            sv = sv._replace(permutations=dict(reversed(list(sv.permutations.items()))))
            used_dt.append(float(dt))
            eigs.append([float(e) for e in eig])
            if record:
                out.append(np.copy(grid))
    return (out if record else grid), used_dt, eigs, sv
