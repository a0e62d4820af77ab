"""Drop-in for the two dissipation helpers of the reference's ``schemes/ppm.py``, run on the GPU.

    eta = apply_flattener(wS, axis, boundary)                       # schemes/ppm.py:111-134
    mu  = apply_artificial_viscosity(wS, axis, sim_variables)       # schemes/ppm.py:138-170

Same names, argument meaning and return shapes as the reference.  ``ppm.run(dissipate=True)`` itself cannot run in the
reference (ppm.py:67 raises a broadcast error), so these are function-level counterparts, not part of a time step.

``apply_artificial_viscosity``: the reference multiplies ``np.minimum(...)`` of shape (N,) by ``lambda_R[..., None]`` of
shape (N, 1) (ppm.py:164), which makes an (N, N) array and then fails against the (N, 8) state for every N != 8; in 2D it
fails earlier (ppm.py:154-156).  The device kernel evaluates the formula cell by cell (nu_i = min(1, ...)_i * lambda_i,
McCorquodale & Colella 2011 eq. 36-38), which is what the line computes once both factors carry the same index; 2D
input raises ValueError like the reference.  Parity for it is against the oracle's cell-wise restatement only
("parity unpinned": the reference has no output to compare with).
"""
import numpy as np

from . import _native as N
from .selectors import make_cfg

_contexts = {}


def _context(shape, boundary, gamma=1.4, dx=1.0, device=0, _lib=None):
    key = (tuple(shape), boundary, float(gamma), float(dx), device, id(_lib))
    ctx = _contexts.get(key)
    if ctx is None:
        dim = len(shape) - 1
        cfg = make_cfg(dimension=dim, nx=shape[0], ny=shape[1] if dim == 2 else 1, boundary=boundary, gamma=gamma, dx=dx, cfl=.5,
                       subgrid="ppm", solver="hllc", timestep="euler", device=device)
        ctx = _contexts[key] = N.Context(cfg, lib=_lib)
    return ctx


def release():
    for ctx in _contexts.values():
        ctx.close()
    _contexts.clear()


def apply_flattener(wS, axis, boundary, slope_determinants=(.33, .75, .85), device=0, _lib=None):
    """Coefficient of the slope flattener [Colella 1990], repeated over the variables like the reference's return value."""
    wS = np.asarray(wS, dtype=np.float64)
    chi = _context(wS.shape, boundary, device=device, _lib=_lib).ppm_flattener(wS, axis, slope_determinants)
    return np.ones_like(wS) * chi[..., None]


def apply_artificial_viscosity(wS, axis, sim_variables, viscosity_determinants=(.3, .3), device=0, _lib=None):
    wS = np.asarray(wS, dtype=np.float64)
    if wS.ndim != 2:
        raise ValueError("operands could not be broadcast together (the reference's 2D branch, ppm.py:154-156, raises this)")
    sv = sim_variables
    ctx = _context(wS.shape, sv.boundary, gamma=sv.gamma, dx=sv.dx, device=device, _lib=_lib)
    return ctx.ppm_viscosity(wS, axis, viscosity_determinants)
