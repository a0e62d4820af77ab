"""Drop-in for the reference's ``num_methods/evolvers.py``: the same two functions, run on the GPU.

    fluxes = evolve_space(grid, sim_variables)              # astrea.py:67   (num_methods/evolvers.py:12-34)
    grid   = evolve_time(grid, fluxes, dt, sim_variables)   # astrea.py:81   (num_methods/evolvers.py:38-206)

``grid`` is the reference's C-contiguous float64 ndarray of conservative cell averages, ``(N, 8)`` or
``(N, N, 8)``; ``fluxes`` is a dict keyed by the permutation tuples of ``sim_variables.permutations`` in their
current iteration order, and ``fluxes[axes]['eigmax']`` is what astrea.py:70-71 reads.  The Riemann fluxes stay
on the device: ``fluxes[axes]['flux']`` is an opaque handle that only ``evolve_time`` understands.

Each call crosses the host/device boundary (upload in ``evolve_space``, download in ``evolve_time``), which is
what a maintainer gets by re-pointing the two lines of astrea.py.  The array ``evolve_time`` returns lives in
page-locked memory (``_native.PinnedPool``): the reference loop passes it straight back to the next ``evolve_space``,
so only the very first upload of a run is a pageable copy.  ``astrea_b200.Simulation`` keeps the grid resident on
the device instead.
"""
import numpy as np

from . import _native as N
from .selectors import cfg_from_sim_variables

_contexts = {}
_pools = {}


class DeviceFlux:
    """Handle of the stage-1 operator result held by a context (flux differences + rate in HBM)."""

    def __init__(self, ctx, token, axes):
        self.ctx, self.token, self.axes = ctx, token, axes

    def __repr__(self):
        return f"<DeviceFlux axes={self.axes} token={self.token}>"


def _key(sv, device):
    return (sv.dimension, sv.cells, sv.boundary, float(sv.gamma), float(sv.dx), float(sv.cfl), sv.subgrid.lower(),
            sv.solver.lower(), sv.timestep.lower(), bool(getattr(sv, "magnetic_2d", False)), device)


def _context(sv, device=0, _lib=None):
    key = _key(sv, device) + (id(_lib),)
    ctx = _contexts.get(key)
    if ctx is None:
        ctx = N.Context(cfg_from_sim_variables(sv, device=device), lib=_lib)
        ctx._token = 0
        _contexts[key] = ctx
    return ctx


def _pool(ctx, device):
    key = (id(ctx.lib), device)
    if key not in _pools:
        _pools[key] = N.PinnedPool(ctx.lib, device)
    return _pools[key]


def release():
    """Free every cached device context and the spare pinned arrays."""
    for ctx in _contexts.values():
        ctx.close()
    _contexts.clear()
    for pool in _pools.values():
        pool.close()
    _pools.clear()


def _parity(sv):
    """astrea.py:85 reverses the order of ``permutations`` every step: first key 0 -> even, else odd."""
    return 0 if next(iter(sv.permutations)) == 0 else 1


def evolve_space(grid, sim_variables, device=0, _lib=None):
    ctx = _context(sim_variables, device, _lib)
    ctx.upload(grid)
    eig = ctx.evolve_space(_parity(sim_variables))    # raises NonFiniteError (a LinAlgError) like fv.py:158
    ctx._token += 1
    fluxes = {}
    for axis, axes in sim_variables.permutations.items():
        fluxes[axes] = {"flux": DeviceFlux(ctx, ctx._token, axes), "eigmax": eig[axis]}
    return fluxes


def evolve_time(grid, fluxes, dt, sim_variables, device=0, _lib=None, out=None):
    ctx = _context(sim_variables, device, _lib)
    handle = next(iter(fluxes.values()))["flux"]
    if not isinstance(handle, DeviceFlux) or handle.ctx is not ctx or handle.token != ctx._token:
        raise ValueError("evolve_time: `fluxes` must come from the latest evolve_space call on this grid")
    ctx.evolve_time(dt)
    if out is None:
        out = _pool(ctx, device).empty(ctx.shape)      # a fresh array, as evolvers.py:206 returns one; page-locked
    out = ctx.download(out=out)      # `out`: optional caller-provided host array for the new grid
    return out.reshape(np.shape(grid))
