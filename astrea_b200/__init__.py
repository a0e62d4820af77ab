"""astrea_b200 — B200-native (sm_100a) per-timestep finite-volume update of mervyzr/astrea.

Host side of the C ABI in ``include/astrea_b200.h``:

* ``astrea_b200.evolvers``  — drop-in for the reference's ``num_methods/evolvers.py`` (``evolve_space`` / ``evolve_time``)
* ``astrea_b200.Simulation`` — device-resident time loop, slab-decomposed over several GPUs
* ``astrea_b200.Context``    — thin wrapper of one ``astrea_ctx``

The native library is loaded on first use; there is no CPU fallback (see ``_native.device_library``).
"""
from ._native import AstreaError, Cfg, Context, NonFiniteError  # noqa: F401
from .selectors import cfg_from_sim_variables, make_cfg  # noqa: F401
from .simulation import Simulation  # noqa: F401
from . import evolvers  # noqa: F401

__all__ = ["AstreaError", "Cfg", "Context", "NonFiniteError", "Simulation", "cfg_from_sim_variables", "evolvers", "make_cfg"]
