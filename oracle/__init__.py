"""CPU oracle for astrea's per-timestep finite-volume update.

TEST INFRASTRUCTURE ONLY.  This package is a numpy restatement of the reference algorithm
(mervyzr/astrea: functions/fv.py, functions/constructor.py, schemes/*.py, num_methods/*.py),
written index-shift style so that it doubles as the specification of the CUDA kernels.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker / the timed CPU arm — the product
package ``astrea_b200`` never imports it and has no CPU fallback.

Parity pinning: the reference has no tests or golden vectors of its own, so the
oracle is pinned by running the unmodified reference in the build container
(``tests/golden/make_golden.py``): every function here is compared bit-for-bit with the
reference on all BASELINE configs, and the resulting states are committed as fixtures under
``tests/golden/`` which ``tests/test_oracle_golden.py`` re-checks wherever the suite runs.
"""
from .config import OracleConfig  # noqa: F401
from .stepper import space_operator, time_update, advance, timestep_from_eigmax  # noqa: F401
