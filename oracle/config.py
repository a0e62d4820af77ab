"""Run configuration of the oracle (test infrastructure; see oracle/__init__.py).

Mirrors the fields of the reference's ``sim_variables`` namedtuple that the hot path reads
(functions/generic.py:159-286, static/tests.py:317-328) and the string rules by which the
reference picks a scheme (num_methods/evolvers.py:14-21), a Riemann solver
(num_methods/solvers.py:13-31), a time integrator (num_methods/evolvers.py:79-81,187,204)
and the pointwise / 4th-order conversions (functions/generic.py:250-255).
"""
from dataclasses import dataclass, field

SOLVER_CATEGORY = {  # static/.db.json rows with type == 'solver'
    "lax": ("lf", "friedrich", "lax-friedrich", "llf", "local lax-friedrich", "lw", "lax-wendroff", "wendroff"),
    "hll": ("hllc", "c", "hlld", "d"),
    "complete": ("os", "osher", "solomon", "osher-solomon", "osher solomon", "es", "entropy", "entropy-stable"),
}
MAGNETIC_2D_CONFIGS = ("orszag-tang", "orszag", "tang", "ot", "mhd rotor", "mhd-rotor", "rotor", "mhd blast",
                       "mhd-blast", "mhd blast wave", "mhd-blast-wave")


def scheme_of(subgrid):
    """evolvers.py:14-21 -> ('weno', order) | ('ppm', 0) | ('plm', 0) | ('pcm', 0)."""
    s = subgrid.lower()
    if s.startswith("w"):
        order = 5
        parts = s.split("weno")  # weno.py:159-165
        if len(parts) == 2:
            try:
                order = int(s.replace("-", "").split("weno")[-1])
            except ValueError:
                order = 5
        if order not in (3, 7):
            order = 5
        return "weno", order
    if s in ("ppm", "parabolic", "p"):
        return "ppm", 0
    if s in ("plm", "linear", "l"):
        return "plm", 0
    return "pcm", 0


def solver_of(solver):
    """solvers.py:13-31 -> 'hllc' | 'hlld' | 'llf' | 'lw' (DOTS / ES are out of scope)."""
    s = solver.lower()
    cat = [k for k, v in SOLVER_CATEGORY.items() if s in v]
    if not cat:
        raise ValueError(f"unknown solver {solver!r}")
    if cat[0] == "hll":
        return "hlld" if s.endswith("d") else "hllc"
    if cat[0] == "complete":
        raise NotImplementedError("DOTS / entropy-stable fluxes are outside the hot-path scope (SURVEY.md §2.1)")
    return "lw" if s.endswith("w") else "llf"


def integrator_of(timestep):
    """evolvers.py:79-81,187,204 -> 'euler' | 'rk4' | 'ssprk22' | 'ssprk33' | 'ssprk43' | 'ssprk53' | 'ssprk54' | 'ssprk104'."""
    t = timestep.lower()
    if t.startswith("ssprk"):
        digits = t.replace(",", "").replace("(", "").replace(")", "").replace("ssprk", "")
        register, order = int(digits[:-1]), int(digits[-1])
        if order == 4:
            return "ssprk104" if register == 10 else "ssprk54"
        if order == 3:
            return "ssprk53" if register == 5 else ("ssprk43" if register == 4 else "ssprk33")
        return "ssprk22"
    if t.startswith("r"):
        return "rk4"
    return "euler"


@dataclass
class OracleConfig:
    config: str = "sod"
    cells: int = 128
    dimension: int = 1
    subgrid: str = "ppm"
    solver: str = "lf"
    timestep: str = "ssprk(3,3)"
    cfl: float = 0.5
    gamma: float = 1.4
    boundary: str = "edge"      # numpy pad mode: 'edge' (outflow) or 'wrap' (periodic)
    dx: float = 1.0
    magnetic_2d: bool = False
    ppm_author: str = "mc"      # evolvers.py:17 always passes 'mc'
    ppm_dissipate: bool = False  # ppm.py:13 default
    slope_limiter: str = "minmod"  # plm.py:27 wires minmod only
    low_mach: bool = False      # solvers.py:92 default
    ct_method: str = "ppm"      # mag_field.py:11 default
    eigen: str = "lapack"       # 'lapack' = np.linalg.eigvals as fv.py:158; 'closed' = |v_n| + c_fast
    step_parity: int = 0        # number of permutation reversals so far, mod 2 (astrea.py:85)
    extra: dict = field(default_factory=dict)

    @property
    def scheme(self):
        return scheme_of(self.subgrid)

    @property
    def riemann(self):
        return solver_of(self.solver)

    @property
    def integrator(self):
        return integrator_of(self.timestep)

    @property
    def solver_category(self):
        return [k for k, v in SOLVER_CATEGORY.items() if self.solver.lower() in v][0]

    @property
    def high_order(self):
        """generic.py:250-255: 4th-order conversions for WENO and PPM, pointwise otherwise."""
        s = self.subgrid.lower()
        return s.startswith("w") or s in ("ppm", "parabolic", "p")

    def sweep_order(self):
        """Sweep axes in the iteration order of sim_variables.permutations (astrea.py:85 reverses it every step)."""
        axes = list(range(self.dimension))
        return axes[::-1] if (self.step_parity % 2) else axes
