"""The reference's string selectors -> enums of the C ABI.

astrea picks its numerics by string matching on fields of the ``sim_variables`` namedtuple
(SURVEY.md §1): ``subgrid`` (num_methods/evolvers.py:14-21, schemes/weno.py:159-165), ``solver`` +
``solver_category`` (num_methods/solvers.py:13-31, static/.db.json), ``timestep``
(evolvers.py:79-81,187,204), ``boundary`` (np.pad mode, static/tests.py) and ``magnetic_2d``
(functions/generic.py:243).  The same rules are applied here so that a run configured for the reference
selects the same algorithm on the device.
"""
from . import _native as N

_LAX = ("lf", "friedrich", "lax-friedrich", "llf", "local lax-friedrich", "lw", "lax-wendroff", "wendroff")
_HLL = ("hllc", "c", "hlld", "d")
_COMPLETE = ("os", "osher", "solomon", "osher-solomon", "osher solomon", "es", "entropy", "entropy-stable")
MAGNETIC_2D = ("orszag-tang", "orszag", "tang", "ot", "mhd rotor", "mhd-rotor", "rotor", "mhd blast", "mhd-blast",
               "mhd blast wave", "mhd-blast-wave")


def scheme_enum(subgrid):
    s = subgrid.lower()
    if s.startswith("w"):
        order = 5
        if len(s.split("weno")) == 2:
            try:
                order = int(s.replace("-", "").split("weno")[-1])
            except ValueError:
                order = 5
        return {3: N.WENO3, 7: N.WENO7}.get(order, N.WENO5)
    if s in ("ppm", "parabolic", "p"):
        return N.PPM
    if s in ("plm", "linear", "l"):
        return N.PLM
    return N.PCM


def solver_category(solver):
    s = solver.lower()
    if s in _HLL:
        return "hll"
    if s in _COMPLETE:
        return "complete"
    if s in _LAX:
        return "lax"
    raise ValueError(f"unknown solver {solver!r}")


def solver_enum(solver, category=None):
    s = solver.lower()
    category = category or solver_category(s)
    if category == "hll":
        return N.HLLD if s.endswith("d") else N.HLLC
    if category == "complete":
        raise NotImplementedError("the DOTS / entropy-stable fluxes (solvers.py:236-391) are outside the device path")
    return N.LW if s.endswith("w") else N.LLF


def integrator_enum(timestep):
    t = timestep.lower()
    if t.startswith("ssprk"):
        digits = t.replace(",", "").replace("(", "").replace(")", "").replace("ssprk", "")
        register, order = int(digits[:-1]), int(digits[-1])
        if order == 4:
            return N.SSPRK104 if register == 10 else N.SSPRK54
        if order == 3:
            return N.SSPRK53 if register == 5 else (N.SSPRK43 if register == 4 else N.SSPRK33)
        return N.SSPRK22
    if t.startswith("r"):
        return N.RK4
    return N.EULER


def ppm_author_enum(author):
    """ppm.py:43,61-66 / limiters.py:89,148: 'mc' (McCorquodale & Colella, what evolvers.py:17 passes), 'c' (Colella
    et al. 2011), 'ph' (Peterson & Hammett 2008)."""
    a = author.lower()
    if "x" in a or "ph" in a or a in ("peterson", "hammett"):
        return N.PPM_PH
    if a == "mc" or "mccorquodale" in a:
        return N.PPM_MC
    return N.PPM_COLELLA


def limiter_enum(name):
    return {"minmod": N.MINMOD, "vanleer": N.VANLEER, "van leer": N.VANLEER, "ospre": N.OSPRE, "vanalbada": N.VANALBADA,
            "van albada": N.VANALBADA, "koren": N.KOREN, "superbee": N.SUPERBEE}[name.lower()]


def boundary_enum(mode):
    return {"edge": N.EDGE, "wrap": N.WRAP}[mode]


def stages_of(integrator):
    """Spatial-operator evaluations per step (1 in core_run + the ones inside evolve_time)."""
    return {N.EULER: 1, N.RK4: 4, N.SSPRK22: 2, N.SSPRK33: 3, N.SSPRK43: 4, N.SSPRK53: 5, N.SSPRK54: 5, N.SSPRK104: 11}[integrator]


def make_cfg(*, dimension, cells=None, nx=None, ny=None, boundary, gamma, dx, cfl, subgrid, solver, timestep,
             solver_category_name=None, magnetic_2d=False, limiter="minmod", low_mach=False, device=0,
             nx_global=None, x_offset=0, threads_2d=0, segment_2d=0, tile_1d=0, general_path=False, ppm_author="mc",
             step_graph=True, recon_bulk=True, flux_block_tile=None, stage_speeds=False):
    cfg = N.Cfg()
    cfg.dimension = int(dimension)
    cfg.boundary = boundary_enum(boundary)
    cfg.nx = int(nx if nx is not None else cells)
    cfg.ny = 1 if dimension == 1 else int(ny if ny is not None else cells)
    cfg.gamma, cfg.dx, cfg.cfl = float(gamma), float(dx), float(cfl)
    cfg.scheme = scheme_enum(subgrid)
    cfg.ppm_author = ppm_author_enum(ppm_author)
    cfg.limiter = limiter_enum(limiter)
    cfg.solver = solver_enum(solver, solver_category_name)
    cfg.low_mach = int(bool(low_mach))
    cfg.integrator = integrator_enum(timestep)
    cfg.magnetic_2d = int(bool(magnetic_2d))
    cfg.device = int(device)
    cfg.nx_global = int(nx_global if nx_global is not None else cfg.nx)
    cfg.x_offset = int(x_offset)
    cfg.threads_2d, cfg.segment_2d, cfg.tile_1d = int(threads_2d), int(segment_2d), int(tile_1d)
    # bit 0: keep the 8-variable kernels for a grid without v_z / B (testing); bit 1: no CUDA-graph replay of small steps;
    # bit 2: reconstruction march with register prefetch instead of the TMA engine's bulk copies (A/B measurements)
    # bit 3 / 4: flux stage with warp-wide / block-wide rows of transverse points (None: by grid width)
    # bit 6: interface wave speeds evaluated in every operator of a step, not only where they are used (A/B measurements)
    cfg.flags = ((1 if general_path else 0) | (0 if step_graph else 2) | (0 if recon_bulk else 4)
                 | (0 if flux_block_tile is None else (16 if flux_block_tile else 8)) | (64 if stage_speeds else 0))
    return cfg


def cfg_from_sim_variables(sv, device=0, **overrides):
    """Build the device configuration from the reference's ``sim_variables`` namedtuple (astrea.py:132-133)."""
    kw = dict(dimension=sv.dimension, cells=sv.cells, boundary=sv.boundary, gamma=sv.gamma, dx=sv.dx, cfl=sv.cfl,
              subgrid=sv.subgrid, solver=sv.solver, timestep=sv.timestep,
              solver_category_name=getattr(sv, "solver_category", None), magnetic_2d=getattr(sv, "magnetic_2d", False),
              device=device)
    kw.update(overrides)
    return make_cfg(**kw)
