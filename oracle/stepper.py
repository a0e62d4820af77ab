"""Spatial operator and Runge-Kutta update of the oracle (test infrastructure; see oracle/__init__.py).

``space_operator`` restates evolvers.evolve_space (num_methods/evolvers.py:12-34) together with the
scheme drivers (schemes/pcm.py:10-39, plm.py:12-63, ppm.py:13-107, weno.py:151-191) and the Riemann
dispatcher (num_methods/solvers.py:10-65); ``time_update`` restates evolvers.evolve_time
(evolvers.py:38-206); ``advance`` is the body of the time loop (astrea.py:67-85).
"""
import numpy as np

from . import ct
from .gridops import (extended, second_difference, ONE_24TH, prim_avg_of_cons_avg, cons_avg_of_prim_avg,
                      centred_of_avg, physical_flux, roe_state, spectral_radius)
from .reconstruct import cell_states_pcm, cell_states_plm, cell_states_ppm, cell_states_weno
from .riemann import llf_flux, lw_flux, lw_flux_closed, hllc_flux, hlld_flux


def _frame(a, ax):
    """Sweep frame of a global (x[,y],8) array: the sweep direction becomes axis 0 (permutations, generic.py:260-264)."""
    return a if ax == 0 else a.transpose(1, 0, 2)


def sweep_data(qbar, ax, cfg):
    """One iteration of the ``for axis, axes in permutations.items()`` loop of a scheme driver."""
    bc, gamma = cfg.boundary, cfg.gamma
    kind, order = cfg.scheme
    qf = _frame(qbar, ax)
    wS = prim_avg_of_cons_avg(qf, cfg, "cell")
    d = {"wS": wS, "sweep_axis": ax}
    if kind == "pcm":
        w_pad = extended(wS, 1, 1, bc)
        q_pad = extended(qf, 1, 1, bc)
        f_pad = physical_flux(w_pad, gamma, ax)
        d.update(wF=wS, wp=w_pad[1:], wm=w_pad[:-1], qp=q_pad[1:], qm=q_pad[:-1], fp=f_pad[1:], fm=f_pad[:-1],
                 avg_pad=w_pad, ct_face=wS)
        return d
    if kind == "plm":
        wL, wR = cell_states_plm(wS, bc, cfg.slope_limiter)
        wF = wR
    elif kind == "ppm":
        wL, wR, wF = cell_states_ppm(wS, ax, cfg)
    else:
        wL, wR = cell_states_weno(wS, bc, order)
        wF = wR
    wp, wm = extended(wL, 0, 1, bc), extended(wR, 1, 0, bc)          # pad(wL)[1:], pad(wR)[:-1]
    if kind == "plm":
        avg = (.5 * (wp + wm))[1:]
    else:
        avg = roe_state(wp, wm)[1:]
    d.update(wF=wF, wp=wp, wm=wm, ct_face=wF,
             qp=cons_avg_of_prim_avg(wp, cfg, "face"), qm=cons_avg_of_prim_avg(wm, cfg, "face"),
             fp=physical_flux(wp, gamma, ax), fm=physical_flux(wm, gamma, ax),
             avg_pad=extended(avg, 1, 1, bc))
    return d


def _solve(cfg, solver_axis, spectrum, local_speed, d):
    kind = cfg.riemann
    if kind == "hllc":
        return hllc_flux(solver_axis, cfg.gamma, d["wp"], d["wm"], d["qp"], d["qm"], d["fp"], d["fm"], cfg.low_mach)
    if kind == "hlld":
        return hlld_flux(solver_axis, cfg.gamma, cfg.boundary, d["wS"], d["wp"], d["wm"], d["qp"], d["qm"], d["fp"], d["fm"])
    if kind == "lw":
        if spectrum is None:      # eigen='closed'
            return lw_flux_closed(d["avg_pad"], cfg.gamma, d["sweep_axis"], d["qp"], d["qm"], d["fp"], d["fm"])
        return lw_flux(spectrum, d["qp"], d["qm"], d["fp"], d["fm"])
    return llf_flux(local_speed, d["qp"], d["qm"], d["fp"], d["fm"])


def space_operator(qbar, cfg):
    """evolve_space: returns {sweep axis: {'flux', 'eigmax'[, 'face_avg', 'emf']}} keyed in iteration order."""
    data = {ax: sweep_data(qbar, ax, cfg) for ax in cfg.sweep_order()}
    if cfg.magnetic_2d:
        for ax in data:
            data[ax]["wT"] = ct.corner_states(data[ax]["ct_face"], cfg)
    out = {}
    for solver_axis, (ax, d) in enumerate(data.items()):
        spectrum, local_speed = spectral_radius(d["avg_pad"], cfg, ax)
        eigmax = np.max(np.maximum(local_speed[:-1], local_speed[1:]))
        if not np.isfinite(eigmax) and cfg.eigen == "closed":
            raise np.linalg.LinAlgError("Array must not contain infs or NaNs")   # what fv.py:158 raises (Q13)
        f_avg = _solve(cfg, solver_axis % 3, spectrum, local_speed, d)
        if cfg.dimension == 2:
            c = dict(d)
            for key in ("wp", "wm", "qp", "qm", "fp", "fm"):
                c[key] = centred_of_avg(d[key], cfg, "face")
            f_cen = _solve(cfg, solver_axis % 3, spectrum, local_speed, c)
            flux = f_cen - ONE_24TH * second_difference(f_avg, cfg.boundary, 1)     # fv.py:147-153
        else:
            flux = f_avg
        out[ax] = {"flux": flux, "eigmax": eigmax}
    if cfg.magnetic_2d:
        emf = ct.corner_emf({ax: data[ax]["wT"] for ax in data}, cfg)
        for ax in out:
            out[ax]["face_avg"] = data[ax]["wF"]
            out[ax]["emf"] = emf
    return out


def rate_of_change(fluxes, cfg):
    """compute_L (evolvers.py:41-60): -(sum_ax dF_ax/dx) with the CT overwrite of the in-plane B rates."""
    total = 0
    for ax, entry in fluxes.items():
        diff = np.diff(entry["flux"], axis=0) / cfg.dx
        total = total + _frame(diff, ax)
    if cfg.magnetic_2d:
        emf = next(iter(fluxes.values()))["emf"]
        d_dy, d_dx = ct.induction_rates(emf, cfg)
        for ax in fluxes:
            total[..., 5 + ax] = d_dy if ax == 0 else d_dx
    return -total


def time_update(grid, fluxes, dt, cfg):
    """evolve_time (evolvers.py:38-206).  Mutates the B slots of ``grid`` in place when magnetic_2d (Q14)."""
    def L(fl):
        return rate_of_change(fl, cfg)

    def refine(g):
        return ct.cell_field_from_faces(g, cfg) if cfg.magnetic_2d else g

    def space(g):
        return space_operator(g, cfg)

    L0 = L(fluxes)
    if cfg.magnetic_2d:
        for ax, entry in fluxes.items():
            grid[..., 5 + ax] = _frame(entry["face_avg"], ax)[..., 5 + ax]
    u, scheme = grid, cfg.integrator
    if scheme == "ssprk104":
        k = np.copy(u)
        fl = fluxes
        for _ in range(5):
            k += refine(1 / 6 * dt * L(fl))
            fl = space(k)
        k5 = refine(3 / 5 * u + 6 / 15 * k + 1 / 15 * dt * L(fl))
        fl = space(k5)
        k2 = np.copy(k5)
        for _ in range(4):
            k2 += refine(1 / 6 * dt * L(fl))
            fl = space(k2)
        return refine(-11 / 35 * u + 5 / 7 * k5 + 3 / 5 * k2 + 1 / 10 * dt * L(fl))
    if scheme == "ssprk54":
        k1 = refine(u + .39175222657189 * dt * L0)
        k2 = refine(.444370493651235 * u + .555629506348765 * k1 + .368410593050371 * dt * L(space(k1)))
        k3 = refine(.620101851488403 * u + .379898148511597 * k2 + .251891774271694 * dt * L(space(k2)))
        L3 = L(space(k3))
        k4 = refine(.178079954393132 * u + .821920045606868 * k3 + .544974750228521 * dt * L3)
        return refine(.517231671970585 * k2 + .096059710526147 * k3 + .06369246866629 * dt * L3
                      + .386708617503269 * k4 + .226007483236906 * dt * L(space(k4)))
    if scheme == "ssprk53":
        k1 = refine(u + .3772689151171 * dt * L0)
        L1 = L(space(k1))
        k2 = refine(k1 + .3772689151171 * dt * L1)
        k3 = refine(.56656131914033 * u + .43343868085967 * k2 + .16352294089771 * dt * L(space(k2)))
        k4 = refine(.09299483444413 * u + .0000209036962 * k1 + .90698426185967 * k3 + .00071997378654 * dt * L0
                    + .34217696850008 * dt * L(space(k3)))
        return refine(.0073613226092 * u + .20127980325145 * k1 + .00182955389682 * k2 + .78952932024253 * k4
                      + (dt * (.0027771981946 * L0 + .00001567934613 * L1 + .29786487010104 * L(space(k4)))))
    if scheme == "ssprk43":
        k1 = refine(u + .5 * dt * L0)
        k2 = refine(k1 + .5 * dt * L(space(k1)))
        k3 = refine(1 / 6 * (4 * u + 2 * k2 + dt * L(space(k2))))
        return refine(k3 + .5 * dt * L(space(k3)))
    if scheme == "ssprk33":
        k1 = refine(u + dt * L0)
        k2 = refine(.25 * (3 * u + k1 + dt * L(space(k1))))
        return refine(1 / 3 * (u + 2 * k2 + 2 * dt * L(space(k2))))
    if scheme == "ssprk22":
        k1 = refine(u + dt * L0)
        return refine(.5 * (u + k1 + dt * L(space(k1))))
    if scheme == "rk4":
        k1 = refine(u + .5 * dt * L0)
        L1 = L(space(k1))
        k2 = refine(u + .5 * dt * L1)
        L2 = L(space(k2))
        k3 = refine(u + dt * L2)
        L3 = L(space(k3))
        return refine(u + 1 / 6 * (dt * (L0 + 2 * L1 + 2 * L2 + L3)))
    return refine(u + dt * L0)


def timestep_from_eigmax(fluxes, cfg):
    """astrea.py:70-71."""
    return cfg.cfl * min(cfg.dx / entry["eigmax"] for entry in fluxes.values())


def advance(grid, cfg, nsteps=1, dts=None, t=0.0, t_end=None):
    """astrea.py:45-85 without I/O: ``nsteps`` full steps.  Returns (grid, dts used); flips cfg.step_parity per step."""
    used = []
    with np.errstate(all="ignore"):
        for n in range(nsteps):
            fluxes = space_operator(grid, cfg)
            dt = timestep_from_eigmax(fluxes, cfg) if dts is None else dts[n]
            if t_end is not None and t + dt >= t_end:     # astrea.py:74-75 with checkpoints == 1
                dt = t_end - t
            grid = time_update(grid, fluxes, dt, cfg)
            t += dt
            cfg.step_parity ^= 1
            used.append(float(dt))
    return grid, used
