"""Riemann fluxes of the oracle (test infrastructure; see oracle/__init__.py).

Interface arrays have N+1 entries along axis 0; ``plus`` is the state on the right of the
interface (left face of cell j), ``minus`` the state on its left (right face of cell j-1).
``axis`` is the *solver* axis, i.e. the reference's private 0,1 counter (solvers.py:34-36,63),
which differs from the sweep axis on odd steps.
"""
import numpy as np

from .gridops import safe_div, length, extended


def pairwise_max(local):
    """fv.py:165 / solvers.py:73-74: max of consecutive entries of the N+2 padded array -> N+1 interfaces."""
    return np.maximum(local[:-1], local[1:])


def llf_flux(local_speed, qp, qm, fp, fm):
    """solvers.py:69-75 — local Lax-Friedrichs."""
    lam = pairwise_max(local_speed)
    return .5 * (fm + fp) - .5 * ((qp - qm) * lam[..., None])


def lw_flux(spectrum, qp, qm, fp, fm):
    """solvers.py:79-88 — 'Lax-Wendroff' with the reference's np.unique column pick."""
    second = np.unique(spectrum, axis=-1)[..., 1]
    coeff = safe_div(second ** 2, np.max(np.abs(spectrum), axis=-1))
    lam = pairwise_max(coeff)
    return .5 * (fm + fp) - .5 * ((qp - qm) * lam[..., None])


def lw_flux_closed(avg_pad, gamma, axis, qp, qm, fp, fm):
    """The same without LAPACK, for states with v_z = 0 and B = 0 (what the device computes).

    There the spectrum of the primitive Jacobian is {u - c, u, u + c, 0} (the value u fills five slots), LAPACK returns
    it in a fixed slot order, and np.unique(axis=-1) sorts the four distinct columns lexicographically over all
    padded points: u - c < u < u + c at the first point, so only the place of the all-zero column is open — it is
    the number of columns whose first non-zero entry is negative.  Column 1 is then u - c, 0 or u.
    """
    from .gridops import closed_spectral_radius
    if np.any(avg_pad[..., 3] != 0) or np.any(avg_pad[..., 5:8] != 0):
        raise ValueError("closed-form Lax-Wendroff needs v_z = 0 and B = 0 (SURVEY Q11)")
    with np.errstate(all="ignore"):
        u = avg_pad[..., 1 + axis]
        c = np.sqrt(gamma * avg_pad[..., 4] / avg_pad[..., 0])
        rank = 0
        for column in (u - c, u, u + c):
            flat = column.reshape(-1)
            nz = np.nonzero(flat)[0]
            if len(nz) and flat[nz[0]] < 0:
                rank += 1
        second = (u - c) if rank == 0 else (np.zeros_like(u) if rank == 1 else u)
        coeff = safe_div(second * second, closed_spectral_radius(avg_pad, gamma, axis))
    lam = pairwise_max(coeff)
    return .5 * (fm + fp) - .5 * ((qp - qm) * lam[..., None])


def hllc_flux(axis, gamma, wp, wm, qp, qm, fp, fm, low_mach=False):
    """solvers.py:92-138 — HLLC with the reference's star-state and selection quirks."""
    rL, uL, pL = wm[..., 0], wm[..., axis + 1], wm[..., 4]
    rR, uR, pR = wp[..., 0], wp[..., axis + 1], wp[..., 4]
    cL, cR = np.sqrt(gamma * safe_div(pL, rL)), np.sqrt(gamma * safe_div(pR, rR))
    u_roe = safe_div(uL * np.sqrt(rL) + uR * np.sqrt(rR), np.sqrt(rL) + np.sqrt(rR))
    c2_roe = (safe_div(np.sqrt(rL) * cL ** 2 + np.sqrt(rR) * cR ** 2, np.sqrt(rL) + np.sqrt(rR))
              + .5 * ((uR - uL) ** 2) * safe_div(np.sqrt(rL) * np.sqrt(rR), (np.sqrt(rL) + np.sqrt(rR)) ** 2))
    sL = np.minimum(uL - cL, u_roe - np.sqrt(c2_roe))
    sR = np.maximum(uR + cR, u_roe + np.sqrt(c2_roe))
    sM = safe_div(pR - pL + rL * uL * (sL - uL) - rR * uR * (sR - uR), rL * (sL - uL) - rR * (sR - uR))
    if low_mach:
        mach = np.maximum(np.abs(safe_div(uL, cL)), np.abs(safe_div(uR, cR)))
        phi = np.sin(.5 * np.pi * np.minimum(1, mach / .1))
        sL, sR = np.copy(phi * sL), np.copy(phi * sR)
    kL, kR = safe_div(sL - uL, sL - sM), safe_div(sR - uR, sR - sM)
    qLs, qRs = qm * kL[..., None], qp * kR[..., None]
    qLs[..., 1] = rL * kL * sM          # Q2: always slot 1
    qRs[..., 1] = rR * kR * sM
    qLs[..., 4] = qLs[..., 4] + kL * (sM - uL) * (rL * sM + safe_div(pL, sL - uL))
    qRs[..., 4] = qRs[..., 4] + kR * (sM - uR) * (rR * sM + safe_div(pR, sR - uR))
    fLs = fm + (qLs - qm) * sL[..., None]
    fRs = fp + (qRs - qp) * sR[..., None]
    out = np.copy(fp)                   # Q3: default is the plus flux
    m = (sL <= 0) & (0 < sM)
    out[m] = fLs[m]
    m = (sM <= 0) & (0 <= sR)
    out[m] = fRs[m]
    m = sR < 0
    out[m] = fp[m]
    return out


def _fast_speed_hlld(w, gamma):
    """solvers.py:144-153 — fast speed with B[...,0] as 'normal' field whatever the axis."""
    rho, P, B = w[..., 0], w[..., 4], w[..., 5:8]
    a = np.sqrt(safe_div(gamma * P, rho))
    b = safe_div(length(B), np.sqrt(rho))
    bx = safe_div(B[..., 0], np.sqrt(rho))
    return np.sqrt(.5 * (a ** 2 + b ** 2 + np.sqrt(((a ** 2 + b ** 2) ** 2) - (4 * (a ** 2) * (bx ** 2)))))


def hlld_flux(axis, gamma, bc, wS, wp, wm, qp, qm, fp, fm):
    """solvers.py:142-232 — HLLD as the reference writes it."""
    n, t1, t2 = axis % 3, (axis + 1) % 3, (axis + 2) % 3
    cell = extended(wS, 0, 1, bc)                     # pad(wS)[1:]: interface j sees cell bc(j)
    Bn = cell[..., n + 5]
    rL, vL, pL, BL = wm[..., 0], wm[..., 1:4], wm[..., 4], wm[..., 5:8]
    rR, vR, pR, BR = wp[..., 0], wp[..., 1:4], wp[..., 4], wp[..., 5:8]
    uL, uR = vL[..., axis], vR[..., axis]
    cfL, cfR = _fast_speed_hlld(wm, gamma), _fast_speed_hlld(wp, gamma)
    sL = np.minimum(uL, uR) - np.maximum(cfL, cfR)
    sR = np.minimum(uL, uR) + np.maximum(cfL, cfR)
    sM = safe_div(pR - pL + rL * uL * (sL - uL) - rR * uR * (sR - uR) + .5 * (length(BR) ** 2) - .5 * (length(BL) ** 2),
                  rL * (sL - uL) - rR * (sR - uR))
    rLs, rRs = rL * safe_div(sL - uL, sL - sM), rR * safe_div(sR - uR, sR - sM)
    sLs, sRs = sM - safe_div(BL[..., axis], np.sqrt(rLs)), sM - safe_div(BR[..., axis], np.sqrt(rRs))
    p_star = safe_div(rL * (pR + .5 * length(BR) ** 2) * (sL - uL) - rR * (pL + .5 * length(BL) ** 2) * (sR - uR)
                      + rR * rL * (sL - uL) * (sR - uR), rL * (sL - uL) - rR * (sR - uR))

    def tangential(r, v, B, s, u, comp):
        vel = v[..., comp] - Bn * B[..., comp] * safe_div(sM - u, r * (s - u) * (s - sM) - Bn ** 2)
        mag = B[..., comp] * safe_div(r * (s - u) ** 2 - Bn ** 2, r * (s - u) * (s - sM) - Bn ** 2)
        return vel, mag

    v1Ls, B1Ls = tangential(rL, vL, BL, sL, uL, t1)
    v1Rs, B1Rs = tangential(rR, vR, BR, sR, uR, t1)
    v2Ls, B2Ls = tangential(rL, vL, BL, sL, uL, t2)
    v2Rs, B2Rs = tangential(rR, vR, BR, sR, uR, t2)

    def dot3(a, b):
        return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]

    def star(q, r, rs, v, B, p, s, u, v1s, v2s, B1s, B2s):
        out = np.zeros_like(q)
        out[..., 0] = rs
        out[..., n + 1] = r * sM          # Q4: rho, not rho*
        out[..., t1 + 1] = r * v1s
        out[..., t2 + 1] = r * v2s
        out[..., n + 5] = np.copy(q[..., n + 5])
        out[..., t1 + 5] = B1s
        out[..., t2 + 5] = B2s
        out[..., 4] = safe_div(q[..., 4] * (s - u) - u * (p + .5 * length(B) ** 2) + p_star * sM
                               + Bn * (dot3(v, B) - dot3(out[..., 1:4], out[..., 5:8])), s - sM)
        return out

    qLs = star(qm, rL, rLs, vL, BL, pL, sL, uL, v1Ls, v2Ls, B1Ls, B2Ls)
    qRs = star(qp, rR, rRs, vR, BR, pR, sR, uR, v1Rs, v2Rs, B1Rs, B2Rs)
    fLs = np.copy(fm) + (qLs - qm) * sL[..., None]
    fRs = np.copy(fp) + (qRs - qp) * sR[..., None]

    sgn = np.sign(Bn)
    sqL, sqR = np.sqrt(rLs), np.sqrt(rRs)
    v1ss = safe_div(v1Rs * sqR + v1Ls * sqL + sgn * (B1Ls - B1Rs), sqL + sqR)
    v2ss = safe_div(v2Rs * sqR + v2Ls * sqL + sgn * (B2Ls - B2Rs), sqL + sqR)
    B1ss = safe_div(B1Ls * sqR + B1Rs * sqL + sgn * (v1Ls - v1Rs) * np.sqrt(rRs * rLs), sqL + sqR)
    B2ss = safe_div(B2Ls * sqR + B2Rs * sqL + sgn * (v2Ls - v2Rs) * np.sqrt(rRs * rLs), sqL + sqR)

    def double_star(qs, rs):
        out = np.zeros_like(qs)
        out[..., 0] = rs
        out[..., n + 1] = sM              # Q4: velocities, no density factor
        out[..., t1 + 1] = v1ss
        out[..., t2 + 1] = v2ss
        out[..., n + 5] = np.copy(qs[..., n + 5])
        out[..., t1 + 5] = B1ss
        out[..., t2 + 5] = B2ss
        out[..., 4] = np.copy(qs[..., 4] - np.sqrt(rs) * sgn * (dot3(qs[..., 1:4], qs[..., 5:8])
                                                                - dot3(out[..., 1:4], out[..., 5:8])))
        return out

    qLss, qRss = double_star(qLs, rLs), double_star(qRs, rRs)
    fLss = np.copy(fm) + (qLss - qLs) * sLs[..., None]     # Q4: built on the minus flux, not on fLs
    fRss = np.copy(fp) + (qRss - qRs) * sRs[..., None]

    out = np.copy(fm)                     # Q4: default is the minus flux
    m = (sL <= 0) & (0 < sLs)
    out[m] = fLs[m]
    m = (sLs <= 0) & (0 < sM)
    out[m] = fLss[m]
    m = (sM <= 0) & (0 < sRs)
    out[m] = fRss[m]
    m = (sRs <= 0) & (0 <= sR)
    out[m] = fRs[m]
    m = sR < 0
    out[m] = fp[m]
    return out
