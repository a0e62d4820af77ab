"""Constrained-transport pieces of the oracle (test infrastructure; see oracle/__init__.py).

Restates num_methods/mag_field.py: transverse reconstruction of face states to cell corners
(:11-121), the upwinded corner electric field (:125-187) and the face->cell "inverse
reconstruction" of the magnetic field (:191-211).
"""
import numpy as np

from .gridops import shifted, extended, safe_div, length, roe_state, primitive_jacobian, centred_of_avg, avg_of_centred
from .reconstruct import ppm_face_value, ppm_face_limiter, ppm_limit_mc, ppm_limit_colella, cell_states_weno


def corner_states(face_state, cfg, method=None, author="mc"):
    """mag_field.py:11-121.  ``face_state`` is in a sweep frame; the result is in the *transposed* frame
    (axis 0 = the direction transverse to that sweep), as the reference returns it."""
    method = cfg.ct_method if method is None else method
    bc = cfg.boundary
    f = np.copy(face_state.transpose(1, 0, 2))            # ortho_axis = (1,0,2) in 2D (generic.py:264)
    if method == "weno":
        return cell_states_weno(f, bc, 5)                 # :84-119 is WENO-5 verbatim (wD = wL, wU = wR)
    up = ppm_face_value(f, bc)
    m1, p1, m2, p2 = shifted(f, -1, bc), shifted(f, 1, bc), shifted(f, -2, bc), shifted(f, 2, bc)
    is_ph = ("x" in author) or ("ph" in author) or author in ("peterson", "hammett")
    if is_ph:
        down = 7 / 12 * (m1 + f) - 1 / 12 * (m2 + p1)
        faceL = ppm_face_limiter(down, m2, m1, f, p1)
        faceR = ppm_face_limiter(up, m1, f, p1, p2)
        pad2 = np.zeros_like(extended(up, 2, 2, bc))
    else:
        if author in ("c", "collela"):
            up = ppm_face_limiter(up, m1, f, p1, p2)
        pad2 = extended(up, 2, 2, bc)
        faceL, faceR = np.copy(pad2[1:-3]), np.copy(pad2[2:-2])
    if author == "mc" or "mccorquodale" in author:
        return ppm_limit_mc(f, faceL, faceR, bc)
    return ppm_limit_colella(f, faceL, faceR, pad2, bc, author)


def corner_wavespeeds(wD, wU, cfg, axis):
    """mag_field.py:128-161 — (a_plus, a_minus) at the Roe average across each corner."""
    bc = cfg.boundary
    plus, minus = extended(wD, 0, 1, bc), extended(wU, 1, 0, bc)
    avg = roe_state(plus, minus)[1:]
    # mag_field.py:152-159 takes max / -min of np.linalg.eigvals for the Lax-type solvers; the spectrum contains 0 (the
    # B_n row of the Jacobian is zero), so that is max(0, v_n + c_f) and -min(0, v_n - c_f) again: with eigen='closed'
    # the oracle (like the device) evaluates the closed form for every solver.
    if cfg.solver_category == "hll" or cfg.eigen == "closed":
        rho, P, B = avg[..., 0], avg[..., 4], avg[..., 5:8]
        vn, Bn = avg[..., 1 + axis % 3], B[..., axis % 3]
        a = np.sqrt(cfg.gamma * safe_div(P, rho))
        b = safe_div(length(B), np.sqrt(rho))
        bn = safe_div(Bn, np.sqrt(rho))
        cf = np.sqrt(.5 * (a ** 2 + b ** 2 + np.sqrt(((a ** 2 + b ** 2) ** 2) - (4 * (a ** 2) * (bn ** 2)))))
        return np.maximum(np.zeros_like(vn), vn + cf), -np.minimum(np.zeros_like(vn), vn - cf)
    spectrum = np.linalg.eigvals(primitive_jacobian(avg, cfg.gamma, axis % 3))
    return np.max(spectrum, axis=-1), -np.min(spectrum, axis=-1)


def corner_emf(transverse, cfg):
    """mag_field.py:163-187.  ``transverse`` maps sweep axis -> (wD, wU) of that sweep (each in its transposed
    frame).  Roles are assigned by *iteration order* of the swapped permutations."""
    order = cfg.sweep_order()                    # keys of sim_variables.permutations in iteration order
    parts, speeds = [], []
    for key in order:
        other = 1 - key                          # swapped_permutations[key] names the other sweep's data
        wD, wU = transverse[other]
        ap, am = corner_wavespeeds(wD, wU, cfg, key)
        if key == 1:                             # alignment_axes = permutations[key]
            ap, am, wD, wU = ap.T, am.T, wD.transpose(1, 0, 2), wU.transpose(1, 0, 2)
        speeds.append((ap, am))
        parts.append((wD, wU))
    (north, south), (east, west) = parts
    (ap_y, am_y), (ap_x, am_x) = speeds
    NE = .5 * (west[..., 2] + south[..., 2]) * south[..., 5] - .5 * (west[..., 1] + south[..., 1]) * west[..., 6]
    NW = .5 * (east[..., 2] + south[..., 2]) * south[..., 5] - .5 * (east[..., 1] + south[..., 1]) * east[..., 6]
    SE = .5 * (west[..., 2] + north[..., 2]) * north[..., 5] - .5 * (west[..., 1] + north[..., 1]) * west[..., 6]
    SW = .5 * (east[..., 2] + north[..., 2]) * north[..., 5] - .5 * (east[..., 1] + north[..., 1]) * east[..., 6]
    return (safe_div(ap_x * ap_y * SW + am_x * ap_y * SE + ap_x * am_y * NW + am_x * am_y * NE, (ap_x + am_x) * (ap_y + am_y))
            - safe_div(ap_y * am_y, ap_y + am_y) * (north[..., 5] - south[..., 5])
            + safe_div(ap_x * am_x, ap_x + am_x) * (east[..., 6] - west[..., 6]))


def induction_rates(emf, cfg):
    """evolvers.py:52-58 — the entries that overwrite total_flux[...,5] and [...,6] (global (x,y) frame)."""
    bc = cfg.boundary
    # axis 0 reads the transposed field and differences it along its axis 0, i.e. along y of ``emf``
    d_dy = (shifted(emf, 1, bc, axis=1) - emf) / cfg.dx
    d_dx = -1 * (shifted(emf, 1, bc, axis=0) - emf) / cfg.dx
    return d_dy, d_dx


def cell_field_from_faces(grid, cfg):
    """mag_field.py:191-211 — face-averaged B (stored in the B slots) -> cell-averaged B, every sweep axis."""
    bc = cfg.boundary
    out = np.copy(grid)
    for ax in cfg.sweep_order():
        frame = grid if ax == 0 else grid.transpose(1, 0, 2)
        fc = centred_of_avg(frame, cfg, "face")
        cc = -1 / 16 * (shifted(fc, -1, bc) + shifted(fc, 2, bc)) + 9 / 16 * (fc + shifted(fc, 1, bc))
        ca = avg_of_centred(cc, cfg, "cell")
        back = ca if ax == 0 else ca.transpose(1, 0, 2)
        out[..., 5 + ax] = back[..., 5 + ax]
    return out
