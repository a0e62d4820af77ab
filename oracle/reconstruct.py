"""Reconstructions and limiters of the oracle (test infrastructure; see oracle/__init__.py).

All functions work in the *sweep frame*: axis 0 of every array is the sweep direction, cells
are ``i = 0..N-1`` and interfaces ``j = 0..N`` with interface ``j`` between cells ``j-1`` and
``j``.  ``bc`` is the numpy pad mode of the run (``'wrap'`` | ``'edge'``); boundary handling is
"pad the derived array", exactly as the reference does.
"""
import numpy as np

from .gridops import shifted, extended, safe_div


# --------------------------------------------------------------------------- slope limiters
def slope_differences(w, bc):
    """Backward / forward differences used by every slope limiter (limiters.py:11,24,30,36,42,48)."""
    return w - shifted(w, -1, bc), shifted(w, 1, bc) - w


def limited_slope(w, bc, name="minmod"):
    """limiters.py:10-49.  Only minmod is wired into plm.py:27; the others are selectable here."""
    a, b = slope_differences(w, bc)
    if name == "minmod":
        out = np.zeros_like(b)
        same = a * b > 0
        pick_a = (np.abs(a) < np.abs(b)) & same
        pick_b = (np.abs(a) >= np.abs(b)) & same
        out[pick_a] = a[pick_a]
        out[pick_b] = b[pick_b]
        return out
    r = safe_div(a, b)
    if name == "vanleer":
        return (r + np.abs(r)) / (1 + np.abs(r)) * b
    if name == "ospre":
        return 1.5 * ((r ** 2 + r) / (r ** 2 + r + 1)) * b
    if name == "vanalbada":
        return (r ** 2 + r) / (r ** 2 + 1) * b
    if name == "koren":
        return np.maximum(np.zeros_like(r), np.minimum(np.minimum(2 * r, (2 + r) / 3), np.full_like(r, 2))) * b
    if name == "superbee":
        return np.maximum(np.zeros_like(r), np.maximum(np.minimum(2 * r, np.ones_like(r)),
                                                       np.minimum(r, np.full_like(r, 2)))) * b
    raise ValueError(name)


# --------------------------------------------------------------------------- PCM / PLM
def cell_states_pcm(wS, bc):
    """pcm.py:28-36: both face states of a cell are the cell average."""
    return np.copy(wS), np.copy(wS)


def cell_states_plm(wS, bc, limiter="minmod"):
    """plm.py:26-36: wL/R = wS -/+ half the limited slope."""
    half = .5 * limited_slope(wS, bc, limiter)
    return wS - half, wS + half


# --------------------------------------------------------------------------- PPM
def ppm_face_value(wS, bc):
    """ppm.py:38 — 4th-order interpolant at face i+1/2."""
    return 7 / 12 * (wS + shifted(wS, 1, bc)) - 1 / 12 * (shifted(wS, -1, bc) + shifted(wS, 2, bc))


def ppm_face_limiter(w_face, w_m1, w_c, w_p1, w_p2):
    """limiters.py:53-78 (authors 'c' / 'ph' only).  The ``.any()`` is a genuine grid-wide switch."""
    C = 5 / 4
    if not ((w_face - w_c) * (w_p1 - w_face) < 0).any():
        return w_face
    dL = w_m1 - 2 * w_c + w_p1
    dC = 3 * (w_c - 2 * w_face + w_p1)
    dR = w_c - 2 * w_p1 + w_p2
    agree = (np.sign(dL) == np.sign(dR)) & (np.sign(dC) == np.sign(dR)) & (np.sign(dC) == np.sign(dL))
    lim = np.sign(dC) * np.minimum(np.abs(dC), np.minimum(np.abs(C * dL), np.abs(C * dR)))
    d2 = np.zeros_like(w_face)
    d2[agree] = lim[agree]
    return .5 * (w_c + w_p1) - d2 / 6


def ppm_limit_mc(wS, faceL, faceR, bc):
    """limiters.py:89-143 — McCorquodale & Colella extrapolant limiter (the default author).

    ``faceL[i]`` / ``faceR[i]`` are the interpolated values at the left / right face of cell i.
    The reference's global ``if cell_extrema.any()`` is reproduced; SURVEY Q6b shows that its
    else-branch equals the if-branch whenever no extremum exists.
    """
    C = 5 / 4
    m1, p1, m2, p2 = shifted(wS, -1, bc), shifted(wS, 1, bc), shifted(wS, -2, bc), shifted(wS, 2, bc)
    dwm, dwp = wS - faceL, faceR - wS
    wL, wR = np.copy(faceL), np.copy(faceR)
    d2f = 6 * (faceL - 2 * wS + faceR)
    d2c = m1 - 2 * wS + p1
    d2c_m1, d2c_p1 = shifted(d2c, -1, bc), shifted(d2c, 1, bc)       # pad of the derived array (Q7)
    d3 = d2c_p1 - d2c                                                   # limiters.py:96
    extremum = (dwm * dwp <= 0) | ((wS - m2) * (p2 - wS) <= 0)

    big_m, big_p = np.abs(dwm) >= 2 * np.abs(dwp), np.abs(dwp) >= 2 * np.abs(dwm)
    if not extremum.any():
        wL[big_m] = (wS - 2 * dwp)[big_m]
        wR[big_p] = (wS + 2 * dwm)[big_p]
        return wL, wR

    curv = np.sign(d2c) * np.minimum(np.minimum(np.abs(d2f), C * np.abs(d2c)),
                                     np.minimum(C * np.abs(d2c_p1), C * np.abs(d2c_m1)))
    d2lim = np.zeros_like(wS)
    d2lim[extremum] = curv[extremum]                                    # Q6: non_monotonic is never applied
    scale = np.maximum(np.abs(wS), np.maximum(np.maximum(np.abs(m1), np.abs(p1)), np.maximum(np.abs(m2), np.abs(p2))))
    sensitive = np.abs(d2f) > 1e-12 * scale
    ratio = safe_div(d2lim, d2f)
    rho = np.zeros_like(wS)
    rho[sensitive] = ratio[sensitive]

    d3_m1, d3_m2, d3_p2 = shifted(d3, -1, bc), shifted(d3, -2, bc), shifted(d3, 2, bc)   # stencil {i-2,i-1,i,i+2} (Q6)
    d3min = np.minimum(np.minimum(d3_m1, d3), np.minimum(d3_m2, d3_p2))
    d3max = np.maximum(np.maximum(d3_m1, d3), np.maximum(d3_m2, d3_p2))
    act = (rho < (1 - 1e-12)) | (.1 * np.maximum(np.abs(d3max), np.abs(d3min)) <= (d3max - d3min))

    m = (dwm * dwp < 0) & act
    wL[m] = (wS - rho * dwm)[m]
    wR[m] = (wS + rho * dwp)[m]
    m = act & big_m
    wL[m] = (wS - 2 * (1 - rho) * dwp - rho * dwm)[m]
    m = act & big_p
    wR[m] = (wS + 2 * (1 - rho) * dwm + rho * dwp)[m]
    return wL, wR


def ppm_limit_colella(wS, faceL, faceR, face_pad2, bc, author):
    """limiters.py:144-201 — Colella et al. 2011 ('c') and Peterson & Hammett ('ph') extrapolant limiters.

    ``face_pad2`` is the pad-2 array of face values the 'c' branch differences (zeros for 'ph', ppm.py:63).
    """
    C = 5 / 4
    ph = ("x" in author) or ("ph" in author) or author in ("peterson", "hammett")
    m1, p1, m2, p2 = shifted(wS, -1, bc), shifted(wS, 1, bc), shifted(wS, -2, bc), shifted(wS, 2, bc)
    dwm, dwp = wS - faceL, faceR - wS
    extremum = dwm * dwp <= 0
    if ph:
        ext2 = (m1 - wS) * (wS - p1) <= 0
    else:
        overshoot = (np.abs(dwm) > 2 * np.abs(dwp)) | (np.abs(dwp) > 2 * np.abs(dwm))
        dfL, dfR = faceL - np.copy(face_pad2[:-4]), np.copy(face_pad2[4:]) - faceR
        dsL, dsR = wS - m1, p1 - wS
        dfm = np.minimum(np.abs(dfL), np.abs(dfR))
        dsm = np.minimum(np.abs(dsL), np.abs(dsR))
        ext2 = ((dfm >= dsm) & (dfL * dfR < 0)) | ((dsm >= dfm) & (dsL * dsR < 0))
    if not (extremum.any() or ext2.any()):
        return faceL, faceR
    D2 = 6 * (faceL - 2 * wS + faceR)
    D2L = m2 - 2 * m1 + wS
    D2C = m1 - 2 * wS + p1
    D2R = wS - 2 * p1 + p2
    sg = np.sign
    agree = ((sg(D2) == sg(D2C)) & (sg(D2) == sg(D2L)) & (sg(D2) == sg(D2R)) & (sg(D2C) == sg(D2L))
             & (sg(D2C) == sg(D2R)) & (sg(D2L) == sg(D2R)))
    curv = sg(D2) * np.minimum(np.minimum(np.abs(D2), np.abs(C * D2C)), np.minimum(np.abs(C * D2L), np.abs(C * D2R)))
    D2lim = np.zeros_like(wS)
    D2lim[extremum & agree] = curv[extremum & agree]
    if ph:
        phi = safe_div(D2lim, D2)
        return wS + phi * (faceL - wS), wS + phi * (faceR - wS)
    D2lim[ext2 & agree] = curv[ext2 & agree]
    phi = safe_div(D2lim, D2)
    duL, duR = np.copy(dwm), np.copy(dwp)
    if overshoot.any():
        mL, mR = np.abs(dwm) > 2 * np.abs(dwp), np.abs(dwp) > 2 * np.abs(dwm)
        duL[mL] = 2 * dwp[mL]
        duR[mR] = 2 * dwm[mR]
    return wS - phi * duL, wS + phi * duR


def ppm_flattener(wS, axis, bc, knobs=(.33, .75, .85)):
    """ppm.py:111-134 — Colella (1990) slope flattener coefficient (off by default, ppm.py:13)."""
    delta, z0, z1 = knobs
    P = wS[..., 4]
    Pm1, Pp1, Pm2, Pp2 = shifted(P, -1, bc), shifted(P, 1, bc), shifted(P, -2, bc), shifted(P, 2, bc)
    vn = wS[..., axis + 1]

    def zeta(z):
        out = np.copy(1 - safe_div(z - z0, z1 - z0))
        out[z > z1] = 0
        out[z < z0] = 1
        return out

    chi_bar = zeta(safe_div(np.abs(Pp1 - Pm1), np.abs(Pp2 - Pm2)))
    chi_bar[((shifted(vn, -1, bc) - shifted(vn, 1, bc)) <= 0)
            & (safe_div(np.abs(Pp1 - Pm1), np.minimum(Pp1, Pm1)) <= delta)] = 0
    sign = np.sign(Pp1 - Pm1)
    chi = np.copy(chi_bar)
    up, dn = np.minimum(chi_bar, shifted(chi_bar, 1, bc)), np.minimum(chi_bar, shifted(chi_bar, -1, bc))
    chi[sign < 0] = up[sign < 0]
    chi[sign > 0] = dn[sign > 0]
    return np.ones_like(wS) * chi[..., None]


def ppm_artificial_viscosity(wS, axis, cfg, knobs=(.3, .3)):
    """ppm.py:138-170 — McCorquodale & Colella artificial-viscosity coefficient ``mu``.

    Computed by the reference when ``dissipate`` is on, but never consumed by any solver
    (SURVEY §8a a14), so it cannot influence a run; kept for function-level parity.
    """
    alpha, beta = knobs
    bc, dx, gamma = cfg.boundary, cfg.dx, cfg.gamma
    w = extended(wS, 1, 1, bc)
    vel, vel_w = wS[..., axis + 1], w[..., axis + 1]
    lam = vel_w[2:] - vel_w[1:-1]
    if vel.ndim != 1:
        for ax in range(1, cfg.dimension):
            pv = extended(vel, 1, 1, bc, axis=ax)
            pw = extended(vel_w, 1, 1, bc, axis=ax)
            lam += .25 * (np.diff(pw.take(range(1, pw.shape[ax]), axis=ax), axis=ax)
                          + np.diff(pv.take(range(1, pv.shape[ax]), axis=ax), axis=ax))
    cs = np.sqrt(safe_div(gamma * w[..., 4], w[..., 0]))
    cmin = np.minimum(cs[1:-1], cs[2:])
    ref = np.copy(lam)
    nu = np.minimum(1, safe_div((dx * lam) ** 2, beta * cmin ** 2)) * lam[..., None]
    nu[ref >= 0] = 0
    return alpha * ((nu * np.ones_like(wS)) * np.diff(w[1:], axis=0))


def ppm_artificial_viscosity_cellwise(wS, axis, cfg, knobs=(.3, .3)):
    """ppm.py:138-170 in 1D, read cell by cell: nu_i = min(1, (dx*lambda_i)^2 / (beta*c_min,i^2)) * lambda_i.

    PARITY UNPINNED: the reference's own line (ppm.py:164) multiplies an (N,) by an (N, 1) array and then fails to
    broadcast against the (N, 8) state for every N != 8 (tests/golden/f_ppm_flattener.json records the errors), so
    there is no reference output.  This is the formula of that line with both factors on the same index
    [McCorquodale & Colella 2011, eq. 36-38]; the device kernel (DissipationKernel) is checked against it.
    """
    alpha, beta = knobs
    bc, dx, gamma = cfg.boundary, cfg.dx, cfg.gamma
    w = extended(wS, 1, 1, bc)
    vel_w = w[..., axis + 1]
    lam = vel_w[2:] - vel_w[1:-1]
    cs = np.sqrt(safe_div(gamma * w[..., 4], w[..., 0]))
    cmin = np.minimum(cs[1:-1], cs[2:])
    nu = np.minimum(1, safe_div((dx * lam) ** 2, beta * cmin ** 2)) * lam
    nu[lam >= 0] = 0
    return alpha * (nu[..., None] * np.diff(w[1:], axis=0))


def cell_states_ppm(wS, axis, cfg):
    """ppm.py:28-79 -> (wL, wR, wF) with wF the (possibly limited / flattened) face-i+1/2 value kept for CT."""
    bc, author = cfg.boundary, cfg.ppm_author.lower()
    wF = ppm_face_value(wS, bc)
    m1, p1, m2, p2 = shifted(wS, -1, bc), shifted(wS, 1, bc), shifted(wS, -2, bc), shifted(wS, 2, bc)
    is_ph = ("x" in author) or ("ph" in author) or author in ("peterson", "hammett")
    is_mc = author == "mc" or "mccorquodale" in author
    if is_ph:
        faceL = 7 / 12 * (m1 + wS) - 1 / 12 * (m2 + p1)
        faceL = ppm_face_limiter(faceL, m2, m1, wS, p1)
        faceR = ppm_face_limiter(wF, m1, wS, p1, p2)
        pad2 = np.zeros_like(extended(wF, 2, 2, bc))
    else:
        if author in ("c", "colella"):
            wF = ppm_face_limiter(wF, m1, wS, p1, p2)
        if is_mc and cfg.ppm_dissipate:
            # ppm.py:66-67 multiplies an (..,8) face array by eta[...,None] of shape (..,8,1): numpy refuses
            # to broadcast that for any N != 8, so the reference cannot run with dissipate=True.
            raise ValueError("ppm dissipate=True: the reference raises a broadcast error at ppm.py:67")
        pad2 = extended(wF, 2, 2, bc)
        faceL, faceR = np.copy(pad2[1:-3]), np.copy(pad2[2:-2])
    if is_mc:
        wL, wR = ppm_limit_mc(wS, faceL, faceR, bc)
    else:
        wL, wR = ppm_limit_colella(wS, faceL, faceR, pad2, bc, author)
    return wL, wR, wF


# --------------------------------------------------------------------------- WENO
def cell_states_weno(wS, bc, order=5):
    """weno.py:22-149 — WENO-3/5/7 face states of every cell (eps = 1e-6, weights d/(beta+eps)^2)."""
    eps = 1e-6
    c0 = wS
    m1, p1 = shifted(wS, -1, bc), shifted(wS, 1, bc)
    if order == 3:
        g0, g1 = 1 / 3, 2 / 3
        b0, b1 = (c0 - m1) ** 2, (p1 - c0) ** 2
        a0 = lambda d: d / (b0 + eps) ** 2  # noqa: E731
        a1 = lambda d: d / (b1 + eps) ** 2  # noqa: E731
        wR = (a0(g0) / (a0(g0) + a1(g1))) * (1.5 * c0 - .5 * m1) + (a1(g1) / (a0(g0) + a1(g1))) * (.5 * c0 + .5 * p1)
        wL = (a1(g0) / (a0(g1) + a1(g0))) * (1.5 * c0 - .5 * p1) + (a0(g1) / (a0(g1) + a1(g0))) * (.5 * c0 + .5 * m1)
        return wL, wR
    m2, p2 = shifted(wS, -2, bc), shifted(wS, 2, bc)
    if order == 7:
        m3, p3 = shifted(wS, -3, bc), shifted(wS, 3, bc)
        g0, g1, g2, g3 = 1 / 35, 12 / 35, 18 / 35, 4 / 35
        b0 = (m3 * (547 * m3 - 3882 * m2 + 4642 * m1 - 1854 * c0) + m2 * (7043 * m2 - 17246 * m1 + 7042 * c0)
              + m1 * (11003 * m1 - 9402 * c0) + c0 * (2107 * c0))
        b1 = (m2 * (267 * m2 - 1642 * m1 + 1602 * c0 - 494 * p1) + m1 * (2843 * m1 - 5966 * c0 + 1922 * p1)
              + c0 * (3443 * c0 - 2522 * p1) + p1 * (547 * p1))
        b2 = (m1 * (547 * m1 - 2522 * c0 + 1922 * p1 - 494 * p2) + c0 * (3443 * c0 - 5966 * p1 + 1602 * p2)
              + p1 * (2843 * p1 - 1642 * p2) + p2 * (267 * p2))
        b3 = (c0 * (2107 * c0 - 9402 * p1 + 7042 * p2 - 1854 * p3) + p1 * (11003 * p1 - 17246 * p2 + 4642 * p3)
              + p2 * (7043 * p2 - 3882 * p3) + p3 * (547 * p3))
        a0 = lambda d: d / (b0 + eps) ** 2  # noqa: E731
        a1 = lambda d: d / (b1 + eps) ** 2  # noqa: E731
        a2 = lambda d: d / (b2 + eps) ** 2  # noqa: E731
        a3 = lambda d: d / (b3 + eps) ** 2  # noqa: E731
        sR = a0(g0) + a1(g1) + a2(g2) + a3(g3)
        wR = ((a0(g0) / sR) * (-1 / 4 * m3 + 13 / 12 * m2 - 23 / 12 * m1 + 25 / 12 * c0)
              + (a1(g1) / sR) * (1 / 12 * m2 - 5 / 12 * m1 + 13 / 12 * c0 + 1 / 4 * p1)
              + (a2(g2) / sR) * (-1 / 12 * m1 + 7 / 12 * c0 + 7 / 12 * p1 - 1 / 12 * p2)
              + (a3(g3) / sR) * (1 / 4 * c0 + 13 / 12 * p1 - 5 / 12 * p2 + 1 / 12 * p3))
        sL = a0(g3) + a1(g2) + a2(g1) + a3(g0)
        wL = ((a0(g3) / sL) * (1 / 4 * c0 + 13 / 12 * m1 - 5 / 12 * m2 + 1 / 12 * m3)
              + (a1(g2) / sL) * (-1 / 12 * p1 + 7 / 12 * c0 + 7 / 12 * m1 - 1 / 12 * m2)
              + (a2(g1) / sL) * (1 / 12 * p2 - 5 / 12 * p1 + 13 / 12 * c0 + 1 / 4 * m1)
              + (a3(g0) / sL) * (-1 / 4 * p3 + 13 / 12 * p2 - 23 / 12 * p1 + 25 / 12 * c0))
        return wL, wR
    g0, g1, g2 = 1 / 10, 3 / 5, 3 / 10
    b0 = 13 / 12 * (m2 - 2 * m1 + c0) ** 2 + 1 / 4 * (m2 - 4 * m1 + 3 * c0) ** 2
    b1 = 13 / 12 * (m1 - 2 * c0 + p1) ** 2 + 1 / 4 * (m1 - p1) ** 2
    b2 = 13 / 12 * (c0 - 2 * p1 + p2) ** 2 + 1 / 4 * (3 * c0 - 4 * p1 + p2) ** 2
    a0 = lambda d: d / (b0 + eps) ** 2  # noqa: E731
    a1 = lambda d: d / (b1 + eps) ** 2  # noqa: E731
    a2 = lambda d: d / (b2 + eps) ** 2  # noqa: E731
    sR = a0(g0) + a1(g1) + a2(g2)
    wR = ((a0(g0) / sR) * (1 / 3 * m2 - 7 / 6 * m1 + 11 / 6 * c0)
          + (a1(g1) / sR) * (-1 / 6 * m1 + 5 / 6 * c0 + 1 / 3 * p1)
          + (a2(g2) / sR) * (1 / 3 * c0 + 5 / 6 * p1 - 1 / 6 * p2))
    sL = a0(g2) + a1(g1) + a2(g0)
    wL = ((a0(g2) / sL) * (1 / 3 * c0 + 5 / 6 * m1 - 1 / 6 * m2)
          + (a1(g1) / sL) * (-1 / 6 * p1 + 5 / 6 * c0 + 1 / 3 * m1)
          + (a2(g0) / sL) * (1 / 3 * p2 - 7 / 6 * p1 + 11 / 6 * c0))
    return wL, wR
