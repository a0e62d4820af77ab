// Reconstructions and limiters: one cell, one variable at a time.
//   slope limiters            num_methods/limiters.py:10-49
//   PLM                       schemes/plm.py:26-36
//   PPM face interpolant      schemes/ppm.py:38
//   PPM-mc extrapolant limiter num_methods/limiters.py:89-143 (the default author, evolvers.py:17)
//   WENO-3/5/7                schemes/weno.py:22-149
//
// Every function takes an accessor ``A`` with
//     double A.s(int64 k)   value of the primitive cell average at *logical* cell index k (already mapped)
//     int64  A.b(int64 k)   boundary map of a logical index: identity where ghost data are genuine
//                           (periodic / interior slab edges), clamp into the domain for 'edge'
// Boundary handling is "pad the derived array" (SURVEY Q7): every derived quantity (d2c, d3, face value) is
// evaluated at a mapped index from mapped neighbours, never from ghost-cell reconstructions.
#pragma once
#include "common.cuh"

namespace astrea {

// ----------------------------------------------------------------------------------------- slope limiters
template <class A, class G = Exact>
HD double limited_slope(const A& acc, int64_t i, int limiter, G&& g = G()) {
    const double c = acc.s(i);
    const double a = c - acc.s(acc.b(i - 1));
    const double b = acc.s(acc.b(i + 1)) - c;
    if (limiter == LIM_MINMOD) {
        if (a * b > 0.0) return (fabs(a) < fabs(b)) ? a : b;
        return 0.0;
    }
    const double r = sdiv(a, b, g);
    switch (limiter) {
        case LIM_VANLEER: return ddiv(r + fabs(r), 1.0 + fabs(r), g) * b;
        case LIM_OSPRE: return 1.5 * ddiv(r * r + r, r * r + r + 1.0, g) * b;
        case LIM_VANALBADA: return ddiv(r * r + r, r * r + 1.0, g) * b;
        case LIM_KOREN: return g.vmax(0.0, g.vmin(g.vmin(2.0 * r, ddiv(2.0 + r, 3.0, g)), 2.0)) * b;
        default: return g.vmax(0.0, g.vmax(g.vmin(2.0 * r, 1.0), g.vmin(r, 2.0))) * b;   // superbee
    }
}

template <class A, class G = Exact>
HD void cell_faces_plm(const A& acc, int64_t i, int limiter, double& wL, double& wR, G&& g = G()) {
    const double half = 0.5 * limited_slope(acc, i, limiter, g);
    const double c = acc.s(i);
    wL = c - half;
    wR = c + half;
}

// ----------------------------------------------------------------------------------------- PPM
// face value at the right face of (mapped) cell k
template <class A>
HD double ppm_face(const A& acc, int64_t k) {
    return 7.0 / 12.0 * (acc.s(k) + acc.s(acc.b(k + 1))) - 1.0 / 12.0 * (acc.s(acc.b(k - 1)) + acc.s(acc.b(k + 2)));
}
template <class A>
HD double ppm_d2c(const A& acc, int64_t k) {   // central second difference at (mapped) cell k
    return acc.s(acc.b(k - 1)) - 2.0 * acc.s(k) + acc.s(acc.b(k + 1));
}
template <class A>
HD double ppm_d3(const A& acc, int64_t k) {    // limiters.py:96 at (mapped) cell k
    return ppm_d2c(acc, acc.b(k + 1)) - ppm_d2c(acc, k);
}

// McCorquodale & Colella limiter (limiters.py:89-143) from the gathered stencil quantities of one cell:
// cell averages c, m1, p1, m2, p2, the two face values, the central second differences at i-1, i, i+1 and the
// third differences at i-2, i-1, i, i+2 (limiters.py:126-129 really skips i+1).
// The reference's grid-wide ``if cell_extrema.any()`` needs no reduction: when no extremum exists anywhere its
// else-branch produces the same numbers as the if-branch, so the if-branch is always taken here.
template <class G = Exact>
HD void ppm_mc_limit(double c, double m1, double p1, double m2, double p2, double faceL, double faceR, double d2c_m1, double d2c,
                     double d2c_p1, double d3_m2, double d3_m1, double d3, double d3_p2, double& wL, double& wR, G&& g = G()) {
    const double C = 5.0 / 4.0;
    const double dwm = c - faceL, dwp = faceR - c;
    const double d2f = 6.0 * (faceL - 2.0 * c + faceR);
    const bool extremum = (dwm * dwp <= 0.0) || ((c - m2) * (p2 - c) <= 0.0);
    double d2lim = 0.0;
    if (extremum) d2lim = g.vsign(d2c) * g.vmin(g.vmin(fabs(d2f), C * fabs(d2c)), g.vmin(C * fabs(d2c_p1), C * fabs(d2c_m1)));
    const double scale = g.vmax(fabs(c), g.vmax(g.vmax(fabs(m1), fabs(p1)), g.vmax(fabs(m2), fabs(p2))));
    const double rho = (fabs(d2f) > 1e-12 * scale) ? sdiv(d2lim, d2f, g) : 0.0;
    const double d3min = g.vmin(g.vmin(d3_m1, d3), g.vmin(d3_m2, d3_p2));
    const double d3max = g.vmax(g.vmax(d3_m1, d3), g.vmax(d3_m2, d3_p2));
    const bool act = (rho < (1.0 - 1e-12)) || (0.1 * g.vmax(fabs(d3max), fabs(d3min)) <= (d3max - d3min));
    wL = faceL;
    wR = faceR;
    if (act) {
        if (dwm * dwp < 0.0) {
            wL = c - rho * dwm;
            wR = c + rho * dwp;
        }
        if (fabs(dwm) >= 2.0 * fabs(dwp)) wL = c - 2.0 * (1.0 - rho) * dwp - rho * dwm;
        if (fabs(dwp) >= 2.0 * fabs(dwm)) wR = c + 2.0 * (1.0 - rho) * dwm + rho * dwp;
    }
}

// Generic form over an accessor (any boundary map).  ``wF`` returns the face value kept for constrained transport (ppm.py:101).
template <class A, class G = Exact>
HD void cell_faces_ppm_mc(const A& acc, int64_t i, double& wL, double& wR, double& wF, G&& g = G()) {
    const double c = acc.s(i);
    const double m1 = acc.s(acc.b(i - 1)), p1 = acc.s(acc.b(i + 1)), m2 = acc.s(acc.b(i - 2)), p2 = acc.s(acc.b(i + 2));
    const double faceR = 7.0 / 12.0 * (c + p1) - 1.0 / 12.0 * (m1 + p2);
    const double faceL = ppm_face(acc, acc.b(i - 1));
    wF = faceR;
    const double d2c = m1 - 2.0 * c + p1;
    const double d2c_m1 = ppm_d2c(acc, acc.b(i - 1)), d2c_p1 = ppm_d2c(acc, acc.b(i + 1));
    const double d3 = d2c_p1 - d2c;
    const double d3_m1 = ppm_d3(acc, acc.b(i - 1)), d3_m2 = ppm_d3(acc, acc.b(i - 2)), d3_p2 = ppm_d3(acc, acc.b(i + 2));
    ppm_mc_limit(c, m1, p1, m2, p2, faceL, faceR, d2c_m1, d2c, d2c_p1, d3_m2, d3_m1, d3, d3_p2, wL, wR, g);
}

// Marching form: the second differences d2c(i-2 .. i+3) and the two face values of the previous cell are carried
// from one cell to the next (identical expressions, identical bits), so each cell evaluates one new second
// difference and one new face value instead of six and two.  ``acc.s(k)`` is the stencil value at offset k from the
// cell (k = -3 .. 4, identity boundary map); ``ppm_mc_march_prime`` supplies what the first cell finds carried.
// The carried values live in circular buffers indexed with the compile-time rotation ROT (the march is unrolled by
// the window length, stages2d.cuh): entry k of the logical window is slot (k + ROT) mod 8, so moving on by one
// cell moves no register.
struct PpmWindow {
    double d2[8];      // logical d2c(i-2), .., d2c(i+3) in slots (0 .. 5 + ROT) mod 8
    double face[2];    // face value at the left / right face of cell i in slots (0 / 1 + ROT) mod 2
};
// what the first cell of a march finds "carried": d2c(i-2 .. i+2) and the left face value, in the slots of rotation ROT
template <int ROT, class A>
HD void ppm_mc_march_prime(const A& acc, PpmWindow& win) {
    constexpr int R = ROT % 8;
#pragma unroll
    for (int k = 0; k < 5; ++k) win.d2[(k + R) % 8] = acc.s(k - 3) - 2.0 * acc.s(k - 2) + acc.s(k - 1);
    win.face[R % 2] = 7.0 / 12.0 * (acc.s(-1) + acc.s(0)) - 1.0 / 12.0 * (acc.s(-2) + acc.s(1));
}
template <int ROT, class A, class G = Exact>
HD void cell_faces_ppm_mc_march(const A& acc, PpmWindow& win, double& wL, double& wR, double& wF, G&& g = G()) {
    auto d2c = [&](int k) { return acc.s(k - 1) - 2.0 * acc.s(k) + acc.s(k + 1); };
    auto face = [&](int k) { return 7.0 / 12.0 * (acc.s(k) + acc.s(k + 1)) - 1.0 / 12.0 * (acc.s(k - 1) + acc.s(k + 2)); };
    constexpr int R = ROT % 8;
    double* d2 = win.d2;
    d2[(5 + R) % 8] = d2c(3);
    win.face[(1 + R) % 2] = face(0);
    const double f0 = win.face[R % 2], f1 = win.face[(1 + R) % 2];
    wF = f1;
    const double a0 = d2[R % 8], a1 = d2[(1 + R) % 8], a2 = d2[(2 + R) % 8], a3 = d2[(3 + R) % 8], a4 = d2[(4 + R) % 8],
                 a5 = d2[(5 + R) % 8];
    ppm_mc_limit(acc.s(0), acc.s(-1), acc.s(1), acc.s(-2), acc.s(2), f0, f1, a1, a2, a3,
                 a1 - a0, a2 - a1, a3 - a2, a5 - a4, wL, wR, g);
}

// ----------------------------------------------------------------------------------------- PPM, authors 'c' / 'ph'
// Colella et al. (2011) and Peterson & Hammett (2008) variants: interface limiter (limiters.py:53-78) + extrapolant
// limiter (limiters.py:144-201).  evolvers.py:17 always passes 'mc', so these are reachable only by calling
// ppm.run(author=...) directly; they are selectable through astrea_cfg.ppm_author.
//
// Both limiters start with a grid-wide ``if mask.any()`` over all cells and variables of the sweep that
// changes the result for EVERY cell, so the kernels run in passes: flag pass 1 -> any_a / any_b (interface limiter),
// flag pass 2 -> any_3 (extrapolant limiter, needs the limited faces), then the reconstruction proper.
struct PpmSwitches { bool any_a, any_b, any_3; };

HD double ppm_limit_face(double w_face, double w_m1, double w_c, double w_p1, double w_p2, bool any) {
    if (!any) return w_face;
    const double C = 5.0 / 4.0;
    const double dL = w_m1 - 2.0 * w_c + w_p1;
    const double dC = 3.0 * (w_c - 2.0 * w_face + w_p1);
    const double dR = w_c - 2.0 * w_p1 + w_p2;
    const double sL = npsign(dL), sC = npsign(dC), sR = npsign(dR);
    const bool agree = (sL == sR) && (sC == sR) && (sC == sL);
    const double lim = sC * npmin(fabs(dC), npmin(fabs(C * dL), fabs(C * dR)));
    const double d2 = agree ? lim : 0.0;
    return 0.5 * (w_c + w_p1) - d2 / 6.0;
}
HD bool ppm_face_extremum(double w_face, double w_c, double w_p1) { return (w_face - w_c) * (w_p1 - w_face) < 0.0; }

// limited value of the face right of (mapped) cell k — author 'c' pads this derived array (ppm.py:69-71)
template <class A>
HD double ppm_face_c(const A& acc, int64_t k, bool any) {
    return ppm_limit_face(ppm_face(acc, k), acc.s(acc.b(k - 1)), acc.s(k), acc.s(acc.b(k + 1)), acc.s(acc.b(k + 2)), any);
}

// what == 0: wL / wR / wF of cell i.  what == 1: only the interface-limiter masks -> (pa, pb).  what == 2: only the
// extrapolant-limiter mask -> p3 (needs sw.any_a / any_b).
template <class A>
HD void cell_faces_ppm_cph(const A& acc, int64_t i, bool ph, PpmSwitches sw, int what, double& wL, double& wR, double& wF,
                           bool& pa, bool& pb, bool& p3) {
    const double C = 5.0 / 4.0;
    const double c = acc.s(i);
    const double m1 = acc.s(acc.b(i - 1)), p1 = acc.s(acc.b(i + 1)), m2 = acc.s(acc.b(i - 2)), p2 = acc.s(acc.b(i + 2));
    const double face_unl = 7.0 / 12.0 * (c + p1) - 1.0 / 12.0 * (m1 + p2);        // ppm.py:38
    double faceL, faceR;
    if (ph) {
        const double faceL0 = 7.0 / 12.0 * (m1 + c) - 1.0 / 12.0 * (m2 + p1);        // ppm.py:52
        if (what == 1) { pa = ppm_face_extremum(faceL0, m1, c); pb = ppm_face_extremum(face_unl, c, p1); return; }
        faceL = ppm_limit_face(faceL0, m2, m1, c, p1, sw.any_a);
        faceR = ppm_limit_face(face_unl, m1, c, p1, p2, sw.any_b);
        wF = face_unl;
    } else {
        if (what == 1) { pa = ppm_face_extremum(face_unl, c, p1); pb = false; return; }
        faceR = ppm_limit_face(face_unl, m1, c, p1, p2, sw.any_a);
        faceL = ppm_face_c(acc, acc.b(i - 1), sw.any_a);
        wF = faceR;
    }
    const double dwm = c - faceL, dwp = faceR - c;
    const bool extremum = dwm * dwp <= 0.0;
    bool ext2;
    if (ph) {
        ext2 = (m1 - c) * (c - p1) <= 0.0;
    } else {
        const double dfL = faceL - ppm_face_c(acc, acc.b(i - 2), sw.any_a), dfR = ppm_face_c(acc, acc.b(i + 2), sw.any_a) - faceR;
        const double dsL = c - m1, dsR = p1 - c;
        const double dfm = npmin(fabs(dfL), fabs(dfR)), dsm = npmin(fabs(dsL), fabs(dsR));
        ext2 = ((dfm >= dsm) && (dfL * dfR < 0.0)) || ((dsm >= dfm) && (dsL * dsR < 0.0));
    }
    if (what == 2) { p3 = extremum || ext2; return; }
    if (!sw.any_3) { wL = faceL; wR = faceR; return; }
    const double D2 = 6.0 * (faceL - 2.0 * c + faceR);
    const double D2L = m2 - 2.0 * m1 + c, D2C = m1 - 2.0 * c + p1, D2R = c - 2.0 * p1 + p2;
    const double s0 = npsign(D2), sC = npsign(D2C), sL = npsign(D2L), sR = npsign(D2R);
    const bool agree = (s0 == sC) && (s0 == sL) && (s0 == sR) && (sC == sL) && (sC == sR) && (sL == sR);
    const double curv = s0 * npmin(npmin(fabs(D2), fabs(C * D2C)), npmin(fabs(C * D2L), fabs(C * D2R)));
    double D2lim = 0.0;
    if (extremum && agree) D2lim = curv;
    if (ph) {
        const double phi = sdiv(D2lim, D2);
        wL = c + phi * (faceL - c);
        wR = c + phi * (faceR - c);
        return;
    }
    if (ext2 && agree) D2lim = curv;
    const double phi = sdiv(D2lim, D2);
    double duL = dwm, duR = dwp;
    if (fabs(dwm) > 2.0 * fabs(dwp)) duL = 2.0 * dwp;
    if (fabs(dwp) > 2.0 * fabs(dwm)) duR = 2.0 * dwm;
    wL = c - phi * duL;
    wR = c + phi * duR;
}

// ----------------------------------------------------------------------------------------- WENO
template <class A, class G = Exact>
HD void cell_faces_weno3(const A& acc, int64_t i, double& wL, double& wR, G&& g = G()) {
    const double eps = 1e-6, g0 = 1.0 / 3.0, g1 = 2.0 / 3.0;
    const double c0 = acc.s(i), m1 = acc.s(acc.b(i - 1)), p1 = acc.s(acc.b(i + 1));
    const double b0 = sq(c0 - m1), b1 = sq(p1 - c0);
    const double e0 = sq(b0 + eps), e1 = sq(b1 + eps);
    const double r0 = ddiv(g0, e0, g), r1 = ddiv(g1, e1, g);         // a0(g0), a1(g1)
    const double l0 = ddiv(g1, e0, g), l1 = ddiv(g0, e1, g);         // a0(g1), a1(g0)
    wR = ddiv(r0, r0 + r1, g) * (1.5 * c0 - 0.5 * m1) + ddiv(r1, r0 + r1, g) * (0.5 * c0 + 0.5 * p1);
    wL = ddiv(l1, l0 + l1, g) * (1.5 * c0 - 0.5 * p1) + ddiv(l0, l0 + l1, g) * (0.5 * c0 + 0.5 * m1);
}

template <class A, class G = Exact>
HD void cell_faces_weno5(const A& acc, int64_t i, double& wL, double& wR, G&& g = G()) {
    const double eps = 1e-6, g0 = 1.0 / 10.0, g1 = 3.0 / 5.0, g2 = 3.0 / 10.0;
    const double c0 = acc.s(i), m1 = acc.s(acc.b(i - 1)), p1 = acc.s(acc.b(i + 1)), m2 = acc.s(acc.b(i - 2)), p2 = acc.s(acc.b(i + 2));
    const double b0 = 13.0 / 12.0 * sq(m2 - 2.0 * m1 + c0) + 1.0 / 4.0 * sq(m2 - 4.0 * m1 + 3.0 * c0);
    const double b1 = 13.0 / 12.0 * sq(m1 - 2.0 * c0 + p1) + 1.0 / 4.0 * sq(m1 - p1);
    const double b2 = 13.0 / 12.0 * sq(c0 - 2.0 * p1 + p2) + 1.0 / 4.0 * sq(3.0 * c0 - 4.0 * p1 + p2);
    const double e0 = sq(b0 + eps), e1 = sq(b1 + eps), e2 = sq(b2 + eps);
    const double r0 = ddiv(g0, e0, g), r1 = ddiv(g1, e1, g), r2 = ddiv(g2, e2, g);
    const double sR = r0 + r1 + r2;
    wR = ddiv(r0, sR, g) * (1.0 / 3.0 * m2 - 7.0 / 6.0 * m1 + 11.0 / 6.0 * c0)
       + ddiv(r1, sR, g) * (-1.0 / 6.0 * m1 + 5.0 / 6.0 * c0 + 1.0 / 3.0 * p1)
       + ddiv(r2, sR, g) * (1.0 / 3.0 * c0 + 5.0 / 6.0 * p1 - 1.0 / 6.0 * p2);
    const double l0 = ddiv(g2, e0, g), l1 = ddiv(g1, e1, g), l2 = ddiv(g0, e2, g);
    const double sL = l0 + l1 + l2;
    wL = ddiv(l0, sL, g) * (1.0 / 3.0 * c0 + 5.0 / 6.0 * m1 - 1.0 / 6.0 * m2)
       + ddiv(l1, sL, g) * (-1.0 / 6.0 * p1 + 5.0 / 6.0 * c0 + 1.0 / 3.0 * m1)
       + ddiv(l2, sL, g) * (1.0 / 3.0 * p2 - 7.0 / 6.0 * p1 + 11.0 / 6.0 * c0);
}

template <class A, class G = Exact>
HD void cell_faces_weno7(const A& acc, int64_t i, double& wL, double& wR, G&& g = G()) {
    const double eps = 1e-6, g0 = 1.0 / 35.0, g1 = 12.0 / 35.0, g2 = 18.0 / 35.0, g3 = 4.0 / 35.0;
    const double c0 = acc.s(i), m1 = acc.s(acc.b(i - 1)), p1 = acc.s(acc.b(i + 1)), m2 = acc.s(acc.b(i - 2)), p2 = acc.s(acc.b(i + 2)),
                 m3 = acc.s(acc.b(i - 3)), p3 = acc.s(acc.b(i + 3));
    const double b0 = m3 * (547.0 * m3 - 3882.0 * m2 + 4642.0 * m1 - 1854.0 * c0) + m2 * (7043.0 * m2 - 17246.0 * m1 + 7042.0 * c0)
                    + m1 * (11003.0 * m1 - 9402.0 * c0) + c0 * (2107.0 * c0);
    const double b1 = m2 * (267.0 * m2 - 1642.0 * m1 + 1602.0 * c0 - 494.0 * p1) + m1 * (2843.0 * m1 - 5966.0 * c0 + 1922.0 * p1)
                    + c0 * (3443.0 * c0 - 2522.0 * p1) + p1 * (547.0 * p1);
    const double b2 = m1 * (547.0 * m1 - 2522.0 * c0 + 1922.0 * p1 - 494.0 * p2) + c0 * (3443.0 * c0 - 5966.0 * p1 + 1602.0 * p2)
                    + p1 * (2843.0 * p1 - 1642.0 * p2) + p2 * (267.0 * p2);
    const double b3 = c0 * (2107.0 * c0 - 9402.0 * p1 + 7042.0 * p2 - 1854.0 * p3) + p1 * (11003.0 * p1 - 17246.0 * p2 + 4642.0 * p3)
                    + p2 * (7043.0 * p2 - 3882.0 * p3) + p3 * (547.0 * p3);
    const double e0 = sq(b0 + eps), e1 = sq(b1 + eps), e2 = sq(b2 + eps), e3 = sq(b3 + eps);
    const double r0 = ddiv(g0, e0, g), r1 = ddiv(g1, e1, g), r2 = ddiv(g2, e2, g), r3 = ddiv(g3, e3, g);
    const double sR = r0 + r1 + r2 + r3;
    wR = ddiv(r0, sR, g) * (-1.0 / 4.0 * m3 + 13.0 / 12.0 * m2 - 23.0 / 12.0 * m1 + 25.0 / 12.0 * c0)
       + ddiv(r1, sR, g) * (1.0 / 12.0 * m2 - 5.0 / 12.0 * m1 + 13.0 / 12.0 * c0 + 1.0 / 4.0 * p1)
       + ddiv(r2, sR, g) * (-1.0 / 12.0 * m1 + 7.0 / 12.0 * c0 + 7.0 / 12.0 * p1 - 1.0 / 12.0 * p2)
       + ddiv(r3, sR, g) * (1.0 / 4.0 * c0 + 13.0 / 12.0 * p1 - 5.0 / 12.0 * p2 + 1.0 / 12.0 * p3);
    const double l0 = ddiv(g3, e0, g), l1 = ddiv(g2, e1, g), l2 = ddiv(g1, e2, g), l3 = ddiv(g0, e3, g);
    const double sL = l0 + l1 + l2 + l3;
    wL = ddiv(l0, sL, g) * (1.0 / 4.0 * c0 + 13.0 / 12.0 * m1 - 5.0 / 12.0 * m2 + 1.0 / 12.0 * m3)
       + ddiv(l1, sL, g) * (-1.0 / 12.0 * p1 + 7.0 / 12.0 * c0 + 7.0 / 12.0 * m1 - 1.0 / 12.0 * m2)
       + ddiv(l2, sL, g) * (1.0 / 12.0 * p2 - 5.0 / 12.0 * p1 + 13.0 / 12.0 * c0 + 1.0 / 4.0 * m1)
       + ddiv(l3, sL, g) * (-1.0 / 4.0 * p3 + 13.0 / 12.0 * p2 - 23.0 / 12.0 * p1 + 25.0 / 12.0 * c0);
}

// Dispatch: wL / wR of cell i (mapped index) for one variable; wF = the face state handed to constrained
// transport (pcm.py:35, plm.py:57, ppm.py:101, weno.py:184).
template <int SCHEME, class A, class G = Exact>
HD void cell_faces(const A& acc, int64_t i, int limiter, double& wL, double& wR, double& wF, G&& g = G()) {
    if (SCHEME == SCH_PCM) {
        wL = wR = wF = acc.s(i);
    } else if (SCHEME == SCH_PLM) {
        cell_faces_plm(acc, i, limiter, wL, wR, g);
        wF = wR;
    } else if (SCHEME == SCH_PPM) {
        cell_faces_ppm_mc(acc, i, wL, wR, wF, g);
    } else if (SCHEME == SCH_WENO3) {
        cell_faces_weno3(acc, i, wL, wR, g);
        wF = wR;
    } else if (SCHEME == SCH_WENO5) {
        cell_faces_weno5(acc, i, wL, wR, g);
        wF = wR;
    } else {
        cell_faces_weno7(acc, i, wL, wR, g);
        wF = wR;
    }
}

}  // namespace astrea
