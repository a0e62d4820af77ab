// Riemann fluxes at one interface in fp64, operation order and quirks as in num_methods/solvers.py.
//   LLF   solvers.py:69-75      HLLC  solvers.py:92-138      HLLD  solvers.py:142-232
// ``plus`` = state on the right of the interface, ``minus`` = state on its left.  SAX is the *solver* axis, the
// reference's private 0,1 counter (solvers.py:34-36,63), which is not the sweep axis on odd steps.
#pragma once
#include "physics.cuh"

namespace astrea {

HD void llf_flux(double lam, const double* qp, const double* qm, const double* fp, const double* fm, double* out) {
#pragma unroll
    for (int v = 0; v < NVAR; ++v) out[v] = 0.5 * (fm[v] + fp[v]) - 0.5 * ((qp[v] - qm[v]) * lam);
}
template <bool H>
HD void llf_flux_t(double lam, const double* qp, const double* qm, const double* fp, const double* fm, double* out) {
#pragma unroll
    for (int k = 0; k < VarSet<H>::N; ++k) { const int v = VarSet<H>::at(k); out[v] = 0.5 * (fm[v] + fp[v]) - 0.5 * ((qp[v] - qm[v]) * lam); }
}

// HLLC in two steps, so that a caller can decide from the waves which side's conservative state and physical flux it has
// to produce at all (FluxStage, face-centred solve): the wave speeds and the branch of solvers.py:124-137 follow from the
// primitive states alone; the flux then reads one side only.
struct HllcWaves {
    double sL, sR, sM;
    int side;          // 0: the plus flux as it is; 1: star state on the plus side; 2: star state on the minus side
};
template <int SAX, class G = Exact>
HD HllcWaves hllc_waves(double gamma, bool low_mach, const double* wp, const double* wm, G&& g = G()) {
    const double rL = wm[0], uL = wm[1 + SAX], pL = wm[4];
    const double rR = wp[0], uR = wp[1 + SAX], pR = wp[4];
    const double cL = dsqrt(gamma * sdiv(pL, rL, g), g), cR = dsqrt(gamma * sdiv(pR, rR, g), g);
    const double sqL = dsqrt(rL, g), sqR = dsqrt(rR, g);
    const double u_roe = sdiv(uL * sqL + uR * sqR, sqL + sqR, g);
    const double c2_roe = sdiv(sqL * (cL * cL) + sqR * (cR * cR), sqL + sqR, g) + 0.5 * sq(uR - uL) * sdiv(sqL * sqR, sq(sqL + sqR), g);
    HllcWaves wv;
    wv.sL = npmin(uL - cL, u_roe - dsqrt(c2_roe, g));
    wv.sR = npmax(uR + cR, u_roe + dsqrt(c2_roe, g));
    wv.sM = sdiv(pR - pL + rL * uL * (wv.sL - uL) - rR * uR * (wv.sR - uR), rL * (wv.sL - uL) - rR * (wv.sR - uR), g);
    if (low_mach) {   // solvers.py:118-122
        const double mach = npmax(fabs(sdiv(uL, cL, g)), fabs(sdiv(uR, cR, g)));
        const double phi = sin(0.5 * 3.141592653589793 * npmin(1.0, ddiv(mach, 0.1, g)));
        wv.sL = phi * wv.sL;
        wv.sR = phi * wv.sR;
    }
    const bool useL = (wv.sL <= 0.0) && (0.0 < wv.sM);
    const bool useR = (wv.sM <= 0.0) && (0.0 <= wv.sR);
    const bool sup = wv.sR < 0.0;
    // later masks override earlier ones (solvers.py:135-137)
    wv.side = (sup || !(useL || useR)) ? 0 : (useR ? 1 : 2);
    return wv;
}
// w, q, f: primitive state, conservative state and physical flux of the side `wv.side` names (plus for 0 and 1, minus for 2);
// side 0 reads f only
template <int SAX, bool H = false, class G = Exact>
HD void hllc_side(const HllcWaves& wv, const double* w, const double* q, const double* f, double* out, G&& g = G()) {
    constexpr int NV = VarSet<H>::N;
    if (wv.side == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) { const int v = VarSet<H>::at(k); out[v] = f[v]; }
        return;
    }
    const double r = w[0], u = w[1 + SAX], p = w[4];
    const double s = wv.side == 1 ? wv.sR : wv.sL, sM = wv.sM;
    const double kk = sdiv(s - u, s - sM, g);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int v = VarSet<H>::at(k);
        double qs = q[v] * kk;
        if (v == 1) qs = r * kk * sM;
        if (v == 4) qs = qs + kk * (sM - u) * (r * sM + sdiv(p, s - u, g));
        out[v] = f[v] + (qs - q[v]) * s;
    }
}
template <int SAX, bool H = false, class G = Exact>
HD void hllc_flux(double gamma, bool low_mach, const double* wp, const double* wm, const double* qp, const double* qm,
                  const double* fp, const double* fm, double* out, G&& g = G()) {
    const HllcWaves wv = hllc_waves<SAX>(gamma, low_mach, wp, wm, g);
    if (wv.side == 2) hllc_side<SAX, H>(wv, wm, qm, fm, out, g);
    else hllc_side<SAX, H>(wv, wp, qp, fp, out, g);
}

template <class G = Exact>
HD double hlld_fast_speed(const double* w, double gamma, G&& g = G()) {   // solvers.py:144-153: B[...,0] whatever the axis
    const double rho = w[0];
    const double a = dsqrt(sdiv(gamma * w[4], rho, g), g);
    const double sr = dsqrt(rho, g);
    const double b = sdiv(norm3(w[5], w[6], w[7], g), sr, g);
    const double bx = sdiv(w[5], sr, g);
    const double inner = dsqrt(sq(a * a + b * b) - (4.0 * (a * a) * (bx * bx)), g);
    return dsqrt(0.5 * (a * a + b * b + inner), g);
}

// bn_cell = normal field of the padded *cell* average on the right of the interface, wS[bc(j)][5+SAX] (solvers.py:168)
template <int SAX, class G = Exact>
HD void hlld_flux(double gamma, double bn_cell, const double* wp, const double* wm, const double* qp, const double* qm,
                  const double* fp, const double* fm, double* out, G&& g = G()) {
    constexpr int n = SAX % 3, t1 = (SAX + 1) % 3, t2 = (SAX + 2) % 3;
    const double Bn = bn_cell;
    const double rL = wm[0], pL = wm[4], rR = wp[0], pR = wp[4];
    const double uL = wm[1 + SAX], uR = wp[1 + SAX];
    const double cfL = hlld_fast_speed(wm, gamma, g), cfR = hlld_fast_speed(wp, gamma, g);
    const double sL = npmin(uL, uR) - npmax(cfL, cfR);
    const double sR = npmin(uL, uR) + npmax(cfL, cfR);
    const double b2L = norm3sq(wm[5], wm[6], wm[7], g), b2R = norm3sq(wp[5], wp[6], wp[7], g);
    const double den = rL * (sL - uL) - rR * (sR - uR);
    const double sM = sdiv(pR - pL + rL * uL * (sL - uL) - rR * uR * (sR - uR) + 0.5 * b2R - 0.5 * b2L, den, g);
    const double rLs = rL * sdiv(sL - uL, sL - sM, g), rRs = rR * sdiv(sR - uR, sR - sM, g);
    const double sLs = sM - sdiv(wm[5 + SAX], dsqrt(rLs, g), g), sRs = sM - sdiv(wp[5 + SAX], dsqrt(rRs, g), g);
    const double p_star = sdiv(rL * (pR + 0.5 * b2R) * (sL - uL) - rR * (pL + 0.5 * b2L) * (sR - uR) + rR * rL * (sL - uL) * (sR - uR), den, g);

    const bool m1 = (sL <= 0.0) && (0.0 < sLs);
    const bool m2 = (sLs <= 0.0) && (0.0 < sM);
    const bool m3 = (sM <= 0.0) && (0.0 < sRs);
    const bool m4 = (sRs <= 0.0) && (0.0 <= sR);
    const bool m5 = sR < 0.0;
    // last matching mask wins (solvers.py:226-231); default is the minus flux
    int sel = 0;
    if (m1) sel = 1;
    if (m2) sel = 2;
    if (m3) sel = 3;
    if (m4) sel = 4;
    if (m5) sel = 5;
    if (sel == 0) {
#pragma unroll
        for (int v = 0; v < NVAR; ++v) out[v] = fm[v];
        return;
    }
    if (sel == 5) {
#pragma unroll
        for (int v = 0; v < NVAR; ++v) out[v] = fp[v];
        return;
    }
    const double dL = rL * (sL - uL) * (sL - sM) - Bn * Bn, dR = rR * (sR - uR) * (sR - sM) - Bn * Bn;
    const double gL = sdiv(sM - uL, dL, g), gR = sdiv(sM - uR, dR, g);
    const double hL = sdiv(rL * sq(sL - uL) - Bn * Bn, dL, g), hR = sdiv(rR * sq(sR - uR) - Bn * Bn, dR, g);
    const double v1Ls = wm[1 + t1] - Bn * wm[5 + t1] * gL, v1Rs = wp[1 + t1] - Bn * wp[5 + t1] * gR;
    const double v2Ls = wm[1 + t2] - Bn * wm[5 + t2] * gL, v2Rs = wp[1 + t2] - Bn * wp[5 + t2] * gR;
    const double B1Ls = wm[5 + t1] * hL, B1Rs = wp[5 + t1] * hR;
    const double B2Ls = wm[5 + t2] * hL, B2Rs = wp[5 + t2] * hR;

    double qLs[NVAR], qRs[NVAR];
    qLs[0] = rLs;                 qRs[0] = rRs;
    qLs[1 + n] = rL * sM;         qRs[1 + n] = rR * sM;            // Q4: rho, not rho*
    qLs[1 + t1] = rL * v1Ls;      qRs[1 + t1] = rR * v1Rs;
    qLs[1 + t2] = rL * v2Ls;      qRs[1 + t2] = rR * v2Rs;
    qLs[5 + n] = qm[5 + n];       qRs[5 + n] = qp[5 + n];
    qLs[5 + t1] = B1Ls;           qRs[5 + t1] = B1Rs;
    qLs[5 + t2] = B2Ls;           qRs[5 + t2] = B2Rs;
    const double vbL = (wm[1] * wm[5] + wm[2] * wm[6]) + wm[3] * wm[7];
    const double vbR = (wp[1] * wp[5] + wp[2] * wp[6]) + wp[3] * wp[7];
    const double mbLs = (qLs[1] * qLs[5] + qLs[2] * qLs[6]) + qLs[3] * qLs[7];
    const double mbRs = (qRs[1] * qRs[5] + qRs[2] * qRs[6]) + qRs[3] * qRs[7];
    qLs[4] = sdiv(qm[4] * (sL - uL) - uL * (pL + 0.5 * b2L) + p_star * sM + Bn * (vbL - mbLs), sL - sM, g);
    qRs[4] = sdiv(qp[4] * (sR - uR) - uR * (pR + 0.5 * b2R) + p_star * sM + Bn * (vbR - mbRs), sR - sM, g);

    if (sel == 1) {
#pragma unroll
        for (int v = 0; v < NVAR; ++v) out[v] = fm[v] + (qLs[v] - qm[v]) * sL;
        return;
    }
    if (sel == 4) {
#pragma unroll
        for (int v = 0; v < NVAR; ++v) out[v] = fp[v] + (qRs[v] - qp[v]) * sR;
        return;
    }
    const double sgn = npsign(Bn);
    const double sqL = dsqrt(rLs, g), sqR = dsqrt(rRs, g);
    const double sden = sqL + sqR;
    const double v1ss = sdiv(v1Rs * sqR + v1Ls * sqL + sgn * (B1Ls - B1Rs), sden, g);
    const double v2ss = sdiv(v2Rs * sqR + v2Ls * sqL + sgn * (B2Ls - B2Rs), sden, g);
    const double srr = dsqrt(rRs * rLs, g);
    const double B1ss = sdiv(B1Ls * sqR + B1Rs * sqL + sgn * (v1Ls - v1Rs) * srr, sden, g);
    const double B2ss = sdiv(B2Ls * sqR + B2Rs * sqL + sgn * (v2Ls - v2Rs) * srr, sden, g);
    const double* qs = (sel == 2) ? qLs : qRs;
    const double rs = (sel == 2) ? rLs : rRs;
    double qss[NVAR];
    qss[0] = rs;
    qss[1 + n] = sM;                                                // Q4: velocities, no density factor
    qss[1 + t1] = v1ss;
    qss[1 + t2] = v2ss;
    qss[5 + n] = qs[5 + n];
    qss[5 + t1] = B1ss;
    qss[5 + t2] = B2ss;
    const double mbs = (sel == 2) ? mbLs : mbRs;
    const double mbss = (qss[1] * qss[5] + qss[2] * qss[6]) + qss[3] * qss[7];
    qss[4] = qs[4] - dsqrt(rs, g) * sgn * (mbs - mbss);
    if (sel == 2) {
#pragma unroll
        for (int v = 0; v < NVAR; ++v) out[v] = fm[v] + (qss[v] - qLs[v]) * sLs;   // Q4: built on the minus flux
    } else {
#pragma unroll
        for (int v = 0; v < NVAR; ++v) out[v] = fp[v] + (qss[v] - qRs[v]) * sRs;
    }
}

}  // namespace astrea
