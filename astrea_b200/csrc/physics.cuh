// Point-wise ideal-MHD physics in fp64, operation order as in the reference.
//   cons<->prim            functions/fv.py:49-53, 89-101
//   physical flux          functions/constructor.py:113-125
//   Roe average            functions/constructor.py:167-176   (B is weighted the other way round)
//   max |eigenvalue|       closed form |v_n| + c_fast of the Jacobian of constructor.py:129-163 (fv.py:157-162)
#pragma once
#include "common.cuh"

namespace astrea {

template <class G = Exact>
HD void prim_of_cons(const double* q, double* w, double gamma, G&& g = G()) {
    const double rho = q[0];
    const double vx = sdiv(q[1], rho, g), vy = sdiv(q[2], rho, g), vz = sdiv(q[3], rho, g);
    const double p = (gamma - 1.0) * (q[4] - 0.5 * (rho * norm3sq(vx, vy, vz, g) + norm3sq(q[5], q[6], q[7], g)));
    w[0] = rho; w[1] = vx; w[2] = vy; w[3] = vz; w[4] = p; w[5] = q[5]; w[6] = q[6]; w[7] = q[7];
}

template <class G = Exact>
HD void cons_of_prim(const double* w, double* q, double gamma, G&& g = G()) {
    const double rho = w[0];
    const double e = ddiv(w[4], gamma - 1.0, g) + 0.5 * (rho * norm3sq(w[1], w[2], w[3], g) + norm3sq(w[5], w[6], w[7], g));
    q[0] = rho; q[1] = w[1] * rho; q[2] = w[2] * rho; q[3] = w[3] * rho; q[4] = e; q[5] = w[5]; q[6] = w[6]; q[7] = w[7];
}

// AX = true sweep axis (0: x, 1: y); components keep their physical meaning under the reference's transposes.
template <int AX, class G = Exact>
HD void physical_flux(const double* w, double* f, double gamma, G&& g = G()) {
    constexpr int n = AX % 3, t1 = (AX + 1) % 3, t2 = (AX + 2) % 3;
    const double rho = w[0], p = w[4];
    const double vn = w[1 + n], bn = w[5 + n];
    f[0] = rho * vn;
    f[1 + n] = rho * (vn * vn) + p + 0.5 * norm3sq(w[5], w[6], w[7], g) - bn * bn;
    f[1 + t1] = rho * vn * w[1 + t1] - bn * w[5 + t1];
    f[1 + t2] = rho * vn * w[1 + t2] - bn * w[5 + t2];
    const double vdotb = (w[1] * w[5] + w[2] * w[6]) + w[3] * w[7];
    f[4] = vn * (0.5 * rho * norm3sq(w[1], w[2], w[3], g) + ddiv(gamma * p, gamma - 1.0, g) + norm3sq(w[5], w[6], w[7], g)) - bn * vdotb;
    f[5 + n] = 0.0;
    f[5 + t1] = w[5 + t1] * vn - bn * w[1 + t1];
    f[5 + t2] = w[5 + t2] * vn - bn * w[1 + t2];
}

// make_Roe_average(first = w_plus, second = w_minus)
template <class G = Exact>
HD void roe_state(const double* first, const double* second, double* out, G&& g = G()) {
    const double s2 = dsqrt(second[0], g), s1 = dsqrt(first[0], g);
    const double den = s2 + s1;
    out[0] = s2 * s1;
    out[1] = sdiv(first[1] * s1 + second[1] * s2, den, g);
    out[2] = sdiv(first[2] * s1 + second[2] * s2, den, g);
    out[3] = sdiv(first[3] * s1 + second[3] * s2, den, g);
    out[4] = sdiv(s1 * first[4] + s2 * second[4], den, g);
    out[5] = sdiv(first[5] * s2 + second[5] * s1, den, g);
    out[6] = sdiv(first[6] * s2 + second[6] * s1, den, g);
    out[7] = sdiv(first[7] * s2 + second[7] * s1, den, g);
}

HD void mean_state(const double* a, const double* b, double* out) {   // plm.py:45
#pragma unroll
    for (int v = 0; v < NVAR; ++v) out[v] = 0.5 * (a[v] + b[v]);
}

// ---------------------------------------------------------------------------------------------- hydro specialisation
// A state whose v_z and B are identically zero keeps them zero under every operation of the path (each such term is
// a product with an exact zero), and dropping those terms does not change a bit of the other components: x + 0 and
// x - 0 are exact.  The H = true variants below touch only [rho, v_x|m_x, v_y|m_y, P|E]; the context selects them
// when the uploaded grid has no v_z / B (api.cu), which is the case for every hydrodynamic 2D configuration.
template <bool H> struct VarSet;
template <> struct VarSet<false> { static constexpr int N = 8; static HD constexpr int at(int a) { return a; } };
template <> struct VarSet<true> { static constexpr int N = 4; static HD constexpr int at(int a) { return a < 3 ? a : 4; } };

template <class G = Exact> HD double norm2sq(double a, double b, G&& g = G()) { double n = dsqrt(a * a + b * b, g); return n * n; }

template <bool H, class G = Exact>
HD void prim_of_cons_t(const double* q, double* w, double gamma, G&& g = G()) {
    if (!H) { prim_of_cons(q, w, gamma, g); return; }
    const double rho = q[0];
    const double vx = sdiv(q[1], rho, g), vy = sdiv(q[2], rho, g);
    w[0] = rho; w[1] = vx; w[2] = vy;
    w[4] = (gamma - 1.0) * (q[4] - 0.5 * (rho * norm2sq(vx, vy, g)));
}

template <bool H, class G = Exact>
HD void cons_of_prim_t(const double* w, double* q, double gamma, G&& g = G()) {
    if (!H) { cons_of_prim(w, q, gamma, g); return; }
    const double rho = w[0];
    q[0] = rho; q[1] = w[1] * rho; q[2] = w[2] * rho;
    q[4] = ddiv(w[4], gamma - 1.0, g) + 0.5 * (rho * norm2sq(w[1], w[2], g));
}

template <int AX, bool H, class G = Exact>
HD void physical_flux_t(const double* w, double* f, double gamma, G&& g = G()) {
    if (!H) { physical_flux<AX>(w, f, gamma, g); return; }
    constexpr int n = AX, t = 1 - AX;          // in-plane normal / transverse component
    const double rho = w[0], p = w[4], vn = w[1 + n];
    f[0] = rho * vn;
    f[1 + n] = rho * (vn * vn) + p;
    f[1 + t] = rho * vn * w[1 + t];
    f[4] = vn * (0.5 * rho * norm2sq(w[1], w[2], g) + ddiv(gamma * p, gamma - 1.0, g));
}

template <bool H, class G = Exact>
HD void roe_state_t(const double* first, const double* second, double* out, G&& g = G()) {
    if (!H) { roe_state(first, second, out, g); return; }
    const double s2 = dsqrt(second[0], g), s1 = dsqrt(first[0], g);
    const double den = s2 + s1;
    out[0] = s2 * s1;
    out[1] = sdiv(first[1] * s1 + second[1] * s2, den, g);
    out[2] = sdiv(first[2] * s1 + second[2] * s2, den, g);
    out[4] = sdiv(s1 * first[4] + s2 * second[4], den, g);
}

template <bool H>
HD void mean_state_t(const double* a, const double* b, double* out) {
#pragma unroll
    for (int k = 0; k < VarSet<H>::N; ++k) { const int v = VarSet<H>::at(k); out[v] = 0.5 * (a[v] + b[v]); }
}

// max |lambda| of the primitive Jacobian (constructor.py:129-163) at state w along AX, in closed form.
// The spectrum is {0, v, v +- sqrt(x)} for x in {c_a^2, c_f^2, c_s^2}, with c_a^2 = Bn^2/rho and c_f^2, c_s^2 the
// roots of x^2 - (a^2 + b^2) x + a^2 c_a^2.  For a physical state all x >= 0 and the maximum is |v| + c_f.  The
// reference feeds unphysical reconstructed states (negative pressure) to np.linalg.eigvals as they are; a root
// x < 0 then gives the complex pair v +- i sqrt(-x) of modulus sqrt(v^2 - x), which is what np.abs returns
// (fv.py:157-162), so those branches are kept.
template <class G = Exact> HD double wave_modulus(double vn, double x, G&& g = G()) { return x >= 0.0 ? vn + dsqrt(x, g) : dsqrt(vn * vn + (-x), g); }

template <int AX, class G = Exact>
HD double spectral_radius(const double* w, double gamma, G&& g = G()) {
    const double rho = w[0];
    const double a2 = ddiv(gamma * w[4], rho, g);
    const double b2 = ddiv((w[5] * w[5] + w[6] * w[6]) + w[7] * w[7], rho, g);
    const double bn2 = ddiv(w[5 + AX] * w[5 + AX], rho, g);
    const double s = a2 + b2;
    const double disc = s * s - 4.0 * (a2 * bn2);
    const double vn = fabs(w[1 + AX]);
    if (disc >= 0.0) {
        const double root = dsqrt(disc, g);
        const double cf2 = 0.5 * (s + root), cs2 = 0.5 * (s - root);
        if (cs2 >= 0.0 && bn2 >= 0.0) return vn + dsqrt(cf2, g);
        return npmax(npmax(wave_modulus(vn, cf2, g), wave_modulus(vn, cs2, g)), wave_modulus(vn, bn2, g));
    }
    if (disc < 0.0) {   // complex conjugate roots x = p +- iq (needs rho < 0)
        const double p = 0.5 * s, q = 0.5 * dsqrt(-disc, g);
        const double m = dsqrt(p * p + q * q, g);
        const double al = dsqrt(0.5 * (m + p), g), be = dsqrt(0.5 * (m - p), g);
        return npmax(dsqrt((vn + al) * (vn + al) + be * be, g), wave_modulus(vn, bn2, g));
    }
    return disc;        // NaN
}

// B == 0: a^2 = gamma P / rho is the only non-zero root (c_f^2 = a^2 if a^2 >= 0, else c_s^2 = a^2 < 0)
template <int AX, bool H, class G = Exact>
HD double spectral_radius_t(const double* w, double gamma, G&& g = G()) {
    if (!H) return spectral_radius<AX>(w, gamma, g);
    const double a2 = ddiv(gamma * w[4], w[0], g);
    const double vn = fabs(w[1 + AX]);
    if (a2 >= 0.0) return vn + dsqrt(a2, g);
    if (a2 < 0.0) return npmax(vn, dsqrt(vn * vn + (-a2), g));
    return a2;          // NaN
}

}  // namespace astrea
