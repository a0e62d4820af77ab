// 2D spatial operator for one sweep direction: a fused "marching" kernel.
//
// Frame: rows of the input plane run along the sweep direction, columns along the transverse direction
// (the x-sweep reads the state as stored, the y-sweep reads its transposed copy, exactly like the reference's
// ``grid.transpose(axes)``).  One thread owns one transverse column; a block marches along the sweep
// direction over ``seg`` cells, keeping only a few rows in shared-memory rings:
//
//     q rows (3)  ->  primitive averages wS (ring of LO+HI+1 rows)  ->  reconstruction + limiter of cell i
//     ->  interface states w+/w- of interface i  ->  4th-order face conversion q+/q- (transverse stencil)
//     ->  Riemann flux of the face averages and of the face-centred states  ->  F = F_c - d2_t(F_avg)/24
//     ->  (F[i] - F[i-1]) / dx written for cell i-1.
//
// Each cell's state therefore crosses HBM once on the way in (plus the transverse halo of the block) and
// one divergence value per variable goes out.  Redundant work is limited to the transverse halo columns and
// the few start-up rows of a segment.
//
// Reference path (dimension == 2): schemes/{pcm,plm,ppm,weno}.py::run, functions/fv.py:105-153,
// num_methods/solvers.py:10-65 (incl. the 4th-order flux assembly :44-57), evolvers.py:41-49.
#pragma once
#include "physics.cuh"
#include "recon.cuh"
#include "riemann.cuh"
#include "runtime.cuh"

namespace astrea {

struct Sweep2DParams {
    Plane q;                   // conservative averages in the sweep frame, ghost filled
    Plane d;                   // out: (F[i+1] - F[i]) / dx, sweep frame
    int64_t ns, nt;            // local cells along the sweep (rows) / transverse (columns)
    int64_t ns_glob, s_off;    // global extent and offset of local row 0 along the sweep direction
    int64_t nt_glob, t_off;    // same for the transverse direction
    double gamma, dx;
    int bc, limiter, low_mach;
    int seg;                   // cells per block along the sweep
    int tt;                    // owned columns per block
    unsigned long long* eigmax_bits;
    int* flag;
};

struct RingAccessor {
    const double* ring;   // first element of this variable/column in ring slot 0
    int64_t slot_stride;  // doubles between ring slots
    int depth;
    int64_t lo_glob, hi_glob, off;   // clamp range (global) and local->global offset
    int bc;
    HD double s(int64_t k) const { return ring[(int64_t)((k + 4096 * (int64_t)depth) % depth) * slot_stride]; }
    HD int64_t b(int64_t k) const { return bc == BC_WRAP ? k : clamp_index(k + off, lo_glob, hi_glob) - off; }
};

// AX = physical sweep axis (flux / eigenvalue direction), SAX = the solver's axis argument
template <int SCHEME, int SOLVER, int AX, int SAX>
struct Sweep2D {
    using Params = Sweep2DParams;
    static constexpr int MAX_THREADS = 256;
    static constexpr bool HO = scheme_high_order(SCHEME);
    static constexpr int LO = recon_lo(SCHEME), HI = recon_hi(SCHEME);
    static constexpr int LAG = (SOLVER == SOL_LLF && SCHEME != SCH_PCM) ? 1 : 0;   // LLF needs lambda of interface j+1
    static constexpr int HT = HO ? 3 : 1;                   // transverse halo columns on each side
    static constexpr int NQ = HO ? 3 : 2;                   // q-row ring depth
    static constexpr int NW = LO + HI + 1;                  // wS-row ring depth
    static constexpr int NI = 1 + LAG;                      // interface-row ring depth
    static constexpr int ROWS = NQ + NW + 4 * NI + 1;       // rows of NVAR x NT doubles
    static int owned_for(int nthreads) { return nthreads - 2 * HT; }
    static size_t smem_bytes(int nthreads) { return sizeof(double) * (size_t)nthreads * (NVAR * ROWS + 2 * NI); }

    struct Tls {
        double wr_prev[NVAR], wl_prev[NVAR], f_prev[NVAR], fc[NVAR], q_prev[NVAR];
        double lam, lam_prev;
        bool bad;
    };

    static HD int pm(int64_t r, int depth) { return (int)((r + 4096 * (int64_t)depth) % depth); }

    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        const int NT = ex.nthreads();
        const int64_t ROW = (int64_t)NVAR * NT;
        double* QR = ex.smem();                 // [NQ][NVAR][NT]
        double* WS = QR + NQ * ROW;             // [NW]
        double* IWP = WS + NW * ROW;            // [NI] w+ of an interface
        double* IWM = IWP + NI * ROW;           // [NI] w-
        double* IQP = IWM + NI * ROW;           // [NI] q+
        double* IQM = IQP + NI * ROW;           // [NI] q-
        double* FA = IQM + NI * ROW;            // [1]  flux of the face averages
        double* LAM = FA + ROW;                 // [NI][NT] wave-speed estimate per interface
        double* BN = LAM + NI * NT;             // [NI][NT] normal field of the cell right of the interface (HLLD)
        const double gamma = p.gamma;
        const double c24 = 1.0 / 24.0;
        const int64_t t0 = (int64_t)bx * p.tt;  // first owned column
        const int64_t s0 = (int64_t)by * p.seg; // first owned cell along the sweep
        const int64_t s1 = (s0 + p.seg < p.ns) ? s0 + p.seg : p.ns;   // one past the last owned cell
        typename Ex::template Local<Tls> tls(ex);

        const int kw0 = HO ? 1 : 0, kw1 = HO ? NT - 1 : NT;       // columns with valid wS / w+-
        const int kq0 = HO ? 2 : 0, kq1 = HO ? NT - 2 : NT;       // columns with valid q+- / F_avg
        const int kf0 = kq0 + 1, kf1 = kq1 - 1;                   // columns with a valid final flux (= owned)

        auto col_of = [&](int tl) { return t0 - HT + tl; };
        // transverse neighbour of column tl at offset o, honouring "pad the derived array" for 'edge'
        auto tnb = [&](int tl, int o) -> int {
            if (p.bc == BC_WRAP) return tl + o;
            const int64_t t = clamp_index(col_of(tl) + o + p.t_off, 0, p.nt_glob - 1) - p.t_off;
            return (int)(t - (t0 - HT));
        };
        auto smap = [&](int64_t r) -> int64_t {   // sweep-direction map of a logical (local) row index
            return p.bc == BC_WRAP ? r : clamp_index(r + p.s_off, 0, p.ns_glob - 1) - p.s_off;
        };

        auto load_q_row = [&](int tl, int64_t r) {
            const int64_t rr = clamp_index(r, -GHOST, p.ns + GHOST - 1);
            const int64_t cc = clamp_index(col_of(tl), -GHOST, p.nt + GHOST - 1);
            double* dst = QR + pm(r, NQ) * ROW + tl;
#pragma unroll
            for (int v = 0; v < NVAR; ++v) dst[v * NT] = *p.q.at(rr, v, cc);
        };
        // primitive average of row r (fv.py:126-143 for the 4th-order schemes, fv.py:97-101 otherwise)
        auto make_w_row = [&](int tl, int64_t r) {
            if (tl < kw0 || tl >= kw1) return;
            double q[NVAR], w[NVAR];
            const double* qc = QR + pm(r, NQ) * ROW;
#pragma unroll
            for (int v = 0; v < NVAR; ++v) q[v] = qc[v * NT + tl];
            if (!HO) {
                prim_of_cons(q, w, gamma);
            } else {
                const double* qd = QR + pm(r - 1, NQ) * ROW;
                const double* qu = QR + pm(r + 1, NQ) * ROW;
                double a[NVAR], b[NVAR], wa[NVAR], wb[NVAR], wc[NVAR], qa[NVAR], ws[NVAR];
                prim_of_cons(q, wc, gamma);
                // axis 0 of the sweep frame first, then the transverse axis (fv.py:134-142)
#pragma unroll
                for (int v = 0; v < NVAR; ++v) { a[v] = qd[v * NT + tl]; b[v] = qu[v * NT + tl]; }
                prim_of_cons(a, wa, gamma);
                prim_of_cons(b, wb, gamma);
#pragma unroll
                for (int v = 0; v < NVAR; ++v) {
                    qa[v] = q[v] - c24 * ((b[v] - q[v]) - (q[v] - a[v]));
                    ws[v] = c24 * ((wb[v] - wc[v]) - (wc[v] - wa[v]));
                }
#pragma unroll
                for (int v = 0; v < NVAR; ++v) { a[v] = qc[v * NT + tl - 1]; b[v] = qc[v * NT + tl + 1]; }
                prim_of_cons(a, wa, gamma);
                prim_of_cons(b, wb, gamma);
#pragma unroll
                for (int v = 0; v < NVAR; ++v) {
                    qa[v] = qa[v] - c24 * ((b[v] - q[v]) - (q[v] - a[v]));
                    ws[v] = ws[v] + c24 * ((wb[v] - wc[v]) - (wc[v] - wa[v]));
                }
                prim_of_cons(qa, w, gamma);
#pragma unroll
                for (int v = 0; v < NVAR; ++v) w[v] = w[v] + ws[v];
            }
            double* dst = WS + pm(r, NW) * ROW + tl;
#pragma unroll
            for (int v = 0; v < NVAR; ++v) dst[v * NT] = w[v];
        };

        // ---------------------------------------------------------------- prologue: fill the rings
        const int64_t i_first = s0 - 1;
        ex.phase([&](int tl) {
            Tls& st = tls[tl];
            st.lam = 0.0; st.lam_prev = 0.0; st.bad = false;
#pragma unroll
            for (int v = 0; v < NVAR; ++v) { st.wr_prev[v] = 0.0; st.wl_prev[v] = 0.0; st.f_prev[v] = 0.0; st.fc[v] = 0.0; st.q_prev[v] = 0.0; }
        });
        // wS rows [i_first-LO, i_first+HI-1] need q rows one further out for the 4th-order conversion
        const int64_t wlo = i_first - LO, whi = i_first + HI - 1;
        for (int64_t r = wlo - (HO ? 1 : 0); r <= whi; ++r) {
            if (HO) {
                ex.phase([&](int tl) { load_q_row(tl, r + 1); if (r == wlo - 1) { load_q_row(tl, r); } });
                if (r >= wlo) ex.phase([&](int tl) { make_w_row(tl, r); });
            } else {
                ex.phase([&](int tl) { load_q_row(tl, r); make_w_row(tl, r); });
            }
        }

        // ---------------------------------------------------------------- march
        const int64_t i_last = s1 + LAG;
        for (int64_t i = i_first; i <= i_last; ++i) {
            // A: next q row;  B: next wS row
            if (HO) {
                ex.phase([&](int tl) { load_q_row(tl, i + HI + 1); });
                ex.phase([&](int tl) { make_w_row(tl, i + HI); });
            } else {
                ex.phase([&](int tl) { load_q_row(tl, i + HI); make_w_row(tl, i + HI); });
            }
            // C: reconstruct cell i, assemble interface j = i
            const int64_t ig = i + p.s_off;                           // global cell index
            const bool edge = p.bc == BC_EDGE;
            const bool cell_valid = !edge || (ig >= 0 && ig < p.ns_glob);
            const int js = pm(i, NI);
            ex.phase([&](int tl) {
                if (tl < kw0 || tl >= kw1) return;
                Tls& st = tls[tl];
                double wl[NVAR], wr[NVAR], wp[NVAR], wm[NVAR];
                if (cell_valid) {
#pragma unroll
                    for (int v = 0; v < NVAR; ++v) {
                        RingAccessor acc{WS + v * NT + tl, ROW, NW, 0, p.ns_glob - 1, p.s_off, p.bc};
                        double wf;
                        cell_faces<SCHEME>(acc, i, p.limiter, wl[v], wr[v], wf);
                    }
                }
                // w_plus[j] = wL[b(j)], w_minus[j] = wR[b(j-1)]   (plm.py:42, ppm.py:82, weno.py:171, pcm.py:33)
#pragma unroll
                for (int v = 0; v < NVAR; ++v) {
                    wp[v] = cell_valid ? wl[v] : st.wl_prev[v];                     // edge: j == N uses cell N-1
                    wm[v] = (edge && ig == 0) ? wr[v] : st.wr_prev[v];              // edge: j == 0 uses cell 0
                }
                if (cell_valid) {
#pragma unroll
                    for (int v = 0; v < NVAR; ++v) { st.wl_prev[v] = wl[v]; st.wr_prev[v] = wr[v]; }
                }
                double* dp = IWP + js * ROW + tl;
                double* dm = IWM + js * ROW + tl;
#pragma unroll
                for (int v = 0; v < NVAR; ++v) { dp[v * NT] = wp[v]; dm[v * NT] = wm[v]; }
                if (SOLVER == SOL_HLLD) BN[js * NT + tl] = WS[pm(smap(i), NW) * ROW + (5 + SAX) * NT + tl];
                // wave-speed estimate: fv.py:157-169 on the pad-1 array of averaged interface states
                const int64_t jg = ig;
                double lam;
                bool counts;
                if (SCHEME == SCH_PCM) {
                    // pcm.py:30: Jacobian at the padded cells; interface j sees cells b(j-1), b(j)
                    lam = spectral_radius<AX>(wp, gamma);
                    counts = cell_valid && i >= 0 && i < p.ns;
                    LAM[js * NT + tl] = npmax(lam, (edge && ig == 0) ? lam : st.lam_prev);
                    if (cell_valid) st.lam_prev = lam;
                    // pcm.py:33-34: q faces are the padded conservative averages themselves
                    double* qp = IQP + js * ROW + tl;
                    double* qm = IQM + js * ROW + tl;
                    const double* qrow = QR + pm(smap(i), NQ) * ROW + tl;
#pragma unroll
                    for (int v = 0; v < NVAR; ++v) {
                        const double qc = qrow[v * NT];
                        qp[v * NT] = qc;
                        qm[v * NT] = (edge && ig == 0) ? qc : st.q_prev[v];
                        if (cell_valid) st.q_prev[v] = qc;
                    }
                } else {
                    double a[NVAR];
                    if (SCHEME == SCH_PLM) mean_state(wp, wm, a); else roe_state(wp, wm, a);
                    lam = spectral_radius<AX>(a, gamma);
                    counts = jg >= 1 && jg <= p.ns_glob && i >= 0 && i <= p.ns;
                    LAM[js * NT + tl] = lam;
                }
                if (counts && i >= s0 && i <= s1 && tl >= kf0 && tl < kf1 && col_of(tl) < p.nt) {
                    if (lam == lam && lam <= 1.7976931348623157e308) st.lam = fmax(st.lam, lam); else st.bad = true;
                }
                if (!HO && SCHEME != SCH_PCM) {   // pointwise face conversion needs no neighbours (fv.py:89-93)
                    double qq[NVAR];
                    cons_of_prim(wp, qq, gamma);
                    double* qp = IQP + js * ROW + tl;
#pragma unroll
                    for (int v = 0; v < NVAR; ++v) qp[v * NT] = qq[v];
                    cons_of_prim(wm, qq, gamma);
                    double* qm = IQM + js * ROW + tl;
#pragma unroll
                    for (int v = 0; v < NVAR; ++v) qm[v * NT] = qq[v];
                }
            });
            // D: 4th-order face conversion w+- -> q+- with the transverse Laplacian (fv.py:105-122, 'face')
            if (HO) {
                ex.phase([&](int tl) {
                    if (tl < kq0 || tl >= kq1) return;
                    const int ta = tnb(tl, -1), tb = tnb(tl, 1);
                    for (int side = 0; side < 2; ++side) {
                        const double* wrow = (side == 0 ? IWP : IWM) + js * ROW;
                        double* qrow = (side == 0 ? IQP : IQM) + js * ROW + tl;
                        double wa[NVAR], wc[NVAR], wb[NVAR], qa[NVAR], qc[NVAR], qb[NVAR], wx[NVAR], qx[NVAR];
#pragma unroll
                        for (int v = 0; v < NVAR; ++v) { wa[v] = wrow[v * NT + ta]; wc[v] = wrow[v * NT + tl]; wb[v] = wrow[v * NT + tb]; }
                        cons_of_prim(wa, qa, gamma);
                        cons_of_prim(wc, qc, gamma);
                        cons_of_prim(wb, qb, gamma);
#pragma unroll
                        for (int v = 0; v < NVAR; ++v) wx[v] = wc[v] - c24 * ((wb[v] - wc[v]) - (wc[v] - wa[v]));
                        cons_of_prim(wx, qx, gamma);
#pragma unroll
                        for (int v = 0; v < NVAR; ++v) qrow[v * NT] = qx[v] + c24 * ((qb[v] - qc[v]) - (qc[v] - qa[v]));
                    }
                });
            }
            // E: Riemann fluxes of interface je = i - LAG: face averages -> FA row, face-centred -> registers
            const int64_t je = i - LAG;
            const int es = pm(je, NI);
            const bool intf_active = je >= s0 && je <= s1;
            ex.phase([&](int tl) {
                if (!intf_active || tl < kq0 || tl >= kq1) return;
                Tls& st = tls[tl];
                const double* WP = IWP + es * ROW;
                const double* WM = IWM + es * ROW;
                const double* QP = IQP + es * ROW;
                const double* QM = IQM + es * ROW;
                double lam = 0.0, bn = 0.0;
                if (SOLVER == SOL_LLF) {
                    if (SCHEME == SCH_PCM) lam = LAM[es * NT + tl];
                    else {
                        // entries j and j+1 of the pad-1 array of interface speeds
                        const int64_t jg = je + p.s_off;
                        const int64_t ja = edge ? clamp_index(jg, 1, p.ns_glob) - p.s_off : je;
                        const int64_t jb = edge ? clamp_index(jg + 1, 1, p.ns_glob) - p.s_off : je + 1;
                        lam = npmax(LAM[pm(ja, NI) * NT + tl], LAM[pm(jb, NI) * NT + tl]);
                    }
                }
                if (SOLVER == SOL_HLLD) bn = BN[es * NT + tl];
                auto solve = [&](const double* wp, const double* wm, const double* qp, const double* qm, const double* fp,
                                 const double* fm, double* out) {
                    if (SOLVER == SOL_HLLC) hllc_flux<SAX>(gamma, p.low_mach != 0, wp, wm, qp, qm, fp, fm, out);
                    else if (SOLVER == SOL_HLLD) hlld_flux<SAX>(gamma, bn, wp, wm, qp, qm, fp, fm, out);
                    else llf_flux(lam, qp, qm, fp, fm, out);
                };
                double wp[NVAR], wm[NVAR], qp[NVAR], qm[NVAR], fp[NVAR], fm[NVAR], out[NVAR];
#pragma unroll
                for (int v = 0; v < NVAR; ++v) { wp[v] = WP[v * NT + tl]; wm[v] = WM[v * NT + tl]; qp[v] = QP[v * NT + tl]; qm[v] = QM[v * NT + tl]; }
                physical_flux<AX>(wp, fp, gamma);
                physical_flux<AX>(wm, fm, gamma);
                solve(wp, wm, qp, qm, fp, fm, out);
#pragma unroll
                for (int v = 0; v < NVAR; ++v) FA[v * NT + tl] = out[v];
                if (tl < kf0 || tl >= kf1) return;
                // face-centred states: x - d2_t(x)/24 for w, q and the physical flux (solvers.py:47-52, fv.py:67-85)
                const int ta = tnb(tl, -1), tb = tnb(tl, 1);
                double na[NVAR], nb[NVAR], fa[NVAR], fb[NVAR];
#pragma unroll
                for (int v = 0; v < NVAR; ++v) { na[v] = WP[v * NT + ta]; nb[v] = WP[v * NT + tb]; }
                physical_flux<AX>(na, fa, gamma);
                physical_flux<AX>(nb, fb, gamma);
#pragma unroll
                for (int v = 0; v < NVAR; ++v) {
                    const double w0 = wp[v], f0 = fp[v], q0 = qp[v];
                    wp[v] = w0 - c24 * ((nb[v] - w0) - (w0 - na[v]));
                    fp[v] = f0 - c24 * ((fb[v] - f0) - (f0 - fa[v]));
                    qp[v] = q0 - c24 * ((QP[v * NT + tb] - q0) - (q0 - QP[v * NT + ta]));
                }
#pragma unroll
                for (int v = 0; v < NVAR; ++v) { na[v] = WM[v * NT + ta]; nb[v] = WM[v * NT + tb]; }
                physical_flux<AX>(na, fa, gamma);
                physical_flux<AX>(nb, fb, gamma);
#pragma unroll
                for (int v = 0; v < NVAR; ++v) {
                    const double w0 = wm[v], f0 = fm[v], q0 = qm[v];
                    wm[v] = w0 - c24 * ((nb[v] - w0) - (w0 - na[v]));
                    fm[v] = f0 - c24 * ((fb[v] - f0) - (f0 - fa[v]));
                    qm[v] = q0 - c24 * ((QM[v * NT + tb] - q0) - (q0 - QM[v * NT + ta]));
                }
                solve(wp, wm, qp, qm, fp, fm, st.fc);
            });
            // F: F = F_c - d2_t(F_avg)/24 (fv.py:147-153), then the flux difference of cell je-1 (evolvers.py:48)
            ex.phase([&](int tl) {
                if (!intf_active || tl < kf0 || tl >= kf1) return;
                Tls& st = tls[tl];
                const int ta = tnb(tl, -1), tb = tnb(tl, 1);
                const int64_t col = col_of(tl);
                const bool write = je - 1 >= s0 && col < p.nt;
#pragma unroll
                for (int v = 0; v < NVAR; ++v) {
                    const double a0 = FA[v * NT + tl];
                    const double f = st.fc[v] - c24 * ((FA[v * NT + tb] - a0) - (a0 - FA[v * NT + ta]));
                    if (write) *p.d.at(je - 1, v, col) = (f - st.f_prev[v]) / p.dx;
                    st.f_prev[v] = f;
                }
            });
        }
        ex.publish_max([&](int k, double& val, bool& bad) { val = tls[k].lam; bad = tls[k].bad; }, p.eigmax_bits, p.flag);
    }
};

}  // namespace astrea
