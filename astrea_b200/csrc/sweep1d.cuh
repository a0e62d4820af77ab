// 1D spatial operator: one fused kernel per RK stage.
//
// A block owns ``tile`` consecutive cells and stages the conservative averages (with halo) in shared memory;
// cons->prim, reconstruction + limiter, interface states, eigenvalue estimate, Riemann flux and the flux
// difference all happen on that tile, so every cell's state is read from HBM once and one divergence value
// per variable is written back.  The block-wide maximum wave speed feeds the CFL reduction through a
// warp-shuffle + atomicMax (runtime.cuh).
//
// Reference path: schemes/{pcm,plm,ppm,weno}.py::run -> num_methods/solvers.py:10-65 -> evolvers.py:41-49,
// for dimension == 1 (permutations = {0: (0, 1)}).
#pragma once
#include "physics.cuh"
#include "recon.cuh"
#include "riemann.cuh"
#include "runtime.cuh"

namespace astrea {

struct Sweep1DParams {
    Plane q;          // conservative cell averages, ghost filled
    Plane d;          // out: (F[i+1] - F[i]) / dx per cell
    int64_t n;        // cells of the (global) domain
    double gamma, dx;
    int bc, limiter, low_mach, tile;
    unsigned long long* eigmax_bits;   // [1]
    unsigned long long* flag;                         // non-finite wave speed seen
    // PPM authors 'c' / 'ph' (recon.cuh): pass 1 / 2 only evaluate the grid-wide switches into ppm_flags[0..2]
    int ppm_author, pass;
    int* ppm_flags;
    // Lax-Wendroff (see FluxStage in stages2d.cuh): lw_pass == 1 searches the first non-zero entry of the columns
    // u - c, u, u + c of the padded spectrum into lw_keys[0..2]
    int lw_pass;
    unsigned long long* lw_keys;
};

struct TileAccessor {
    const double* row;   // shared-memory row of one variable
    int64_t base;        // logical index of row[0]
    int64_t n;
    int bc;
    HD double s(int64_t k) const { return row[k - base]; }
    HD int64_t b(int64_t k) const { return bc == BC_WRAP ? k : clamp_index(k, 0, n - 1); }
};

template <int SCHEME, int SOLVER>
struct Sweep1D {
    using Params = Sweep1DParams;
    static constexpr int MAX_THREADS = 256;
    static constexpr bool HO = scheme_high_order(SCHEME);
    static constexpr int LO = recon_lo(SCHEME), HI = recon_hi(SCHEME);
    static constexpr int HL = LO + 2;          // halo cells in front of the tile
    static constexpr int HH = HI + 3;          // halo cells behind it
    static int threads_for(int tile) { return tile + HL + HH; }
    static size_t smem_bytes(int nthreads) { return sizeof(double) * (size_t)nthreads * (NVAR * 5 + 1); }

    struct Tls { double lam; bool bad; unsigned long long key[3]; };

    template <class Ex>
    static HD void block(const Params& p, int bx, int, Ex& ex) {
        const int NT = ex.nthreads();
        double* Q = ex.smem();              // [NVAR][NT] conservative averages
        double* W = Q + NVAR * NT;          // primitive averages
        double* WL = W + NVAR * NT;         // left-face state of each cell
        double* WR = WL + NVAR * NT;        // right-face state of each cell
        double* F = WR + NVAR * NT;         // interface flux
        double* LAM = F + NVAR * NT;        // wave-speed estimate per interface (per cell for PCM)
        const int64_t c0 = (int64_t)bx * p.tile;
        const int64_t base = c0 - HL;       // logical cell index of thread 0
        const double gamma = p.gamma;
        typename Ex::template Local<Tls> tls(ex);

        ex.phase([&](int k) {
            tls[k].lam = 0.0;
            tls[k].bad = false;
            tls[k].key[0] = tls[k].key[1] = tls[k].key[2] = ~0ull;
            const int64_t c = clamp_index(base + k, -GHOST, p.n + GHOST - 1);
#pragma unroll
            for (int v = 0; v < NVAR; ++v) Q[v * NT + k] = *p.q.at(0, v, c);
        });
        ex.phase([&](int k) {
            double q[NVAR], w[NVAR];
            if (!HO) {
#pragma unroll
                for (int v = 0; v < NVAR; ++v) q[v] = Q[v * NT + k];
                prim_of_cons(q, w, gamma);
            } else {
                if (k < 1 || k >= NT - 1) return;
                // fv.py:126-143 with one spatial axis
                double qm[NVAR], qp[NVAR], wm[NVAR], wc[NVAR], wp[NVAR], qa[NVAR];
#pragma unroll
                for (int v = 0; v < NVAR; ++v) { qm[v] = Q[v * NT + k - 1]; q[v] = Q[v * NT + k]; qp[v] = Q[v * NT + k + 1]; }
                prim_of_cons(qm, wm, gamma);
                prim_of_cons(q, wc, gamma);
                prim_of_cons(qp, wp, gamma);
#pragma unroll
                for (int v = 0; v < NVAR; ++v) qa[v] = q[v] - (1.0 / 24.0) * ((qp[v] - q[v]) - (q[v] - qm[v]));
                prim_of_cons(qa, w, gamma);
#pragma unroll
                for (int v = 0; v < NVAR; ++v) w[v] = w[v] + (1.0 / 24.0) * ((wp[v] - wc[v]) - (wc[v] - wm[v]));
            }
#pragma unroll
            for (int v = 0; v < NVAR; ++v) W[v * NT + k] = w[v];
        });
        const int kw0 = HO ? 1 : 0, kw1 = HO ? NT - 1 : NT;      // threads holding a valid W
        ex.phase([&](int k) {
            if (k - LO < kw0 || k + HI >= kw1) return;
            const int64_t c = base + k;
            if (p.bc == BC_EDGE && (c < 0 || c >= p.n)) return;    // never read: interfaces use mapped cells
#pragma unroll
            for (int v = 0; v < NVAR; ++v) {
                TileAccessor acc{W + v * NT, base, p.n, p.bc};
                double wl, wr, wf;
                if (SCHEME == SCH_PPM && p.ppm_author != PPM_MC) {
                    const PpmSwitches sw{p.ppm_flags[0] != 0, p.ppm_flags[1] != 0, p.ppm_flags[2] != 0};
                    bool pa = false, pb = false, p3 = false;
                    cell_faces_ppm_cph(acc, c, p.ppm_author == PPM_PH, sw, p.pass, wl, wr, wf, pa, pb, p3);
                    if (p.pass != 0) {      // the switches look at the cells of the domain only
                        if (c >= 0 && c < p.n) {
                            if (p.pass == 1) { if (pa) p.ppm_flags[0] = 1; if (pb) p.ppm_flags[1] = 1; }
                            else if (p3) p.ppm_flags[2] = 1;
                        }
                        continue;
                    }
                } else {
                    cell_faces<SCHEME>(acc, c, p.limiter, wl, wr, wf);
                }
                WL[v * NT + k] = wl;
                WR[v * NT + k] = wr;
            }
        });
        if (p.pass != 0) return;
        const int kr0 = kw0 + LO, kr1 = kw1 - HI;                 // threads holding valid WL / WR
        auto bmap = [&](int64_t c) { return p.bc == BC_WRAP ? c : clamp_index(c, 0, p.n - 1); };
        // wave-speed estimate (fv.py:157-169) at the averaged interface state, or at the cells for PCM
        ex.phase([&](int k) {
            const int64_t j = base + k;
            double a[NVAR];
            if (SCHEME == SCH_PCM) {
#pragma unroll
                for (int v = 0; v < NVAR; ++v) a[v] = W[v * NT + k];
            } else {
                if (k - 1 < kr0 || k >= kr1) return;
                if (p.bc == BC_EDGE && (j < 1 || j > p.n)) return;
                double wp[NVAR], wm[NVAR];
                const int kp = (int)(bmap(j) - base), km = (int)(bmap(j - 1) - base);
#pragma unroll
                for (int v = 0; v < NVAR; ++v) { wp[v] = WL[v * NT + kp]; wm[v] = WR[v * NT + km]; }
                if (SCHEME == SCH_PLM) mean_state(wp, wm, a); else roe_state(wp, wm, a);
            }
            const double lam = spectral_radius<0>(a, gamma);
            LAM[k] = lam;
            const bool counts = (SCHEME == SCH_PCM) ? (j >= 0 && j < p.n) : (j >= 1 && j <= p.n);
            if (counts && j >= c0 && j <= c0 + p.tile) {
                if (lam == lam && lam <= 1.7976931348623157e308) tls[k].lam = lam; else tls[k].bad = true;
            }
            if (SOLVER == SOL_LW) {
                // solvers.py:84-87: second^2 / max|lambda| with second = np.unique(characteristics, axis=-1)[..., 1]
                const double u = a[1], c = sqrt(gamma * a[4] / a[0]);
                if (p.lw_pass == 1) {
                    const int64_t first = SCHEME == SCH_PCM ? 0 : 1, last = SCHEME == SCH_PCM ? p.n - 1 : p.n;
                    if (!(counts && j >= c0 && j <= c0 + p.tile)) return;
                    const bool wrap = p.bc == BC_WRAP;
                    const int64_t rows[3] = {SCHEME == SCH_PCM ? j + 1 : j, j == (wrap ? last : first) ? 0 : -1,
                                             j == (wrap ? first : last) ? p.n + 1 : -1};
                    const double col[3] = {u - c, u, u + c};
                    for (int q = 0; q < 3; ++q) {
                        if (!(col[q] != 0.0)) continue;
                        for (int m = 0; m < 3; ++m) {
                            if (rows[m] < 0) continue;
                            const unsigned long long kk = ((unsigned long long)rows[m] << 1) | (col[q] > 0.0 ? 1ull : 0ull);
                            if (kk < tls[k].key[q]) tls[k].key[q] = kk;
                        }
                    }
                    return;
                }
                int rank = 0;
                for (int q = 0; q < 3; ++q) rank += (p.lw_keys[q] != ~0ull && (p.lw_keys[q] & 1ull) == 0) ? 1 : 0;
                const double second = rank == 0 ? u - c : (rank == 1 ? 0.0 : u);
                LAM[k] = sdiv(second * second, lam);
                // negative averaged pressure: the reference continues in complex arithmetic, the device reports it
                if (counts && j >= c0 && j <= c0 + p.tile && !(LAM[k] == LAM[k])) tls[k].bad = true;
            }
        });
        if (SOLVER == SOL_LW && p.lw_pass == 1) {
            ex.publish_min3([&](int k, unsigned long long* key) { for (int q = 0; q < 3; ++q) key[q] = tls[k].key[q]; }, p.lw_keys);
            return;
        }
        // Riemann flux at interface j = base + k
        ex.phase([&](int k) {
            const int64_t j = base + k;
            if (j < c0 || j > c0 + p.tile || j > p.n) return;
            double wp[NVAR], wm[NVAR], qp[NVAR], qm[NVAR], fp[NVAR], fm[NVAR], out[NVAR];
            const int kp = (int)(bmap(j) - base), km = (int)(bmap(j - 1) - base);
            if (SCHEME == SCH_PCM) {   // pcm.py:33-36: faces are the padded cell arrays themselves
#pragma unroll
                for (int v = 0; v < NVAR; ++v) { wp[v] = W[v * NT + kp]; wm[v] = W[v * NT + km]; qp[v] = Q[v * NT + kp]; qm[v] = Q[v * NT + km]; }
            } else {
#pragma unroll
                for (int v = 0; v < NVAR; ++v) { wp[v] = WL[v * NT + kp]; wm[v] = WR[v * NT + km]; }
                cons_of_prim(wp, qp, gamma);
                cons_of_prim(wm, qm, gamma);
            }
            physical_flux<0>(wp, fp, gamma);
            physical_flux<0>(wm, fm, gamma);
            if (SOLVER == SOL_HLLC) {
                hllc_flux<0>(gamma, p.low_mach != 0, wp, wm, qp, qm, fp, fm, out);
            } else if (SOLVER == SOL_HLLD) {
                hlld_flux<0>(gamma, W[(5 + 0) * NT + kp], wp, wm, qp, qm, fp, fm, out);
            } else {
                // solvers.py:73-74 on the pad-1 array of averaged states: entries j and j+1
                int ka, kb;
                if (SCHEME == SCH_PCM) { ka = km; kb = kp; }
                else if (p.bc == BC_WRAP) { ka = k; kb = k + 1; }
                else { ka = (int)(clamp_index(j, 1, p.n) - base); kb = (int)(clamp_index(j + 1, 1, p.n) - base); }
                llf_flux(npmax(LAM[ka], LAM[kb]), qp, qm, fp, fm, out);
            }
#pragma unroll
            for (int v = 0; v < NVAR; ++v) F[v * NT + k] = out[v];
        });
        ex.phase([&](int k) {
            const int64_t c = base + k;
            if (c < c0 || c >= c0 + p.tile || c >= p.n) return;
#pragma unroll
            for (int v = 0; v < NVAR; ++v) *p.d.at(0, v, c) = (F[v * NT + k + 1] - F[v * NT + k]) / p.dx;
        });
        ex.publish_max([&](int k, double& val, bool& bad) { val = tls[k].lam; bad = tls[k].bad; }, p.eigmax_bits, p.flag);
    }
};

}  // namespace astrea
