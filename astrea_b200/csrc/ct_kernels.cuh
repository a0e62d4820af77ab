// Constrained transport of the in-plane magnetic field (num_methods/mag_field.py, evolvers.py:26-32,52-58,73-76).
//
//   face states (ReconStage ``wf``, written in the other sweep's frame)  --halo fill = "pad the derived array"-->
//   PPM-mc along the transverse direction (ReconStage<PPM>, cell aligned; the bundle that marches in the y frame writes
//   its corner states transposed, so no transpose kernel is left on this path)             mag_field.py:11-82
//   CornerEmfKernel: Roe-averaged HLL wave speeds across each corner + upwinded E_z        mag_field.py:125-187
//   RateKernel (aux_kernels.cuh): dBx/dt = -dE_z/dy, dBy/dt = +dE_z/dx                      evolvers.py:52-58
//   FaceFieldKernel: B slots of the grid <- face averages, once per step        evolvers.py:73-76
//   RefineFieldKernel: face-averaged B -> cell-averaged B after every register update        mag_field.py:191-211
//
// All kernels here work in the x frame ([x][var][y] planes).
#pragma once
#include "physics.cuh"
#include "runtime.cuh"

namespace astrea {

// ------------------------------------------------------------------------------------------------ corner E_z
struct CornerEmfParams {
    // (wD, wU) of mag_field.reconstruct_transverse, both bundles in the x frame:
    //   bundle 0: face states of the y sweep reconstructed along x  (swapped_permutations key 0, wave speeds along x)
    //   bundle 1: face states of the x sweep reconstructed along y  (key 1, wave speeds along y)
    Plane d0, u0, d1, u1;
    double* emf;               // out: [nrow][ncol], pitch ncol
    int64_t nrow, ncol;
    int64_t nx_glob, x_off;    // 'edge' clamp of the +1 neighbour along x
    double gamma;
    int bc, parity;
};
struct CornerEmfKernel {
    using Params = CornerEmfParams;
    static constexpr int MAX_THREADS = 128;
    // mag_field.py:128-149 ('hll' branch): a+ = max(0, v_n + c_f), a- = -min(0, v_n - c_f) at the Roe average.
    // The Lax-type branch (:152-159) takes max / -min of np.linalg.eigvals of the primitive Jacobian, whose spectrum
    // contains 0, v_n +- c_f: the same two numbers up to LAPACK round-off, so this closed form serves every solver.
    template <int AXIS>
    static HD void speeds(const double* plus, const double* minus, double gamma, double& ap, double& am) {
        double avg[NVAR];
        roe_state(plus, minus, avg);
        const double rho = avg[0], P = avg[4];
        const double vn = avg[1 + AXIS], Bn = avg[5 + AXIS];
        const double a = sqrt(gamma * sdiv(P, rho));
        const double sr = sqrt(rho);
        const double b = sdiv(norm3(avg[5], avg[6], avg[7]), sr);
        const double bn = sdiv(Bn, sr);
        const double cf = sqrt(0.5 * (a * a + b * b + sqrt(sq(a * a + b * b) - (4.0 * (a * a) * (bn * bn)))));
        ap = npmax(0.0, vn + cf);
        am = -npmin(0.0, vn - cf);
    }
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        const int NT = ex.nthreads();
        ex.phase([&](int tid) {
            const int64_t j = (int64_t)bx * NT + tid, i = by;
            if (j >= p.ncol) return;
            const bool edge = p.bc == BC_EDGE;
            // pad(wD)[1:]: the +1 neighbour along the reconstruction direction (ghost data when periodic / interior)
            const int64_t in = edge ? clamp_index(i + 1 + p.x_off, 0, p.nx_glob - 1) - p.x_off : i + 1;
            const int64_t jn = edge ? clamp_index(j + 1, 0, p.ncol - 1) : j + 1;
            double D0[NVAR], U0[NVAR], D1[NVAR], U1[NVAR], nb[NVAR];
            double ap0, am0, ap1, am1;
#pragma unroll
            for (int v = 0; v < NVAR; ++v) { U0[v] = *p.u0.at(i, v, j); nb[v] = *p.d0.at(in, v, j); }
            speeds<0>(nb, U0, p.gamma, ap0, am0);
#pragma unroll
            for (int v = 0; v < NVAR; ++v) { U1[v] = *p.u1.at(i, v, j); nb[v] = *p.d1.at(i, v, jn); }
            speeds<1>(nb, U1, p.gamma, ap1, am1);
#pragma unroll
            for (int v = 0; v < NVAR; ++v) { D0[v] = *p.d0.at(i, v, j); D1[v] = *p.d1.at(i, v, j); }
            // mag_field.py:164-185 unpacks by iteration order of the (reversed every step) permutations
            const double* north = p.parity ? D1 : D0;
            const double* south = p.parity ? U1 : U0;
            const double* east = p.parity ? D0 : D1;
            const double* west = p.parity ? U0 : U1;
            const double ap_y = p.parity ? ap1 : ap0, am_y = p.parity ? am1 : am0;
            const double ap_x = p.parity ? ap0 : ap1, am_x = p.parity ? am0 : am1;
            const double NE = 0.5 * (west[2] + south[2]) * south[5] - 0.5 * (west[1] + south[1]) * west[6];
            const double NW = 0.5 * (east[2] + south[2]) * south[5] - 0.5 * (east[1] + south[1]) * east[6];
            const double SE = 0.5 * (west[2] + north[2]) * north[5] - 0.5 * (west[1] + north[1]) * west[6];
            const double SW = 0.5 * (east[2] + north[2]) * north[5] - 0.5 * (east[1] + north[1]) * east[6];
            const double e = sdiv(ap_x * ap_y * SW + am_x * ap_y * SE + ap_x * am_y * NW + am_x * am_y * NE, (ap_x + am_x) * (ap_y + am_y))
                           - sdiv(ap_y * am_y, ap_y + am_y) * (north[5] - south[5])
                           + sdiv(ap_x * am_x, ap_x + am_x) * (east[6] - west[6]);
            p.emf[i * p.ncol + j] = e;
        });
    }
};

// ------------------------------------------------------------------------------------------------ face field
struct FaceFieldParams {
    Plane grid;        // in/out: components 5 and 6 are overwritten
    Plane wfx;         // face states of the x sweep, kept as a y-frame plane [y][v][x] (api.cu corner_field)
    Plane wfy;         // face states of the y sweep, kept as an x-frame plane [x][v][y]
    int64_t nrow, ncol;
};
struct FaceFieldKernel {
    using Params = FaceFieldParams;
    static constexpr int MAX_THREADS = 256;
    static constexpr int TILE = 32;
    static size_t smem_bytes() { return sizeof(double) * TILE * (TILE + 1); }
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        double* tile = ex.smem();
        const int64_t c0 = (int64_t)bx * TILE, r0 = (int64_t)by * TILE;
        ex.phase([&](int tid) {
            const int tx = tid % TILE;
            for (int ty = tid / TILE; ty < TILE; ty += MAX_THREADS / TILE) {
                const int64_t yr = c0 + ty, xc = r0 + tx;
                if (yr < p.ncol && xc < p.nrow) tile[ty * (TILE + 1) + tx] = *p.wfx.at(yr, 5, xc);
            }
        });
        ex.phase([&](int tid) {
            const int tx = tid % TILE;
            for (int ty = tid / TILE; ty < TILE; ty += MAX_THREADS / TILE) {
                const int64_t r = r0 + ty, c = c0 + tx;
                if (r >= p.nrow || c >= p.ncol) continue;
                *p.grid.at(r, 5, c) = tile[tx * (TILE + 1) + ty];
                *p.grid.at(r, 6, c) = *p.wfy.at(r, 6, c);
            }
        });
    }
};

// ------------------------------------------------------------------------------------------------ refine (inverse reconstruct)
struct RefineFieldParams {
    Plane in;          // register with face-averaged Bx, By in components 5, 6 (ghost filled)
    Plane out;         // components 5, 6 receive the cell averages (a scratch plane; copied back by CopyFieldKernel)
    int64_t nrow, ncol;
    int64_t nx_glob, x_off;
    int bc;
    int copy_back;     // 1: this launch copies out -> in instead
};
// Tiled: a block refines TX x TY cells.  The three levels of the reconstruction (face-centred value, 4-point centre
// interpolation, centre -> average) are built once per tile in shared memory instead of being re-derived per cell (the
// per-cell form evaluates ~60 loads and ~20 face conversions per cell and component; 4.3 ms per step at 4096^2).  Every
// level is read at the MAPPED index ("pad the derived array": at a physical 'edge' boundary the neighbour beyond the
// domain is the boundary entry of the derived array itself, not a value derived from ghost cells), so a level only has
// to exist at positions of the domain, and a mapped index never leaves the tile's halo.
struct RefineFieldKernel {
    using Params = RefineFieldParams;
    static constexpr int MAX_THREADS = 256;
    static constexpr int TX = 32, TY = 8, H = 3;
    static constexpr int GW = TX + 2 * H, GH = TY + 2 * H;     // every level is stored on the tile + H halo grid
    static size_t smem_bytes() { return sizeof(double) * 6 * GW * GH; }
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        const int64_t c0 = (int64_t)bx * TX, r0 = (int64_t)by * TY;
        if (p.copy_back) {
            ex.phase([&](int tid) {
                const int64_t j = c0 + tid % TX, i = r0 + tid / TX;
                if (j >= p.ncol || i >= p.nrow) return;
                *p.in.at(i, 5, j) = *p.out.at(i, 5, j);
                *p.in.at(i, 6, j) = *p.out.at(i, 6, j);
            });
            return;
        }
        double* G5 = ex.smem();            // face averages of Bx / By
        double* G6 = G5 + GW * GH;
        double* F5 = G6 + GW * GH;         // face-centred values
        double* F6 = F5 + GW * GH;
        double* C5 = F6 + GW * GH;         // cell-centred values
        double* C6 = C5 + GW * GH;
        const bool edge = p.bc == BC_EDGE;
        const double c24 = 1.0 / 24.0;
        // boundary maps of the reference's np.pad on the derived arrays; tile-local addressing
        auto mx = [&](int64_t r) -> int64_t { return edge ? clamp_index(r + p.x_off, 0, p.nx_glob - 1) - p.x_off : r; };
        auto my = [&](int64_t c) -> int64_t { return edge ? clamp_index(c, 0, p.ncol - 1) : c; };
        auto at = [&](const double* a, int64_t r, int64_t c) -> double { return a[(r - r0 + H) * GW + (c - c0 + H)]; };
        auto inside = [&](int64_t r, int64_t c) -> bool {      // positions at which a derived array exists
            return !edge || (r + p.x_off >= 0 && r + p.x_off <= p.nx_glob - 1 && c >= 0 && c <= p.ncol - 1);
        };
        // visit the tile grown by (hr, hc) cells
        auto region = [&](int tid, int hr_lo, int hr_hi, int hc_lo, int hc_hi, auto&& f) {
            const int w = TX + hc_lo + hc_hi, h = TY + hr_lo + hr_hi;
            for (int e = tid; e < w * h; e += MAX_THREADS) f(r0 - hr_lo + e / w, c0 - hc_lo + e % w);
        };
        ex.phase([&](int tid) {
            region(tid, H, H, H, H, [&](int64_t r, int64_t c) {
                const int k = (int)((r - r0 + H) * GW + (c - c0 + H));
                G5[k] = *p.in.at(r, 5, c);
                G6[k] = *p.in.at(r, 6, c);
            });
        });
        ex.phase([&](int tid) {
            // fv.high_order_convert('avg', ., 'face'): x - d2_transverse / 24      (mag_field.py:199-207)
            region(tid, 2, 3, 1, 1, [&](int64_t r, int64_t c) {          // Bx: sweep x, transverse y
                if (!inside(r, c)) return;
                const double a = at(G5, r, c);
                F5[(r - r0 + H) * GW + (c - c0 + H)] = a - c24 * ((at(G5, r, my(c + 1)) - a) - (a - at(G5, r, my(c - 1))));
            });
            region(tid, 1, 1, 2, 3, [&](int64_t r, int64_t c) {          // By: sweep y, transverse x
                if (!inside(r, c)) return;
                const double a = at(G6, r, c);
                F6[(r - r0 + H) * GW + (c - c0 + H)] = a - c24 * ((at(G6, mx(r + 1), c) - a) - (a - at(G6, mx(r - 1), c)));
            });
        });
        ex.phase([&](int tid) {
            // 4-point face -> centre interpolation along the sweep
            region(tid, 1, 1, 1, 1, [&](int64_t r, int64_t c) {
                if (!inside(r, c)) return;
                const int k = (int)((r - r0 + H) * GW + (c - c0 + H));
                C5[k] = -1.0 / 16.0 * (at(F5, mx(r - 1), c) + at(F5, mx(r + 2), c)) + 9.0 / 16.0 * (at(F5, r, c) + at(F5, mx(r + 1), c));
                C6[k] = -1.0 / 16.0 * (at(F6, r, my(c - 1)) + at(F6, r, my(c + 2))) + 9.0 / 16.0 * (at(F6, r, c) + at(F6, r, my(c + 1)));
            });
        });
        ex.phase([&](int tid) {
            const int64_t j = c0 + tid % TX, i = r0 + tid / TX;
            if (j >= p.ncol || i >= p.nrow) return;
            {   // centre -> average, axis 0 of the component's own frame first: x then y for Bx
                const double a = at(C5, i, j);
                double ca = a + c24 * ((at(C5, mx(i + 1), j) - a) - (a - at(C5, mx(i - 1), j)));
                ca = ca + c24 * ((at(C5, i, my(j + 1)) - a) - (a - at(C5, i, my(j - 1))));
                *p.out.at(i, 5, j) = ca;
            }
            {   // y then x for By
                const double a = at(C6, i, j);
                double ca = a + c24 * ((at(C6, i, my(j + 1)) - a) - (a - at(C6, i, my(j - 1))));
                ca = ca + c24 * ((at(C6, mx(i + 1), j) - a) - (a - at(C6, mx(i - 1), j)));
                *p.out.at(i, 6, j) = ca;
            }
        });
    }
};

}  // namespace astrea
