// Data-movement and Runge-Kutta kernels around the sweep kernels.
//   pack / unpack     host AoS (N[,N],8) C-order  <->  ghost-padded planes [row][var][col]   (+ cons->prim on download,
//                     astrea.py:47 snapshots are primitive: fv.py:97-101 / :126-143)
//   halo fill         fv.add_boundary (fv.py:57-61) as ghost cells of the *state* array: 'wrap' | 'edge'
//   transpose         grid.transpose(axes) of the reference's y sweep
//   assemble rate     compute_L (evolvers.py:41-60): L = -(dF_x/dx + dF_y/dx)
//   combine           one Runge-Kutta register update, evaluation order as written in evolvers.py:79-206
#pragma once
#include "physics.cuh"
#include "runtime.cuh"

namespace astrea {

// ------------------------------------------------------------------------------------------------ pack / unpack
struct PackParams {
    Plane plane;
    double* aos;          // device staging buffer, (nrow, ncol, 8) C-order
    int64_t nrow, ncol;
    int to_plane;         // 1: aos -> plane, 0: plane -> aos
    int* mhd_flag;        // to_plane: set to 1 if any v_z / B component is non-zero (nullptr: do not check)
};
struct PackKernel {
    using Params = PackParams;
    static constexpr int MAX_THREADS = 256;
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        const int NT = ex.nthreads();
        ex.phase([&](int tid) {
            const int64_t c = (int64_t)bx * NT + tid, r = by;
            if (c >= p.ncol) return;
            double* a = p.aos + (r * p.ncol + c) * NVAR;
#pragma unroll
            for (int v = 0; v < NVAR; ++v) {
                if (p.to_plane) *p.plane.at(r, v, c) = a[v]; else a[v] = *p.plane.at(r, v, c);
            }
            if (p.to_plane && p.mhd_flag != nullptr && (a[3] != 0.0 || a[5] != 0.0 || a[6] != 0.0 || a[7] != 0.0)) *p.mhd_flag = 1;
        });
    }
};

// plane -> AoS with the two grid axes swapped: aos[(c * nrow + r) * 8 + v] = plane(r, v, c), the layout astrea.py:47
// stores (``.transpose(ortho_axis)``, ortho_axis = (1, 0, 2) in 2D).  32 x 32 cell tiles through shared memory so that
// both the plane reads (along c) and the AoS writes (along r, 64 B per cell) are contiguous.
struct PackTransposedParams {
    Plane plane;
    double* aos;          // device buffer (ncol, nrow, 8) C-order
    int64_t nrow, ncol;
};
struct PackTransposedKernel {
    using Params = PackTransposedParams;
    static constexpr int MAX_THREADS = 256;
    static constexpr int TILE = 32;
    static size_t smem_bytes() { return sizeof(double) * NVAR * TILE * (TILE + 1); }
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        double* tile = ex.smem();          // [v][r][c + pad]
        const int64_t c0 = (int64_t)bx * TILE, r0 = (int64_t)by * TILE;
        ex.phase([&](int tid) {
            const int tx = tid % TILE;
            for (int ty = tid / TILE; ty < TILE; ty += MAX_THREADS / TILE) {
                const int64_t r = r0 + ty, c = c0 + tx;
                if (r < p.nrow && c < p.ncol) {
#pragma unroll
                    for (int v = 0; v < NVAR; ++v) tile[(v * TILE + ty) * (TILE + 1) + tx] = *p.plane.at(r, v, c);
                }
            }
        });
        ex.phase([&](int tid) {
            // consecutive threads write consecutive doubles of the output: (c fixed, r running, v fastest)
            for (int e = tid; e < TILE * TILE * NVAR; e += MAX_THREADS) {
                const int v = e % NVAR, ty = (e / NVAR) % TILE, tx = e / (NVAR * TILE);
                const int64_t r = r0 + ty, c = c0 + tx;
                if (r < p.nrow && c < p.ncol) p.aos[(c * p.nrow + r) * NVAR + v] = tile[(v * TILE + ty) * (TILE + 1) + tx];
            }
        });
    }
};

// ------------------------------------------------------------------------------------------------ halo fill
struct HaloParams {
    Plane plane;
    int64_t nrow, ncol;
    int bc;
    int phase;            // 0: ghost columns of interior rows, 1: ghost rows (full padded width, corners included)
    int fill_lo, fill_hi; // phase 1: which row ghosts to fill locally (0 when a neighbour rank provides them)
    VarList vars;
    int64_t row0;         // phase 0: first interior row to treat (blockIdx.y counts from here)
    int skip_col_lo, skip_col_hi;   // phase 0: leave the low / high ghost columns alone (they hold genuine neighbour data)
};
struct HaloKernel {
    using Params = HaloParams;
    static constexpr int MAX_THREADS = 256;
    static HD int64_t src(int64_t g, int64_t n, int bc) { return bc == BC_WRAP ? wrap_index(g, n) : clamp_index(g, 0, n - 1); }
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        const int NT = ex.nthreads();
        ex.phase([&](int tid) {
            if (p.phase == 0) {
                // by = row, threads over the 2*GHOST ghost columns x NVAR
                const int64_t r = p.row0 + by;
                for (int e = tid; e < 2 * GHOST * p.vars.n; e += NT) {
                    const int v = p.vars.v[e / (2 * GHOST)], g = e % (2 * GHOST);
                    if ((g < GHOST && p.skip_col_lo) || (g >= GHOST && p.skip_col_hi)) continue;
                    const int64_t c = g < GHOST ? g - GHOST : p.ncol + (g - GHOST);
                    *p.plane.at(r, v, c) = *p.plane.at(r, v, src(c, p.ncol, p.bc));
                }
            } else {
                // by = ghost row id (0..2*GHOST-1) x var, threads over padded columns
                const int g = by / NVAR;
                if (by % NVAR >= p.vars.n) return;
                const int v = p.vars.v[by % NVAR];
                const bool lo = g < GHOST;
                if ((lo && !p.fill_lo) || (!lo && !p.fill_hi)) return;
                const int64_t r = lo ? g - GHOST : p.nrow + (g - GHOST);
                const int64_t rs = src(r, p.nrow, p.bc);
                const int64_t c = (int64_t)bx * NT + tid - GHOST;
                if (c < p.ncol + GHOST) *p.plane.at(r, v, c) = *p.plane.at(rs, v, c);
            }
        });
    }
};

// ------------------------------------------------------------------------------------------------ transpose
struct TransposeParams {
    Plane src, dst;       // dst(row = c, col = r) = src(row = r, col = c), ghosts included
    int64_t r_lo, r_hi, c_lo, c_hi;   // half-open ranges of src rows / cols to move
    VarList vars;
};
struct TransposeKernel {
    using Params = TransposeParams;
    static constexpr int MAX_THREADS = 256;
    static constexpr int TILE = 32;
    static size_t smem_bytes() { return sizeof(double) * TILE * (TILE + 1); }
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        double* tile = ex.smem();
        const int64_t c0 = p.c_lo + (int64_t)bx * TILE, r0 = p.r_lo + (int64_t)by * TILE;
        for (int a = 0; a < p.vars.n; ++a) {
            const int v = p.vars.v[a];
            ex.phase([&](int tid) {
                const int tx = tid % TILE;
                for (int ty = tid / TILE; ty < TILE; ty += MAX_THREADS / TILE) {
                    const int64_t r = r0 + ty, c = c0 + tx;
                    if (r < p.r_hi && c < p.c_hi) tile[ty * (TILE + 1) + tx] = *p.src.at(r, v, c);
                }
            });
            ex.phase([&](int tid) {
                const int tx = tid % TILE;
                for (int ty = tid / TILE; ty < TILE; ty += MAX_THREADS / TILE) {
                    const int64_t r = r0 + tx, c = c0 + ty;     // dst row = c, dst col = r
                    if (r < p.r_hi && c < p.c_hi) *p.dst.at(c, v, r) = tile[tx * (TILE + 1) + ty];
                }
            });
        }
    }
};

// ------------------------------------------------------------------------------------------------ assemble L
struct RateParams {
    Plane f0;             // Riemann flux of the x sweep at interface rows 0..nrow, [x][v][y]  (1D: unused, see d0)
    Plane f1t;            // same of the y sweep in its own (transposed) frame, [y][v][x]
    Plane d0;             // 1D only: (F[i+1]-F[i])/dx written by the fused 1D sweep kernel
    Plane out;            // L = -(dF_x/dx + dF_y/dx)   (evolvers.py:41-60)
    int64_t nrow, ncol;
    int dimension;
    // constrained transport (evolvers.py:52-58): overwrite the in-plane field rates with emf differences
    const double* emf;    // corner field [x][y] (pitch ncol), or nullptr
    int64_t emf_rows;     // rows of emf that hold data: nrow, or nrow + 1 when the row behind the slab was computed from ghost data
    int64_t nx_glob, x_off;
    double dx;
    double inv_dx_exact;  // 1 / dx when dx is a power of two (x / dx == x * inv_dx_exact bit for bit), else 0
    int bc;
    VarList vars;
    int64_t row_lo, row_hi;   // half-open range of rows to process (slab hosts update the edge rows first)
};
struct RateKernel {
    using Params = RateParams;
    static constexpr int MAX_THREADS = 256;
    static constexpr int TILE = 32;
    static size_t smem_bytes() { return sizeof(double) * TILE * (TILE + 1); }
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        double* tile = ex.smem();
        const int64_t c0 = (int64_t)bx * TILE, r0 = p.row_lo + (int64_t)by * TILE;
        for (int a = 0; a < p.vars.n; ++a) {
            const int v = p.vars.v[a];
            if (p.dimension == 2) {
                ex.phase([&](int tid) {     // flux difference of the y sweep, read coalesced along its own columns (= x)
                    const int tx = tid % TILE;
                    for (int ty = tid / TILE; ty < TILE; ty += MAX_THREADS / TILE) {
                        const int64_t yr = c0 + ty, xc = r0 + tx;
                        if (yr < p.ncol && xc < p.row_hi)
                            tile[ty * (TILE + 1) + tx] = (*p.f1t.at(yr + 1, v, xc) - *p.f1t.at(yr, v, xc)) / p.dx;
                    }
                });
            }
            ex.phase([&](int tid) {
                const int tx = tid % TILE;
                for (int ty = tid / TILE; ty < TILE; ty += MAX_THREADS / TILE) {
                    const int64_t r = r0 + ty, c = c0 + tx;
                    if (r >= p.row_hi || c >= p.ncol) continue;
                    double total;
                    if (p.dimension == 2) {
                        // compute_L sums the sweeps in iteration order (evolvers.py:45-49); the sum of two terms
                        // does not depend on that order
                        total = (*p.f0.at(r + 1, v, c) - *p.f0.at(r, v, c)) / p.dx;
                        total = total + tile[tx * (TILE + 1) + ty];
                    } else {
                        total = *p.d0.at(r, v, c);
                    }
                    if (p.emf != nullptr && (v == 5 || v == 6)) {
                        // diff(pad(E_z)[1:]) (evolvers.py:56-57): the +1 neighbour wraps or clamps
                        const bool wrap = p.bc == BC_WRAP;
                        const double e0 = p.emf[r * p.ncol + c];
                        if (v == 5) {
                            const int64_t cn = c + 1 < p.ncol ? c + 1 : (wrap ? 0 : p.ncol - 1);
                            total = (p.emf[r * p.ncol + cn] - e0) / p.dx;                 // (-1)^0 dE/dy
                        } else {
                            const int64_t rn = r + 1 < p.emf_rows ? r + 1 : (wrap ? 0 : p.nrow - 1);
                            total = (-1.0 * (p.emf[rn * p.ncol + c] - e0)) / p.dx;        // (-1)^1 dE/dx
                        }
                    }
                    *p.out.at(r, v, c) = -total;
                }
            });
        }
    }
};

// ------------------------------------------------------------------------------------------------ RK combine
// out = scale * ( sum_k coef_k * X_k )              terms evaluated and added left to right;
//   X_k is a state register (coef used as is) or a rate buffer (coef multiplied by dt first: (b*dt)*L),
// or, for the two closing formulas that bracket the rates (evolvers.py:146 and :202),
// out = sum_regs a_i R_i + scale * (dt * (sum_rates b_j L_j)).
constexpr int MAX_TERMS = 8;
struct CombineParams {
    Plane out;
    Plane term[MAX_TERMS];
    double coef[MAX_TERMS];
    int is_rate[MAX_TERMS];
    int nterms;
    int bracket_rates;    // 0: interleaved form, 1: bracketed form
    double scale;         // 1.0 = no scaling
    const double* dt;     // device scalar
    int64_t nrow, ncol;
    VarList vars;
    int64_t row_lo;       // first row to process (blockIdx.y counts from here)
};
struct CombineKernel {
    using Params = CombineParams;
    static constexpr int MAX_THREADS = 256;
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        const int NT = ex.nthreads();
        ex.phase([&](int tid) {
            const int64_t c = (int64_t)bx * NT + tid;
            const int64_t r = p.row_lo + by;
            if (c >= p.ncol) return;
            const double dt = *p.dt;
            for (int a = 0; a < p.vars.n; ++a) {
                const int v = p.vars.v[a];
                double acc = 0.0;
                if (!p.bracket_rates) {
                    for (int k = 0; k < p.nterms; ++k) {
                        const double x = *p.term[k].at(r, v, c);
                        const double t = p.is_rate[k] ? (p.coef[k] * dt) * x : p.coef[k] * x;
                        acc = (k == 0) ? t : acc + t;
                    }
                    if (p.scale != 1.0) acc = p.scale * acc;
                } else {
                    double regs = 0.0, rates = 0.0;
                    bool fr = true, fl = true;
                    for (int k = 0; k < p.nterms; ++k) {
                        const double x = *p.term[k].at(r, v, c);
                        if (p.is_rate[k]) { const double t = p.coef[k] * x; rates = fl ? t : rates + t; fl = false; }
                        else { const double t = p.coef[k] * x; regs = fr ? t : regs + t; fr = false; }
                    }
                    double tail = dt * rates;
                    if (p.scale != 1.0) tail = p.scale * tail;
                    acc = regs + tail;
                }
                *p.out.at(r, v, c) = acc;
            }
        });
    }
};

// ------------------------------------------------------------------------------------------------ fused rate + RK update
// A register update whose newest rate L is used nowhere else (SSPRK(2,2), (3,3), (4,3), (10,4), Euler, and the
// last use in the others) does not need L in memory: it is assembled from the flux planes on the fly (term with
// is_rate == 2) inside the update, which saves one plane write and one plane read per stage.
struct UpdateParams {
    RateParams rate;      // rate.out is unused
    CombineParams comb;   // comb.term[k] with is_rate[k] == 2 is the rate assembled on the fly
    Plane rate_store;     // base != nullptr: also store the assembled rate (it is re-used by a later formula)
};
// NTERMS / BRACKET are compile-time so that the term loop unrolls and every pointer and coefficient stays in a
// register (a run-time term count costs ~370 instructions per cell and variable, mostly 64-bit address arithmetic).
template <int NTERMS, bool BRACKET>
struct UpdateKernel {
    using Params = UpdateParams;
    static constexpr int MAX_THREADS = 256;
    static constexpr int TILE = 32;
#ifndef ASTREA_UPDATE_MIN_BLOCKS
#define ASTREA_UPDATE_MIN_BLOCKS 4
#endif
    static constexpr int MIN_BLOCKS = ASTREA_UPDATE_MIN_BLOCKS;   // 64 registers: 1024 threads per SM keep enough loads in flight
#ifndef ASTREA_UPDATE_VG
#define ASTREA_UPDATE_VG 4
#endif
    static constexpr int VG = ASTREA_UPDATE_VG;           // variables per pass through the tile (4: a hydro state is one pass)
    static constexpr int ROWS = TILE * TILE / MAX_THREADS; // tile rows per thread
    static size_t smem_bytes() { return sizeof(double) * VG * TILE * (TILE + 1); }
    // Division by dx: when dx is a power of two (every BASELINE configuration: unit-length boxes with 2^k cells) the
    // IEEE quotient x / dx equals the product x * (1 / dx) for every x (scaling by a power of two is exact), so the
    // kernel multiplies; otherwise it divides.  The register update is in place (``out`` is one of the terms), so it
    // cannot be repeated the way the Fast-guarded kernels are.
    struct DivideByDx { double dx; HD double operator()(double x) const { return x / dx; } };
    struct TimesInvDx { double inv; HD double operator()(double x) const { return x * inv; } };
    template <class Ex>
    static HD void block(const Params& pp, int bx, int by, Ex& ex) {
        if (pp.rate.inv_dx_exact != 0.0) body(pp, bx, by, ex, TimesInvDx{pp.rate.inv_dx_exact});
        else body(pp, bx, by, ex, DivideByDx{pp.rate.dx});
    }
    template <class Ex, class D>
    static HD void body(const Params& pp, int bx, int by, Ex& ex, D over_dx) {
        const RateParams& p = pp.rate;
        const CombineParams& cb = pp.comb;
        double* tile = ex.smem();
        const int64_t c0 = (int64_t)bx * TILE, r0 = p.row_lo + (int64_t)by * TILE;
        const bool two_d = p.dimension == 2;
        // every plane of a context has the same geometry: one offset addresses them all
        const int64_t rp = cb.out.row_pitch, cp = cb.out.col_pitch;
        if (!two_d) {
            // 1D: one row, a thread per cell (blocks of MAX_THREADS cells), every variable in one go — no tile, no barrier
            ex.phase([&](int tid) {
                const int64_t c = (int64_t)bx * MAX_THREADS + tid;
                if (c >= p.ncol) return;
                const double dt = *cb.dt;
                for (int a = 0; a < p.vars.n; ++a) {
                    const int64_t off = (int64_t)p.vars.v[a] * cp + c;
                    const double L = -p.d0.base[off];
                    if (pp.rate_store.base != nullptr) pp.rate_store.base[off] = L;
                    double acc = 0.0;
                    if (!BRACKET) {
#pragma unroll
                        for (int k = 0; k < NTERMS; ++k) {
                            const double x = cb.is_rate[k] == 2 ? L : cb.term[k].base[off];
                            const double t = (cb.is_rate[k] ? cb.coef[k] * dt : cb.coef[k]) * x;
                            acc = (k == 0) ? t : acc + t;
                        }
                        if (cb.scale != 1.0) acc = cb.scale * acc;
                    } else {
                        double regs = 0.0, rates = 0.0;
                        bool fr = true, fl = true;
#pragma unroll
                        for (int k = 0; k < NTERMS; ++k) {
                            const double x = cb.is_rate[k] == 2 ? L : cb.term[k].base[off];
                            const double t = cb.coef[k] * x;
                            if (cb.is_rate[k]) { rates = fl ? t : rates + t; fl = false; }
                            else { regs = fr ? t : regs + t; fr = false; }
                        }
                        double tail = dt * rates;
                        if (cb.scale != 1.0) tail = cb.scale * tail;
                        acc = regs + tail;
                    }
                    cb.out.base[off] = acc;
                }
            });
            return;
        }
        // 2D: the variables go through the tile in groups of VG, so that one barrier covers a whole group and every
        // load of a group is in flight at once (the update is bound by memory latency, not by arithmetic)
        const double dt = *cb.dt;
        double cf[NTERMS];
        int kind[NTERMS];
#pragma unroll
        for (int k = 0; k < NTERMS; ++k) {
            kind[k] = cb.is_rate[k];
            cf[k] = (!BRACKET && kind[k]) ? cb.coef[k] * dt : cb.coef[k];
        }
        const bool wrap = p.bc == BC_WRAP;
        for (int a0 = 0; a0 < p.vars.n; a0 += VG) {
            ex.phase([&](int tid) {     // flux difference of the y sweep, read coalesced along its own columns (= x)
                const int tx = tid % TILE;
                const int64_t xc = r0 + tx;
#pragma unroll
                for (int gi = 0; gi < VG; ++gi) {
                    if (a0 + gi >= p.vars.n) continue;
                    const int v = p.vars.v[a0 + gi];
                    double lo[ROWS], hi[ROWS];
#pragma unroll
                    for (int i = 0; i < ROWS; ++i) {
                        const int ty = tid / TILE + i * (MAX_THREADS / TILE);
                        const int64_t yr = c0 + ty;
                        const bool ok = yr < p.ncol && xc < p.row_hi;
                        const double* f = p.f1t.at(ok ? yr : 0, v, ok ? xc : 0);
                        lo[i] = f[0];
                        hi[i] = f[p.f1t.row_pitch];
                    }
#pragma unroll
                    for (int i = 0; i < ROWS; ++i) {
                        const int ty = tid / TILE + i * (MAX_THREADS / TILE);
                        tile[(gi * TILE + ty) * (TILE + 1) + tx] = over_dx(hi[i] - lo[i]);
                    }
                }
            });
            ex.phase([&](int tid) {
                const int tx = tid % TILE;
                const int64_t c = c0 + tx;
#pragma unroll 1
                for (int i = 0; i < ROWS; ++i) {
                    const int ty = tid / TILE + i * (MAX_THREADS / TILE);
                    const int64_t r = r0 + ty;
                    if (r >= p.row_hi || c >= p.ncol) continue;
                    const int64_t off = r * rp + c;
                    double fl[VG], fh[VG], x[VG][NTERMS];
#pragma unroll
                    for (int gi = 0; gi < VG; ++gi) {
                        if (a0 + gi >= p.vars.n) continue;
                        const int64_t o = off + (int64_t)p.vars.v[a0 + gi] * cp;
                        fl[gi] = p.f0.base[o];
                        fh[gi] = p.f0.base[o + rp];
#pragma unroll
                        for (int k = 0; k < NTERMS; ++k) x[gi][k] = kind[k] == 2 ? 0.0 : cb.term[k].base[o];
                    }
#pragma unroll
                    for (int gi = 0; gi < VG; ++gi) {
                        if (a0 + gi >= p.vars.n) continue;
                        const int v = p.vars.v[a0 + gi];
                        const int64_t o = off + (int64_t)v * cp;
                        double total = over_dx(fh[gi] - fl[gi]);
                        total = total + tile[(gi * TILE + tx) * (TILE + 1) + ty];
                        if (p.emf != nullptr && (v == 5 || v == 6)) {
                            // diff(pad(E_z)[1:]) (evolvers.py:56-57): the +1 neighbour wraps or clamps
                            const double e0 = p.emf[r * p.ncol + c];
                            if (v == 5) {
                                const int64_t cn = c + 1 < p.ncol ? c + 1 : (wrap ? 0 : p.ncol - 1);
                                total = over_dx(p.emf[r * p.ncol + cn] - e0);                 // (-1)^0 dE/dy
                            } else {
                                const int64_t rn = r + 1 < p.emf_rows ? r + 1 : (wrap ? 0 : p.nrow - 1);
                                total = over_dx(-1.0 * (p.emf[rn * p.ncol + c] - e0));        // (-1)^1 dE/dx
                            }
                        }
                        const double L = -total;
                        if (pp.rate_store.base != nullptr) pp.rate_store.base[o] = L;
                        double acc = 0.0;
                        if (!BRACKET) {
#pragma unroll
                            for (int k = 0; k < NTERMS; ++k) {
                                const double t = cf[k] * (kind[k] == 2 ? L : x[gi][k]);
                                acc = (k == 0) ? t : acc + t;
                            }
                            if (cb.scale != 1.0) acc = cb.scale * acc;
                        } else {
                            double regs = 0.0, rates = 0.0;
                            bool fr = true, fl_ = true;
#pragma unroll
                            for (int k = 0; k < NTERMS; ++k) {
                                const double t = cf[k] * (kind[k] == 2 ? L : x[gi][k]);
                                if (kind[k]) { rates = fl_ ? t : rates + t; fl_ = false; }
                                else { regs = fr ? t : regs + t; fr = false; }
                            }
                            double tail = dt * rates;
                            if (cb.scale != 1.0) tail = cb.scale * tail;
                            acc = regs + tail;
                        }
                        cb.out.base[o] = acc;
                    }
                }
            });
        }
    }
};

// ------------------------------------------------------------------------------------------------ PPM dissipation
// schemes/ppm.py:111-170 — the slope flattener coefficient [Colella 1990] and the artificial-viscosity term
// [McCorquodale & Colella 2011, eq. 35-38] that ppm.run(dissipate=True) asks for.  The reference cannot run with
// dissipate=True (ppm.py:67 multiplies arrays whose shapes do not broadcast), so neither is on the time-step path;
// both functions run standalone and are reproduced at function level: wS (primitive cell averages in the sweep frame,
// axis 0 = sweep direction) in, coefficient out.  ``axis`` picks the velocity component w[..., axis + 1].
// Boundary handling as everywhere: ghost cells of wS are np.pad copies ('wrap' | 'edge', filled by HaloKernel); the
// padded *derived* array chi_bar (ppm.py:126) is evaluated at the mapped index.
struct DissipationParams {
    Plane w;              // primitive cell averages, ghost cells filled
    Plane out;            // flattener: chi in variable slot 0; viscosity: mu in all 8 slots
    int64_t nrow, ncol;
    int dimension;        // 1: the sweep runs along the columns of the single row; 2: along the rows
    int axis;             // velocity component (the reference's permutation key)
    int bc;
    int what;             // 0: apply_flattener, 1: apply_artificial_viscosity
    double delta, z0, z1; // slope_determinants  (ppm.py:112)
    double alpha, beta;   // viscosity_determinants (ppm.py:139)
    double gamma, dx;
};
struct DissipationKernel {
    using Params = DissipationParams;
    static constexpr int MAX_THREADS = 128;
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        const int NT = ex.nthreads();
        ex.phase([&](int tid) {
            const int64_t c = (int64_t)bx * NT + tid, r = by;
            if (c >= p.ncol) return;
            const bool along_rows = p.dimension == 2;
            const int64_t n = along_rows ? p.nrow : p.ncol, i = along_rows ? r : c;
            // value of variable v at sweep index i + k (ghost cells hold the padded values)
            auto at = [&](int64_t k, int v) -> double { return along_rows ? *p.w.at(r + k, v, c) : *p.w.at(r, v, c + k); };
            if (p.what == 0) {
                // chi_bar of the cell at sweep offset k from this one (ppm.py:124-125)
                auto chi_bar = [&](int64_t k) -> double {
                    const double pp1 = at(k + 1, 4), pm1 = at(k - 1, 4), pp2 = at(k + 2, 4), pm2 = at(k - 2, 4);
                    const double d1 = fabs(pp1 - pm1), d2 = fabs(pp2 - pm2);
                    const double z = sdiv(d1, d2);
                    double zeta = 1.0 - sdiv(z - p.z0, p.z1 - p.z0);
                    if (z > p.z1) zeta = 0.0;
                    if (z < p.z0) zeta = 1.0;
                    const bool compressive_weak = ((at(k - 1, 1 + p.axis) - at(k + 1, 1 + p.axis)) <= 0.0) && (sdiv(d1, npmin(pp1, pm1)) <= p.delta);
                    return compressive_weak ? 0.0 : zeta;
                };
                // pad of the derived array: 'wrap' ghost data are genuine, 'edge' repeats the boundary cell's value
                auto mapped = [&](int64_t k) -> int64_t { return p.bc == BC_WRAP ? k : clamp_index(i + k, 0, n - 1) - i; };
                const double own = chi_bar(0);
                const double sign = npsign(at(1, 4) - at(-1, 4));
                double chi = own;
                if (sign < 0.0) chi = npmin(own, chi_bar(mapped(1)));
                if (sign > 0.0) chi = npmin(own, chi_bar(mapped(-1)));
                *p.out.at(r, 0, c) = chi;
            } else {
                // ppm.py:138-170 read cell by cell (1D): lambda_R, c_min, nu, mu of the face right of cell i
                const double lam = at(1, 1 + p.axis) - at(0, 1 + p.axis);
                const double cs0 = dsqrt(sdiv(p.gamma * at(0, 4), at(0, 0))), cs1 = dsqrt(sdiv(p.gamma * at(1, 4), at(1, 0)));
                const double cmin = npmin(cs0, cs1);
                double nu = npmin(1.0, sdiv(sq(p.dx * lam), p.beta * sq(cmin))) * lam;
                if (lam >= 0.0) nu = 0.0;
#pragma unroll
                for (int v = 0; v < NVAR; ++v) *p.out.at(r, v, c) = p.alpha * (nu * (at(1, v) - at(0, v)));
            }
        });
    }
};

// ------------------------------------------------------------------------------------------------ device-side clock
// astrea.py:70-78 without a host round trip: dt = cfl * min(dx / eigmax), clipped so that t + dt does not pass
// t_stop, written where the register updates read it; t and the step count advance on the device.
struct ClockParams {
    double* clock;                       // [0] t, [1] t_stop, [2] steps taken, [3] last dt
    double* dt;                          // the scalar the register updates read
    const unsigned long long* eig_bits;  // [2] per-axis max wave speed of operator 0 (bit patterns)
    double* history;                     // dt of step n at history[n % history_len]
    int history_len;
    int mode;                            // 0: set t, t_stop and reset the step count; 1: take dt from the wave speeds
    int dimension;
    double cfl, dx, t, t_stop;
};
struct ClockKernel {
    using Params = ClockParams;
    static constexpr int MAX_THREADS = 32;
    template <class Ex>
    static HD void block(const Params& p, int, int, Ex& ex) {
        ex.phase([&](int tid) {
            if (tid != 0) return;
            if (p.mode == 0) {
                p.clock[0] = p.t; p.clock[1] = p.t_stop; p.clock[2] = 0.0; p.clock[3] = 0.0;
                return;
            }
            double e0, e1;
            memcpy(&e0, &p.eig_bits[0], sizeof(double));
            memcpy(&e1, &p.eig_bits[1], sizeof(double));
            double m = p.dx / e0;
            if (p.dimension == 2) { const double m1 = p.dx / e1; m = m1 < m ? m1 : m; }
            double dt = p.cfl * m;
            const double t = p.clock[0], t_stop = p.clock[1];
            if (t_stop > t && t + dt >= t_stop) dt = t_stop - t;
            *p.dt = dt;
            const long long n = (long long)p.clock[2];
            p.history[n % p.history_len] = dt;
            p.clock[0] = t + dt;
            p.clock[2] = (double)(n + 1);
            p.clock[3] = dt;
        });
    }
};

// ------------------------------------------------------------------------------------------------ cons -> prim for download
struct PrimParams {
    Plane q, w;           // w may alias a scratch register; ghosts of q must be valid for the 4th-order conversion
    int64_t nrow, ncol;
    int dimension, high_order;
    double gamma;
};
struct PrimKernel {
    using Params = PrimParams;
    static constexpr int MAX_THREADS = 256;
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        const int NT = ex.nthreads();
        ex.phase([&](int tid) {
            const int64_t c = (int64_t)bx * NT + tid, r = by;
            if (c >= p.ncol) return;
            const double c24 = 1.0 / 24.0;
            double q[NVAR], w[NVAR];
            auto ld = [&](int64_t rr, int64_t cc, double* dst) {
#pragma unroll
                for (int v = 0; v < NVAR; ++v) dst[v] = *p.q.at(rr, v, cc);
            };
            ld(r, c, q);
            if (!p.high_order) {
                prim_of_cons(q, w, p.gamma);
            } else {
                double a[NVAR], b[NVAR], wa[NVAR], wb[NVAR], wc[NVAR], qa[NVAR], ws[NVAR];
                prim_of_cons(q, wc, p.gamma);
#pragma unroll
                for (int v = 0; v < NVAR; ++v) { qa[v] = q[v]; ws[v] = 0.0; }
                for (int ax = 0; ax < p.dimension; ++ax) {
                    // 1D data live in a single row: its only axis is the column axis
                    const bool along_rows = (p.dimension == 2 && ax == 0);
                    ld(along_rows ? r - 1 : r, along_rows ? c : c - 1, a);
                    ld(along_rows ? r + 1 : r, along_rows ? c : c + 1, b);
                    prim_of_cons(a, wa, p.gamma);
                    prim_of_cons(b, wb, p.gamma);
#pragma unroll
                    for (int v = 0; v < NVAR; ++v) {
                        qa[v] = qa[v] - c24 * ((b[v] - q[v]) - (q[v] - a[v]));
                        ws[v] = ws[v] + c24 * ((wb[v] - wc[v]) - (wc[v] - wa[v]));
                    }
                }
                prim_of_cons(qa, w, p.gamma);
#pragma unroll
                for (int v = 0; v < NVAR; ++v) w[v] = w[v] + ws[v];
            }
#pragma unroll
            for (int v = 0; v < NVAR; ++v) *p.w.at(r, v, c) = w[v];
        });
    }
};

// ------------------------------------------------------------------------------------------------ fp64 peak probe
// Measurement aid for bench.py: dependent-free chains of DFMA, so that the roofline discussion can quote the fp64
// rate this GPU actually sustains (the path is bound by the fp64 pipe).  8 independent accumulators per
// thread, `iters` x 8 FMAs each; the result is stored so that nothing is optimised away.
struct Fp64ProbeParams {
    double* out;
    int iters;
    double a, b;
};
struct Fp64ProbeKernel {
    using Params = Fp64ProbeParams;
    static constexpr int MAX_THREADS = 256;
    template <class Ex>
    static HD void block(const Params& p, int bx, int, Ex& ex) {
        const int NT = ex.nthreads();
        ex.phase([&](int tid) {
            double x0 = tid, x1 = tid + 1, x2 = tid + 2, x3 = tid + 3, x4 = tid + 4, x5 = tid + 5, x6 = tid + 6, x7 = tid + 7;
            const double a = p.a, b = p.b;
            for (int i = 0; i < p.iters; ++i) {
                x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
                x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
            }
            p.out[(int64_t)bx * NT + tid] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
        });
    }
};

// ------------------------------------------------------------------------------------------------ initial conditions
// constructor.initialise(sim_variables, convert=True) (functions/constructor.py:11-109) for problems whose pointwise
// primitive state is piecewise constant: the grid starts as ``initial_right`` and a list of regions is painted over
// it in order, like the reference's sequence of ``grid[np.where(...)] = state`` assignments (:46-60, :72-73).  The
// pointwise primitives become conservative cell averages as in :105-109 with generic.py:250-255: 4th-order
// (fv.py:126-143) or pointwise conversion, then ``+ laplacian / 24`` (fv.py:67-85), every sum in the reference's order.
// The square N x N problem is evaluated on the fly at base cell ((row + x_off) mod N, col), which also tiles it
// periodically along x for the slab-decomposed weak-scaling runs (initial.initial_slab).
constexpr int MAX_REGIONS = 8;
constexpr int MAX_PROFILES = 4;
enum RegionKind : int { REG_X_LT = 0, REG_X_LE = 1, REG_Y_LE = 2, REG_X_LE_Y_GE = 3, REG_X_GT_Y_GE = 4, REG_DISC_LE = 5 };
struct InitParams {
    Plane out;
    int64_t nrow, ncol, x_off;
    int64_t n;                 // cells per side of the square base problem
    double start, step;        // np.linspace(lo - half, hi + half, n + 2): point k is k * step + start, cell i is point i + 1
    double gamma;
    int bc, high_order, nregions;
    int kind[MAX_REGIONS];
    double a[MAX_REGIONS], b[MAX_REGIONS];     // threshold(s): shock position, or (centre, radius^2) for a disc
    double state[MAX_REGIONS + 1][NVAR];       // state[0]: background (initial_right); state[k + 1]: region k
    int* mhd_flag;
    // separable profiles painted over the regions (constructor.py:44, :64-67): variable prof_var[k] of the point (i, j)
    // becomes prof_tab[k][prof_along[k] == 0 ? i : j]; the n-entry tables are evaluated on the host with numpy, so the
    // transcendental function is the reference's own
    int nprofiles;
    int prof_var[MAX_PROFILES], prof_along[MAX_PROFILES];
    const double* prof_tab[MAX_PROFILES];
};
struct InitKernel {
    using Params = InitParams;
    static constexpr int MAX_THREADS = 128;
    static HD int64_t nb(int64_t i, int64_t n, int bc) { return bc == BC_WRAP ? wrap_index(i, n) : clamp_index(i, 0, n - 1); }
    static HD void point_prim(const Params& p, int64_t i, int64_t j, double* w) {
        const double x = (double)(i + 1) * p.step + p.start, y = (double)(j + 1) * p.step + p.start;
        int pick = 0;
        for (int k = 0; k < p.nregions; ++k) {
            bool in;
            switch (p.kind[k]) {
                case REG_X_LT: in = x < p.a[k]; break;
                case REG_X_LE: in = x <= p.a[k]; break;
                case REG_Y_LE: in = y <= p.a[k]; break;
                case REG_X_LE_Y_GE: in = x <= p.a[k] && y >= p.a[k]; break;
                case REG_X_GT_Y_GE: in = x > p.a[k] && y >= p.a[k]; break;
                default: in = ((x - p.a[k]) * (x - p.a[k]) + (y - p.a[k]) * (y - p.a[k])) <= p.b[k]; break;
            }
            if (in) pick = k + 1;
        }
#pragma unroll
        for (int v = 0; v < NVAR; ++v) w[v] = p.state[pick][v];
        for (int k = 0; k < p.nprofiles; ++k) w[p.prof_var[k]] = p.prof_tab[k][p.prof_along[k] == 0 ? i : j];
    }
    // conservative point / 4th-order value at base cell (i, j) before the final Laplacian (initial._cons_from_point_prim)
    static HD void cons_at(const Params& p, int64_t i, int64_t j, double* q) {
        const double c24 = 1.0 / 24.0;
        double w[NVAR];
        point_prim(p, i, j, w);
        if (!p.high_order) { cons_of_prim(w, q, p.gamma); return; }
        double wacc[NVAR], qacc[NVAR], qc[NVAR];
        cons_of_prim(w, qc, p.gamma);
#pragma unroll
        for (int v = 0; v < NVAR; ++v) { wacc[v] = w[v]; qacc[v] = 0.0; }
        for (int ax = 0; ax < 2; ++ax) {
            double wu[NVAR], wd[NVAR], qu[NVAR], qd[NVAR];
            const int64_t iu = ax == 0 ? nb(i + 1, p.n, p.bc) : i, id = ax == 0 ? nb(i - 1, p.n, p.bc) : i;
            const int64_t ju = ax == 1 ? nb(j + 1, p.n, p.bc) : j, jd = ax == 1 ? nb(j - 1, p.n, p.bc) : j;
            point_prim(p, iu, ju, wu);
            point_prim(p, id, jd, wd);
            cons_of_prim(wu, qu, p.gamma);
            cons_of_prim(wd, qd, p.gamma);
#pragma unroll
            for (int v = 0; v < NVAR; ++v) {
                wacc[v] = wacc[v] - c24 * ((wu[v] - w[v]) - (w[v] - wd[v]));
                qacc[v] = qacc[v] + c24 * ((qu[v] - qc[v]) - (qc[v] - qd[v]));
            }
        }
        double qa[NVAR];
        cons_of_prim(wacc, qa, p.gamma);
#pragma unroll
        for (int v = 0; v < NVAR; ++v) q[v] = qa[v] + qacc[v];
    }
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        const int NT = ex.nthreads();
        ex.phase([&](int tid) {
            const int64_t c = (int64_t)bx * NT + tid, r = by;
            if (c >= p.ncol) return;
            const double c24 = 1.0 / 24.0;
            const int64_t i = wrap_index(r + p.x_off, p.n), j = c;
            double q[NVAR], out[NVAR];
            cons_at(p, i, j, q);
#pragma unroll
            for (int v = 0; v < NVAR; ++v) out[v] = q[v];
            for (int ax = 0; ax < 2; ++ax) {
                double qu[NVAR], qd[NVAR];
                cons_at(p, ax == 0 ? nb(i + 1, p.n, p.bc) : i, ax == 1 ? nb(j + 1, p.n, p.bc) : j, qu);
                cons_at(p, ax == 0 ? nb(i - 1, p.n, p.bc) : i, ax == 1 ? nb(j - 1, p.n, p.bc) : j, qd);
#pragma unroll
                for (int v = 0; v < NVAR; ++v) out[v] = out[v] + c24 * ((qu[v] - q[v]) - (q[v] - qd[v]));
            }
#pragma unroll
            for (int v = 0; v < NVAR; ++v) *p.out.at(r, v, c) = out[v];
            if (p.mhd_flag != nullptr && (out[3] != 0.0 || out[5] != 0.0 || out[6] != 0.0 || out[7] != 0.0)) *p.mhd_flag = 1;
        });
    }
};

// ------------------------------------------------------------------------------------------------ arithmetic self-check
// Fast (common.cuh) against the compiler's IEEE division and square root on generated operands.  Operand classes,
// by the low bits of the sample index: ordinary magnitudes (exponents within +-40 of 1), magnitudes across the
// whole of Fast's range (+-400), arbitrary bit patterns (subnormals, huge values, Inf, NaN), and hand-picked
// specials (+-0, +-1, +-Inf, NaN, smallest / largest normals).  counts[0]: operations Fast accepted (ok stayed
// true); counts[1]: of those, results whose bits differ from IEEE (NaN matches NaN); counts[2]: operations Fast
// declined.  counts[1] must be 0; counts[0] must cover the ordinary classes.
struct ArithCheckParams {
    unsigned long long* counts;
    unsigned long long seed;
    int per_thread;
};
struct ArithCheckKernel {
    using Params = ArithCheckParams;
    static constexpr int MAX_THREADS = 256;
    static HD unsigned long long mix(unsigned long long z) {
        z += 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    static HD double operand(unsigned long long bits, int cls) {
        double v;
        if (cls == 2) { memcpy(&v, &bits, 8); return v; }
        if (cls == 3) {
            const double special[12] = {0.0, -0.0, 1.0, -1.0, INFINITY, -INFINITY, NAN, 2.2250738585072014e-308,
                                        1.7976931348623157e308, 4.9e-324, 3.0, 0.1};
            return special[bits % 12];
        }
        const int span = cls == 0 ? 40 : 400;
        const long long e = 1023 + (long long)((bits >> 52) % (2 * span + 1)) - span;
        const unsigned long long b = (bits & 0x800FFFFFFFFFFFFFull) | ((unsigned long long)e << 52);
        memcpy(&v, &b, 8);
        return v;
    }
    static HD bool same(double a, double b) {
        unsigned long long x, y;
        memcpy(&x, &a, 8);
        memcpy(&y, &b, 8);
        return x == y || (a != a && b != b);
    }
    template <class Ex>
    static HD void block(const Params& p, int bx, int, Ex& ex) {
        const int NT = ex.nthreads();
        ex.phase([&](int tid) {
            unsigned long long accepted = 0, wrong = 0, declined = 0;
            unsigned long long state = p.seed + ((unsigned long long)bx * NT + tid) * 0x632BE59BD9B4E019ull;
            for (int i = 0; i < p.per_thread; ++i) {
                const unsigned long long a = mix(state), b = mix(a);
                state = b;
                const int cls = (int)(i & 3), cls_y = (cls == 3) ? (int)((i >> 2) & 3) : cls;
                const double x = operand(a, cls), y = operand(b, cls_y);
                Exact exact;
#ifdef ASTREA_DEVICE_BUILD
                { Fast f; const double r = f.div(x, y); if (f.good()) { ++accepted; wrong += same(r, exact.div(x, y)) ? 0 : 1; } else ++declined; }
                { Fast f; const double r = f.safe_div(x, y); if (f.good()) { ++accepted; wrong += same(r, exact.safe_div(x, y)) ? 0 : 1; } else ++declined; }
                { FastT<false> f; const double r = f.div(x, y); if (f.good()) { ++accepted; wrong += same(r, exact.div(x, y)) ? 0 : 1; } else ++declined; }
                { FastT<false> f; const double r = f.safe_div(x, y); if (f.good()) { ++accepted; wrong += same(r, exact.safe_div(x, y)) ? 0 : 1; } else ++declined; }
                { Fast f; const double r = f.root(x); if (f.good()) { ++accepted; wrong += same(r, exact.root(x)) ? 0 : 1; } else ++declined; }
                { Fast f; const double r = f.root(fabs(y)); if (f.good()) { ++accepted; wrong += same(r, exact.root(fabs(y))) ? 0 : 1; } else ++declined; }
#else
                (void)x; (void)y; (void)exact; accepted += 6;
#endif
            }
#ifdef __CUDA_ARCH__
            atomicAdd(p.counts + 0, accepted); atomicAdd(p.counts + 1, wrong); atomicAdd(p.counts + 2, declined);
#else
            p.counts[0] += accepted; p.counts[1] += wrong; p.counts[2] += declined;
#endif
        });
    }
};

// ------------------------------------------------------------------------------------------------ diagnostics
// Device-side reductions of the post-processing the reference does on its HDF5 snapshots (functions/analytic.py):
//   conservation (:66-77)    sum over the grid of every conservative variable
//   total variation (:48-62) sum of |np.diff along every axis in turn| of the primitive snapshot
// One partial result per block and variable, combined on the host in block order: deterministic.
struct DiagParams {
    Plane q;              // conservative cell averages
    Plane w;              // primitive snapshot (PrimKernel output)
    int64_t nrow, ncol;
    int dimension;
    double* partial;      // [gridDim.y * gridDim.x][2 * NVAR]: sums, then total variations
};
struct DiagKernel {
    using Params = DiagParams;
    static constexpr int MAX_THREADS = 256;
    static size_t smem_bytes() { return sizeof(double) * MAX_THREADS; }
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        const int NT = ex.nthreads();
        double* red = ex.smem();
        const int64_t r = by;
        for (int k = 0; k < 2 * NVAR; ++k) {
            const int v = k % NVAR;
            ex.phase([&](int tid) {
                const int64_t c = (int64_t)bx * NT + tid;
                double x = 0.0;
                if (c < p.ncol) {
                    if (k < NVAR) {
                        x = *p.q.at(r, v, c);
                    } else if (p.dimension == 1) {
                        if (c + 1 < p.ncol) x = fabs(*p.w.at(r, v, c + 1) - *p.w.at(r, v, c));
                    } else if (c + 1 < p.ncol && r + 1 < p.nrow) {
                        // np.diff along axis 0, then along axis 1
                        const double d0 = *p.w.at(r + 1, v, c) - *p.w.at(r, v, c);
                        const double d1 = *p.w.at(r + 1, v, c + 1) - *p.w.at(r, v, c + 1);
                        x = fabs(d1 - d0);
                    }
                }
                red[tid] = x;
            });
            for (int stride = NT / 2; stride > 0; stride >>= 1) {
                ex.phase([&](int tid) { if (tid < stride) red[tid] = red[tid] + red[tid + stride]; });
            }
            ex.phase([&](int tid) {
                if (tid == 0) p.partial[((int64_t)by * ((p.ncol + NT - 1) / NT) + bx) * 2 * NVAR + k] = red[0];
            });
        }
    }
};

// ------------------------------------------------------------------------------------------------ solution error
// functions/analytic.py:24-44 calculate_solution_error: per-channel norm of (numerical - theoretical) primitive cell
// averages over the grid, for the 8 primitive variables plus the specific total energy E_tot = e(P -> e_tot) / rho and
// the specific internal energy E_int = P / (rho (gamma - 1)) (:33-37).  One partial result per block and channel,
// combined on the host in block order (deterministic; the summation order differs from numpy's pairwise sum).
constexpr int ERR_CHANNELS = NVAR + 2;
struct ErrorParams {
    Plane num, theo;      // primitive snapshots: of the current grid (PrimKernel), of the initial state (uploaded)
    int64_t nrow, ncol;
    double gamma, norm;   // norm > 10: maximum; norm <= 0: plain sum of |d|; else sum of |d|^norm
    double* partial;      // [gridDim.y * gridDim.x][ERR_CHANNELS]
};
struct ErrorKernel {
    using Params = ErrorParams;
    static constexpr int MAX_THREADS = 256;
    static size_t smem_bytes() { return sizeof(double) * MAX_THREADS; }
    static HD void channels(const double* w, double gamma, double* out) {
#pragma unroll
        for (int v = 0; v < NVAR; ++v) out[v] = w[v];
        const double e = w[4] / (gamma - 1.0) + 0.5 * (w[0] * norm3sq(w[1], w[2], w[3]) + norm3sq(w[5], w[6], w[7]));   // fv.py:50
        out[NVAR] = sdiv(e, w[0]);
        out[NVAR + 1] = sdiv(w[4], w[0] * (gamma - 1.0));
    }
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        const int NT = ex.nthreads();
        double* red = ex.smem();
        const int64_t r = by;
        const bool use_max = p.norm > 10.0;
        for (int k = 0; k < ERR_CHANNELS; ++k) {
            ex.phase([&](int tid) {
                const int64_t c = (int64_t)bx * NT + tid;
                double x = 0.0;
                if (c < p.ncol) {
                    double a[NVAR], b[NVAR], ca[ERR_CHANNELS], cb[ERR_CHANNELS];
#pragma unroll
                    for (int v = 0; v < NVAR; ++v) { a[v] = *p.num.at(r, v, c); b[v] = *p.theo.at(r, v, c); }
                    channels(a, p.gamma, ca);
                    channels(b, p.gamma, cb);
                    const double d = fabs(ca[k] - cb[k]);
                    x = (use_max || p.norm <= 0.0 || p.norm == 1.0) ? d : (p.norm == 2.0 ? d * d : pow(d, p.norm));
                }
                red[tid] = x;
            });
            for (int stride = NT / 2; stride > 0; stride >>= 1) {
                ex.phase([&](int tid) {
                    if (tid < stride) red[tid] = use_max ? npmax(red[tid], red[tid + stride]) : red[tid] + red[tid + stride];
                });
            }
            ex.phase([&](int tid) {
                if (tid == 0) p.partial[((int64_t)by * ((p.ncol + NT - 1) / NT) + bx) * ERR_CHANNELS + k] = red[0];
            });
        }
    }
};

}  // namespace astrea
