// 2D spatial operator for one sweep direction as three small, high-occupancy kernels.
//
// Frame: rows of every plane run along the sweep direction (s), columns along the transverse direction (t); the
// x sweep reads the state as stored, the y sweep its transposed copy (the reference's ``grid.transpose(axes)``).
//
//   PrimStage   q  -> wS       primitive cell averages: pointwise (PCM/PLM, fv.py:97-101) or 4th-order
//                              (PPM/WENO, fv.py:126-143).  Thread per cell, 34x10 shared-memory tile so that each
//                              pointwise conversion is evaluated once.
//   ReconStage  wS -> w+, w-   reconstruction + limiter (schemes/*.py, limiters.py).  Thread per (column, variable)
//                              marching along the sweep with the stencil in a rotating register window (the march is
//                              unrolled by the window length): one load and two stores per cell and variable, no
//                              block barrier.  The rows ahead arrive through the TMA engine's bulk copies into a per-warp
//                              shared-memory ring (PLM, PPM: BULK) or through a register prefetch queue (WENO); outputs
//                              wanted in the other frame by constrained transport leave through a shared-memory tile
//                              (STAGED).
//   FluxStage   w+- -> F       face conversion (fv.py:105-122 'face'), physical fluxes (constructor.py:113-125),
//                              averaged-state wave speed (fv.py:157-169), Riemann flux of the face averages and of
//                              the face-centred states, F = F_c - d2_t(F_avg)/24 (solvers.py:44-57, fv.py:147-153).
//                              One warp per 32 transverse points of one interface row — or, on wide grids, one block
//                              per row of NT points (BTILE); transverse neighbours are exchanged through shared-memory
//                              slots (four-variable hydro states) or by warp shuffle (eight-variable states),
//                              everything else lives in registers.  Work whose result nobody reads is left out: the
//                              wave speed in the operators after the first of a step (need_speed), and for HLLC on
//                              four-variable states the conversions of the side the waves of a row do not pick (LAZY).
//
// Every division and square root goes through a guard (common.cuh): the kernels run a branch-free pass first and a
// warp (block, for the tiled primitive stages) repeats its work with the compiler's IEEE routines if an operand was
// outside the range the branch-free sequences cover.  Results never depend on which pass produced them.
//
// The flux difference and the Runge-Kutta update follow in aux_kernels.cuh.  HBM traffic per cell and sweep:
// 64 B (q) + 64 B (wS) + 64 B + 128 B (w+-) + 128 B + 64 B (F); see DESIGN.md for the roofline discussion (the
// path is bound by the fp64 pipe, not by HBM).
#pragma once
#include <type_traits>
#include "physics.cuh"
#include "recon.cuh"
#include "riemann.cuh"
#include "runtime.cuh"

namespace astrea {

// ------------------------------------------------------------------------------------------------ PrimStage
struct PrimStageParams {
    Plane q, w;
    int64_t r_lo, r_hi, c_lo, c_hi;   // half-open output range (may reach into the ghost region)
    int64_t r_min, r_max, c_min, c_max;   // allocated index range of the planes (inclusive) for clamped halo loads
    double gamma;
    int high_order;
};
template <bool HYDRO>
struct PrimStage {
    using Params = PrimStageParams;
    using VS = VarSet<HYDRO>;
    static constexpr int MAX_THREADS = 256;
    static constexpr int TX = 32, TY = 8, SX = TX + 2, SY = TY + 2;
    static size_t smem_bytes(int high_order) { return high_order ? sizeof(double) * 2 * NVAR * SX * SY : 0; }
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
#ifdef ASTREA_TWO_PASS
        FirstGuardPlain first;                      // see FluxStage::block; here the unit of repetition is the block
        body(p, bx, by, ex, first);
        if (!ex.block_any(!first.good())) return;
#endif
        Exact exact;
        body(p, bx, by, ex, exact);
    }
    template <class Ex, class G>
    static HD void body(const Params& p, int bx, int by, Ex& ex, G& g) {
        const int64_t c0 = p.c_lo + (int64_t)bx * TX, r0 = p.r_lo + (int64_t)by * TY;
        const double gamma = p.gamma;
        if (!p.high_order) {
            ex.phase([&](int tid) {
                const int64_t c = c0 + tid % TX, r = r0 + tid / TX;
                if (c >= p.c_hi || r >= p.r_hi) return;
                double q[NVAR], w[NVAR];
#pragma unroll
                for (int k = 0; k < VS::N; ++k) { const int v = VS::at(k); q[v] = *p.q.at(r, v, c); }
                prim_of_cons_t<HYDRO>(q, w, gamma, g);
#pragma unroll
                for (int k = 0; k < VS::N; ++k) { const int v = VS::at(k); *p.w.at(r, v, c) = w[v]; }
            });
            return;
        }
        double* Q = ex.smem();                 // [NVAR][SY][SX] conservative averages of the tile + 1 halo
        double* W = Q + NVAR * SX * SY;        // pointwise primitives of the same cells
        ex.phase([&](int tid) {
            for (int e = tid; e < SX * SY; e += MAX_THREADS) {
                const int x = e % SX, y = e / SX;
                const int64_t c = clamp_index(c0 - 1 + x, p.c_min, p.c_max), r = clamp_index(r0 - 1 + y, p.r_min, p.r_max);
                double q[NVAR], w[NVAR];
#pragma unroll
                for (int k = 0; k < VS::N; ++k) { const int v = VS::at(k); q[v] = *p.q.at(r, v, c); }
                prim_of_cons_t<HYDRO>(q, w, gamma, g);
#pragma unroll
                for (int k = 0; k < VS::N; ++k) { const int v = VS::at(k); Q[(v * SY + y) * SX + x] = q[v]; W[(v * SY + y) * SX + x] = w[v]; }
            }
        });
        ex.phase([&](int tid) {
            const int x = tid % TX + 1, y = tid / TX + 1;
            const int64_t c = c0 + x - 1, r = r0 + y - 1;
            if (c >= p.c_hi || r >= p.r_hi) return;
            const double c24 = 1.0 / 24.0;
            double qa[NVAR], w[NVAR], ws[NVAR];
            // fv.py:134-142: axis 0 of the sweep frame first, then the transverse axis
#pragma unroll
            for (int k = 0; k < VS::N; ++k) {
                const int v = VS::at(k);
                const double* q = Q + v * SY * SX;
                const double* w0 = W + v * SY * SX;
                const double qc = q[y * SX + x], wc = w0[y * SX + x];
                double a = qc - c24 * ((q[(y + 1) * SX + x] - qc) - (qc - q[(y - 1) * SX + x]));
                double s = c24 * ((w0[(y + 1) * SX + x] - wc) - (wc - w0[(y - 1) * SX + x]));
                a = a - c24 * ((q[y * SX + x + 1] - qc) - (qc - q[y * SX + x - 1]));
                s = s + c24 * ((w0[y * SX + x + 1] - wc) - (wc - w0[y * SX + x - 1]));
                qa[v] = a;
                ws[v] = s;
            }
            prim_of_cons_t<HYDRO>(qa, w, gamma, g);
#pragma unroll
            for (int k = 0; k < VS::N; ++k) { const int v = VS::at(k); *p.w.at(r, v, c) = w[v] + ws[v]; }
        });
    }
};

// ------------------------------------------------------------------------------------------------ PrimBothStage
// The primitive cell averages of BOTH sweep frames from one read of q: the pointwise conversions are shared, the
// 4th-order correction is summed in each frame's own axis order (fv.py:134-142 sums axis 0 of the frame first, so
// the two frames differ in the last bits), and the y-frame result is written transposed through shared memory.
// Replaces transpose(q) + two PrimStage launches for every scheme whose flux stage does not read q itself (all but PCM).
struct PrimBothParams {
    Plane q;                           // x frame
    Plane wx, wy;                      // out: primitive averages in the x frame / in the y frame (transposed)
    int64_t r_lo, r_hi, c_lo, c_hi;    // half-open output range in x-frame coordinates (rows = x, cols = y)
    int64_t r_min, r_max, c_min, c_max;
    double gamma;
    int high_order;
};
template <bool HYDRO>
struct PrimBothStage {
    using Params = PrimBothParams;
    using VS = VarSet<HYDRO>;
    static constexpr int MAX_THREADS = 256;
#ifndef ASTREA_PRIM_TY
#define ASTREA_PRIM_TY 16
#endif
#ifndef ASTREA_PRIM_MIN_BLOCKS
#define ASTREA_PRIM_MIN_BLOCKS 0
#endif
    static constexpr int MIN_BLOCKS = ASTREA_PRIM_MIN_BLOCKS;      // 0: let ptxas choose
    static constexpr int TX = 32, TY = ASTREA_PRIM_TY, SX = TX + 2, SY = TY + 2, PT = TY + 1;
    static size_t smem_bytes() { return sizeof(double) * VS::N * (2 * SX * SY + TX * PT); }
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
#ifdef ASTREA_TWO_PASS
        FirstGuardPlain first;                      // see FluxStage::block; here the unit of repetition is the block
        body(p, bx, by, ex, first);
        if (!ex.block_any(!first.good())) return;
#endif
        Exact exact;
        body(p, bx, by, ex, exact);
    }
    template <class Ex, class G>
    static HD void body(const Params& p, int bx, int by, Ex& ex, G& g) {
        const int64_t c0 = p.c_lo + (int64_t)bx * TX, r0 = p.r_lo + (int64_t)by * TY;
        const double gamma = p.gamma, c24 = 1.0 / 24.0;
        double* Q = ex.smem();                     // [N][SY][SX] conservative averages of the tile + 1 halo
        double* W = Q + VS::N * SX * SY;           // pointwise primitives of the same cells
        double* T = W + VS::N * SX * SY;           // [N][TX][PT] y-frame result, staged for the transposed write
        ex.phase([&](int tid) {
            // the tile (with its ring) is PER cells per thread: all their loads are issued before the first conversion, so
            // that a thread has PER x N loads in flight instead of N (the stage is bound by load latency and its barriers)
            constexpr int PER = (SX * SY + MAX_THREADS - 1) / MAX_THREADS;
            double qs[PER][VS::N];
#pragma unroll
            for (int m = 0; m < PER; ++m) {
                const int e = tid + m * MAX_THREADS;
                const int x = e % SX, y = e / SX;
                const int64_t c = clamp_index(c0 - 1 + x, p.c_min, p.c_max), r = clamp_index(r0 - 1 + (y < SY ? y : SY - 1), p.r_min, p.r_max);
#pragma unroll
                for (int k = 0; k < VS::N; ++k) qs[m][k] = *p.q.at(r, VS::at(k), c);
            }
#pragma unroll
            for (int m = 0; m < PER; ++m) {
                const int e = tid + m * MAX_THREADS;
                if (e >= SX * SY) continue;
                const int x = e % SX, y = e / SX;
                double q[NVAR], w[NVAR];
#pragma unroll
                for (int k = 0; k < VS::N; ++k) q[VS::at(k)] = qs[m][k];
                prim_of_cons_t<HYDRO>(q, w, gamma, g);
#pragma unroll
                for (int k = 0; k < VS::N; ++k) { Q[(k * SY + y) * SX + x] = q[VS::at(k)]; W[(k * SY + y) * SX + x] = w[VS::at(k)]; }
            }
        });
        ex.phase([&](int tid) {
            const int x = tid % TX + 1;
            for (int y = tid / TX + 1; y <= TY; y += MAX_THREADS / TX) {
                const int64_t c = c0 + x - 1, r = r0 + y - 1;
                const bool inside = c < p.c_hi && r < p.r_hi;
                double wx[NVAR], wy[NVAR];
                if (!p.high_order) {
#pragma unroll
                    for (int k = 0; k < VS::N; ++k) { wx[VS::at(k)] = W[(k * SY + y) * SX + x]; wy[VS::at(k)] = wx[VS::at(k)]; }
                } else {
                    double qx[NVAR], qy[NVAR], sx[NVAR], sy[NVAR];
#pragma unroll
                    for (int k = 0; k < VS::N; ++k) {
                        const int v = VS::at(k);
                        const double* q = Q + k * SY * SX;
                        const double* w0 = W + k * SY * SX;
                        const double qc = q[y * SX + x], wc = w0[y * SX + x];
                        const double dq_r = (q[(y + 1) * SX + x] - qc) - (qc - q[(y - 1) * SX + x]);     // along x (rows)
                        const double dq_c = (q[y * SX + x + 1] - qc) - (qc - q[y * SX + x - 1]);         // along y (cols)
                        const double dw_r = (w0[(y + 1) * SX + x] - wc) - (wc - w0[(y - 1) * SX + x]);
                        const double dw_c = (w0[y * SX + x + 1] - wc) - (wc - w0[y * SX + x - 1]);
                        qx[v] = (qc - c24 * dq_r) - c24 * dq_c;       // x frame: axis 0 = x first
                        sx[v] = c24 * dw_r + c24 * dw_c;
                        qy[v] = (qc - c24 * dq_c) - c24 * dq_r;       // y frame: axis 0 = y first
                        sy[v] = c24 * dw_c + c24 * dw_r;
                    }
                    prim_of_cons_t<HYDRO>(qx, wx, gamma, g);
                    prim_of_cons_t<HYDRO>(qy, wy, gamma, g);
#pragma unroll
                    for (int k = 0; k < VS::N; ++k) { const int v = VS::at(k); wx[v] = wx[v] + sx[v]; wy[v] = wy[v] + sy[v]; }
                }
#pragma unroll
                for (int k = 0; k < VS::N; ++k) {
                    if (inside) *p.wx.at(r, VS::at(k), c) = wx[VS::at(k)];
                    T[(k * TX + (x - 1)) * PT + (y - 1)] = wy[VS::at(k)];
                }
            }
        });
        ex.phase([&](int tid) {
            // y frame: row = y (this tile's columns), col = x (this tile's rows): TY contiguous doubles per row
            const int xx = tid % TY;
            for (int yy = tid / TY; yy < TX; yy += MAX_THREADS / TY) {
                const int64_t c = c0 + yy, r = r0 + xx;
                if (c >= p.c_hi || r >= p.r_hi) continue;
#pragma unroll
                for (int k = 0; k < VS::N; ++k) *p.wy.at(c, VS::at(k), r) = T[(k * TX + yy) * PT + xx];
            }
        });
    }
};

// ------------------------------------------------------------------------------------------------ ReconStage
struct ReconStageParams {
    Plane w;                   // primitive cell averages, valid on rows [-(LO+2) .. ns+HI+1] where data are genuine
    Plane wp, wm;              // out: state on the right / left of interface j (rows 0 .. ns [+1])
    Plane wf;                  // out (optional, base == nullptr to skip): face state handed to constrained transport
    int64_t ns;                // local cells along the sweep
    int64_t ns_glob, s_off;    // global extent and offset of local row 0 (for the 'edge' clamp)
    int64_t c_lo, c_hi;        // half-open column range to process
    int64_t i_lo, i_hi;        // inclusive range of cells to reconstruct (local indices)
    int bc, limiter, seg;
    int cell_aligned;          // 1: wp / wm receive wL / wR of cell i at row i (transverse reconstruction of mag_field.py)
    int nvar;                  // variables to reconstruct (8, or 4 for a hydro state) and their indices
    int vars[NVAR];
    // PPM authors 'c' / 'ph' (recon.cuh): pass 1 / 2 only evaluate the grid-wide switches into ppm_flags[0..2]
    int ppm_author, pass, force_any3;
    int* ppm_flags;
    int64_t nt;                // interior columns (the switches look at genuine cells only)
    int bulk;                  // 1: march fed by bulk asynchronous copies (device, aligned row segments; ReconStage BULK)
    // Outputs written in the OTHER frame, element (row = column t, col = cell i): a marching thread then writes consecutive
    // addresses cell after cell, which the L2 merges into full sectors, so the transposed copy costs no extra pass.
    // Constrained transport reads the face states of a sweep along the transverse direction (mag_field.py:15) and the
    // corner states of both bundles in one frame (:164-185).
    int wf_t;                  // wf is written transposed
    int out_t;                 // cell_aligned: wp / wm are written transposed
};

// accessor over the register stencil: logical offset k relative to the cell, identity boundary map.  The window is a
// circular buffer of NW registers; ROT is the compile-time rotation of the current step of the unrolled march, so
// every index folds to a constant and advancing by one cell moves no register.
template <int LO, int NW, int ROT>
struct StencilAccessor {
    const double* r;
    HD double s(int64_t k) const { return r[(k + LO + ROT) % NW]; }
    HD int64_t b(int64_t k) const { return k; }
};
template <int B, int E, class F>
HD void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}
// accessor over a plane column with the clamp of an 'edge' boundary (rows near the physical boundary only)
struct ColumnAccessor {
    const double* col;         // address of (row 0, var, column)
    int64_t row_pitch, lo_glob, hi_glob, off;
    HD double s(int64_t k) const { return col[k * row_pitch]; }
    HD int64_t b(int64_t k) const { return clamp_index(k + off, lo_glob, hi_glob) - off; }
};

// CPH: the PPM authors 'c' / 'ph' (a separate instantiation keeps the default 'mc' march lean)
// BULK (device only): the rows ahead of the march are brought into a per-warp shared-memory ring by the TMA engine's
// bulk copies (cp.async.bulk, one 256-byte row segment per copy, NW rows per mbarrier) instead of being requested
// PF cells ahead into registers: a warp then has 2-3 x NW rows in flight whatever its register budget, and the march
// reads them with LDS.  Needs 16-byte aligned row segments (even column pitch, even first column): the launcher checks.
// STAGED (device only): a launch with transposed outputs (constrained transport); a separate instantiation keeps the
// staging logic out of the marches of every other configuration (it cost them 8-10 % when it was a run-time switch).
template <int SCHEME, bool CPH = false, bool BULK = false, bool STAGED = false>
struct ReconStage {
    using Params = ReconStageParams;
    static constexpr int MAX_THREADS = 128;
#ifndef ASTREA_RECON_MIN_BLOCKS
#define ASTREA_RECON_MIN_BLOCKS 4
#endif
    // register-prefetch march: 4 blocks (16 warps, 128 registers, no spills).  Round 1 measured 5 blocks best (PPM 2048^2:
    // 4 / 5 / 6 blocks 0.99 / 0.91 / 1.11 ms per step); with the branch-free limiters and the unrolled window of this round
    // WENO5 4096^2 takes 4.25 / 4.52 / 4.40 ms at 4 / 5 / 6 blocks.
#ifndef ASTREA_RECON_MIN_BLOCKS_BULK
#define ASTREA_RECON_MIN_BLOCKS_BULK 4
#endif
    // with the bulk-copy ring the rows in flight no longer depend on the number of warps: 4 blocks (128 registers, no
    // spills) beat 5 (PPM 8192^2: 11.1 vs 12.5 ms per step; the register-prefetch march takes 12.0 at 5 blocks)
    static constexpr int MIN_BLOCKS = BULK ? ASTREA_RECON_MIN_BLOCKS_BULK : ASTREA_RECON_MIN_BLOCKS;
    static constexpr int LO = recon_lo(SCHEME), HI = recon_hi(SCHEME), NW = LO + HI + 1;
#ifndef ASTREA_RECON_PREFETCH
#define ASTREA_RECON_PREFETCH 3
#endif
    static constexpr int PF = ASTREA_RECON_PREFETCH;     // rows in flight beyond the one the next cell needs
#ifndef ASTREA_RECON_BULK_SLOTS
#define ASTREA_RECON_BULK_SLOTS 3
#endif
    static constexpr int NSLOT = ASTREA_RECON_BULK_SLOTS; // BULK: groups of NW rows in the ring of a warp
    // shared memory of a block, in doubles: [mbarriers + rings of the warps (BULK)] [staging tiles of the warps, only
    // for launches with transposed outputs].  A staging tile holds SD cells x 32 columns (+ 1 pad) per output array: the
    // cells of GROUPS unrolled groups, at least 8, so that a flush writes runs of >= 64 bytes per row of the other frame.
    static constexpr int GROUPS = NW >= 8 ? 1 : (NW >= 4 ? 2 : 4);
    static constexpr int SD = GROUPS * NW;
    static constexpr int STG = 2 * SD * 33;      // per warp: wf alone, or wp and wm
    static HD size_t ring_doubles(int nthreads) {
        if (!BULK) return 0;
        const size_t nwarp = nthreads / 32;
        return (nwarp * NSLOT * sizeof(uint64_t) + 15) / 16 * 2 + nwarp * NSLOT * NW * 32;
    }
    static size_t smem_bytes(int nthreads) {
#ifdef ASTREA_DEVICE_BUILD
        return sizeof(double) * (ring_doubles(nthreads) + (STAGED ? (size_t)(nthreads / 32) * STG : 0));
#else
        (void)nthreads;
        return 0;
#endif
    }
    static HD bool staged_outputs(const Params& p) {
        return (p.wf_t != 0 && p.wf.base != nullptr && !p.cell_aligned) || (p.cell_aligned && p.out_t != 0);
    }
    // how far the limiter of a cell can reach through nested boundary maps (recon.cuh): stay on the generic path there
    static constexpr int REACH = HI + 2;
    // The threads are independent; on the device a warp marches with the Fast division first and repeats its
    // segment with Exact if any lane met an operand Fast does not cover (see FluxStage::block, common.cuh).  The
    // authors 'c' / 'ph' write grid-wide switches on the way and stay with Exact.
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
#ifdef ASTREA_TWO_PASS
        if constexpr (!CPH && SCHEME != SCH_PCM) {
            // PLM / PPM divide by way of their limiters, whose numerators are -0 routinely: sign fix on.  The WENO weights
            // g / (beta + eps)^2 and their normalisation have positive numerators: no sign fix (a -0 would be flagged).
            std::conditional_t<(SCHEME >= SCH_WENO3), FirstGuardPlain, FirstGuard> first;
            march(p, bx, by, ex, first);
            if (!ex.warp_any(!first.good())) return;
        }
#endif
        Exact exact;
        march(p, bx, by, ex, exact);
    }
    template <class Ex, class G>
    static HD void march(const Params& p, int bx, int by, Ex& ex, G& g) {
        const int NT = ex.nthreads();
        ex.wphase([&](int tid) {
            const int64_t t_lane = p.c_lo + (int64_t)bx * NT + tid;
#ifdef ASTREA_DEVICE_BUILD
            // Transposed outputs (wf_t / out_t) leave the warp through a shared-memory tile, NW cells at a time, so that every
            // store instruction writes full 32-byte sectors of the other frame's rows (a lane writing its own column cell by cell
            // touches 32 sectors per instruction and was measured 2.5 x slower).  All lanes of a warp that has any column to do
            // stay for the flush; a lane beyond the range repeats the last column and stores nothing itself.
            constexpr bool staged = STAGED;
            const int64_t t_warp = t_lane - (tid & 31);
            if (staged ? t_warp >= p.c_hi : t_lane >= p.c_hi) return;
            const bool alive = !staged || t_lane < p.c_hi;
            const int64_t t = alive ? t_lane : p.c_hi - 1;
            double* stg = ex.smem() + ring_doubles(NT) + (size_t)(tid >> 5) * STG;
#else
            constexpr bool staged = false, alive = true;
            if (t_lane >= p.c_hi) return;
            const int64_t t = t_lane;
#endif
            const int v = p.vars[by % p.nvar];
            const int64_t first = p.i_lo + (int64_t)(by / p.nvar) * p.seg;
            int64_t last = first + p.seg - 1;
            if (last > p.i_hi) last = p.i_hi;
            if (first > last) return;
            const double* col = p.w.at(0, v, t);
            const int64_t rp = p.w.row_pitch;
            const bool edge = p.bc == BC_EDGE;
            // what to do with the faces of cell i (ig: its global index); in flag passes of the authors 'c' / 'ph' nothing is stored
            // transposed outputs exist in the STAGED instantiations only (device; the launcher picks them), so that the other
            // marches keep compile-time strides; the host simulation, one instantiation for both, decides at run time
#ifdef ASTREA_DEVICE_BUILD
            const bool wf_t = STAGED && p.wf_t != 0, out_t = STAGED && p.out_t != 0;
#else
            const bool wf_t = p.wf_t != 0, out_t = p.out_t != 0;
#endif
            auto face_slot = [&](int64_t i) -> double* { return wf_t ? p.wf.at(t, v, i) : p.wf.at(i, v, t); };
            auto store = [&](int64_t i, int64_t ig, double wl, double wr, double wf) {
                if (!alive) return;
                if (p.cell_aligned) {
                    *(out_t ? p.wp.at(t, v, i) : p.wp.at(i, v, t)) = wl;
                    *(out_t ? p.wm.at(t, v, i) : p.wm.at(i, v, t)) = wr;
                    return;
                }
                // w_plus[j] = wL[b(j)], w_minus[j] = wR[b(j-1)]  (plm.py:42, ppm.py:82, weno.py:171)
                *p.wp.at(i, v, t) = wl;
                *p.wm.at(i + 1, v, t) = wr;
                if (edge && ig == 0) *p.wm.at(i, v, t) = wr;                      // j = 0 sees cell 0 on both sides
                if (edge && ig == p.ns_glob - 1) *p.wp.at(i + 1, v, t) = wl;      // j = N sees cell N-1 on both sides
                if (p.wf.base != nullptr) *face_slot(i) = wf;
            };
            // one cell with accessor ``acc`` centred on offset ``at``; CPH: the authors 'c' / 'ph' with their flag passes
            auto cph_cell = [&](const auto& acc, int64_t at, int64_t i, int64_t ig) {
                const bool interior = i >= 0 && i < p.ns && t >= 0 && t < p.nt;
                if (p.pass != 0 && !interior) return;
                const PpmSwitches sw{p.ppm_flags[0] != 0, p.ppm_flags[1] != 0, p.ppm_flags[2] != 0 || p.force_any3 != 0};
                const bool ph = p.ppm_author == PPM_PH;
                bool pa = false, pb = false, p3 = false;
                double wl, wr, wf;
                cell_faces_ppm_cph(acc, at, ph, sw, p.pass, wl, wr, wf, pa, pb, p3);
                if (p.pass == 1) { if (pa) p.ppm_flags[0] = 1; if (pb) p.ppm_flags[1] = 1; return; }
                if (p.pass == 2) { if (p3) p.ppm_flags[2] = 1; return; }
                if (!alive) return;
                *p.wp.at(i, v, t) = wl;
                *p.wm.at(i + 1, v, t) = wr;
                if (edge && ig == 0) *p.wm.at(i, v, t) = wr;
                if (edge && ig == p.ns_glob - 1) *p.wp.at(i + 1, v, t) = wl;
                if (p.wf.base != nullptr) *face_slot(i) = wf;
            };
            // cells within REACH of a physical 'edge' boundary: generic accessor with the clamp map, straight from memory
            auto near_boundary = [&](int64_t lo_i, int64_t hi_i) {
                for (int64_t i = lo_i; i <= hi_i; ++i) {
                    const int64_t ig = i + p.s_off;
                    if (ig < 0 || ig > p.ns_glob - 1) continue;                  // no such cell
                    ColumnAccessor acc{col, rp, 0, p.ns_glob - 1, p.s_off};
                    if constexpr (SCHEME == SCH_PPM && CPH) {
                        cph_cell(acc, i, i, ig);
                    } else {
                        for (int k = -(LO + 1); k <= HI + 1; ++k) g.input(acc.s(acc.b(i + k)));     // everything the cell can read
                        double wl, wr, wf;
                        cell_faces<SCHEME>(acc, i, p.limiter, wl, wr, wf, g);
                        store(i, ig, wl, wr, wf);
                    }
                }
            };
            // the segment splits into [first, mid_lo) near the lower boundary, [mid_lo, mid_hi] away from both, (mid_hi, last]
            int64_t mid_lo = first, mid_hi = last;
            if (edge) {
                const int64_t a = REACH - p.s_off, b = p.ns_glob - 1 - REACH - p.s_off;     // first / last cell away from the boundaries
                if (mid_lo < a) mid_lo = a;
                if (mid_hi > b) mid_hi = b;
                if (mid_lo > last + 1) mid_lo = last + 1;
                if (mid_hi < mid_lo - 1) mid_hi = mid_lo - 1;
                near_boundary(first, mid_lo - 1);
            }
            if (mid_lo <= mid_hi) {
                double r[NW];
                PpmWindow win;                                   // PPM: second differences / face values carried along the march
#pragma unroll
                for (int k = 0; k < NW; ++k) { r[k] = col[(mid_lo - LO + k) * rp]; g.input(r[k]); }
                if constexpr (SCHEME == SCH_PPM && !CPH) ppm_mc_march_prime<0>(StencilAccessor<LO, NW, 0>{r}, win);
                const bool tr = p.cell_aligned && out_t;
                double* out_p = tr ? p.wp.at(t, v, mid_lo) : p.wp.at(mid_lo, v, t);
                double* out_m = tr ? p.wm.at(t, v, mid_lo) : p.wm.at(p.cell_aligned ? mid_lo : mid_lo + 1, v, t);
#ifdef ASTREA_DEVICE_BUILD
                // on the device face states are wanted by constrained transport only, whose launches are the STAGED ones
                double* out_f = (STAGED && p.wf.base != nullptr && !p.cell_aligned) ? face_slot(mid_lo) : nullptr;
#else
                double* out_f = (p.wf.base != nullptr && !p.cell_aligned) ? face_slot(mid_lo) : nullptr;
#endif
                const int64_t rp_p = tr ? 1 : p.wp.row_pitch, rp_m = tr ? 1 : p.wm.row_pitch, rp_f = wf_t ? 1 : p.wf.row_pitch;
                int pending = 0;                 // staged outputs: cells in the staging tile, first of them is cell flush_i0
                int64_t flush_i0 = mid_lo;
                // one cell of the march: `ahead` is row (i + 1) + HI, which replaces the oldest window entry afterwards
                auto step = [&](auto uc, int64_t i, double ahead) {
                    constexpr int U = decltype(uc)::value;
                    g.input(ahead);
                    StencilAccessor<LO, NW, U> acc{r};
                    if constexpr (SCHEME == SCH_PPM && CPH) {
                        cph_cell(acc, 0, i, i + p.s_off);
                    } else {
                        double wl, wr, wf;
                        if constexpr (SCHEME == SCH_PPM) cell_faces_ppm_mc_march<U>(acc, win, wl, wr, wf, g);
                        else cell_faces<SCHEME>(acc, 0, p.limiter, wl, wr, wf, g);
                        // away from the physical boundaries: w_plus[i] = wL, w_minus[i + 1] = wR (cell aligned: both at i)
#ifdef ASTREA_DEVICE_BUILD
                        if constexpr (staged) {
                            const int lane_ = tid & 31, slot = pending + U;      // cells staged since the last flush
                            if (tr) { stg[slot * 33 + lane_] = wl; stg[(SD + slot) * 33 + lane_] = wr; }
                            else {
                                if (alive) { *out_p = wl; *out_m = wr; }
                                out_p += rp_p;
                                out_m += rp_m;
                                stg[slot * 33 + lane_] = wf;
                            }
                        } else
#endif
                        {
                            *out_p = wl;
                            *out_m = wr;
                            if (out_f != nullptr) { *out_f = wf; out_f += rp_f; }
                            out_p += rp_p;
                            out_m += rp_m;
                        }
                    }
                    r[U % NW] = ahead;       // the oldest entry makes room for row (i + 1) + HI
                };
                // After every unrolled group: once GROUPS groups are staged (or the march ends) the staged cells flush_i0 ..
                // leave the warp — 4 lanes per row of the other frame, 8 rows per instruction, consecutive lanes on consecutive
                // addresses, so that the stores fill whole sectors.
                auto flush = [&](int64_t i0, bool last_group) {
#ifdef ASTREA_DEVICE_BUILD
                    if constexpr (staged && !(SCHEME == SCH_PPM && CPH)) {
                        pending += (int)(mid_hi - i0 + 1 < NW ? mid_hi - i0 + 1 : NW);
                        if (pending + NW <= SD && !last_group) return;
                        warp_sync(0xffffffffu);
                        const int lane_ = tid & 31;
                        const int64_t rows = p.c_hi - t_warp < 32 ? p.c_hi - t_warp : 32;
#pragma unroll
                        for (int pass = 0; pass < 4; ++pass) {
                            const int row = pass * 8 + (lane_ >> 2);
                            if (row >= rows) continue;
                            if (tr) {
                                double* dp = p.wp.at(t_warp + row, v, flush_i0);
                                double* dm = p.wm.at(t_warp + row, v, flush_i0);
                                for (int e = lane_ & 3; e < pending; e += 4) { dp[e] = stg[e * 33 + row]; dm[e] = stg[(SD + e) * 33 + row]; }
                            } else {
                                double* df = p.wf.at(t_warp + row, v, flush_i0);
                                for (int e = lane_ & 3; e < pending; e += 4) df[e] = stg[e * 33 + row];
                            }
                        }
                        warp_sync(0xffffffffu);
                        flush_i0 += pending;
                        pending = 0;
                    }
#else
                    (void)i0; (void)last_group;
#endif
                };
#ifdef ASTREA_DEVICE_BUILD
                if constexpr (BULK) {
                    // ring of NSLOT groups of NW row segments (the 32 columns of this warp) fed by cp.async.bulk; group k
                    // holds the `ahead` rows of the cells mid_lo + k NW .. + NW - 1 and completes on its own mbarrier
                    const int lane = tid & 31, warp = tid >> 5, nwarp = NT / 32;
                    const unsigned mask = warp_mask();
                    uint64_t* bars = reinterpret_cast<uint64_t*>(ex.smem()) + warp * NSLOT;
                    double* ring = ex.smem() + ((size_t)nwarp * NSLOT * sizeof(uint64_t) + 15) / 16 * 2 + (size_t)warp * NSLOT * NW * 32;
                    const int64_t t0 = t_warp;                                       // first column of the warp
                    const int64_t c_end = p.w.col_pitch - GHOST;                      // columns of a plane row: [-GHOST, c_end)
                    const uint32_t rowbytes = (uint32_t)(8 * (c_end - t0 < 32 ? c_end - t0 : 32));
                    const double* src0 = p.w.at(0, v, t0);
                    auto arm = [&](int64_t k) {
                        const int64_t first_i = mid_lo + k * NW;
                        const int64_t n = mid_hi - first_i < NW ? mid_hi - first_i : NW;      // cells first_i .. mid_hi - 1 need a row
                        if (n <= 0) return;
                        uint64_t* b = bars + k % NSLOT;
                        mbar_expect_tx(b, (uint32_t)n * rowbytes);
                        for (int64_t u = 0; u < n; ++u)
                            bulk_load(ring + ((k % NSLOT) * NW + u) * 32, src0 + (first_i + 1 + HI + u) * rp, rowbytes, b);
                    };
                    if (lane == 0) {
#pragma unroll
                        for (int sl = 0; sl < NSLOT; ++sl) mbar_init(bars + sl, 1);
                        mbar_init_fence();
                        for (int k = 0; k < NSLOT; ++k) arm(k);
                    }
                    warp_sync(mask);
                    int64_t k = 0;
                    for (int64_t i0 = mid_lo; i0 <= mid_hi; i0 += NW, ++k) {
                        if (k > 0) {                          // every lane has read the previous group: its slot takes group k - 1 + NSLOT
                            warp_sync(mask);
                            if (lane == 0) arm(k - 1 + NSLOT);
                        }
                        if (i0 < mid_hi) mbar_wait(bars + k % NSLOT, (uint32_t)((k / NSLOT) & 1));
                        const double* grp = ring + (k % NSLOT) * NW * 32 + (t - t0);   // a lane beyond the range repeats the last column
                        if (i0 + NW <= mid_hi) {          // a full group: every cell exists and has a row ahead of it
                            static_for<0, NW>([&](auto uc) { step(uc, i0 + decltype(uc)::value, grp[decltype(uc)::value * 32]); });
                        } else {
                            static_for<0, NW>([&](auto uc) {
                                constexpr int U = decltype(uc)::value;
                                const int64_t i = i0 + U;
                                if (i > mid_hi) return;
                                step(uc, i, i < mid_hi ? grp[U * 32] : 0.0);
                            });
                        }
                        flush(i0, i0 + NW > mid_hi);
                    }
                    warp_sync(mask);
                    if (lane == 0) {
#pragma unroll
                        for (int sl = 0; sl < NSLOT; ++sl) mbar_inval(bars + sl);
                    }
                    warp_sync(mask);
                } else
#endif
                {
                    // rows requested PF + 1 cells before their first use, so that a warp has several loads in flight (the march
                    // is bound by load latency at the few warps its registers allow): queue[k] holds row (i + 1) + HI + k
                    double queue[PF];
#pragma unroll
                    for (int k = 0; k < PF; ++k) queue[k] = (mid_lo + 1 + k <= mid_hi) ? col[(mid_lo + 1 + HI + k) * rp] : 0.0;
                    // running address of the march: one pointer increment per cell instead of an index product
                    const double* in = col + (mid_lo + 1 + HI + PF) * rp;
                    // The march is unrolled by the window length: in step U the stencil value at offset k sits in register
                    // (k + LO + U) mod NW, the oldest one is replaced by the row requested one cell early.
                    for (int64_t i0 = mid_lo; i0 <= mid_hi; i0 += NW) {
                        if (i0 + NW + PF <= mid_hi) {      // a full group whose prefetches all stay inside the segment
                            static_for<0, NW>([&](auto uc) {
                                const double ahead = queue[0];                      // row (i + 1) + HI
#pragma unroll
                                for (int k = 0; k + 1 < PF; ++k) queue[k] = queue[k + 1];
                                queue[PF - 1] = *in;                               // row (i + 1 + PF) + HI
                                in += rp;
                                step(uc, i0 + decltype(uc)::value, ahead);
                            });
                        } else {
                            static_for<0, NW>([&](auto uc) {
                                constexpr int U = decltype(uc)::value;
                                const int64_t i = i0 + U;
                                if (i > mid_hi) return;
                                const double ahead = queue[0];                      // row (i + 1) + HI
#pragma unroll
                                for (int k = 0; k + 1 < PF; ++k) queue[k] = queue[k + 1];
                                queue[PF - 1] = (i + 1 + PF <= mid_hi) ? *in : 0.0;    // row (i + 1 + PF) + HI
                                in += rp;
                                step(uc, i, ahead);
                            });
                        }
                        flush(i0, i0 + NW > mid_hi);
                    }
                }
            }
            if (edge) near_boundary(mid_hi + 1, last);
        });
    }
};

// ------------------------------------------------------------------------------------------------ FluxStage
struct FluxStageParams {
    Plane wp, wm;              // interface states from ReconStage (unused for PCM)
    Plane ws, q;               // primitive / conservative cell averages (PCM faces, HLLD normal field)
    Plane f;                   // out: Riemann flux at interface rows 0 .. ns
    int64_t ns, nt;
    int64_t ns_glob, s_off, nt_glob, t_off;
    double gamma;
    int bc, low_mach;
    unsigned long long* eigmax_bits;
    unsigned long long* flag;
    // Lax-Wendroff (solvers.py:79-88): lw_pass == 1 only searches, for each of the spectrum columns
    // u - c, u, u + c, the first entry of the (N+2, N) array of padded averaged states that is non-zero
    // (lw_keys[k] = 2 * flat index + (entry > 0)); the flux pass ranks the all-zero column among them.
    int lw_pass;
    unsigned long long* lw_keys;
    int block_tile;            // 1: the hydro kernels of LLF / HLLC take the block-wide row of transverse points (FluxStage BT)
    // 1: the interface speeds are wanted (first operator of a step: CFL reduction, astrea.py:70-71).  0: they only serve the
    // reference's non-finite check (fv.py:158) — HLLC / HLLD then skip them for states that cannot give a non-finite speed
    int need_speed;
};

// Entries of both states in [1e-25, 1e25] (density) / [-1e25, 1e25] (the rest): the Roe or mean state and the closed-form
// spectral radius of it are finite whatever the values (square roots of positive numbers, denominators >= 1e-25,
// magnitudes far from overflow: |B|^2 / rho <= 3e125, its square 1e251), so the evaluation can say nothing the
// non-finite check would report.  NaN fails every comparison.
template <bool H>
HD bool speed_surely_finite(const double* a, const double* b) {
    bool ok = a[0] >= 1e-25 && a[0] <= 1e25 && b[0] >= 1e-25 && b[0] <= 1e25;
#pragma unroll
    for (int k = 1; k < VarSet<H>::N; ++k) { const int v = VarSet<H>::at(k); ok = ok && fabs(a[v]) <= 1e25 && fabs(b[v]) <= 1e25; }
    return ok;
}

// KIND: 0 = PCM (faces are the padded cell arrays), 1 = pointwise face conversion (PLM), 2 = 4th-order (PPM/WENO)
// MEAN: arithmetic mean (PLM) instead of the Roe average for the wave-speed state.  AX = physical sweep axis,
// SAX = the solver's axis argument.
// HYDRO: v_z and B are identically zero, only [rho, v_x, v_y, P] are processed (physics.cuh).
// EDGE: the boundary mode ('edge' needs index clamps and the "own value" rule, 'wrap' reads ghost data as is).
// BTILE: the block, not the warp, is the row of transverse points (see BT below)
template <int KIND, int SOLVER, int AX, int SAX, bool HYDRO = false, bool EDGE = true, bool BTILE = false>
struct FluxStage {
    using Params = FluxStageParams;
    using VS = VarSet<HYDRO>;
    static constexpr int MAX_THREADS = 128;
#ifndef ASTREA_FLUX_SMEM_EXCHANGE
#define ASTREA_FLUX_SMEM_EXCHANGE 1
#endif
#ifndef ASTREA_FLUX_PREFETCH_ROWS
#define ASTREA_FLUX_PREFETCH_ROWS 16
#endif
#ifndef ASTREA_FLUX_MIN_BLOCKS
#define ASTREA_FLUX_MIN_BLOCKS 3
#endif
#ifndef ASTREA_FLUX_MIN_BLOCKS_HYDRO
#define ASTREA_FLUX_MIN_BLOCKS_HYDRO 4
#endif
    // 8-variable kernels: 128 threads x 3 blocks = 12 warps per SM at 168 registers (the HLLD kernel spills 1.4 KB per
    // thread at 128 registers; Orszag-Tang 4096^2, flux stages per step at 4 / 3 / 2 blocks: 15.3 / 14.4 / 16.2 ms);
    // hydro kernels: 4 blocks = 16 warps at 128 registers too.  With the branch-free division the compiler interleaves
    // independent chains, so registers buy more than warps: 7 blocks (72 registers, ~70 spill instructions per thread)
    // 2.06 ms, 6 blocks 2.04 ms, 5 blocks 2.01 ms per step of flux stages at 2048^2 (r1l); round 2 (same kernels with the
    // shared-memory exchange): 5 / 4 / 3 blocks 1.85 / 1.78 / 1.94 ms at 2048^2 and 27.5 / 26.8 / 29.1 ms at 8192^2.
    // Walking several interface rows per warp with the next row's loads issued early was tried and lost: the
    // per-thread state then lives across a loop and ptxas spills it (flux stage 2.6 -> 3.5 ms per step).
    // The LAZY instantiations (HLLC, four variables, 4th order; see below) hold one side's q / f instead of two: 5 blocks
    // at 96 registers (80-104 B of spills) against 4 at 116: 8192^2 23.45 vs 23.88 ms of flux stages per step (6 blocks at 80
    // registers, 160 B of spills: 24.0).
#ifndef ASTREA_FLUX_MIN_BLOCKS_LAZY
#define ASTREA_FLUX_MIN_BLOCKS_LAZY 5
#endif
    static constexpr int MIN_BLOCKS = !HYDRO ? ASTREA_FLUX_MIN_BLOCKS
        : ((SOLVER == SOL_HLLC && KIND == 2 && ASTREA_FLUX_SMEM_EXCHANGE != 0) ? ASTREA_FLUX_MIN_BLOCKS_LAZY : ASTREA_FLUX_MIN_BLOCKS_HYDRO);
    static constexpr bool HO = KIND == 2, PCM = KIND == 0;
    static constexpr int H = HO ? 2 : 1;            // halo lanes on each side of a warp
    static constexpr int OWN = 32 - 2 * H;          // transverse points a warp owns
    static constexpr bool LW = SOLVER == SOL_LW;
    static constexpr bool LLF = SOLVER == SOL_LLF || LW;     // Lax-Wendroff has LLF's form with another coefficient
    // transverse exchange through shared memory instead of shuffles (runtime.cuh put / nbr): hydro kernels only, the
    // nine exchanged arrays of a four-variable state are 9 KB per warp (37 KB per block, 5 blocks per SM).  Measured at
    // 2048^2 PPM+HLLC: 2.01 -> 1.89 ms of flux stages per step (-250 of ~2150 static instructions per thread)
    // (the eight-variable kernels keep the shuffles: with 72 slots per warp in shared memory the HLLD flux stages of
    // Orszag-Tang 4096^2 took 17.1 instead of 14.4 ms per step, both at 3 blocks per SM)
    static constexpr bool XS = HYDRO && !LW && (ASTREA_FLUX_SMEM_EXCHANGE != 0);
#ifndef ASTREA_FLUX_BLOCK_TILE
#define ASTREA_FLUX_BLOCK_TILE 1
#endif
    // BT: with the shared-memory exchange the row of transverse points a group of lanes works on is the whole block
    // (one interface row per block, NT - 2H owned points) instead of one warp (32 - 2H): the halo lanes, whose work is
    // redundant, drop from 4 of 32 to 4 of NT (12.5 % -> 3.1 % at 128 threads); the phases then end in a block barrier,
    // which costs more than the lanes save on small grids (measured, PPM + HLLC flux stages per step: 2048^2 1.85 ms
    // with warp rows / 1.97 ms with block rows; 8192^2 28.5 / 27.5 ms), so the launcher picks it for wide grids only.
    static constexpr bool BT = XS && BTILE && (ASTREA_FLUX_BLOCK_TILE != 0);
    static constexpr int NS = 9 * VarSet<HYDRO>::N;
    enum Slot : int { S_WP = 0, S_WM = 1, S_QP = 2, S_QM = 3, S_FP = 4, S_FM = 5, S_AP = 6, S_AM = 7, S_FA = 8 };
    static size_t smem_bytes(int nthreads) {
        return BT ? sizeof(double) * NS * (nthreads + 2) : (XS ? sizeof(double) * (nthreads / 32) * NS * 32 : 0);
    }
    // launch geometry: points along the transverse direction a block owns, interface rows it covers
    static int own_points(int nthreads) { return (BT ? nthreads : 32) - 2 * H; }
    static int rows_per_block(int nthreads) { return BT ? 1 : nthreads / 32; }

    // Per-thread values; a member read by the neighbouring lanes in phase n is never written in phase n.
    struct Tls {
        double wp[NVAR], wm[NVAR];     // face-averaged primitive states            (A)
        double qp[NVAR], qm[NVAR];     // conservative states of wp / wm, pointwise  (A)
        double fp[NVAR], fm[NVAR];     // physical flux of wp / wm                   (A)
        double xp[NVAR], xm[NVAR];     // w - d2_t(w)/24: face-centred primitives    (B)
        double ap[NVAR], am[NVAR];     // face-averaged conservative states          (B)
        double fa[NVAR];               // Riemann flux of the face averages          (B)
        double fc[NVAR];               // Riemann flux of the face-centred states    (C)
        double lam, bn, lam_max;
        int64_t j, t;
        bool live, bad;
        HllcWaves wva, wvc;            // LAZY: waves of the face-averaged (A) and of the face-centred states (B)
    };
    // LAZY (HLLC on four-variable states, 4th order): the waves of both Riemann problems of an interface follow from the
    // primitive states alone and name the one side (plus or minus) whose conservative state and physical flux the solver
    // reads (riemann.cuh, hllc_waves / hllc_side).  The group of lanes that forms a row votes on the sides any of its
    // lanes needs, and cons_of_prim / physical_flux / the face conversion of q are evaluated for those only: one side in
    // smooth flow, both where the lanes of a row disagree.  Same values, fewer of them.
    static constexpr bool LAZY = SOLVER == SOL_HLLC && XS && HO;

    // The warps of a block are independent (warp phases, warp-scope reduction).  On the device the work is done with
    // the branch-free Fast division / square root first; a warp in which any lane met an operand outside the range
    // where that sequence is known to be IEEE (common.cuh) repeats its work with Exact before anything is published.
    // (The audit build of the host simulation runs the same two passes with the counting Audit guard.)
    struct Pub { double lam_max; bool bad; };
    template <class Ex>
    static HD void block(const Params& p, int bx, int by, Ex& ex) {
        typename Ex::template Local<Pub> pub(ex);
        const bool search = LW && p.lw_pass == 1;
        bool done = false;
#ifdef ASTREA_TWO_PASS
        if constexpr (!LW) {
            FirstGuardPlain first;
            body(p, bx, by, ex, pub, first);
            done = !ex.template group_any<BT>(!first.good());
        }
#endif
        if (!done) { Exact exact; body(p, bx, by, ex, pub, exact); }
        if (!search) ex.publish_max([&](int k, double& val, bool& bad) { val = pub[k].lam_max; bad = pub[k].bad; }, p.eigmax_bits, p.flag);
    }

    template <class Ex, class L, class G>
    static HD void body(const Params& p, int bx, int by, Ex& ex, L& pub, G& g) {
        typename Ex::template Local<Tls> tls(ex);
        const int NT = ex.nthreads();
        const int nwarp = BT ? 1 : NT / 32;          // interface rows per block
        const int width = BT ? NT : 32;              // lanes that form one row of transverse points
        const int own = width - 2 * H;
        const double gamma = p.gamma, c24 = 1.0 / 24.0;
        constexpr bool edge = EDGE;      // the launcher picks the instantiation from cfg.boundary
        auto smap = [&](int64_t r) -> int64_t { return edge ? clamp_index(r + p.s_off, 0, p.ns_glob - 1) - p.s_off : r; };
        auto solve = [&](const Tls& st, const double* wp, const double* wm, const double* qp, const double* qm, const double* fp,
                         const double* fm, double* out) {
            if (SOLVER == SOL_HLLC) hllc_flux<SAX, HYDRO>(gamma, p.low_mach != 0, wp, wm, qp, qm, fp, fm, out, g);
            else if (SOLVER == SOL_HLLD) hlld_flux<SAX>(gamma, st.bn, wp, wm, qp, qm, fp, fm, out, g);
            else llf_flux_t<HYDRO>(st.lam, qp, qm, fp, fm, out);
        };
        // transverse second difference of a per-thread array member produced in an earlier phase; the neighbour of
        // a point on a physical 'edge' boundary is the point itself ("pad the derived array")
        auto d2t = [&](int tid, const Tls& st, double own, int slot, auto get) -> double {
            double a = ex.template nbr<XS, NS, BT>(tid, slot, -1, get), b = ex.template nbr<XS, NS, BT>(tid, slot, 1, get);
            if (edge) {
                const int64_t tg = st.t + p.t_off;
                if (tg - 1 < 0) a = own;
                if (tg + 1 > p.nt_glob - 1) b = own;
            }
            return (b - own) - (own - a);
        };
        // Lax-Wendroff: which column np.unique(characteristics, axis=-1)[..., 1] is.  For a state without v_z / B the
        // LAPACK spectrum has the four distinct columns u - c < u < u + c and 0; they are sorted lexicographically
        // over all padded points, so only the place of the zero column has to be found: the number of columns whose
        // first non-zero entry is negative.  0: second = u - c, 1: second = 0, else: second = u.
        int lw_rank = 0;
        if (LW) {
#pragma unroll
            for (int k = 0; k < 3; ++k) lw_rank += (p.lw_keys[k] != ~0ull && (p.lw_keys[k] & 1ull) == 0) ? 1 : 0;
        }
        // LLF: max |eigenvalue|; LW: second^2 / max |eigenvalue| (solvers.py:84-87) of an averaged (or PCM cell) state
        auto dissipation = [&](const double* a) -> double {
            const double lam = spectral_radius_t<AX, HYDRO>(a, gamma, g);
            if (!LW) return lam;
            const double u = a[1 + AX], c = dsqrt(ddiv(gamma * a[4], a[0], g), g);
            const double second = lw_rank == 0 ? u - c : (lw_rank == 1 ? 0.0 : u);
            return sdiv(second * second, lam, g);
        };
        // the state of padded entry `jj` (interface row, or cell row for PCM) at this thread's column (fv.py:157-169)
        auto state_at = [&](int64_t jj, int64_t tc, double* a) {
            double x[NVAR], y[NVAR];
            if (PCM) {
#pragma unroll
                for (int kv = 0; kv < VS::N; ++kv) { const int v = VS::at(kv); a[v] = *p.ws.at(jj, v, tc); }
            } else {
#pragma unroll
                for (int kv = 0; kv < VS::N; ++kv) {
                const int v = VS::at(kv); x[v] = *p.wp.at(jj, v, tc); y[v] = *p.wm.at(jj, v, tc); }
                if (KIND == 1) mean_state_t<HYDRO>(x, y, a); else roe_state_t<HYDRO>(x, y, a, g);
            }
        };
        auto speed_at = [&](int64_t jj, int64_t tc) -> double {
            double a[NVAR];
            state_at(jj, tc, a);
            return dissipation(a);
        };

        if (LW && p.lw_pass == 1) {
            // search pass: rows of the padded array are m = 0 .. N+1; entry m holds interface bc(m) (cells bc(m-1) for PCM)
            ex.publish_min3([&](int tid, unsigned long long* key) {
                key[0] = key[1] = key[2] = ~0ull;
                const int lane_id = tid & 31, w = tid >> 5;
                const int64_t j = (int64_t)by * nwarp + w, t = (int64_t)bx * OWN - H + lane_id;
                const int64_t n = p.ns_glob;
                const int64_t first = PCM ? 0 : 1, last = PCM ? n - 1 : n;     // interfaces 1..N, or cells 0..N-1
                // a slab searches the entries it holds; positions in the array are those of the whole grid
                const int64_t jg = j + p.s_off, tg = t + p.t_off;
                if (lane_id < H || lane_id >= 32 - H || t < 0 || t >= p.nt || j > p.ns || jg < first || jg > last) return;
                double a[NVAR];
                state_at(j, t, a);
                const double u = a[1 + AX], c = dsqrt(ddiv(gamma * a[4], a[0], g), g);
                const double col[3] = {u - c, u, u + c};
                // padded rows this entry appears in
                int64_t rows[3] = {PCM ? jg + 1 : jg, -1, -1};
                const bool wrap = p.bc == BC_WRAP;
                if (jg == (wrap ? last : first)) rows[1] = 0;            // pad in front: np.pad wraps / repeats the edge
                if (jg == (wrap ? first : last)) rows[2] = n + 1;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (!(col[k] != 0.0)) continue;
                    for (int q = 0; q < 3; ++q) {
                        if (rows[q] < 0) continue;
                        const unsigned long long kk = ((unsigned long long)(rows[q] * p.nt_glob + tg) << 1) | (col[k] > 0.0 ? 1ull : 0ull);
                        if (kk < key[k]) key[k] = kk;
                    }
                }
            }, p.lw_keys);
            return;
        }

        // A: load the interface states, pointwise conversions, wave speed
        ex.template xphase<BT>([&](int tid) {
            Tls& st = tls[tid];
            const int lane_id = BT ? tid : (tid & 31), w = BT ? 0 : (tid >> 5);
            st.j = (int64_t)by * nwarp + w;
            st.t = (int64_t)bx * own - H + lane_id;
            st.live = st.j <= p.ns;
            st.lam = 0.0; st.bn = 0.0; st.lam_max = 0.0; st.bad = false;
            pub[tid].lam_max = 0.0; pub[tid].bad = false;
            if (!st.live) return;
            const int64_t j = st.j;
            // columns the earlier stages filled: [-H, nt + H); lanes beyond them (last warp of a row) feed nobody
            const int64_t tc = clamp_index(st.t, -(int64_t)H, p.nt + H - 1);
            if (PCM) {
                const int64_t rp_ = smap(j), rm_ = smap(j - 1);
#pragma unroll
                for (int kv = 0; kv < VS::N; ++kv) {
                const int v = VS::at(kv);
                    st.wp[v] = *p.ws.at(rp_, v, tc); st.wm[v] = *p.ws.at(rm_, v, tc);
                    st.qp[v] = *p.q.at(rp_, v, tc);  st.qm[v] = *p.q.at(rm_, v, tc);
                }
            } else {
#pragma unroll
                for (int kv = 0; kv < VS::N; ++kv) {
                const int v = VS::at(kv); st.wp[v] = *p.wp.at(j, v, tc); st.wm[v] = *p.wm.at(j, v, tc); }
#if defined(ASTREA_DEVICE_BUILD) && ASTREA_FLUX_PREFETCH_ROWS > 0
                // a block starts with these loads and has nothing to overlap them with (long scoreboard is the largest stall
                // of the stage): the blocks that run a few waves later find their interface states in L2
                if (LAZY && j + ASTREA_FLUX_PREFETCH_ROWS <= p.ns) {
#pragma unroll
                    for (int kv = 0; kv < VS::N; ++kv) {
                        const int v = VS::at(kv);
                        prefetch_l2(p.wp.at(j + ASTREA_FLUX_PREFETCH_ROWS, v, tc)); prefetch_l2(p.wm.at(j + ASTREA_FLUX_PREFETCH_ROWS, v, tc));
                    }
                }
#endif
                if constexpr (!LAZY) {
                    cons_of_prim_t<HYDRO>(st.wp, st.qp, gamma, g);
                    cons_of_prim_t<HYDRO>(st.wm, st.qm, gamma, g);
                }
            }
            if constexpr (LAZY) {
                st.wva = hllc_waves<SAX>(gamma, p.low_mach != 0, st.wp, st.wm, g);
            } else {
                physical_flux_t<AX, HYDRO>(st.wp, st.fp, gamma, g);
                physical_flux_t<AX, HYDRO>(st.wm, st.fm, gamma, g);
            }
            if (SOLVER == SOL_HLLD) st.bn = *p.ws.at(smap(j), 5 + SAX, tc);
            // wave speeds: the per-interface estimate feeds the CFL reduction, LLF also uses it as its dissipation
            const int64_t jg = j + p.s_off;
            double lam_here = 0.0;
            bool counts;
            // the later operators of a step use the speed for nothing but the non-finite check: skipped where it cannot fire
            const bool speed = LLF || p.need_speed != 0 || !speed_surely_finite<HYDRO>(st.wp, st.wm);
            if (PCM) {
                // pcm.py:30: Jacobian at the padded cells; interface j sees cells b(j-1) and b(j)
                if (speed) lam_here = spectral_radius_t<AX, HYDRO>(st.wp, gamma, g);
                counts = jg >= 0 && jg < p.ns_glob && j < p.ns;
                if (LLF) st.lam = npmax(dissipation(st.wm), dissipation(st.wp));
            } else if (!speed) {
                counts = jg >= 1 && jg <= p.ns_glob && j >= 1;
            } else {
                double a[NVAR];
                if (KIND == 1) mean_state_t<HYDRO>(st.wp, st.wm, a); else roe_state_t<HYDRO>(st.wp, st.wm, a, g);
                lam_here = spectral_radius_t<AX, HYDRO>(a, gamma, g);
                counts = jg >= 1 && jg <= p.ns_glob && j >= 1;
                if (LLF) {
                    // entries j and j+1 of the pad-1 array of interface speeds (solvers.py:73-74)
                    const int64_t ja = edge ? clamp_index(jg, 1, p.ns_glob) - p.s_off : j;
                    const int64_t jb = edge ? clamp_index(jg + 1, 1, p.ns_glob) - p.s_off : j + 1;
                    const double here = LW ? dissipation(a) : lam_here;
                    const double la = (ja == j) ? here : speed_at(ja, tc);
                    const double lb = (jb == j) ? here : speed_at(jb, tc);
                    st.lam = npmax(la, lb);
                }
            }
            const bool owned = lane_id >= H && lane_id < width - H && st.t >= 0 && st.t < p.nt;
            if (counts && owned) {
                if (lam_here == lam_here && lam_here <= 1.7976931348623157e308) st.lam_max = lam_here; else st.bad = true;
                // Lax-Wendroff with a negative averaged pressure: the reference carries on in complex arithmetic
                // (complex eigenvalues); the device path does not follow it there and reports the step as non-finite
                if (LW && !(st.lam == st.lam)) st.bad = true;
            }
            pub[tid].lam_max = st.lam_max; pub[tid].bad = st.bad;
#pragma unroll
            for (int kv = 0; kv < VS::N; ++kv) {
                const int v = VS::at(kv);
                ex.template put<XS, NS, BT>(S_WP * VS::N + kv, st.wp[v]); ex.template put<XS, NS, BT>(S_WM * VS::N + kv, st.wm[v]);
                if constexpr (!LAZY) {
                    ex.template put<XS, NS, BT>(S_QP * VS::N + kv, st.qp[v]); ex.template put<XS, NS, BT>(S_QM * VS::N + kv, st.qm[v]);
                    ex.template put<XS, NS, BT>(S_FP * VS::N + kv, st.fp[v]); ex.template put<XS, NS, BT>(S_FM * VS::N + kv, st.fm[v]);
                }
            }
        });
        if constexpr (LAZY) { lazy_phases(p, ex, tls, g, d2t); return; }
        // B: w - d2_t(w)/24, face conversion of q (fv.py:105-122), Riemann flux of the face averages
        ex.template xphase<BT>([&](int tid) {
            Tls& st = tls[tid];
            if (!st.live) return;                 // the interface row of a warp: all of its lanes leave together
            double qx[NVAR];
#pragma unroll
            for (int kv = 0; kv < VS::N; ++kv) {
                const int v = VS::at(kv);
                st.xp[v] = st.wp[v] - c24 * d2t(tid, st, st.wp[v], S_WP * VS::N + kv, [&](int k) { return tls[k].wp[v]; });
                st.xm[v] = st.wm[v] - c24 * d2t(tid, st, st.wm[v], S_WM * VS::N + kv, [&](int k) { return tls[k].wm[v]; });
            }
            if (HO) {
                cons_of_prim_t<HYDRO>(st.xp, qx, gamma, g);
#pragma unroll
                for (int kv = 0; kv < VS::N; ++kv) { const int v = VS::at(kv); st.ap[v] = qx[v] + c24 * d2t(tid, st, st.qp[v], S_QP * VS::N + kv, [&](int k) { return tls[k].qp[v]; }); }
                cons_of_prim_t<HYDRO>(st.xm, qx, gamma, g);
#pragma unroll
                for (int kv = 0; kv < VS::N; ++kv) { const int v = VS::at(kv); st.am[v] = qx[v] + c24 * d2t(tid, st, st.qm[v], S_QM * VS::N + kv, [&](int k) { return tls[k].qm[v]; }); }
            } else {
#pragma unroll
                for (int kv = 0; kv < VS::N; ++kv) {
                const int v = VS::at(kv); st.ap[v] = st.qp[v]; st.am[v] = st.qm[v]; }
            }
            solve(st, st.wp, st.wm, st.ap, st.am, st.fp, st.fm, st.fa);
#pragma unroll
            for (int kv = 0; kv < VS::N; ++kv) {
                const int v = VS::at(kv);
                ex.template put<XS, NS, BT>(S_AP * VS::N + kv, st.ap[v]); ex.template put<XS, NS, BT>(S_AM * VS::N + kv, st.am[v]);
                ex.template put<XS, NS, BT>(S_FA * VS::N + kv, st.fa[v]);
            }
        });
        // C: face-centred q and physical flux (solvers.py:47-52), Riemann flux of the centred states
        ex.template xphase<BT>([&](int tid) {
            Tls& st = tls[tid];
            if (!st.live) return;
            double cqp[NVAR], cqm[NVAR], cfp[NVAR], cfm[NVAR];
#pragma unroll
            for (int kv = 0; kv < VS::N; ++kv) {
                const int v = VS::at(kv);
                cqp[v] = st.ap[v] - c24 * d2t(tid, st, st.ap[v], S_AP * VS::N + kv, [&](int k) { return tls[k].ap[v]; });
                cqm[v] = st.am[v] - c24 * d2t(tid, st, st.am[v], S_AM * VS::N + kv, [&](int k) { return tls[k].am[v]; });
                cfp[v] = st.fp[v] - c24 * d2t(tid, st, st.fp[v], S_FP * VS::N + kv, [&](int k) { return tls[k].fp[v]; });
                cfm[v] = st.fm[v] - c24 * d2t(tid, st, st.fm[v], S_FM * VS::N + kv, [&](int k) { return tls[k].fm[v]; });
            }
            // (solving both Riemann problems of the interface here, side by side, was measured slower twice: round 1 at 96
            // registers 2.05 vs 2.01 ms at 2048^2, round 2 at 128 registers 27.6 vs 26.8 ms at 8192^2)
            solve(st, st.xp, st.xm, cqp, cqm, cfp, cfm, st.fc);
        });
        // D: F = F_c - d2_t(F_avg)/24 (fv.py:147-153)
        ex.template xphase<BT>([&](int tid) {
            Tls& st = tls[tid];
            if (!st.live) return;
            const int lane_id = BT ? tid : (tid & 31);
            const bool owned = st.live && lane_id >= H && lane_id < width - H && st.t >= 0 && st.t < p.nt;
#pragma unroll
            for (int kv = 0; kv < VS::N; ++kv) {
                const int v = VS::at(kv);
                const double f = st.fc[v] - c24 * d2t(tid, st, st.fa[v], S_FA * VS::N + kv, [&](int k) { return tls[k].fa[v]; });
                if (owned) *p.f.at(st.j, v, st.t) = f;
            }
        });
    }

    // Phases B .. D of the LAZY instantiations (see LAZY above); phase A has loaded and published wp / wm and holds the waves
    // of the face-averaged problem.
    template <class Ex, class L, class G, class D2>
    static HD void lazy_phases(const Params& p, Ex& ex, L& tls, G& g, D2& d2t) {
        const int NT = ex.nthreads();
        const int width = BT ? NT : 32;
        const double gamma = p.gamma, c24 = 1.0 / 24.0;
        // B: face-centred primitives of both sides, their waves
        ex.template xphase<BT>([&](int tid) {
            Tls& st = tls[tid];
            if (!st.live) return;
#pragma unroll
            for (int kv = 0; kv < VS::N; ++kv) {
                const int v = VS::at(kv);
                st.xp[v] = st.wp[v] - c24 * d2t(tid, st, st.wp[v], S_WP * VS::N + kv, [&](int k) { return tls[k].wp[v]; });
                st.xm[v] = st.wm[v] - c24 * d2t(tid, st, st.wm[v], S_WM * VS::N + kv, [&](int k) { return tls[k].wm[v]; });
            }
            st.wvc = hllc_waves<SAX>(gamma, p.low_mach != 0, st.xp, st.xm, g);
        });
        // the sides any lane of the row reads: plus for sides 0 and 1, minus for side 2
        const bool plus = ex.template group_or<BT>([&](int k) { return tls[k].live && (tls[k].wva.side != 2 || tls[k].wvc.side != 2); });
        const bool minus = ex.template group_or<BT>([&](int k) { return tls[k].live && (tls[k].wva.side == 2 || tls[k].wvc.side == 2); });
        // B': pointwise conversions of the sides in use
        ex.template xphase<BT>([&](int tid) {
            Tls& st = tls[tid];
            if (!st.live) return;
            if (plus) {
                cons_of_prim_t<HYDRO>(st.wp, st.qp, gamma, g);
                physical_flux_t<AX, HYDRO>(st.wp, st.fp, gamma, g);
#pragma unroll
                for (int kv = 0; kv < VS::N; ++kv) {
                    const int v = VS::at(kv);
                    ex.template put<XS, NS, BT>(S_QP * VS::N + kv, st.qp[v]); ex.template put<XS, NS, BT>(S_FP * VS::N + kv, st.fp[v]);
                }
            }
            if (minus) {
                cons_of_prim_t<HYDRO>(st.wm, st.qm, gamma, g);
                physical_flux_t<AX, HYDRO>(st.wm, st.fm, gamma, g);
#pragma unroll
                for (int kv = 0; kv < VS::N; ++kv) {
                    const int v = VS::at(kv);
                    ex.template put<XS, NS, BT>(S_QM * VS::N + kv, st.qm[v]); ex.template put<XS, NS, BT>(S_FM * VS::N + kv, st.fm[v]);
                }
            }
        });
        // B'': face conversion of q (fv.py:105-122), Riemann flux of the face averages
        ex.template xphase<BT>([&](int tid) {
            Tls& st = tls[tid];
            if (!st.live) return;
            double qx[NVAR];
            if (plus) {
                cons_of_prim_t<HYDRO>(st.xp, qx, gamma, g);
#pragma unroll
                for (int kv = 0; kv < VS::N; ++kv) {
                    const int v = VS::at(kv);
                    st.ap[v] = qx[v] + c24 * d2t(tid, st, st.qp[v], S_QP * VS::N + kv, [&](int k) { return tls[k].qp[v]; });
                    ex.template put<XS, NS, BT>(S_AP * VS::N + kv, st.ap[v]);
                }
            }
            if (minus) {
                cons_of_prim_t<HYDRO>(st.xm, qx, gamma, g);
#pragma unroll
                for (int kv = 0; kv < VS::N; ++kv) {
                    const int v = VS::at(kv);
                    st.am[v] = qx[v] + c24 * d2t(tid, st, st.qm[v], S_QM * VS::N + kv, [&](int k) { return tls[k].qm[v]; });
                    ex.template put<XS, NS, BT>(S_AM * VS::N + kv, st.am[v]);
                }
            }
            if (st.wva.side == 2) hllc_side<SAX, HYDRO>(st.wva, st.wm, st.am, st.fm, st.fa, g);
            else hllc_side<SAX, HYDRO>(st.wva, st.wp, st.ap, st.fp, st.fa, g);
#pragma unroll
            for (int kv = 0; kv < VS::N; ++kv) ex.template put<XS, NS, BT>(S_FA * VS::N + kv, st.fa[VS::at(kv)]);
        });
        // C: face-centred q and physical flux of the side the centred waves pick (solvers.py:47-52), its Riemann flux
        ex.template xphase<BT>([&](int tid) {
            Tls& st = tls[tid];
            if (!st.live) return;
            double cq[NVAR], cf[NVAR];
            if (st.wvc.side == 2) {
#pragma unroll
                for (int kv = 0; kv < VS::N; ++kv) {
                    const int v = VS::at(kv);
                    cq[v] = st.am[v] - c24 * d2t(tid, st, st.am[v], S_AM * VS::N + kv, [&](int k) { return tls[k].am[v]; });
                    cf[v] = st.fm[v] - c24 * d2t(tid, st, st.fm[v], S_FM * VS::N + kv, [&](int k) { return tls[k].fm[v]; });
                }
                hllc_side<SAX, HYDRO>(st.wvc, st.xm, cq, cf, st.fc, g);
            } else {
#pragma unroll
                for (int kv = 0; kv < VS::N; ++kv) {
                    const int v = VS::at(kv);
                    cq[v] = 0.0;
                    if (st.wvc.side == 1) cq[v] = st.ap[v] - c24 * d2t(tid, st, st.ap[v], S_AP * VS::N + kv, [&](int k) { return tls[k].ap[v]; });
                    cf[v] = st.fp[v] - c24 * d2t(tid, st, st.fp[v], S_FP * VS::N + kv, [&](int k) { return tls[k].fp[v]; });
                }
                hllc_side<SAX, HYDRO>(st.wvc, st.xp, cq, cf, st.fc, g);
            }
        });
        // D: F = F_c - d2_t(F_avg)/24 (fv.py:147-153)
        ex.template xphase<BT>([&](int tid) {
            Tls& st = tls[tid];
            if (!st.live) return;
            const int lane_id = BT ? tid : (tid & 31);
            const bool owned = lane_id >= H && lane_id < width - H && st.t >= 0 && st.t < p.nt;
#pragma unroll
            for (int kv = 0; kv < VS::N; ++kv) {
                const int v = VS::at(kv);
                const double f = st.fc[v] - c24 * d2t(tid, st, st.fa[v], S_FA * VS::N + kv, [&](int k) { return tls[k].fa[v]; });
                if (owned) *p.f.at(st.j, v, st.t) = f;
            }
        });
    }
};

// launch geometry of the flux stage the dispatcher will pick (dispatch.cuh): transverse points a block owns and
// interface rows it covers.  Mirrors FluxStage::BT: the block is the tile for the hydro kernels of LLF / HLLC.
inline bool flux_stage_block_tile(int solver, int hydro, int wanted) {
    return wanted && hydro && (solver == SOL_LLF || solver == SOL_HLLC) && (ASTREA_FLUX_SMEM_EXCHANGE != 0) && (ASTREA_FLUX_BLOCK_TILE != 0);
}
inline void flux_stage_geometry(int kind, int solver, int hydro, int wanted, int nthreads, int& own, int& rows) {
    const int h = kind == 2 ? 2 : 1;
    const bool bt = flux_stage_block_tile(solver, hydro, wanted);
    own = (bt ? nthreads : 32) - 2 * h;
    rows = bt ? 1 : nthreads / 32;
}

}  // namespace astrea
