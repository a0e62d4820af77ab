// Run-time -> compile-time dispatch of the sweep kernels.  Each reconstruction scheme is instantiated in its
// own translation unit (inst_*.cu) so that the ~80 fused kernels build in parallel.
#pragma once
#include "sweep1d.cuh"
#include "sweep2d.cuh"

namespace astrea {

struct SweepGeometry {
    int nthreads;     // threads per block
    int seg;          // 2D: cells per block along the sweep
};

// return 0 on success, a cudaError_t (> 0) on launch failure, -1 for an unsupported combination
int launch_sweep1d(int scheme, int solver, const Sweep1DParams& p, int nthreads, Stream st);
int launch_sweep2d(int scheme, int solver, int ax, int sax, Sweep2DParams p, int nthreads, Stream st);
// threads -> owned columns for a scheme (2D), tile -> threads (1D)
int sweep2d_owned_columns(int scheme, int nthreads);
int sweep1d_threads(int scheme, int tile);

#define ASTREA_DECLARE_SCHEME(NAME)                                                                   \
    int launch_sweep1d_##NAME(int solver, const Sweep1DParams& p, int nthreads, Stream st);          \
    int launch_sweep2d_##NAME(int solver, int ax, int sax, Sweep2DParams p, int nthreads, Stream st);
ASTREA_DECLARE_SCHEME(pcm)
ASTREA_DECLARE_SCHEME(plm)
ASTREA_DECLARE_SCHEME(ppm)
ASTREA_DECLARE_SCHEME(weno3)
ASTREA_DECLARE_SCHEME(weno5)
ASTREA_DECLARE_SCHEME(weno7)

// Body of one inst_*.cu
#define ASTREA_DEFINE_SCHEME(NAME, SCH)                                                                              \
    template <int SOL>                                                                                               \
    static int run1d_##NAME(const Sweep1DParams& p, int nthreads, Stream st) {                                       \
        using K = Sweep1D<SCH, SOL>;                                                                                 \
        const int gx = (int)((p.n + p.tile - 1) / p.tile);                                                           \
        return launch<K>(p, gx, 1, nthreads, K::smem_bytes(nthreads), st);                                           \
    }                                                                                                                \
    template <int SOL, int AX, int SAX>                                                                              \
    static int run2d_##NAME(Sweep2DParams p, int nthreads, Stream st) {                                              \
        using K = Sweep2D<SCH, SOL, AX, SAX>;                                                                        \
        p.tt = K::owned_for(nthreads);                                                                               \
        const int gx = (int)((p.nt + p.tt - 1) / p.tt), gy = (int)((p.ns + p.seg - 1) / p.seg);                      \
        return launch<K>(p, gx, gy, nthreads, K::smem_bytes(nthreads), st);                                          \
    }                                                                                                                \
    int launch_sweep1d_##NAME(int solver, const Sweep1DParams& p, int nthreads, Stream st) {                         \
        switch (solver) {                                                                                            \
            case SOL_LLF: return run1d_##NAME<SOL_LLF>(p, nthreads, st);                                             \
            case SOL_HLLC: return run1d_##NAME<SOL_HLLC>(p, nthreads, st);                                           \
            case SOL_HLLD: return run1d_##NAME<SOL_HLLD>(p, nthreads, st);                                           \
            default: return -1;                                                                                      \
        }                                                                                                            \
    }                                                                                                                \
    int launch_sweep2d_##NAME(int solver, int ax, int sax, Sweep2DParams p, int nthreads, Stream st) {               \
        const int key = ax * 2 + sax;                                                                                \
        switch (solver) {                                                                                            \
            case SOL_LLF: /* LLF ignores the solver axis (solvers.py:69) */                                          \
                return ax == 0 ? run2d_##NAME<SOL_LLF, 0, 0>(p, nthreads, st) : run2d_##NAME<SOL_LLF, 1, 1>(p, nthreads, st); \
            case SOL_HLLC:                                                                                           \
                switch (key) {                                                                                       \
                    case 0: return run2d_##NAME<SOL_HLLC, 0, 0>(p, nthreads, st);                                    \
                    case 1: return run2d_##NAME<SOL_HLLC, 0, 1>(p, nthreads, st);                                    \
                    case 2: return run2d_##NAME<SOL_HLLC, 1, 0>(p, nthreads, st);                                    \
                    default: return run2d_##NAME<SOL_HLLC, 1, 1>(p, nthreads, st);                                   \
                }                                                                                                    \
            case SOL_HLLD:                                                                                           \
                switch (key) {                                                                                       \
                    case 0: return run2d_##NAME<SOL_HLLD, 0, 0>(p, nthreads, st);                                    \
                    case 1: return run2d_##NAME<SOL_HLLD, 0, 1>(p, nthreads, st);                                    \
                    case 2: return run2d_##NAME<SOL_HLLD, 1, 0>(p, nthreads, st);                                    \
                    default: return run2d_##NAME<SOL_HLLD, 1, 1>(p, nthreads, st);                                   \
                }                                                                                                    \
            default: return -1;                                                                                      \
        }                                                                                                            \
    }

}  // namespace astrea
