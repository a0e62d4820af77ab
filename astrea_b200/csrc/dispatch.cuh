// Run-time -> compile-time dispatch of the templated kernels.  The instantiations are spread over several
// translation units (inst_*.cu) so that they build in parallel.
#pragma once
#include "stages2d.cuh"
#include "sweep1d.cuh"

namespace astrea {

// return 0 on success, a cudaError_t (> 0) on launch failure, -1 for an unsupported combination
int launch_sweep1d(int scheme, int solver, const Sweep1DParams& p, int nthreads, Stream st);
int launch_recon(int scheme, const ReconStageParams& p, int gx, int gy, int nthreads, Stream st);
// kind: 0 = PCM, 1 = pointwise face conversion (PLM), 2 = 4th-order face conversion (PPM / WENO)
// hydro: 1 = the state has no v_z / B (LLF and HLLC only), see physics.cuh
int launch_flux(int kind, int solver, int ax, int sax, int hydro, const FluxStageParams& p, int gx, int gy, int nthreads, Stream st);
int launch_flux_pcm(int solver, int ax, int sax, int hydro, const FluxStageParams& p, int gx, int gy, int nthreads, Stream st);
int launch_flux_plm(int solver, int ax, int sax, int hydro, const FluxStageParams& p, int gx, int gy, int nthreads, Stream st);
int launch_flux_ho(int solver, int ax, int sax, int hydro, const FluxStageParams& p, int gx, int gy, int nthreads, Stream st);

template <class K>
inline int launch_flux_kernel(const FluxStageParams& p, int gx, int gy, int nthreads, Stream st) {
    return launch<K>(p, gx, gy, nthreads, K::smem_bytes(nthreads), st);
}

// Body of one inst_flux_*.cu
#define ASTREA_DEFINE_FLUX(NAME, KIND)                                                                               \
    template <int SOL, int AX, int SAX>                                                                              \
    static int runflux_##NAME(int hydro, const FluxStageParams& p, int gx, int gy, int nthreads, Stream st) {        \
        constexpr bool HY = SOL != SOL_HLLD;                                                                         \
        const bool bt = flux_stage_block_tile(SOL, hydro && HY, p.block_tile);                                       \
        if (p.bc == BC_EDGE) {                                                                                       \
            if (hydro && bt) return launch_flux_kernel<FluxStage<KIND, SOL, AX, SAX, HY, true, HY>>(p, gx, gy, nthreads, st); \
            if (hydro) return launch_flux_kernel<FluxStage<KIND, SOL, AX, SAX, HY, true>>(p, gx, gy, nthreads, st);  \
            return launch_flux_kernel<FluxStage<KIND, SOL, AX, SAX, false, true>>(p, gx, gy, nthreads, st);          \
        }                                                                                                            \
        if (hydro && bt) return launch_flux_kernel<FluxStage<KIND, SOL, AX, SAX, HY, false, HY>>(p, gx, gy, nthreads, st); \
        if (hydro) return launch_flux_kernel<FluxStage<KIND, SOL, AX, SAX, HY, false>>(p, gx, gy, nthreads, st);     \
        return launch_flux_kernel<FluxStage<KIND, SOL, AX, SAX, false, false>>(p, gx, gy, nthreads, st);             \
    }                                                                                                                \
    int launch_flux_##NAME(int solver, int ax, int sax, int hydro, const FluxStageParams& p, int gx, int gy, int nthreads, Stream st) { \
        const int key = ax * 2 + sax;                                                                                \
        switch (solver) {                                                                                            \
            case SOL_LW: /* Lax-Wendroff: states without v_z / B only (the launcher checks), no solver axis */         \
                if (!hydro) return -1;                                                                               \
                if (p.bc == BC_EDGE)                                                                                 \
                    return ax == 0 ? launch_flux_kernel<FluxStage<KIND, SOL_LW, 0, 0, true, true>>(p, gx, gy, nthreads, st)   \
                                   : launch_flux_kernel<FluxStage<KIND, SOL_LW, 1, 1, true, true>>(p, gx, gy, nthreads, st);  \
                return ax == 0 ? launch_flux_kernel<FluxStage<KIND, SOL_LW, 0, 0, true, false>>(p, gx, gy, nthreads, st)      \
                               : launch_flux_kernel<FluxStage<KIND, SOL_LW, 1, 1, true, false>>(p, gx, gy, nthreads, st);     \
            case SOL_LLF: /* LLF ignores the solver axis (solvers.py:69) */                                          \
                return ax == 0 ? runflux_##NAME<SOL_LLF, 0, 0>(hydro, p, gx, gy, nthreads, st)                       \
                               : runflux_##NAME<SOL_LLF, 1, 1>(hydro, p, gx, gy, nthreads, st);                      \
            case SOL_HLLC:                                                                                           \
                switch (key) {                                                                                       \
                    case 0: return runflux_##NAME<SOL_HLLC, 0, 0>(hydro, p, gx, gy, nthreads, st);                   \
                    case 1: return runflux_##NAME<SOL_HLLC, 0, 1>(hydro, p, gx, gy, nthreads, st);                   \
                    case 2: return runflux_##NAME<SOL_HLLC, 1, 0>(hydro, p, gx, gy, nthreads, st);                   \
                    default: return runflux_##NAME<SOL_HLLC, 1, 1>(hydro, p, gx, gy, nthreads, st);                  \
                }                                                                                                    \
            case SOL_HLLD:                                                                                           \
                switch (key) {                                                                                       \
                    case 0: return runflux_##NAME<SOL_HLLD, 0, 0>(0, p, gx, gy, nthreads, st);                       \
                    case 1: return runflux_##NAME<SOL_HLLD, 0, 1>(0, p, gx, gy, nthreads, st);                       \
                    case 2: return runflux_##NAME<SOL_HLLD, 1, 0>(0, p, gx, gy, nthreads, st);                       \
                    default: return runflux_##NAME<SOL_HLLD, 1, 1>(0, p, gx, gy, nthreads, st);                      \
                }                                                                                                    \
            default: return -1;                                                                                      \
        }                                                                                                            \
    }

}  // namespace astrea
