// Sweep1D instantiations: every reconstruction scheme x Riemann solver of the 1D operator.
#include "dispatch.cuh"
namespace astrea {
template <int SCH>
static int run1d(int solver, const Sweep1DParams& p, int nthreads, Stream st) {
    const int gx = (int)((p.n + p.tile - 1) / p.tile);
    switch (solver) {
        case SOL_LLF: return launch<Sweep1D<SCH, SOL_LLF>>(p, gx, 1, nthreads, Sweep1D<SCH, SOL_LLF>::smem_bytes(nthreads), st);
        case SOL_HLLC: return launch<Sweep1D<SCH, SOL_HLLC>>(p, gx, 1, nthreads, Sweep1D<SCH, SOL_HLLC>::smem_bytes(nthreads), st);
        case SOL_LW: return launch<Sweep1D<SCH, SOL_LW>>(p, gx, 1, nthreads, Sweep1D<SCH, SOL_LW>::smem_bytes(nthreads), st);
        case SOL_HLLD: return launch<Sweep1D<SCH, SOL_HLLD>>(p, gx, 1, nthreads, Sweep1D<SCH, SOL_HLLD>::smem_bytes(nthreads), st);
        default: return -1;
    }
}
int launch_sweep1d(int scheme, int solver, const Sweep1DParams& p, int nthreads, Stream st) {
    switch (scheme) {
        case SCH_PCM: return run1d<SCH_PCM>(solver, p, nthreads, st);
        case SCH_PLM: return run1d<SCH_PLM>(solver, p, nthreads, st);
        case SCH_PPM: return run1d<SCH_PPM>(solver, p, nthreads, st);
        case SCH_WENO3: return run1d<SCH_WENO3>(solver, p, nthreads, st);
        case SCH_WENO5: return run1d<SCH_WENO5>(solver, p, nthreads, st);
        case SCH_WENO7: return run1d<SCH_WENO7>(solver, p, nthreads, st);
        default: return -1;
    }
}
}
