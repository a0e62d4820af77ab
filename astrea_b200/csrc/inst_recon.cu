// ReconStage instantiations, one per reconstruction scheme (PCM needs none: its faces are the cell averages).
#include "dispatch.cuh"
namespace astrea {
int launch_recon(int scheme, const ReconStageParams& p, int gx, int gy, int nthreads, Stream st) {
    switch (scheme) {
        case SCH_PLM: return launch<ReconStage<SCH_PLM>>(p, gx, gy, nthreads, 0, st);
        case SCH_PPM:
            if (p.ppm_author != PPM_MC) return launch<ReconStage<SCH_PPM, true>>(p, gx, gy, nthreads, 0, st);
            return launch<ReconStage<SCH_PPM>>(p, gx, gy, nthreads, 0, st);
        case SCH_WENO3: return launch<ReconStage<SCH_WENO3>>(p, gx, gy, nthreads, 0, st);
        case SCH_WENO5: return launch<ReconStage<SCH_WENO5>>(p, gx, gy, nthreads, 0, st);
        case SCH_WENO7: return launch<ReconStage<SCH_WENO7>>(p, gx, gy, nthreads, 0, st);
        default: return -1;
    }
}
int launch_flux(int kind, int solver, int ax, int sax, int hydro, const FluxStageParams& p, int gx, int gy, int nthreads, Stream st) {
    switch (kind) {
        case 0: return launch_flux_pcm(solver, ax, sax, hydro, p, gx, gy, nthreads, st);
        case 1: return launch_flux_plm(solver, ax, sax, hydro, p, gx, gy, nthreads, st);
        default: return launch_flux_ho(solver, ax, sax, hydro, p, gx, gy, nthreads, st);
    }
}
}
