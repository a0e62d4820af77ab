// ReconStage instantiations, one per reconstruction scheme (PCM needs none: its faces are the cell averages).
#include "dispatch.cuh"
namespace astrea {
#ifdef ASTREA_DEVICE_BUILD
template <int SCHEME>
static int launch_bulk(const ReconStageParams& p, int gx, int gy, int nthreads, Stream st) {
    using K = ReconStage<SCHEME, false, true>;
    return launch<K>(p, gx, gy, nthreads, K::smem_bytes(nthreads), st);
}
#endif
int launch_recon(int scheme, const ReconStageParams& p, int gx, int gy, int nthreads, Stream st) {
#ifdef ASTREA_DEVICE_BUILD
    // The march fed by the TMA engine's bulk copies (stages2d.cuh): aligned row segments, default author.  PLM and PPM
    // only: measured on a B200 (ms of reconstruction per step, bulk at 4 blocks per SM / register prefetch at 5) PPM
    // 8192^2 11.1 / 12.0, PLM + transverse PPM 4096^2 10.9 / 12.6, but WENO5 4096^2 4.86 / 4.41 — the WENO march is bound
    // by its divisions, not by load latency, and loses more to the lower occupancy than the ring gains.
    if (p.bulk && (p.c_lo & 1) == 0 && (p.w.col_pitch & 1) == 0 && (scheme != SCH_PPM || p.ppm_author == PPM_MC)) {
        switch (scheme) {
            case SCH_PLM: return launch_bulk<SCH_PLM>(p, gx, gy, nthreads, st);
            case SCH_PPM: return launch_bulk<SCH_PPM>(p, gx, gy, nthreads, st);
            default: break;
        }
    }
#endif
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
