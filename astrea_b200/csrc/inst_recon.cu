// ReconStage instantiations, one per reconstruction scheme (PCM needs none: its faces are the cell averages).
#include "dispatch.cuh"
namespace astrea {
template <class K>
static int run(const ReconStageParams& p, int gx, int gy, int nthreads, Stream st) {
    return launch<K>(p, gx, gy, nthreads, K::smem_bytes(nthreads), st);
}
// STAGED: transposed outputs (constrained transport) leave the warps through shared-memory tiles (device only)
template <int SCHEME, bool BULK>
static int by_outputs(const ReconStageParams& p, int gx, int gy, int nthreads, Stream st) {
#ifdef ASTREA_DEVICE_BUILD
    if (ReconStage<SCHEME>::staged_outputs(p)) return run<ReconStage<SCHEME, false, BULK, true>>(p, gx, gy, nthreads, st);
#endif
    return run<ReconStage<SCHEME, false, BULK, false>>(p, gx, gy, nthreads, st);
}
int launch_recon(int scheme, const ReconStageParams& p, int gx, int gy, int nthreads, Stream st) {
#ifdef ASTREA_DEVICE_BUILD
    // The march fed by the TMA engine's bulk copies (stages2d.cuh): aligned row segments, default author.  PLM and PPM
    // only: measured on a B200 (ms of reconstruction per step, bulk at 4 blocks per SM / register prefetch at 5) PPM
    // 8192^2 11.1 / 12.0, PLM + transverse PPM 4096^2 10.9 / 12.6, but WENO5 4096^2 4.86 / 4.41 — the WENO march is bound
    // by its divisions, not by load latency, and loses more to the lower occupancy than the ring gains.
    if (p.bulk && (p.c_lo & 1) == 0 && (p.w.col_pitch & 1) == 0 && (scheme != SCH_PPM || p.ppm_author == PPM_MC)) {
        switch (scheme) {
            case SCH_PLM: return by_outputs<SCH_PLM, true>(p, gx, gy, nthreads, st);
            case SCH_PPM: return by_outputs<SCH_PPM, true>(p, gx, gy, nthreads, st);
            default: break;
        }
    }
#endif
    switch (scheme) {
        case SCH_PLM: return by_outputs<SCH_PLM, false>(p, gx, gy, nthreads, st);
        case SCH_PPM:
            if (p.ppm_author != PPM_MC) {
#ifdef ASTREA_DEVICE_BUILD
                // authors 'c' / 'ph' with constrained transport: transposed face states, stored lane by lane (rare path)
                if (ReconStage<SCH_PPM, true>::staged_outputs(p)) return run<ReconStage<SCH_PPM, true, false, true>>(p, gx, gy, nthreads, st);
#endif
                return run<ReconStage<SCH_PPM, true>>(p, gx, gy, nthreads, st);
            }
            return by_outputs<SCH_PPM, false>(p, gx, gy, nthreads, st);
        case SCH_WENO3: return by_outputs<SCH_WENO3, false>(p, gx, gy, nthreads, st);
        case SCH_WENO5: return by_outputs<SCH_WENO5, false>(p, gx, gy, nthreads, st);
        case SCH_WENO7: return by_outputs<SCH_WENO7, false>(p, gx, gy, nthreads, st);
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
