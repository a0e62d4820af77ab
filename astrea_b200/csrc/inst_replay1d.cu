// Persistent 1D replay (replay1d.cuh): dispatch over the reconstruction scheme; the kernels live in inst_replay1d_*.cu.
#include "replay1d.cuh"
namespace astrea {
int launch_replay1d(int scheme, int solver, const ReplayOp* ops, int nops, const unsigned char* blob, int nsteps, size_t smem_bytes, Stream st) {
    switch (scheme) {
        case SCH_PCM: return launch_replay1d_pcm(solver, ops, nops, blob, nsteps, smem_bytes, st);
        case SCH_PLM: return launch_replay1d_plm(solver, ops, nops, blob, nsteps, smem_bytes, st);
        case SCH_PPM: return launch_replay1d_ppm(solver, ops, nops, blob, nsteps, smem_bytes, st);
        case SCH_WENO3: return launch_replay1d_weno3(solver, ops, nops, blob, nsteps, smem_bytes, st);
        case SCH_WENO5: return launch_replay1d_weno5(solver, ops, nops, blob, nsteps, smem_bytes, st);
        case SCH_WENO7: return launch_replay1d_weno7(solver, ops, nops, blob, nsteps, smem_bytes, st);
        default: return -1;
    }
}
}
