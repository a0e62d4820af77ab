// Persistent replay of the recorded launch list of a 1D time step (runtime.cuh "launch recording").
//
// BASELINE config 1 (Sod tube, 1024 cells) is 7 launches per step of a few microseconds of work each; even as a CUDA
// graph a step costs ~45 us, almost all of it launch and dependency latency.  Here ONE thread block of 256 threads
// executes the recorded operations in order — the blocks of each recorded launch one after the other, a block barrier
// between operations — for as many steps as asked, without returning to the host: the grid (64 KB per register at 1024
// cells) lives in L2 / L1, dt is computed by the recorded ClockKernel.  Same kernel bodies, same arithmetic, same
// results as the launched path; selected by astrea_run_steps for 1D grids up to REPLAY_MAX_CELLS.
#pragma once
#include "aux_kernels.cuh"
#include "sweep1d.cuh"

namespace astrea {

constexpr int REPLAY_THREADS = 256;
constexpr int64_t REPLAY_MAX_CELLS = 16384;

template <class K> struct type_tag { using type = K; };

// one recorded operation, executed by the calling block (device) / by a host-simulated block per recorded block (host)
template <int SCH, int SOL, class Run>
HD void replay_dispatch(const ReplayOp& o, Run&& run) {
    switch (o.kind) {
        case RK_HALO: run(type_tag<HaloKernel>{}); break;
        case RK_SWEEP1D: run(type_tag<Sweep1D<SCH, SOL>>{}); break;
        case RK_COMBINE: run(type_tag<CombineKernel>{}); break;
        case RK_RATE: run(type_tag<RateKernel>{}); break;
        case RK_CLOCK: run(type_tag<ClockKernel>{}); break;
        case RK_UPDATE:
            switch (o.sub) {
                case 1 * 2: run(type_tag<UpdateKernel<1, false>>{}); break;
                case 2 * 2: run(type_tag<UpdateKernel<2, false>>{}); break;
                case 3 * 2: run(type_tag<UpdateKernel<3, false>>{}); break;
                case 4 * 2: run(type_tag<UpdateKernel<4, false>>{}); break;
                case 5 * 2: run(type_tag<UpdateKernel<5, false>>{}); break;
                case 5 * 2 + 1: run(type_tag<UpdateKernel<5, true>>{}); break;
                case 7 * 2 + 1: run(type_tag<UpdateKernel<7, true>>{}); break;
                default: break;
            }
            break;
        default: break;
    }
}

#ifdef ASTREA_DEVICE_BUILD
// the execution layer seen by a kernel body that was recorded with fewer threads than the replay block has
struct SubExec : DeviceExec {
    int n;
    __device__ explicit SubExec(int nthreads) : n(nthreads) {}
    __device__ int nthreads() const { return n; }
    template <class F>
    __device__ __forceinline__ void phase(F&& f) {
        if ((int)threadIdx.x < n) f((int)threadIdx.x);
        __syncthreads();
    }
    template <class F>
    __device__ __forceinline__ void wphase(F&& f) {
        if ((int)threadIdx.x < n) f((int)threadIdx.x);
        __syncwarp();
    }
    template <class G>
    __device__ __forceinline__ void publish_max(G&& get, unsigned long long* dst, unsigned long long* flag) {
        DeviceExec::publish_max([&](int tid, double& val, bool& bad) { if (tid < n) get(tid, val, bad); }, dst, flag);
    }
    template <class G>
    __device__ void publish_min3(G&& get, unsigned long long* dst) {
        DeviceExec::publish_min3([&](int tid, unsigned long long* key) {
            key[0] = key[1] = key[2] = ~0ull;
            if (tid < n) get(tid, key);
        }, dst);
    }
};

// one recorded launch: its blocks one after the other.  Not inlined: every kernel body is compiled once per translation
// unit and keeps its own register allocation instead of being folded into one 255-register monolith.
template <class K>
__device__ __noinline__ void replay_run(const ReplayOp& o, const unsigned char* blob) {
    const typename K::Params& p = *reinterpret_cast<const typename K::Params*>(blob + o.offset);
    SubExec ex(o.nthreads);
    for (int by = 0; by < o.gy; ++by)
        for (int bx = 0; bx < o.gx; ++bx) K::block(p, bx, by, ex);
}

template <int SCH, int SOL>
__global__ void __launch_bounds__(REPLAY_THREADS, 1) replay1d_entry(const ReplayOp* ops, int nops, const unsigned char* blob, int nsteps) {
    for (int step = 0; step < nsteps; ++step) {
        for (int k = 0; k < nops; ++k) {
            const ReplayOp o = ops[k];
            if (o.kind == RK_MEMSET) {
                unsigned char* dst = reinterpret_cast<unsigned char*>((uintptr_t)o.ptr);
                for (unsigned long long b = threadIdx.x; b < o.bytes; b += REPLAY_THREADS) dst[b] = (unsigned char)o.sub;
                __syncthreads();
                continue;
            }
            replay_dispatch<SCH, SOL>(o, [&](auto tag) { replay_run<typename decltype(tag)::type>(o, blob); });
            __syncthreads();          // the next operation reads what this one wrote (global memory, same block)
        }
    }
}
#endif

template <int SCH, int SOL>
inline int run_replay(const ReplayOp* ops, int nops, const unsigned char* blob, int nsteps, size_t smem_bytes, Stream st) {
#ifdef ASTREA_DEVICE_BUILD
    static size_t configured = 48 * 1024;
    if (smem_bytes > configured) {
        cudaError_t e = cudaFuncSetAttribute(replay1d_entry<SCH, SOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return (int)e;
        configured = smem_bytes;
    }
    replay1d_entry<SCH, SOL><<<1, REPLAY_THREADS, smem_bytes, st.s>>>(ops, nops, blob, nsteps);
    return (int)cudaGetLastError();
#else
    // host simulation: the recorded blocks, one host-simulated block each, in recorded order
    for (int step = 0; step < nsteps; ++step)
        for (int k = 0; k < nops; ++k) {
            const ReplayOp& o = ops[k];
            if (o.kind == RK_MEMSET) { std::memset(reinterpret_cast<void*>((uintptr_t)o.ptr), o.sub, (size_t)o.bytes); continue; }
            replay_dispatch<SCH, SOL>(o, [&](auto tag) {
                using K = typename decltype(tag)::type;
                const typename K::Params& p = *reinterpret_cast<const typename K::Params*>(blob + o.offset);
                for (int by = 0; by < o.gy; ++by)
                    for (int bx = 0; bx < o.gx; ++bx) {
                        HostExec ex(o.nthreads, smem_bytes);
                        K::block(p, bx, by, ex);
                    }
            });
        }
    (void)st;
    return 0;
#endif
}

// Body of one inst_replay1d_*.cu: the four solvers of one reconstruction scheme
#define ASTREA_DEFINE_REPLAY(NAME, SCH)                                                                                            \
    int launch_replay1d_##NAME(int solver, const ReplayOp* ops, int nops, const unsigned char* blob, int nsteps, size_t smem_bytes, Stream st) { \
        switch (solver) {                                                                                                          \
            case SOL_LLF: return run_replay<SCH, SOL_LLF>(ops, nops, blob, nsteps, smem_bytes, st);                                \
            case SOL_LW: return run_replay<SCH, SOL_LW>(ops, nops, blob, nsteps, smem_bytes, st);                                  \
            case SOL_HLLC: return run_replay<SCH, SOL_HLLC>(ops, nops, blob, nsteps, smem_bytes, st);                              \
            case SOL_HLLD: return run_replay<SCH, SOL_HLLD>(ops, nops, blob, nsteps, smem_bytes, st);                              \
            default: return -1;                                                                                                    \
        }                                                                                                                          \
    }
int launch_replay1d_pcm(int solver, const ReplayOp* ops, int nops, const unsigned char* blob, int nsteps, size_t smem_bytes, Stream st);
int launch_replay1d_plm(int solver, const ReplayOp* ops, int nops, const unsigned char* blob, int nsteps, size_t smem_bytes, Stream st);
int launch_replay1d_ppm(int solver, const ReplayOp* ops, int nops, const unsigned char* blob, int nsteps, size_t smem_bytes, Stream st);
int launch_replay1d_weno3(int solver, const ReplayOp* ops, int nops, const unsigned char* blob, int nsteps, size_t smem_bytes, Stream st);
int launch_replay1d_weno5(int solver, const ReplayOp* ops, int nops, const unsigned char* blob, int nsteps, size_t smem_bytes, Stream st);
int launch_replay1d_weno7(int solver, const ReplayOp* ops, int nops, const unsigned char* blob, int nsteps, size_t smem_bytes, Stream st);

// 0 on success, > 0 a CUDA error, -1 for an unknown scheme / solver
int launch_replay1d(int scheme, int solver, const ReplayOp* ops, int nops, const unsigned char* blob, int nsteps, size_t smem_bytes, Stream st);

}  // namespace astrea
