// Execution layer: one kernel body, two back ends.
//
// A kernel is a struct ``K`` with
//     template <class Ex> static HD void block(const K::Params& p, int bx, int by, Ex& ex);
// whose body is a sequence of ``ex.phase([&](int tid) {...});`` calls.  ``phase`` runs the lambda for every
// thread of the block and ends with a block-wide barrier.
//   * device build : ``phase`` calls the lambda with threadIdx.x and then __syncthreads()
//   * hostsim build: ``phase`` loops tid = 0..nthreads-1 (test infrastructure, see common.cuh)
// Per-thread values that live across phases are declared as ``typename Ex::template Local<T>`` and indexed
// by tid (a register-resident T on the device, an array on the host).
#pragma once
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef ASTREA_DEVICE_BUILD
#include <cuda_runtime.h>
#endif

namespace astrea {

#ifdef ASTREA_DEVICE_BUILD
// ------------------------------------------------------------------------------------------- device back end
struct DeviceExec {
    template <class T>
    struct Local {
        T v;
        __device__ explicit Local(DeviceExec&) {}
        __device__ T& operator[](int) { return v; }
    };
    __device__ int nthreads() const { return blockDim.x; }
    __device__ double* smem() const {
        extern __shared__ __align__(16) double astrea_dyn_smem[];
        return astrea_dyn_smem;
    }
    template <class F>
    __device__ __forceinline__ void phase(F&& f) {
        f((int)threadIdx.x);
        __syncthreads();
    }
    // warp-scope phase: the threads of a warp run f, then re-converge (no block barrier)
    template <class F>
    __device__ __forceinline__ void wphase(F&& f) {
        f((int)threadIdx.x);
        __syncwarp();
    }
    // Value that lane (lane + delta) mod 32 of the same warp produced in an EARLIER phase: ``get(t)`` returns the
    // value held by thread t.  On the device every lane evaluates get() on itself and the value moves by shuffle.
    template <class G>
    __device__ __forceinline__ double lane(int tid, int delta, G&& get) {
        const double own = get(tid);
        return __shfl_sync(0xffffffffu, own, ((tid & 31) + delta) & 31);
    }
    // Exchange of per-lane values inside a warp through shared memory (XS) instead of shuffles: ``put`` publishes this
    // lane's value in one of the warp's NS slots, ``nbr`` reads the value lane (lane + delta) mod 32 published in an
    // EARLIER warp phase.  A 64-bit shuffle is two SHFL plus the moves that re-pair the halves; through shared
    // memory the exchange of one value with both neighbours is one STS.64 and two LDS.64.  With XS = false ``put`` does
    // nothing and ``nbr`` is ``lane``.
    // BT ("block tile"): the lanes of the whole block form one row of transverse points, so the neighbour of a lane
    // may sit in another warp; the slots are then block-wide arrays (one pad element at either end: the outermost
    // lanes are halo lanes whose neighbour values are never used) and the phases are separated by block barriers.
    template <bool XS, int NS, bool BT = false>
    __device__ __forceinline__ void put(int slot, double v) {
        if constexpr (XS && BT) smem()[slot * ((int)blockDim.x + 2) + 1 + (int)threadIdx.x] = v;
        else if constexpr (XS) smem()[(((int)threadIdx.x >> 5) * NS + slot) * 32 + ((int)threadIdx.x & 31)] = v;
    }
    template <bool XS, int NS, bool BT = false, class G>
    __device__ __forceinline__ double nbr(int tid, int slot, int delta, G&& get) {
        if constexpr (XS && BT) return smem()[slot * ((int)blockDim.x + 2) + 1 + tid + delta];
        else if constexpr (XS) return smem()[((tid >> 5) * NS + slot) * 32 + (((tid & 31) + delta) & 31)];
        else return lane(tid, delta, get);
    }
    // phase of a kernel whose unit of cooperation is the warp (BLOCK = false) or the block (BLOCK = true)
    template <bool BLOCK, class F>
    __device__ __forceinline__ void xphase(F&& f) {
        f((int)threadIdx.x);
        if constexpr (BLOCK) __syncthreads(); else __syncwarp();
    }
    template <bool BLOCK>
    __device__ __forceinline__ bool group_any(bool b) const { if constexpr (BLOCK) return block_any(b); else return warp_any(b); }
    // OR over the lanes of the group of ``get(thread)``, the same answer in every lane (called by all threads of the block)
    template <bool BLOCK, class G>
    __device__ __forceinline__ bool group_or(G&& get) const { return group_any<BLOCK>(get((int)threadIdx.x)); }
    __device__ __forceinline__ bool warp_any(bool b) const { return __any_sync(0xffffffffu, b) != 0; }
    __device__ __forceinline__ bool block_any(bool b) const { return __syncthreads_or(b) != 0; }
    // Max of a non-negative, finite per-thread value -> atomicMax on the bit pattern of *dst (for such values the
    // bit pattern orders like the value); ``bad`` (non-finite seen) sets *flag to the bit pattern of 1.0.
    // Warp scope: two 32-bit REDUX steps (high word, then the low words of the lanes that hold the maximal high
    // word) and one vote; lane 0 touches memory only when its warp raises the running maximum.  No block barrier.
    template <class G>
    __device__ __forceinline__ void publish_max(G&& get, unsigned long long* dst, unsigned long long* flag) {
        double val = 0.0;
        bool bad = false;
        get((int)threadIdx.x, val, bad);
        const unsigned hi = (unsigned)__double2hiint(val), lo = (unsigned)__double2loint(val);
        const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
        const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
        const bool any_bad = __any_sync(0xffffffffu, bad) != 0;
        if ((threadIdx.x & 31) == 0) {
            const unsigned long long bits = ((unsigned long long)mhi << 32) | mlo;
            if (bits > *(volatile unsigned long long*)dst) atomicMax(dst, bits);
            if (any_bad) atomicMax(flag, 0x3FF0000000000000ull);      // bit pattern of 1.0: the flag is all-reduced (MAX) as a double
        }
    }
    // Three running minima (64-bit keys) of per-thread values -> atomicMin on dst[0..2]; warp shuffle first.
    template <class G>
    __device__ void publish_min3(G&& get, unsigned long long* dst) {
        unsigned long long v[3];
        get((int)threadIdx.x, v);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, v[k], o);
                v[k] = other < v[k] ? other : v[k];
            }
            if ((threadIdx.x & 31) == 0 && v[k] != ~0ull) atomicMin(dst + k, v[k]);
        }
    }
};

// ---- bulk asynchronous copies (the TMA engine's 1-D mode: cp.async.bulk, SASS UBLKCP) completing on an mbarrier.
// Declared __host__ __device__ because the kernel bodies are host-device lambdas; the host pass of nvcc never runs them.
#ifdef __CUDA_ARCH__
#define ASTREA_DEVICE_ONLY(...) __VA_ARGS__
#else
#define ASTREA_DEVICE_ONLY(...)
#endif
__host__ __device__ __forceinline__ uint32_t smem_addr(const void* p) {
    ASTREA_DEVICE_ONLY(return (uint32_t)__cvta_generic_to_shared(p);)
    return 0;
}
__host__ __device__ __forceinline__ unsigned warp_mask() {
    ASTREA_DEVICE_ONLY(return __activemask();)
    return 0;
}
__host__ __device__ __forceinline__ void warp_sync(unsigned mask) {
    ASTREA_DEVICE_ONLY(__syncwarp(mask);)
    (void)mask;
}
__host__ __device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    ASTREA_DEVICE_ONLY(asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");)
    (void)bar; (void)count;
}
__host__ __device__ __forceinline__ void mbar_inval(uint64_t* bar) {
    ASTREA_DEVICE_ONLY(asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");)
    (void)bar;
}
// make the initialised barriers visible to the async proxy before the first copy names them
__host__ __device__ __forceinline__ void mbar_init_fence() {
    ASTREA_DEVICE_ONLY(asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");)
}
__host__ __device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    ASTREA_DEVICE_ONLY(asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");)
    (void)bar; (void)bytes;
}
__host__ __device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    ASTREA_DEVICE_ONLY(asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");)
    (void)bar; (void)parity;
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is counted on `bar`
__host__ __device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_global, uint32_t bytes, uint64_t* bar) {
    ASTREA_DEVICE_ONLY(asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                    ::"r"(smem_addr(dst_smem)), "l"(src_global), "r"(bytes), "r"(smem_addr(bar)) : "memory");)
    (void)dst_smem; (void)src_global; (void)bytes; (void)bar;
}

// hint: bring the line of `p` into L2 (no register, no dependency; an address outside any allocation must not be passed)
__host__ __device__ __forceinline__ void prefetch_l2(const void* p) {
    ASTREA_DEVICE_ONLY(asm volatile("prefetch.global.L2 [%0];" ::"l"(p));)
    (void)p;
}

// resident blocks per SM the register allocation should allow: K::MIN_BLOCKS when the kernel declares it, else 1
template <class K, class = void>
struct min_blocks_of { static constexpr int value = 0; };   // 0 = let ptxas choose (same as omitting the argument)
template <class K>
struct min_blocks_of<K, decltype((void)K::MIN_BLOCKS)> { static constexpr int value = K::MIN_BLOCKS; };

template <class K>
__global__ void __launch_bounds__(K::MAX_THREADS, min_blocks_of<K>::value) kernel_entry(const typename K::Params p) {
    DeviceExec ex;
    K::block(p, (int)blockIdx.x, (int)blockIdx.y, ex);
}

struct Stream { cudaStream_t s; };

template <class K>
inline int launch(const typename K::Params& p, int gx, int gy, int nthreads, size_t smem_bytes, Stream st) {
    // per kernel and device (function attributes belong to the device's context; the entry points select the context's
    // device before they launch): the opt-in dynamic shared-memory size set so far
    if (smem_bytes > 48 * 1024) {
        static size_t configured_bytes[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        size_t& have = configured_bytes[dev & 63];
        if (smem_bytes > have) {
            cudaError_t e = cudaFuncSetAttribute(kernel_entry<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            if (e != cudaSuccess) return (int)e;
            have = smem_bytes;
        }
    }
    kernel_entry<K><<<dim3(gx, gy, 1), dim3(nthreads, 1, 1), smem_bytes, st.s>>>(p);
    return (int)cudaGetLastError();
}

inline void* dev_alloc(size_t bytes) { void* p = nullptr; return cudaMalloc(&p, bytes) == cudaSuccess ? p : nullptr; }
inline void dev_free(void* p) { if (p) cudaFree(p); }
inline int dev_zero(void* p, size_t bytes, Stream st) { return (int)cudaMemsetAsync(p, 0, bytes, st.s); }
inline int dev_ones(void* p, size_t bytes, Stream st) { return (int)cudaMemsetAsync(p, 0xFF, bytes, st.s); }
inline int copy_h2d(void* d, const void* h, size_t bytes, Stream st) { return (int)cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st.s); }
inline int copy_d2h(void* h, const void* d, size_t bytes, Stream st) { return (int)cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st.s); }
inline int copy_d2d(void* d, const void* s, size_t bytes, Stream st) { return (int)cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, st.s); }
inline int copy_d2h_2d(void* h, size_t hpitch, const void* d, size_t dpitch, size_t width, size_t height, Stream st) {
    return (int)cudaMemcpy2DAsync(h, hpitch, d, dpitch, width, height, cudaMemcpyDeviceToHost, st.s);
}
inline int stream_sync(Stream st) { return (int)cudaStreamSynchronize(st.s); }

#else
// ------------------------------------------------------------------------------------------- hostsim back end
struct HostExec {
    int nthr;
    std::vector<double> shared;
    template <class T>
    struct Local {
        std::vector<T> v;
        explicit Local(HostExec& ex) : v(ex.nthr) {}
        T& operator[](int tid) { return v[tid]; }
    };
    HostExec(int n, size_t smem_bytes) : nthr(n), shared(smem_bytes / sizeof(double) + 1) {}
    int nthreads() const { return nthr; }
    double* smem() { return shared.data(); }
    template <class F>
    void phase(F&& f) {
        for (int t = 0; t < nthr; ++t) f(t);
    }
    template <class F>
    void wphase(F&& f) {
        for (int t = 0; t < nthr; ++t) f(t);
    }
    template <bool XS, int NS, bool BT = false>
    void put(int, double) {}
    template <bool XS, int NS, bool BT = false, class G>
    double nbr(int tid, int, int delta, G&& get) {
        if (XS && BT) { const int k = tid + delta; return get(k < 0 ? 0 : (k >= nthr ? nthr - 1 : k)); }
        return lane(tid, delta, get);
    }
    template <bool BLOCK, class F>
    void xphase(F&& f) {
        for (int t = 0; t < nthr; ++t) f(t);
    }
    template <bool BLOCK>
    bool group_any(bool b) const { return b; }
    template <bool BLOCK, class G>
    bool group_or(G&& get) const {              // the whole block stands in for the group (a superset of a warp)
        bool any = false;
        for (int t = 0; t < nthr; ++t) any = any || get(t);
        return any;
    }
    bool warp_any(bool b) const { return b; }      // the host simulation runs a block thread by thread: one guard per block
    bool block_any(bool b) const { return b; }
    template <class G>
    double lane(int tid, int delta, G&& get) {
        return get((tid & ~31) + (((tid & 31) + delta) & 31));
    }
    template <class G>
    void publish_max(G&& get, unsigned long long* dst, unsigned long long* flag) {
        for (int t = 0; t < nthr; ++t) {
            double val = 0.0, cur;
            bool bad = false;
            get(t, val, bad);
            std::memcpy(&cur, dst, 8);
            if (val > cur) std::memcpy(dst, &val, 8);
            if (bad) *flag = 0x3FF0000000000000ull;
        }
    }
    template <class G>
    void publish_min3(G&& get, unsigned long long* dst) {
        for (int t = 0; t < nthr; ++t) {
            unsigned long long v[3];
            get(t, v);
            for (int k = 0; k < 3; ++k) if (v[k] < dst[k]) dst[k] = v[k];
        }
    }
};

struct Stream { int s; };

template <class K>
inline int launch(const typename K::Params& p, int gx, int gy, int nthreads, size_t smem_bytes, Stream) {
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            HostExec ex(nthreads, smem_bytes);
            K::block(p, bx, by, ex);
        }
    return 0;
}

inline void* dev_alloc(size_t bytes) { return std::malloc(bytes); }
inline void dev_free(void* p) { std::free(p); }
inline int dev_zero(void* p, size_t bytes, Stream) { std::memset(p, 0, bytes); return 0; }
inline int dev_ones(void* p, size_t bytes, Stream) { std::memset(p, 0xFF, bytes); return 0; }
inline int copy_h2d(void* d, const void* h, size_t bytes, Stream) { std::memcpy(d, h, bytes); return 0; }
inline int copy_d2h(void* h, const void* d, size_t bytes, Stream) { std::memcpy(h, d, bytes); return 0; }
inline int copy_d2d(void* d, const void* s, size_t bytes, Stream) { std::memcpy(d, s, bytes); return 0; }
inline int copy_d2h_2d(void* h, size_t hpitch, const void* d, size_t dpitch, size_t width, size_t height, Stream) {
    for (size_t r = 0; r < height; ++r) std::memcpy((char*)h + r * hpitch, (const char*)d + r * dpitch, width);
    return 0;
}
inline int stream_sync(Stream) { return 0; }
#endif

}  // namespace astrea
