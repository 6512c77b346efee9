// Shared infrastructure for the cmbl_b200 kernels.
//
// Every kernel in this library is written in "phase style": a body functor whose work is a sequence of
// block-wide thread loops (CMBL_FOR_THREADS) separated by CMBL_SYNC().  Compiled by nvcc for sm_100a the loop
// collapses to `tid = threadIdx.x` and the sync to `__syncthreads()`.  Compiled by g++ with -DCMBL_EMU the very
// same source runs the threads of a block one after another on the host; that build (tests/_emu) exists ONLY so
// the kernel index arithmetic can be unit-tested on machines without a GPU.  It is not shipped and the product
// package never loads it.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <stdexcept>
#include <vector>
#ifdef CMBL_EMU
#include <thread>
#include <atomic>
#endif

#ifdef CMBL_EMU
#define HD inline
#define DEV inline
#define DEV_NOINLINE inline
typedef void* cmblStream_t;
#define CMBL_FOR_THREADS(tid, NT) for (int tid = 0; tid < (NT); ++tid)
#define CMBL_FOR_GROUP(tid, NG, goff) for (int tid = 0; tid < (NG); ++tid)
#define CMBL_SYNC() ((void)0)
#define CMBL_LDG(p) ::cmbl::ldg(p)
#else
#include <cuda_runtime.h>
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
#define DEV_NOINLINE __device__ __noinline__
typedef cudaStream_t cmblStream_t;
#define CMBL_FOR_THREADS(tid, NT) for (int tid = threadIdx.x, _once = 1; _once; _once = 0)
// a thread GROUP of a warp-specialised block: NG consecutive threads starting at thread `goff` (tid counts from 0 inside the group)
#define CMBL_FOR_GROUP(tid, NG, goff) for (int tid = (int)threadIdx.x - (goff), _once = 1; _once; _once = 0)
#define CMBL_SYNC() __syncthreads()
#define CMBL_LDG(p) ::cmbl::ldg(p)
#endif

namespace cmbl {

// ---------------------------------------------------------------------------------------------------------------
// complex pair, aligned so the compiler emits one 64/128-bit access
// ---------------------------------------------------------------------------------------------------------------
template <class T> struct alignas(2 * sizeof(T)) C2 { T x, y; };
template <class T> HD C2<T> mk(T x, T y) { C2<T> r; r.x = x; r.y = y; return r; }
template <class T> HD C2<T> operator+(C2<T> a, C2<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <class T> HD C2<T> operator-(C2<T> a, C2<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <class T> HD C2<T> cmul(C2<T> a, C2<T> b) { return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <class T> HD C2<T> cmulc(C2<T> a, C2<T> b) { /* a * conj(b) */ return mk<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
template <class T> HD C2<T> cconj(C2<T> a) { return mk<T>(a.x, -a.y); }
template <class T> HD C2<T> cscale(C2<T> a, T s) { return mk<T>(a.x * s, a.y * s); }

// read-only (non-coherent) global loads
template <class U> HD U ldg(const U* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
template <> HD C2<float> ldg<C2<float>>(const C2<float>* p) {
#ifdef __CUDA_ARCH__
    float2 v = __ldg(reinterpret_cast<const float2*>(p)); return mk<float>(v.x, v.y);
#else
    return *p;
#endif
}
template <> HD C2<double> ldg<C2<double>>(const C2<double>* p) {
#ifdef __CUDA_ARCH__
    double2 v = __ldg(reinterpret_cast<const double2*>(p)); return mk<double>(v.x, v.y);
#else
    return *p;
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------
struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
#define CMBL_REQUIRE(cond, msg) do { if (!(cond)) throw ::cmbl::Error(std::string(msg) + "  [" #cond "]"); } while (0)

#ifndef CMBL_EMU
#define CMBL_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) \
    throw ::cmbl::Error(std::string(#call) + ": " + cudaGetErrorString(_e)); } while (0)
#endif

// ---------------------------------------------------------------------------------------------------------------
// device memory / copies (host malloc in the emulator build)
// ---------------------------------------------------------------------------------------------------------------
inline void* dev_alloc(size_t bytes) {
    if (bytes == 0) bytes = 16;
#ifdef CMBL_EMU
    void* p = nullptr; if (posix_memalign(&p, 256, bytes)) throw Error("host alloc failed"); return p;
#else
    void* p = nullptr; CMBL_CUDA(cudaMalloc(&p, bytes)); return p;
#endif
}
inline void dev_free(void* p) {
    if (!p) return;
#ifdef CMBL_EMU
    free(p);
#else
    cudaFree(p);
#endif
}
inline void dev_upload(void* dst, const void* src, size_t bytes, cmblStream_t st) {
#ifdef CMBL_EMU
    (void)st; memcpy(dst, src, bytes);
#else
    CMBL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    CMBL_CUDA(cudaStreamSynchronize(st));          // src may be a temporary
#endif
}
inline void dev_download(void* dst, const void* src, size_t bytes, cmblStream_t st) {
#ifdef CMBL_EMU
    (void)st; memcpy(dst, src, bytes);
#else
    CMBL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    CMBL_CUDA(cudaStreamSynchronize(st));
#endif
}
inline void dev_copy(void* dst, const void* src, size_t bytes, cmblStream_t st) {
#ifdef CMBL_EMU
    (void)st; memmove(dst, src, bytes);
#else
    CMBL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
#endif
}
inline void dev_zero(void* dst, size_t bytes, cmblStream_t st) {
#ifdef CMBL_EMU
    (void)st; memset(dst, 0, bytes);
#else
    CMBL_CUDA(cudaMemsetAsync(dst, 0, bytes, st));
#endif
}

// RAII device buffer that can grow
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { dev_free(p); }
    void* reserve(size_t bytes) {
        if (bytes > cap) {
#ifndef CMBL_EMU
            if (p) cudaDeviceSynchronize();
#endif
            dev_free(p); p = dev_alloc(bytes); cap = bytes;
        }
        return p;
    }
    template <class U> U* as() const { return reinterpret_cast<U*>(p); }
};

// ---------------------------------------------------------------------------------------------------------------
// kernel launch: Body has `static constexpr int NT` and `void operator()(int blk, unsigned char* smem) const`
// ---------------------------------------------------------------------------------------------------------------
extern long long g_launch_count;      // number of kernels of this library launched (bench.py reports it)

// optional per-kernel timing with CUDA events on the launching stream (cmbl_profile_begin / cmbl_profile_end)
extern bool g_profiling;
void prof_before(const char* name, cmblStream_t st);
void prof_after(cmblStream_t st);
template <class Body> struct KernelName { static const char* get() { return Body::name(); } };

#ifndef CMBL_EMU
template <class Body, class = void> struct MinBlocks { static constexpr int value = 1; };
template <class Body> struct MinBlocks<Body, decltype((void)Body::MINB)> { static constexpr int value = Body::MINB; };
template <class Body, class = void> struct UsesPdl { static constexpr bool value = false; };
template <class Body> struct UsesPdl<Body, decltype((void)Body::PDL)> { static constexpr bool value = Body::PDL; };
// programmatic dependent launch of a kernel that supports it (Body::PDL): CMBL_PDL=0/1 forces it off/on, otherwise the launcher's hint decides
inline int pdl_env() { static const int v = [] { const char* e = getenv("CMBL_PDL"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }(); return v; }
inline bool pdl_enabled(int hint) { const int e = pdl_env(); return e >= 0 ? e == 1 : hint > 0; }
template <class Body> __global__ void __launch_bounds__(Body::NT, MinBlocks<Body>::value) kern(const Body b) {
    extern __shared__ __align__(1024) unsigned char cmbl_smem[];   // 1024: tensor-map copies with the 128-byte swizzle
    b((int)blockIdx.x, cmbl_smem);
}
#endif

// cluster > 1: the blocks of the grid are launched as thread-block clusters of that many consecutive blocks (co-scheduled by the hardware,
// so a cluster barrier between them can never wait for a block that is not resident)
template <class Body> void launch(const Body& b, int grid, size_t smem, cmblStream_t st, int cluster = 1, int pdl_hint = 0) {
    if (grid <= 0) return;
    ++g_launch_count;
#ifdef CMBL_EMU
    (void)st; (void)cluster; (void)pdl_hint;
    int nthr = (int)std::thread::hardware_concurrency(); if (nthr < 1) nthr = 1; if (nthr > grid) nthr = grid;
    if (const char* e = getenv("CMBL_EMU_THREADS")) { nthr = atoi(e); if (nthr < 1) nthr = 1; }
    if (grid <= 64) nthr = grid;       // small (persistent) grids: every block gets its own host thread — blocks of a persistent
                                       // kernel may wait for each other (flags), so all of them must be running
    std::atomic<int> next(0);
    auto worker = [&]() {
        std::vector<unsigned char> buf(smem + 64);
        unsigned char* sm = buf.data(); sm += (64 - ((uintptr_t)sm & 63)) & 63;
        for (;;) { int blk = next.fetch_add(1); if (blk >= grid) break; b(blk, sm); }
    };
    if (nthr == 1) worker();
    else { std::vector<std::thread> th; for (int i = 0; i < nthr; ++i) th.emplace_back(worker); for (auto& t : th) t.join(); }
#else
    static thread_local bool configured = false;              // one attribute call per instantiation per thread
    CMBL_REQUIRE(smem <= 227 * 1024, "kernel tile exceeds 227 KB of shared memory");
    if (smem > 48 * 1024 && !configured) {
        CMBL_CUDA(cudaFuncSetAttribute(kern<Body>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        configured = true;
    }
    if (g_profiling) prof_before(Body::name(), st);
    if (cluster > 1) {
        CMBL_REQUIRE(grid % cluster == 0, "grid must be a multiple of the cluster size");
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)Body::NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CMBL_CUDA(cudaLaunchKernelEx(&cfg, kern<Body>, b));
    } else if (UsesPdl<Body>::value && pdl_enabled(pdl_hint) && !g_profiling) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)Body::NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CMBL_CUDA(cudaLaunchKernelEx(&cfg, kern<Body>, b));
    } else kern<Body><<<grid, Body::NT, smem, st>>>(b);
    if (g_profiling) prof_after(st);
    CMBL_CUDA(cudaGetLastError());
#endif
}

// grid of a persistent kernel: every block that can be resident on the device at once (cached per instantiation)
template <class Body> int persistent_blocks(size_t smem) {
#ifdef CMBL_EMU
    (void)smem;
    static int v = [] { const char* e = getenv("CMBL_EMU_PERSISTENT"); int n = e ? atoi(e) : 6; return n < 1 ? 1 : n; }();
    return v;
#else
    static thread_local int cached = 0;
    if (!cached) {
        CMBL_CUDA(cudaFuncSetAttribute(kern<Body>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        int per_sm = 0, dev = 0, sms = 0;
        CMBL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern<Body>, Body::NT, smem));
        CMBL_CUDA(cudaGetDevice(&dev));
        CMBL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CMBL_REQUIRE(per_sm >= 1, "persistent kernel does not fit on an SM");
        cached = per_sm * sms;
    }
    return cached;
#endif
}

// blocks of a persistent kernel launched in clusters of `cluster` blocks that can be resident at once
template <class Body> int persistent_blocks_clustered(size_t smem, int cluster) {
#ifdef CMBL_EMU
    const int n = persistent_blocks<Body>(smem); return n - n % cluster;
#else
    static thread_local int cached = 0, cached_for = 0;
    if (!cached || cached_for != cluster) {
        CMBL_CUDA(cudaFuncSetAttribute(kern<Body>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(cluster * 1024)); cfg.blockDim = dim3((unsigned)Body::NT); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nclusters = 0;
        CMBL_CUDA(cudaOccupancyMaxActiveClusters(&nclusters, kern<Body>, &cfg));
        CMBL_REQUIRE(nclusters >= 1, "persistent kernel does not fit on the device in clusters");
        cached = nclusters * cluster; cached_for = cluster;
    }
    return cached;
#endif
}
// barrier over the blocks of this block's cluster (a no-op for an unclustered launch: the implicit cluster is the block itself)
DEV void cluster_sync() {
#ifdef __CUDA_ARCH__
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}

// named barriers of a warp-specialised block (bar.sync / bar.arrive / bar.red with an id and a participant count); id 0 with the
// whole block is __syncthreads().  The host emulator runs the roles of a block one after another, so they are no-ops there.
DEV void group_sync(int bar, int count) {
#ifdef __CUDA_ARCH__
    if (bar == 0) __syncthreads(); else asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(count) : "memory");
#else
    (void)bar; (void)count;
#endif
}
DEV void group_arrive(int bar, int count) {
#ifdef __CUDA_ARCH__
    asm volatile("bar.arrive %0, %1;" ::"r"(bar), "r"(count) : "memory");
#else
    (void)bar; (void)count;
#endif
}
DEV int group_or(int bar, int count, int pred) {          // barrier + OR of `pred` over the group
#ifdef __CUDA_ARCH__
    if (bar == 0) return __syncthreads_or(pred);
    int out;
    asm volatile("{\n .reg .pred p, q;\n setp.ne.s32 q, %1, 0;\n bar.red.or.pred p, %2, %3, q;\n selp.s32 %0, 1, 0, p;\n}"
                 : "=r"(out) : "r"(pred), "r"(bar), "r"(count) : "memory");
    return out;
#else
    (void)bar; (void)count; return pred;
#endif
}

HD int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }

// inter-block signalling (device atomics; GCC builtins in the emulator, whose blocks run on several host threads)
DEV int atomic_inc_int(int* p) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, 1);
#else
    return __atomic_fetch_add(p, 1, __ATOMIC_ACQ_REL);
#endif
}
DEV void mem_fence() {
#ifdef __CUDA_ARCH__
    __threadfence();
#else
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
#endif
}
template <class U> DEV U ld_cg(const U* p) {          // read data another block wrote during this launch: bypass L1
#ifdef __CUDA_ARCH__
    return __ldcg(p);
#else
    return *reinterpret_cast<const volatile U*>(p);
#endif
}

// 128-bit vector of T for coalesced global access
template <class T> struct alignas(16) Vec { static constexpr int N = 16 / sizeof(T); T v[16 / sizeof(T)]; };
template <class T> HD Vec<T> vload(const T* p) { return *reinterpret_cast<const Vec<T>*>(p); }
template <class T> HD void vstore(T* p, const Vec<T>& x) { *reinterpret_cast<Vec<T>*>(p) = x; }

}  // namespace cmbl
