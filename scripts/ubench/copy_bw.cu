// Micro-benchmark (development aid): global -> shared streaming bandwidth of persistent blocks that fetch 32 KB tiles,
// with (0) cp.async 16 B per thread, (1) one 1-D bulk copy (TMA engine, mbarrier completion) per tile, (2) plain LDG.128 + STS.
// Double-buffered like the stage kernels; nothing is computed.  usage: copy_bw <mode> <blocks_per_sm> <tile_kb> <nbuf>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdint>

__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned phase) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}

template <int MODE> __global__ void __launch_bounds__(128) k(const char* in, float* out, long ntiles, int tile_bytes, int nbuf, int contiguous, int scatter, int sweeps, char* gout) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);            // up to 8 barriers
    unsigned char* bufs = smem + 128;
    const int tid = threadIdx.x;
    if (MODE == 1 && tid == 0) { for (int i = 0; i < nbuf; ++i) mbar_init(bars + i, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    float acc = 0;
    long t = blockIdx.x; int it = 0;
    long tstep = gridDim.x, tend = ntiles;
    if (contiguous) { t = (long)blockIdx.x * ntiles / gridDim.x; tend = (long)(blockIdx.x + 1) * ntiles / gridDim.x; tstep = 1; }
    auto issue = [&](long tile, int b) {
        const char* src = in + tile * tile_bytes; unsigned char* dst = bufs + (size_t)b * tile_bytes;
        if (MODE == 0) { for (int o = tid * 16; o < tile_bytes; o += 128 * 16) { int oo = o; if (scatter) { int kk = o >> 4, x = kk >> 1, cl = kk & 1; x ^= (x >> 4) & 7; oo = (cl * (tile_bytes >> 5) + x) << 4; } unsigned d = (unsigned)__cvta_generic_to_shared(dst + oo); asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + o) : "memory"); } asm volatile("cp.async.commit_group;" ::: "memory"); }
        else if (MODE == 1) { if (tid == 0) { mbar_expect(bars + b, tile_bytes); bulk_g2s(dst, src, tile_bytes, bars + b); } }
        else { for (int o = tid * 16; o < tile_bytes; o += 128 * 16) *reinterpret_cast<float4*>(dst + o) = __ldg(reinterpret_cast<const float4*>(src + o)); }
    };
    // prologue: nbuf-1 tiles in flight
    for (int i = 0; i < nbuf - 1; ++i) if (t + (long)i * tstep < tend) issue(t + (long)i * tstep, i);
    for (; t < tend; t += tstep, ++it) {
        const int b = it % nbuf;
        const long nx = t + (long)(nbuf - 1) * tstep;
        if (MODE == 0) { if (nbuf == 2) asm volatile("cp.async.wait_group 0;" ::: "memory"); else if (nbuf == 3) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 2;" ::: "memory"); }
        if (MODE == 1) mbar_wait(bars + b, (it / nbuf) & 1);
        __syncthreads();
        if (nx < tend) issue(nx, (it + nbuf - 1) % nbuf);
        else if (MODE == 0) asm volatile("cp.async.commit_group;" ::: "memory");
        acc += reinterpret_cast<float*>(bufs + (size_t)b * tile_bytes)[tid];   // touch
        for (int sw = 0; sw < sweeps; ++sw) {
            float4* t4 = reinterpret_cast<float4*>(bufs + (size_t)b * tile_bytes);
            float4 v[16];
            for (int i = 0; i < 16; ++i) v[i] = t4[tid + 128 * i];
            for (int i = 0; i < 16; ++i) { v[i].x += 1.f; t4[tid + 128 * ((i + sw + 1) & 15)] = v[i]; }
            __syncthreads();
        }
        if (gout) { const float4* t4 = reinterpret_cast<const float4*>(bufs + (size_t)b * tile_bytes); float4* g4 = reinterpret_cast<float4*>(gout + t * tile_bytes);
            for (int i = 0; i < 16; ++i) g4[tid + 128 * i] = t4[tid + 128 * i]; }
        __syncthreads();
    }
    if (acc == 1.2345f) out[0] = acc;
}

int main(int argc, char** argv) {
    int mode = atoi(argv[1]), bps = atoi(argv[2]), tile_kb = atoi(argv[3]), nbuf = atoi(argv[4]);
    size_t bytes = (size_t)(argc > 5 ? atoi(argv[5]) : 1024) << 20; int contiguous = argc > 6 ? atoi(argv[6]) : 0; int scatter = argc > 7 ? atoi(argv[7]) : 0; int sweeps = argc > 8 ? atoi(argv[8]) : 0; int dostore = argc > 9 ? atoi(argv[9]) : 0; char* gout = nullptr; if (dostore) cudaMalloc(&gout, bytes); int tile_bytes = tile_kb * 1024; long ntiles = bytes / tile_bytes;
    char* in; float* out; cudaMalloc(&in, bytes); cudaMalloc(&out, 4); cudaMemset(in, 1, bytes);
    size_t smem = 128 + (size_t)nbuf * tile_bytes;
    auto fn = mode == 0 ? k<0> : mode == 1 ? k<1> : k<2>;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    int grid = 148 * bps;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); fn<<<grid, 128, smem>>>(in, out, ntiles, tile_bytes, nbuf, contiguous, scatter, sweeps, gout); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r == 2) printf("scatter %d sweeps %d store %d MB %zu contiguous %d mode %d blocks/SM %d tile %dKB nbuf %d: %.1f us  %.0f GB/s  (%s)\n", scatter, sweeps, dostore, bytes >> 20, contiguous, mode, bps, tile_kb, nbuf, ms * 1e3, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
