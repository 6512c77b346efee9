// Micro-benchmark (development aid): what does the memory system deliver for the ACCESS PATTERN of the column kernel's memory phase, with
// the arithmetic and the shared-memory sweeps taken away and the bytes in flight per SM as a free parameter?
// Row-grouped planes (G = 4 rows per group, Nx = Ny = 1024, fp64), C = 16 maps + 2 p planes per batch item; persistent blocks take column
// tiles (M = 4 columns: 256 runs of 128 B at a 32 KB stride per operand) round-robin, polarisation fastest, like FastColBody.  Per 32-byte
// unit a middle RK4 stage reads u, tmp, p1, p2, y, acc and writes acc, u' (8 streams).  UNR units per thread are loaded before the first use.
// usage: stream_pattern  (prints GB/s for several UNR x blocks/SM)      build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_pattern stream_pattern.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int N = 1024, G = 4, M = 4, NT = 128, NB = 8, NPOL = 2, C = NB * NPOL;
constexpr int UNITS = M * N / G;                       // 32-byte units per tile and operand (1024)
constexpr int TPP = N / M;                             // column tiles per plane

struct V4 { double a, b, c, d; };
__device__ __forceinline__ V4 ld(const double* p) { V4 r; asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p)); return r; }
__device__ __forceinline__ void st(double* p, const V4& v) { asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "d"(v.d) : "memory"); }

template <int UNR, int MINB, int PF> __global__ void __launch_bounds__(NT, MINB) k(const double* u, const double* tmp, const double* p, const double* y, double* acc, double* uo, int ntiles, int kind) {
    const size_t nmap = (size_t)N * N;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int pol = t % NPOL, ct = (t / NPOL) % TPP, item = (t / NPOL) / TPP, c = item * NPOL + pol, x0 = ct * M;
        const double* p1 = p + (size_t)(item * 2) * nmap; const double* p2 = p1 + nmap;
        const size_t base = (size_t)c * nmap;
        if (PF && t + (int)gridDim.x < ntiles) {               // the whole next tile of this block into L2: bytes in flight that cost no registers
            const int tn = t + gridDim.x, poln = tn % NPOL, ctn = (tn / NPOL) % TPP, itemn = (tn / NPOL) / TPP;
            const size_t basen = (size_t)(itemn * NPOL + poln) * nmap;
            const double* p1n = p + (size_t)(itemn * 2) * nmap;
            for (int r = threadIdx.x; r < N / G; r += NT) {    // one 128-byte run per row group and operand
                const size_t gn = ((size_t)r * N + ctn * M) * G;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(u + basen + gn)); asm volatile("prefetch.global.L2 [%0];" ::"l"(tmp + basen + gn));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p1n + gn)); asm volatile("prefetch.global.L2 [%0];" ::"l"(p1n + nmap + gn));
                if (kind != 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(y + basen + gn));
                if (kind != 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(acc + basen + gn));
            }
        }
        for (int it = 0; it < UNITS / NT; it += UNR) {
            V4 a[UNR], b[UNR], q1[UNR], q2[UNR], yy[UNR], ac[UNR]; size_t g[UNR];
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int kk = threadIdx.x + (it + j) * NT, col = kk & (M - 1), yb = kk >> 2;
                g[j] = ((size_t)yb * N + x0 + col) * G;
                a[j] = ld(u + base + g[j]); b[j] = ld(tmp + base + g[j]); q1[j] = ld(p1 + g[j]); q2[j] = ld(p2 + g[j]);
                if (kind != 2) yy[j] = ld(y + base + g[j]);
                if (kind != 0) ac[j] = ld(acc + base + g[j]);
            }
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                V4 kx; kx.a = q1[j].a * b[j].a + q2[j].a * a[j].a; kx.b = q1[j].b * b[j].b + q2[j].b * a[j].b; kx.c = q1[j].c * b[j].c + q2[j].c * a[j].c; kx.d = q1[j].d * b[j].d + q2[j].d * a[j].d;
                V4 y0 = (kind != 2) ? yy[j] : V4{0, 0, 0, 0}, a0 = (kind != 0) ? ac[j] : y0, o1, o2;
                o1.a = a0.a + 0.1 * kx.a; o1.b = a0.b + 0.1 * kx.b; o1.c = a0.c + 0.1 * kx.c; o1.d = a0.d + 0.1 * kx.d;
                o2.a = y0.a + 0.2 * kx.a; o2.b = y0.b + 0.2 * kx.b; o2.c = y0.c + 0.2 * kx.c; o2.d = y0.d + 0.2 * kx.d;
                st(acc + base + g[j], o1);
                if (kind != 2) st(uo + base + g[j], o2);
            }
        }
    }
}

template <int UNR, int MINB, int PF> static void run(const double* u, const double* tmp, const double* p, const double* y, double* acc, double* uo, int sms, int kind) {
    const int ntiles = C * TPP, grid = MINB * sms;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) k<UNR, MINB, PF><<<grid, NT>>>(u, tmp, p, y, acc, uo, ntiles, kind);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) k<UNR, MINB, PF><<<grid, NT>>>(u, tmp, p, y, acc, uo, ntiles, kind);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const double plane = (double)N * N * 8;
    // bytes: u, tmp (+ y unless last, + acc unless first) per map, p once per item (second polarisation from L2), acc out (+ u' unless last)
    const double reads = C * (2 + (kind != 2) + (kind != 0)) + NB * 2, writes = C * (1 + (kind != 2));
    const double bytes = (reads + writes) * plane;
    printf("kind %d  UNR %d  blocks/SM %d  L2 prefetch of the next tile %d  (%5.1f KB of loads in flight per SM): %7.1f us  %6.0f GB/s  (%s)\n", kind, UNR, MINB, PF,
           UNR * 32 * (4 + (kind != 2) + (kind != 0)) * NT * MINB / 1024.0, ms * 1e3, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t nmap = (size_t)N * N;
    double *u, *tmp, *p, *y, *acc, *uo;
    cudaMalloc(&u, nmap * C * 8); cudaMalloc(&tmp, nmap * C * 8); cudaMalloc(&y, nmap * C * 8); cudaMalloc(&acc, nmap * C * 8); cudaMalloc(&uo, nmap * C * 8); cudaMalloc(&p, nmap * NB * 2 * 8);
    cudaMemset(u, 0, nmap * C * 8); cudaMemset(tmp, 0, nmap * C * 8); cudaMemset(y, 0, nmap * C * 8); cudaMemset(acc, 0, nmap * C * 8); cudaMemset(uo, 0, nmap * C * 8); cudaMemset(p, 0, nmap * NB * 2 * 8);
    for (int kind = 0; kind < 3; ++kind) {
        run<2, 3, 0>(u, tmp, p, y, acc, uo, sms, kind);
        run<2, 3, 1>(u, tmp, p, y, acc, uo, sms, kind);
        run<2, 4, 0>(u, tmp, p, y, acc, uo, sms, kind);
        run<2, 4, 1>(u, tmp, p, y, acc, uo, sms, kind);
        run<1, 8, 0>(u, tmp, p, y, acc, uo, sms, kind);
        run<1, 8, 1>(u, tmp, p, y, acc, uo, sms, kind);
        run<4, 2, 1>(u, tmp, p, y, acc, uo, sms, kind);
    }
    return 0;
}
