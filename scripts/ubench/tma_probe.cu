// Probe (development aid): does a 3-D tensor-map copy of one column of a row-grouped plane land as a planar column with the
// chunk swizzle ch ^ ((ch >> 3) & 7)?  Prints the first mismatches.   usage: tma_probe <Ny> <Nx> <G> <elem 4|8>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const void* tmap, int x, int c2, int bytes, float* out, int mode, const CUtensorMap __grid_constant__ pm) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        printf("smem base %u\n", (unsigned)__cvta_generic_to_shared(smem));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)), "r"(bytes) : "memory");
        const void* t = mode ? (const void*)&pm : tmap;
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(t), "r"(0), "r"(x), "r"(c2), "r"((unsigned)__cvta_generic_to_shared(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"((unsigned)__cvta_generic_to_shared(&bar)) : "memory");
    for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}
int main(int argc, char** argv) {
    int Ny = atoi(argv[1]), Nx = atoi(argv[2]), G = atoi(argv[3]), es = atoi(argv[4]), mode = argc > 5 ? atoi(argv[5]) : 0, swz = argc > 6 ? atoi(argv[6]) : 3;
    int C = 3; size_t nel = (size_t)Ny * Nx * C;
    std::vector<float> h32(nel); std::vector<double> h64(nel);
    for (size_t i = 0; i < nel; ++i) { h32[i] = (float)i; h64[i] = (double)i; }
    void* d; cudaMalloc(&d, nel * es); cudaMemcpy(d, es == 4 ? (void*)h32.data() : (void*)h64.data(), nel * es, cudaMemcpyHostToDevice);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t gdim[3] = {(cuuint64_t)G, (cuuint64_t)Nx, (cuuint64_t)(Ny / G) * C};
    cuuint64_t gstr[2] = {(cuuint64_t)G * es, (cuuint64_t)Nx * G * es};
    cuuint32_t box[3] = {(cuuint32_t)G, 1u, (cuuint32_t)(Ny / G)}; cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)fp)(&m, es == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc %d  (Ny %d Nx %d G %d es %d mode %d swz %d)\n", (int)r, Ny, Nx, G, es, mode, swz); fflush(stdout);
    void* dm; cudaMalloc(&dm, 128); cudaMemcpy(dm, &m, 128, cudaMemcpyHostToDevice);
    int bytes = Ny * es; float* out; cudaMalloc(&out, bytes);
    int x = 5, plane = 1;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<<<1, 128, bytes>>>(dm, x, plane * (Ny / G), bytes, out, mode, m);
    cudaError_t e = cudaDeviceSynchronize(); printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<unsigned char> ho(bytes); cudaMemcpy(ho.data(), out, bytes, cudaMemcpyDeviceToHost);
    int V = 16 / es, CH = Ny / V, bad = 0;
    for (int ch = 0; ch < CH; ++ch) for (int e2 = 0; e2 < V; ++e2) {
        int y = ch * V + e2; size_t gi = (size_t)plane * Ny * Nx + ((size_t)(y / G) * Nx + x) * G + y % G;
        int pos = (swz == 3 ? (ch ^ ((ch >> 3) & 7)) : ch) * V + e2;
        double got = es == 4 ? ((float*)ho.data())[pos] : ((double*)ho.data())[pos];
        if (got != (double)gi) { if (bad < 6) printf("  y %d: expected %zu at pos %d, got %.0f\n", y, gi, pos, got); ++bad; }
    }
    printf("Ny %d Nx %d G %d es %d mode %d: %d mismatches of %d\n", Ny, Nx, G, es, mode, bad, Ny);
    return 0;
}
