/* abi_smoke.c — a plain C program against include/cmbl_b200.h: proves the header is valid C, that every entry point links from C, and
 * (on a GPU box) runs the hot path with nothing but the C ABI and the CUDA runtime: plan, rfft2/irfft2 round trip, LenseFlow(phi)*f and
 * L\(L*f), dot.  Without a GPU it checks the error contract instead: a negative status and a message from cmbl_last_error().
 *   gcc -std=c99 -Iinclude tests/abi_smoke.c -o abi_smoke -Lcmblensing.jl_b200 -lcmbl_b200 -L$CUDA/lib64 -lcudart -lm
 * (tests/test_abi_c.py builds and runs it). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cmbl_b200.h"

/* the few CUDA runtime calls a C host needs (declared here so that no CUDA header is required to compile this file) */
extern int cudaGetDeviceCount(int*);
extern int cudaMalloc(void**, size_t);
extern int cudaFree(void*);
extern int cudaMemcpy(void*, const void*, size_t, int);
extern int cudaDeviceSynchronize(void);

#define CHECK(call) do { int rc_ = (call); if (rc_ != CMBL_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, cmbl_last_error()); return 1; } } while (0)

int main(void) {
    int ndev = 0;
    printf("%s\n", cmbl_version());
    if (cudaGetDeviceCount(&ndev) != 0 || ndev == 0) {
        cmbl_plan* p = NULL;
        int rc = cmbl_plan_create(&p, 0, 12, 8, 1.0, 1);              /* not a power of two: rejected before any CUDA call */
        if (rc >= 0 || strlen(cmbl_last_error()) == 0) { fprintf(stderr, "expected a negative status and a message\n"); return 1; }
        rc = cmbl_rfft2(NULL, NULL, NULL, 1, NULL);
        if (rc >= 0) { fprintf(stderr, "NULL plan accepted\n"); return 1; }
        printf("no GPU: error contract ok (%s)\n", cmbl_last_error());
        return 0;
    }
    const int Ny = 256, Nx = 256, Npol = 2, Nb = 2, C = Npol * Nb, nmap = Ny * Nx, nf = (Ny / 2 + 1) * Nx;
    cmbl_plan* plan = NULL; cmbl_flow* L = NULL;
    CHECK(cmbl_plan_create(&plan, 0, Ny, Nx, 2.0, 1));
    double* hf = (double*)malloc(sizeof(double) * nmap * C); double* hphi = (double*)malloc(sizeof(double) * nmap * Nb); double* hout = (double*)malloc(sizeof(double) * nmap * C);
    unsigned s = 12345u;
    for (int i = 0; i < nmap * C; ++i) { s = s * 1664525u + 1013904223u; hf[i] = (double)(s >> 8) / 16777216.0 - 0.5; }
    for (int b = 0; b < Nb; ++b) for (int x = 0; x < Nx; ++x) for (int y = 0; y < Ny; ++y)          /* a smooth potential, arcminute deflections */
        hphi[(b * Nx + x) * Ny + y] = 2e-6 * sin(6.2831853 * (x + 3 * b) / Nx) * cos(6.2831853 * 2 * y / Ny);
    void *df, *dphi, *dout, *dfour;
    if (cudaMalloc(&df, sizeof(double) * nmap * C) || cudaMalloc(&dout, sizeof(double) * nmap * C) || cudaMalloc(&dphi, sizeof(double) * nmap * Nb) ||
        cudaMalloc(&dfour, 2 * sizeof(double) * nf * C)) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
    cudaMemcpy(df, hf, sizeof(double) * nmap * C, 1); cudaMemcpy(dphi, hphi, sizeof(double) * nmap * Nb, 1);
    /* FFT round trip */
    CHECK(cmbl_rfft2(plan, df, dfour, C, NULL)); CHECK(cmbl_irfft2(plan, dfour, dout, C, NULL));
    cudaDeviceSynchronize(); cudaMemcpy(hout, dout, sizeof(double) * nmap * C, 2);
    double e = 0; for (int i = 0; i < nmap * C; ++i) e = fmax(e, fabs(hout[i] - hf[i]));
    printf("irfft2(rfft2(f)) max error %.2e\n", e); if (e > 1e-12) return 1;
    /* LenseFlow: L*f, then L\(L*f) ~ f; adjoint identity through dot */
    CHECK(cmbl_lenseflow_create(&L, plan, 7, Npol, Nb, Nb));
    CHECK(cmbl_lenseflow_precompute(L, dphi, CMBL_MAP, 0, NULL));
    printf("kernel path %d\n", cmbl_lenseflow_kernel_path(L));
    CHECK(cmbl_lenseflow_apply(L, CMBL_OP_L, df, dout, NULL));
    double d0[2], d1[2];
    CHECK(cmbl_dot(plan, CMBL_MAP, df, df, Npol, Nb, d0, NULL)); CHECK(cmbl_dot(plan, CMBL_MAP, dout, dout, Npol, Nb, d1, NULL));
    printf("|f|^2 = %.6e %.6e   |L f|^2 = %.6e %.6e\n", d0[0], d0[1], d1[0], d1[1]);
    if (!(fabs(d1[0] / d0[0] - 1) < 0.05)) return 1;                  /* lensing moves power around, nearly conserving it */
    CHECK(cmbl_lenseflow_apply(L, CMBL_OP_LINV, dout, dout, NULL));
    cudaDeviceSynchronize(); cudaMemcpy(hout, dout, sizeof(double) * nmap * C, 2);
    double num = 0, den = 0; for (int i = 0; i < nmap * C; ++i) { num += (hout[i] - hf[i]) * (hout[i] - hf[i]); den += hf[i] * hf[i]; }
    printf("|L\\(L f) - f| / |f| = %.2e\n", sqrt(num / den)); if (sqrt(num / den) > 1e-3) return 1;
    /* host-buffer entry point */
    CHECK(cmbl_lenseflow_apply_host(L, CMBL_OP_L, hf, hout, NULL));
    const double eta_scale = 1e3;
    for (int i = 0; i < nmap * Nb; ++i) hphi[i] *= eta_scale;
    void* deta; cudaMalloc(&deta, sizeof(double) * nmap * Nb); cudaMemcpy(deta, hphi, sizeof(double) * nmap * Nb, 1);
    double am[2]; CHECK(cmbl_max_lensing_step(plan, dphi, CMBL_MAP, deta, CMBL_MAP, Nb, am, NULL));
    if (!(am[0] > 0 && am[1] > 0)) return 1;
    printf("get_max_lensing_step %.4f %.4f\n", am[0], am[1]);
    CHECK(cmbl_lenseflow_destroy(L)); CHECK(cmbl_plan_destroy(plan));
    cudaFree(df); cudaFree(dout); cudaFree(dphi); cudaFree(dfour); cudaFree(deta); free(hf); free(hphi); free(hout);
    printf("abi smoke ok\n");
    return 0;
}
