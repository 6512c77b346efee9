// Plan = flat-sky grid metadata + FFT tables for one (Ny, Nx, θpix, T).
// Mirrors what the reference memoises per field type: ProjLambert (src/proj_lambert.jl:24-75) and the FFT plan
// (src/util_fft.jl:32-39).
#pragma once
#include "fft_core.cuh"
#include <memory>

namespace cmbl {

constexpr int RED_BLOCKS = 64;      // partial results per batch item of every reduction (fixed → deterministic)

// per-axis device tables
template <class T> struct AxisTables {
    Fft1D<T> fft;
    const T* mult_deriv = nullptr;   // [N] tile order: ℓ_herm[k(p)] / N   (Nyquist entry 0)
    const T* mult_sign = nullptr;    // [N] tile order: sign(k(p)) / N      (DC and Nyquist 0)
    T ell_nyq = 0;                   // ℓ of the Nyquist frequency (negative, src/proj_lambert.jl:63-64)
    int nyq_pos = 0;                 // tile position of the Nyquist frequency after the forward passes
    // fast path (flow_fast.cuh): a three-sweep schedule [R1, R2, 16] — the generic schedule up to N = 1024, [8, 16, 16] at N = 2048 (where
    // the generic kernels run four sweeps) — with planar twiddle tables
    //   ftw1[(q-1)][re|im][j]  = W_N^(j q),        j < N/R1, q = 1..R1-1     (first forward / last inverse pass)
    //   ftw2[(q-1)][re|im][jj] = W_{N/R1}^(jj q),  jj < 16,  q = 1..R2-1     (second forward / first inverse pass)
    // and the multiplier tables in ITS tile order (the same arrays as mult_deriv / mult_sign when the schedules coincide)
    const T* ftw1 = nullptr; const T* ftw2 = nullptr;
    const T* fmult_deriv = nullptr; const T* fmult_sign = nullptr;
};

struct PlanBase {
    int device = 0, Ny = 0, Nx = 0, Nyh = 0, dtype = 0;
    double theta_pix = 0;
    virtual ~PlanBase() {}
};

template <class T> struct PlanT : PlanBase {
    AxisTables<T> ax, ay;
    // natural-order grids (device): lx[Nx], ly[Nyh], lam[Nyh] (λ_rfft), sin2phi/cos2phi [Nx][Nyh]
    const T *lx = nullptr, *ly = nullptr, *lam = nullptr, *sin2phi = nullptr, *cos2phi = nullptr;
    T dx = 0, dlx = 0, dly = 0, omega_pix = 0, nyquist = 0;
    std::vector<void*> owned;                  // device allocations freed with the plan
    DevBuf scratch_four;                       // half-plane scratch for irfft2 (input is never clobbered)
    DevBuf scratch_red;                        // partial sums of reductions
    // host copies (for the host mirror / tests)
    std::vector<T> h_lx, h_ly, h_lam, h_sin2phi, h_cos2phi;
    ~PlanT() override { for (void* p : owned) dev_free(p); }
    size_t map_elems() const { return (size_t)Ny * Nx; }
    size_t four_elems() const { return (size_t)Nyh * Nx; }
};

std::unique_ptr<PlanBase> make_plan(int device, int Ny, int Nx, double theta_pix, int dtype);

// radix schedule for a power-of-two length (last pass radix 8 whenever N >= 8)
void fft_schedule(int N, int& npass, int* radix);
// pos[k] for a schedule
std::vector<int> fft_positions(int N, int npass, const int* radix);

}  // namespace cmbl
