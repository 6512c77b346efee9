// Plan construction: ℓ-grids computed in type T in the order of src/proj_lambert.jl:58-71, FFT schedules,
// twiddle / position / multiplier tables.
#include "plan.cuh"

namespace cmbl {

long long g_launch_count = 0;
bool g_profiling = false;

#ifndef CMBL_EMU
namespace {
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
std::vector<ProfRec> g_prof;
}
void prof_before(const char* name, cmblStream_t st) {
    ProfRec r; r.name = name;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
    g_prof.push_back(r);
}
void prof_after(cmblStream_t st) { cudaEventRecord(g_prof.back().e1, st); }
// returns "name count total_ms\n" lines
std::string prof_report() {
    cudaDeviceSynchronize();
    std::vector<std::string> names; std::vector<double> tot; std::vector<long long> cnt;
    for (auto& r : g_prof) {
        float ms = 0; cudaEventElapsedTime(&ms, r.e0, r.e1);
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
        size_t i = 0; for (; i < names.size(); ++i) if (names[i] == r.name) break;
        if (i == names.size()) { names.push_back(r.name); tot.push_back(0); cnt.push_back(0); }
        tot[i] += ms; cnt[i] += 1;
    }
    g_prof.clear();
    std::string out;
    for (size_t i = 0; i < names.size(); ++i) out += names[i] + " " + std::to_string(cnt[i]) + " " + std::to_string(tot[i]) + "\n";
    return out;
}
#else
void prof_before(const char*, cmblStream_t) {}
void prof_after(cmblStream_t) {}
std::string prof_report() { return ""; }
#endif

void fft_schedule(int N, int& npass, int* radix) {
    CMBL_REQUIRE(N >= 4 && (N & (N - 1)) == 0, "FFT length must be a power of two >= 4");
    for (int i = 0; i < MAX_PASSES; ++i) radix[i] = 0;
    if (N <= 16) { npass = 1; radix[0] = N; return; }
    // The LAST pass is twiddle-free and fuses with the first inverse pass in registers, so it gets the largest radix (16);
    // the leading passes carry twiddles (kept in registers for radix <= 8).
    int m = ilog2(N) - 4;
    npass = 0;
    while (m > 0) {
        int bits = (m >= 3 && m != 4) ? 3 : (m == 4 ? 2 : m);      // 8s, finishing with 4·4 / 4 / 2
        radix[npass++] = 1 << bits;
        m -= bits;
        CMBL_REQUIRE(npass < MAX_PASSES, "FFT length too large");
    }
    radix[npass++] = 16;
}

std::vector<int> fft_positions(int N, int npass, const int* radix) {
    std::vector<int> pos(N);
    for (int k = 0; k < N; ++k) {
        int rem = k, p = 0, stride = N;
        for (int i = 0; i < npass; ++i) {
            int q = rem % radix[i]; rem /= radix[i];
            stride /= radix[i];
            p += q * stride;
        }
        pos[k] = p;
    }
    return pos;
}

template <class U> static const U* upload_vec(std::vector<void*>& owned, const std::vector<U>& v) {
    void* d = dev_alloc(v.size() * sizeof(U));
    owned.push_back(d);
    dev_upload(d, v.data(), v.size() * sizeof(U), 0);
    return reinterpret_cast<const U*>(d);
}

template <class T>
static void build_axis(std::vector<void*>& owned, AxisTables<T>& a, int N, T dl) {
    Fft1D<T>& f = a.fft;
    f.N = N; f.logN = ilog2(N);
    fft_schedule(N, f.npass, f.radix);
    f.sk = ilog2(f.radix[f.npass - 1]);
    std::vector<int> pos = fft_positions(N, f.npass, f.radix);
    std::vector<C2<T>> W(N);
    const long double tau = 6.283185307179586476925286766559005768L;
    for (int t = 0; t < N; ++t) {
        long double ang = -tau * (long double)t / (long double)N;
        W[t].x = (T)cosl(ang); W[t].y = (T)sinl(ang);
    }
    // exact values on the axes
    W[0].x = 1; W[0].y = 0;
    if (N % 4 == 0) { W[N / 4].x = 0; W[N / 4].y = -1; W[N / 2].x = -1; W[N / 2].y = 0; W[3 * N / 4].x = 0; W[3 * N / 4].y = 1; }
    f.W = upload_vec(owned, W);
    f.pos = upload_vec(owned, pos);
    std::vector<T> md(N), ms(N);
    for (int k = 0; k < N; ++k) {
        int ks = (k < N / 2) ? k : k - N;                      // ifftshift order: 0..N/2-1, -N/2..-1
        T ell = (T)ks * dl;                                    // computed in T like the reference
        T herm = (k == N / 2) ? (T)0 : ell;
        md[pos[k]] = herm / (T)N;
        ms[pos[k]] = (k == 0 || k == N / 2) ? (T)0 : (T)(ks > 0 ? 1 : -1) / (T)N;
    }
    a.mult_deriv = upload_vec(owned, md);
    a.mult_sign = upload_vec(owned, ms);
    a.ell_nyq = (T)(-(N / 2)) * dl;
    a.nyq_pos = pos[N / 2];
    // fast-path schedule and tables (plan.cuh)
    int fr[MAX_PASSES] = {0, 0, 0, 0, 0, 0};
    bool fast = false;
    if (f.npass == 3 && f.radix[2] == 16) { fast = true; for (int i = 0; i < 3; ++i) fr[i] = f.radix[i]; }
    else if (N == 2048) { fast = true; fr[0] = 8; fr[1] = 16; fr[2] = 16; }
    if (fast) {
        const int R1 = fr[0], R2 = fr[1], S1 = N / R1;
        std::vector<T> t1((size_t)(R1 - 1) * 2 * S1), t2((size_t)(R2 - 1) * 2 * 16);
        for (int q = 1; q < R1; ++q)
            for (int j = 0; j < S1; ++j) { t1[((size_t)(q - 1) * 2) * S1 + j] = W[j * q].x; t1[((size_t)(q - 1) * 2 + 1) * S1 + j] = W[j * q].y; }
        for (int q = 1; q < R2; ++q)
            for (int jj = 0; jj < 16; ++jj) { t2[((size_t)(q - 1) * 2) * 16 + jj] = W[jj * q * R1].x; t2[((size_t)(q - 1) * 2 + 1) * 16 + jj] = W[jj * q * R1].y; }
        a.ftw1 = upload_vec(owned, t1);
        a.ftw2 = upload_vec(owned, t2);
        a.fmult_deriv = a.mult_deriv; a.fmult_sign = a.mult_sign;
        if (fr[1] != f.radix[1]) {                             // its own tile order
            std::vector<int> fpos = fft_positions(N, 3, fr);
            CMBL_REQUIRE(fpos[N / 2] == 8, "the fast kernels read the Nyquist coefficient at tile position 8");
            std::vector<T> fmd(N), fms(N);
            for (int k = 0; k < N; ++k) { fmd[fpos[k]] = md[pos[k]]; fms[fpos[k]] = ms[pos[k]]; }
            a.fmult_deriv = upload_vec(owned, fmd);
            a.fmult_sign = upload_vec(owned, fms);
        }
    }
}

template <class T> static std::unique_ptr<PlanBase> make_plan_t(int device, int Ny, int Nx, double theta_pix) {
    auto P = std::make_unique<PlanT<T>>();
    P->device = device; P->Ny = Ny; P->Nx = Nx; P->Nyh = Ny / 2 + 1; P->theta_pix = theta_pix;
    P->dtype = sizeof(T) == 4 ? 0 : 1;
    const double pi = 3.14159265358979323846;
    // src/proj_lambert.jl:58-62 (2π is Float64 in Julia; the quotient is rounded to T)
    T dx = (T)(theta_pix / 60.0 * (pi / 180.0));
    P->dx = dx;
    P->dlx = (T)(2 * pi / (double)((T)Nx * dx));
    P->dly = (T)(2 * pi / (double)((T)Ny * dx));
    P->nyquist = (T)(2 * pi / (double)((T)2 * dx));
    P->omega_pix = dx * dx;
    build_axis<T>(P->owned, P->ax, Nx, P->dlx);
    build_axis<T>(P->owned, P->ay, Ny, P->dly);
    const int Nyh = P->Nyh;
    P->h_lx.resize(Nx); P->h_ly.resize(Nyh); P->h_lam.resize(Nyh);
    for (int k = 0; k < Nx; ++k) P->h_lx[k] = (T)((k < Nx / 2) ? k : k - Nx) * P->dlx;          // :64
    for (int k = 0; k < Nyh; ++k) P->h_ly[k] = (T)((k < Ny / 2) ? k : k - Ny) * P->dly;          // :63 (last entry negative)
    for (int k = 0; k < Nyh; ++k) P->h_lam[k] = (k == 0 || (Ny % 2 == 0 && k == Ny / 2)) ? (T)1 : (T)2;   // util_fft.jl:137-143
    P->h_sin2phi.resize((size_t)Nx * Nyh); P->h_cos2phi.resize((size_t)Nx * Nyh);
    for (int ix = 0; ix < Nx; ++ix)
        for (int iy = 0; iy < Nyh; ++iy) {
            T phi = (T)std::atan2(P->h_ly[iy], P->h_lx[ix]);                                     // :66
            P->h_sin2phi[(size_t)ix * Nyh + iy] = (T)std::sin((T)2 * phi);                       // :67
            P->h_cos2phi[(size_t)ix * Nyh + iy] = (T)std::cos((T)2 * phi);
        }
    if (Ny % 2 == 0)                                                                             // :69-71
        for (int j = 1; j < Nx / 2; ++j)
            P->h_sin2phi[(size_t)(Nx - j) * Nyh + (Nyh - 1)] = P->h_sin2phi[(size_t)j * Nyh + (Nyh - 1)];
    P->lx = upload_vec(P->owned, P->h_lx);
    P->ly = upload_vec(P->owned, P->h_ly);
    P->lam = upload_vec(P->owned, P->h_lam);
    P->sin2phi = upload_vec(P->owned, P->h_sin2phi);
    P->cos2phi = upload_vec(P->owned, P->h_cos2phi);
    return P;
}

std::unique_ptr<PlanBase> make_plan(int device, int Ny, int Nx, double theta_pix, int dtype) {
    CMBL_REQUIRE(Ny >= 4 && (Ny & (Ny - 1)) == 0 && Nx >= 4 && (Nx & (Nx - 1)) == 0,
                 "Ny and Nx must be powers of two >= 4");
    CMBL_REQUIRE(Ny <= 8192 && Nx <= 8192, "Ny, Nx up to 8192 supported");
    CMBL_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (Float32) or 1 (Float64)");
#ifndef CMBL_EMU
    CMBL_CUDA(cudaSetDevice(device));
#endif
    if (dtype == 0) return make_plan_t<float>(device, Ny, Nx, theta_pix);
    return make_plan_t<double>(device, Ny, Nx, theta_pix);
}

}  // namespace cmbl
