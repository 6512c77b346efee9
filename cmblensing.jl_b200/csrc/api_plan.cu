// extern "C": plan, FFT entry points, error reporting.
#include "api_common.cuh"
#include "fft2d.cuh"
#include "pointwise.cuh"

namespace cmbl {
std::string prof_report();
static thread_local std::string t_last_error;
static std::string g_prof_text;
void set_last_error(const std::string& s) { t_last_error = s; }
}

extern "C" {

const char* cmbl_last_error(void) { return cmbl::t_last_error.c_str(); }

const char* cmbl_version(void) {
#ifdef CMBL_EMU
    return "cmbl_b200 0.1 (host kernel-logic emulator: tests only)";
#else
    return "cmbl_b200 0.1 (sm_100a)";
#endif
}

long long cmbl_launch_count(void) { return cmbl::g_launch_count; }

int cmbl_profile_begin(void) { cmbl::g_profiling = true; return CMBL_OK; }

const char* cmbl_profile_end(void) {
    cmbl::g_profiling = false;
    cmbl::g_prof_text = cmbl::prof_report();
    return cmbl::g_prof_text.c_str();
}

int cmbl_plan_create(cmbl_plan** plan, int device, int Ny, int Nx, double theta_pix_arcmin, int dtype) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(plan != nullptr, "plan out-pointer is NULL");
    auto h = std::make_unique<cmbl_plan>();
    h->p = cmbl::make_plan(device, Ny, Nx, theta_pix_arcmin, dtype);
    *plan = h.release();
    CMBL_API_END
}

int cmbl_plan_destroy(cmbl_plan* plan) {
    CMBL_API_BEGIN
    delete plan;
    CMBL_API_END
}

int cmbl_plan_grids(cmbl_plan* plan, void* lx, void* ly, void* lam, void* s2, void* c2, double* scalars) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(plan && plan->p, "NULL plan");
    CMBL_DISPATCH(plan->p.get(), {
        if (lx) memcpy(lx, P.h_lx.data(), P.h_lx.size() * sizeof(T));
        if (ly) memcpy(ly, P.h_ly.data(), P.h_ly.size() * sizeof(T));
        if (lam) memcpy(lam, P.h_lam.data(), P.h_lam.size() * sizeof(T));
        if (s2) memcpy(s2, P.h_sin2phi.data(), P.h_sin2phi.size() * sizeof(T));
        if (c2) memcpy(c2, P.h_cos2phi.data(), P.h_cos2phi.size() * sizeof(T));
        if (scalars) { scalars[0] = P.dx; scalars[1] = P.dlx; scalars[2] = P.dly; scalars[3] = P.omega_pix; scalars[4] = P.nyquist; }
    });
    CMBL_API_END
}

int cmbl_cl_to_cov(cmbl_plan* plan, const double* ell_host, const double* cl_host, int n, double units, void* out, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(plan && plan->p && ell_host && cl_host && out, "NULL argument");
    CMBL_REQUIRE(n >= 2, "the Cl table needs at least two points");
    for (int i = 1; i < n; ++i) CMBL_REQUIRE(ell_host[i] > ell_host[i - 1], "ell must be strictly ascending");
    CMBL_DISPATCH(plan->p.get(), {
        cmbl::DevBuf tab;
        double* d = (double*)tab.reserve(sizeof(double) * 2 * (size_t)n);
        cmbl::dev_upload(d, ell_host, sizeof(double) * n, as_stream(stream));
        cmbl::dev_upload(d + n, cl_host, sizeof(double) * n, as_stream(stream));
        cmbl::ClTo2DBody<T> b{P.Nx, P.Nyh, n, P.lx, P.ly, d, d + n, units == 0 ? P.omega_pix : (T)units, (T*)out};
        cmbl::launch(b, (int)((P.four_elems() + b.NT - 1) / b.NT), 0, as_stream(stream));
#ifndef CMBL_EMU
        CMBL_CUDA(cudaStreamSynchronize(as_stream(stream)));                  // the table buffer is released on return
#endif
    });
    CMBL_API_END
}

int cmbl_rfft2(cmbl_plan* plan, const void* map, void* four, int C, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(plan && plan->p, "NULL plan");
    CMBL_REQUIRE(C >= 0 && (C == 0 || (map && four)), "NULL buffer");
    CMBL_DISPATCH(plan->p.get(), cmbl::rfft2<T>(P, (const T*)map, (cmbl::C2<T>*)four, C, as_stream(stream)));
    CMBL_API_END
}

int cmbl_irfft2(cmbl_plan* plan, const void* four, void* map, int C, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(plan && plan->p, "NULL plan");
    CMBL_REQUIRE(C >= 0 && (C == 0 || (map && four)), "NULL buffer");
    CMBL_DISPATCH(plan->p.get(), cmbl::irfft2<T>(P, (const cmbl::C2<T>*)four, (T*)map, C, as_stream(stream)));
    CMBL_API_END
}

}  // extern "C"
