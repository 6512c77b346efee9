// extern "C": diagonal/basis operators, dot, CG Wiener filter.
#include "api_common.cuh"
#include "cg.cuh"

CMBL_FLOW_STRUCT;
struct cmbl_cg { std::unique_ptr<cmbl::CgBase> g; cmbl_flow* flow; };

#define CG_T(cg) (*static_cast<cmbl::CgT<T>*>((cg)->g.get()))

extern "C" {

int cmbl_diag_mul(cmbl_plan* plan, int basis, const void* diag, int Cd, const void* in, void* out, int C, int ldiv, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(plan && plan->p && diag && in && out, "NULL argument");
    CMBL_REQUIRE(basis == CMBL_MAP || basis == CMBL_FOURIER, "basis must be Map or Fourier");
    CMBL_DISPATCH(plan->p.get(), cmbl::diag_mul<T>(P, basis, (const T*)diag, Cd, in, out, C, ldiv != 0, as_stream(stream)));
    CMBL_API_END
}

int cmbl_field_axpby(cmbl_plan* plan, int basis, const double* a_host, int na, const void* x, const double* b_host, int nb, const void* y_or_null,
                     void* out, int Npol, int Nb, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(plan && plan->p && a_host && x && out, "NULL argument");
    CMBL_REQUIRE(basis == CMBL_MAP || basis == CMBL_FOURIER, "basis must be Map or Fourier");
    CMBL_REQUIRE(Nb >= 1 && Nb <= cmbl::AXPBY_MAX_NB && Npol >= 1, "cmbl_field_axpby handles 1..64 batch items");
    CMBL_REQUIRE((na == 1 || na == Nb) && (!y_or_null || (b_host && (nb == 1 || nb == Nb))), "scalars must have length 1 or Nb (batch sizes must broadcast)");
    CMBL_DISPATCH(plan->p.get(), {
        cmbl::AxpbyBody<T> k;
        k.per_batch = (basis == CMBL_MAP ? P.map_elems() : 2 * P.four_elems()) * (size_t)Npol;
        k.total = k.per_batch * (size_t)Nb; k.x = (const T*)x; k.y = (const T*)y_or_null; k.out = (T*)out; k.na = na; k.nb = y_or_null ? nb : 1;
        for (int i = 0; i < na; ++i) k.a[i] = a_host[i];
        if (y_or_null) for (int i = 0; i < nb; ++i) k.b[i] = b_host[i];
        cmbl::launch(k, (int)((k.total + k.NT - 1) / k.NT), 0, as_stream(stream));
    });
    CMBL_API_END
}

int cmbl_qu_eb(cmbl_plan* plan, int dir, const void* in, void* out, int Nb, int pair_stride_planes, int first_plane, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(plan && plan->p && in && out, "NULL argument");
    CMBL_DISPATCH(plan->p.get(), cmbl::qu_eb<T>(P, dir, (const cmbl::C2<T>*)in, (cmbl::C2<T>*)out, Nb, pair_stride_planes, first_plane, as_stream(stream)));
    CMBL_API_END
}

int cmbl_blockdiag_ieb(cmbl_plan* plan, int mode, const void* block, const void* in, void* out, int Nb, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(plan && plan->p && block && in && out, "NULL argument");
    CMBL_DISPATCH(plan->p.get(), cmbl::blockdiag_ieb<T>(P, mode, (const T*)block, (const cmbl::C2<T>*)in, (cmbl::C2<T>*)out, Nb, as_stream(stream)));
    CMBL_API_END
}

int cmbl_dot(cmbl_plan* plan, int basis, const void* a, const void* b, int Npol, int Nb, double* out_host, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(plan && plan->p && a && b && out_host, "NULL argument");
    CMBL_REQUIRE(basis == CMBL_MAP || basis == CMBL_FOURIER, "basis must be Map or Fourier");
    CMBL_DISPATCH(plan->p.get(), {
        double* part = (double*)P.scratch_red.reserve(sizeof(double) * (size_t)Nb * (cmbl::RED_BLOCKS + 1));
        cmbl::dot_partials<T>(P, basis, a, b, Npol, Nb, part, as_stream(stream));
        cmbl::SumPartialsBody s{Nb, part, part + (size_t)Nb * cmbl::RED_BLOCKS};
        cmbl::launch(s, 1, 0, as_stream(stream));
        cmbl::dev_download(out_host, part + (size_t)Nb * cmbl::RED_BLOCKS, sizeof(double) * Nb, as_stream(stream));
    });
    CMBL_API_END
}

int cmbl_cg_create(cmbl_cg** cg, cmbl_flow* flow, const cmbl_dataset_desc* ds, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(cg && flow && flow->f && ds, "NULL argument");
    auto h = std::make_unique<cmbl_cg>();
    h->flow = flow;
    CMBL_DISPATCH(flow->f->plan, {
        auto G = std::make_unique<cmbl::CgT<T>>();
        G->plan = &P; G->P = &P; G->F = static_cast<cmbl::FlowT<T>*>(flow->f.get());
        cmbl::cg_setup<T>(*G, *ds, as_stream(stream));
        h->g = std::move(G);
    });
    *cg = h.release();
    CMBL_API_END
}

int cmbl_cg_destroy(cmbl_cg* cg) {
    CMBL_API_BEGIN
    delete cg;
    CMBL_API_END
}

int cmbl_cg_begin(cmbl_cg* cg, const void* fstart_or_null, int offset, double* res_host, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(cg && cg->g, "NULL argument");
    CMBL_DISPATCH(cg->g->plan, cmbl::cg_begin<T>(CG_T(cg), (const cmbl::C2<T>*)fstart_or_null, offset != 0, res_host, as_stream(stream)));
    CMBL_API_END
}

int cmbl_cg_step(cmbl_cg* cg, double* res_host, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(cg && cg->g, "NULL argument");
    CMBL_DISPATCH(cg->g->plan, cmbl::cg_step<T>(CG_T(cg), res_host, as_stream(stream)));
    CMBL_API_END
}

int cmbl_cg_mark_best(cmbl_cg* cg, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(cg && cg->g, "NULL argument");
    CMBL_DISPATCH(cg->g->plan, { auto& G = CG_T(cg); cmbl::dev_copy(G.bestx.p, G.x.p, sizeof(cmbl::C2<T>) * G.nf() * G.C, as_stream(stream)); });
    CMBL_API_END
}

int cmbl_cg_result(cmbl_cg* cg, int which, void* f_out, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(cg && cg->g && f_out, "NULL argument");
    CMBL_REQUIRE(which == 0 || which == 1, "which must be 0 (bestx) or 1 (x)");
    CMBL_DISPATCH(cg->g->plan, { auto& G = CG_T(cg); cmbl::dev_copy(f_out, which == 0 ? G.bestx.p : G.x.p, sizeof(cmbl::C2<T>) * G.nf() * G.C, as_stream(stream)); });
    CMBL_API_END
}

int cmbl_wiener_cg(cmbl_cg* cg, const void* fstart_or_null, void* f_out, int nsteps, double tol, int offset,
                   int* iters_out, double* res_hist_host, void* stream) {
    return cmbl_wiener_cg_sharded(cg, nullptr, fstart_or_null, f_out, nsteps, tol, offset, iters_out, res_hist_host, stream);
}

int cmbl_gradientf_logpdf(cmbl_cg* cg, const void* f, const void* d_or_null, int d_is_zero, void* out, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(cg && cg->g && f && out, "NULL argument");
    CMBL_REQUIRE(f != out, "out must not alias f");
    CMBL_DISPATCH(cg->g->plan, cmbl::cg_gradientf<T>(CG_T(cg), (const cmbl::C2<T>*)f, (const cmbl::C2<T>*)d_or_null, d_is_zero != 0, (cmbl::C2<T>*)out, as_stream(stream)));
    CMBL_API_END
}

}  // extern "C"
