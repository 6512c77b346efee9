#include "cg.cuh"

namespace cmbl {

// elementwise helpers for the derived diagonals --------------------------------------------------------------
template <class T> struct DerivedDiagBody {
    static constexpr int NT = 256;
    static const char* name() { return "derived_diag"; }
    size_t n; const T *Cf, *Cn, *Cnhat, *B, *Bhat, *Mf;
    T *inv_Cf, *inv_Cn, *inv_Cn_Mf, *precond;
    HD static T pinv(T v) { return v == (T)0 ? (T)0 : (T)1 / v; }
    DEV void operator()(int blk, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) {
            size_t e = (size_t)blk * NT + tid;
            if (e < n) {
                T icf = pinv(Cf[e]), icn = pinv(Cn[e]);
                inv_Cf[e] = icf; inv_Cn[e] = icn;
                inv_Cn_Mf[e] = Mf[e] * icn;                    // M' starts with Mf' right after pinv(Cn)
                // Hessian_logpdf_preconditioner(:f): pinv(Cf) + B̂'M̂'pinv(Cn̂)M̂B̂   (src/dataset.jl:129-132)
                precond[e] = icf + Bhat[e] * Mf[e] * pinv(Cnhat[e]) * Mf[e] * Bhat[e];
            }
        }
    }
};

// Npol = 3 (pol = :IP): every harmonic operator is a BlockDiagIEB of 4 planes [ΣTE[1,1], ΣTE[2,1], ΣTE[2,2], ΣB]
// (src/specialops.jl:61-118).  2×2 pinv as src/field_vectors.jl:74-78 (the off-diagonal is read from [2,1] twice);
// the preconditioner pinv(Cf) + B̂'M̂'pinv(Cn̂)M̂B̂ is formed with full 2×2 products, left to right (src/specialops.jl:101-102).
template <class T> struct DerivedBlockBody {
    static constexpr int NT = 256;
    static const char* name() { return "derived_block"; }
    size_t nf; const T *Cf, *Cn, *Cnhat, *Bhat, *Mf;
    T *inv_Cf, *inv_Cn, *inv_precond;
    struct M2 { T a, b, c, d, e; };                   // [a b; c d] ⊕ e
    HD static T pinv(T v) { return v == (T)0 ? (T)0 : (T)1 / v; }
    HD M2 ld(const T* A, size_t r) const { return M2{A[r], A[nf + r], A[nf + r], A[2 * nf + r], A[3 * nf + r]}; }
    HD void st4(T* A, size_t r, const M2& m) const { A[r] = m.a; A[nf + r] = m.c; A[2 * nf + r] = m.d; A[3 * nf + r] = m.e; }
    HD static M2 inv(const M2& m) { const T id = pinv(m.a * m.d - m.c * m.c); return M2{m.d * id, -(m.c * id), -(m.c * id), m.a * id, pinv(m.e)}; }
    HD static M2 mul(const M2& x, const M2& y) { return M2{x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d, x.e * y.e}; }
    DEV void operator()(int blk, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) {
            size_t r = (size_t)blk * NT + tid;
            if (r < nf) {
                const M2 icf = inv(ld(Cf, r)), icn = inv(ld(Cn, r)), icnh = inv(ld(Cnhat, r)), bh = ld(Bhat, r), mf = ld(Mf, r);
                st4(inv_Cf, r, icf); st4(inv_Cn, r, icn);
                const M2 h = mul(mul(mul(mul(bh, mf), icnh), mf), bh);
                st4(inv_precond, r, inv(M2{icf.a + h.a, icf.b + h.b, icf.c + h.c, icf.d + h.d, icf.e + h.e}));
            }
        }
    }
};

template <class T> void cg_setup(CgT<T>& G, const cmbl_dataset_desc& ds, cmblStream_t st) {
    CMBL_REQUIRE(ds.Npol >= 1 && ds.Npol <= 3, "CG Wiener filter supports Npol = 1 (I), 2 (P) or 3 (IP, BlockDiagIEB operators)");
    CMBL_REQUIRE(ds.Npol == G.F->Npol && ds.Nb == G.F->Nb, "dataset Npol/Nb must match the LenseFlow handle");
    CMBL_REQUIRE(ds.Cf && ds.Cn && ds.Cnhat && ds.B && ds.Bhat && ds.Mf && ds.d, "NULL dataset diagonal");
    G.Npol = ds.Npol; G.Nb = ds.Nb; G.C = ds.Npol * ds.Nb;
    G.Cf = (const T*)ds.Cf; G.Cn = (const T*)ds.Cn; G.Cnhat = (const T*)ds.Cnhat; G.B = (const T*)ds.B; G.Bhat = (const T*)ds.Bhat;
    G.Mf = (const T*)ds.Mf; G.mask = (const T*)ds.mask_pix; G.d = (const C2<T>*)ds.d;
    if (G.Npol == 3) {
        const size_t n4 = G.nf() * 4;
        DerivedBlockBody<T> b{G.nf(), G.Cf, G.Cn, G.Cnhat, G.Bhat, G.Mf,
            (T*)G.inv_Cf.reserve(n4 * sizeof(T)), (T*)G.inv_Cn.reserve(n4 * sizeof(T)), (T*)G.precond.reserve(n4 * sizeof(T))};
        launch(b, (int)((G.nf() + b.NT - 1) / b.NT), 0, st);
    } else {
        const size_t n = G.nf() * G.Npol;
        DerivedDiagBody<T> b{n, G.Cf, G.Cn, G.Cnhat, G.B, G.Bhat, G.Mf,
            (T*)G.inv_Cf.reserve(n * sizeof(T)), (T*)G.inv_Cn.reserve(n * sizeof(T)), (T*)G.inv_Cn_Mf.reserve(n * sizeof(T)),
            (T*)G.precond.reserve(n * sizeof(T))};
        launch(b, (int)((n + b.NT - 1) / b.NT), 0, st);
    }
    const size_t vb = sizeof(C2<T>) * G.nf() * G.C;
    for (DevBuf* v : {&G.x, &G.r, &G.z, &G.p, &G.Ap, &G.b, &G.bestx, &G.w1, &G.w2}) v->reserve(vb);
    G.m1.reserve(sizeof(T) * G.nmap() * G.C);
    G.scal.reserve(sizeof(double) * ((size_t)2 * G.Nb + 2 * (size_t)G.Nb * RED_BLOCKS));
    G.h_res.assign(G.Nb, 0.0);
    G.begun = false;
}

template <class T>
static void chain(CgT<T>& G, const C2<T>* in, const T* din, const C2<T>* d, bool neg, const T* pre, int rot, const T* post,
                  const T* sdiag, const C2<T>* sub, C2<T>* out, cmblStream_t st, const T* pre2 = nullptr, int rot2 = 0) {
    FourierChainBody<T> b;
    b.Npol = G.Npol; b.Nb = G.Nb; b.nf = G.nf(); b.rot = (G.Npol >= 2) ? rot : 0; b.rot2 = (G.Npol >= 2) ? rot2 : 0; b.neg = neg;
    b.sin2phi = G.P->sin2phi; b.cos2phi = G.P->cos2phi;
    b.in = in; b.din = din; b.d = d; b.pre = pre; b.pre2 = pre2; b.post = post; b.sdiag = sdiag; b.sub = sub; b.out = out;
    size_t threads = (G.Npol >= 2) ? b.nf * G.Nb : b.nf * G.Nb * G.Npol;
    launch(b, (int)((threads + b.NT - 1) / b.NT), 0, st);
}

// Lϕ'*(B'*(M'*(pinv(Cn)*(d − M*(B*(Lϕ*f)))))) − pinv(Cf)*f
template <class T> void cg_gradientf(CgT<T>& G, const C2<T>* f, const C2<T>* d, bool d_zero, C2<T>* out, cmblStream_t st) {
    PlanT<T>& P = *G.P; FlowT<T>& F = *G.F;
    C2<T>* w1 = (C2<T>*)G.w1.p; C2<T>* w2 = (C2<T>*)G.w2.p; T* m1 = (T*)G.m1.p;
    const int C = G.C, n = F.nsteps;
    const T* none = nullptr; const C2<T>* cnone = nullptr;
    if (!d && !d_zero) d = G.d;
    // Mf'·pinv(Cn): one fused diagonal for Npol ≤ 2, two block operators (pinv(Cn) then Mf) for Npol = 3
    const T* const icn1 = G.Npol == 3 ? (const T*)G.inv_Cn.p : (const T*)G.inv_Cn_Mf.p;
    const T* const icn2 = G.Npol == 3 ? G.Mf : nullptr;
    // Ł(f): EB→QU, irfft2 ; Lϕ* in map space
    chain<T>(G, f, none, cnone, false, none, 1, none, none, cnone, w1, st);
    if (const int Gd = flow_rg_direct(F)) {         // the two transforms hand the integrator's row-grouped buffer over directly
        T* yrg = reinterpret_cast<T*>(F.yrg.reserve(sizeof(T) * P.map_elems() * F.C));
        irfft2<T>(P, w1, yrg, C, st, nullptr, 1, Gd);
        flow_integrate<T>(F, false, yrg, 0, 2 * n, st, true);
        rfft2<T>(P, yrg, w1, C, st, Gd);
    } else {
        irfft2<T>(P, w1, m1, C, st);
        flow_integrate<T>(F, false, m1, 0, 2 * n, st);
        rfft2<T>(P, m1, w1, C, st);
    }
    if (G.mask) {
        // B then M = Mf∘Mpix:   QU→EB, ×B, EB→QU (one pass) | irfft2 with ×Mpix on its store | rfft2 | QU→EB ×Mf
        chain<T>(G, w1, none, cnone, false, none, 2, G.B, none, cnone, w2, st, none, 1);
        irfft2<T>(P, w2, m1, C, st, G.mask, G.Npol);
        rfft2<T>(P, m1, w1, C, st);
        chain<T>(G, w1, none, cnone, false, none, 2, G.Mf, none, cnone, w2, st);
        // pinv(Cn)(d − ·), then M' = Mpix'∘Mf': ×(Mf·pinv(Cn)), EB→QU | irfft2 | ×Mpix | rfft2 | QU→EB, ×B', EB→QU
        chain<T>(G, w2, none, d, d == nullptr, icn1, 1, none, none, cnone, w1, st, icn2);
        irfft2<T>(P, w1, m1, C, st, G.mask, G.Npol);
        rfft2<T>(P, m1, w1, C, st);
        chain<T>(G, w1, none, cnone, false, none, 2, G.B, none, cnone, w1, st, none, 1);
    } else {
        // everything between Lϕ and Lϕ' is diagonal in the harmonic basis: B'·Mf·pinv(Cn)·(d − Mf·B·f̃)
        chain<T>(G, w1, none, cnone, false, none, 2, G.B, none, cnone, w2, st);
        chain<T>(G, w2, G.Mf, d, d == nullptr, icn1, 0, G.B, none, cnone, w1, st, icn2);
        chain<T>(G, w1, none, cnone, false, none, 1, none, none, cnone, w1, st);
    }
    // Lϕ' (Fourier QU state), back to the harmonic basis, − pinv(Cf) f
    flow_apply<T>(F, CMBL_OP_LH, w1, w1, st);
    chain<T>(G, w1, none, cnone, false, none, 2, none, (const T*)G.inv_Cf.p, f, out, st);
}

static void read_scalars(double* host, const double* dev, int n, cmblStream_t st) { dev_download(host, dev, sizeof(double) * n, st); }

template <class T> void cg_begin(CgT<T>& G, const C2<T>* fstart, bool offset, double* res_host, cmblStream_t st) {
    PlanT<T>& P = *G.P;
    const size_t vb = sizeof(C2<T>) * G.nf() * G.C;
    C2<T>* b = (C2<T>*)G.b.p; C2<T>* x = (C2<T>*)G.x.p; C2<T>* Ap = (C2<T>*)G.Ap.p;
    // b = −gradientf_logpdf(f=0, d):  with f = 0 the first half of the chain is identically zero, but it is evaluated
    // the same way the reference does (maximization.jl:34) so that b carries the same rounding.
    dev_zero(x, vb, st);
    cg_gradientf<T>(G, x, nullptr, false, b, st);
    {   // b = −b  (a₀ = gradientf_logpdf(0, 0) is exactly zero for this linear model, so `offset` adds nothing)
        FourierChainBody<T> k{};
        k.Npol = 1; k.Nb = G.C; k.nf = G.nf(); k.rot = 0; k.neg = true; k.in = b; k.out = b;
        launch(k, (int)((k.nf * G.C + k.NT - 1) / k.NT), 0, st);
    }
    (void)offset;
    const C2<T>* Ax = nullptr;
    if (fstart) {
        dev_copy(x, fstart, vb, st);
        cg_gradientf<T>(G, x, nullptr, true, Ap, st);        // A x = gradientf_logpdf(x, d=0) − a₀
        Ax = Ap;
    }
    CgInitBody<T> k{G.nf() * G.Npol, P.Nyh, P.lam, 1.0 / ((double)P.Ny * (double)P.Nx), (const T*)G.precond.p,
                    b, Ax, (C2<T>*)G.r.p, (C2<T>*)G.z.p, (C2<T>*)G.p.p, G.res_part()};
    if (G.Npol == 3) k.nf_block = G.nf();
    launch(k, G.Nb * RED_BLOCKS, sizeof(double) * k.NT, st);
    G.flip = 0;
    SumPartialsBody s{G.Nb, G.res_part(), G.res_cur()};
    launch(s, 1, 0, st);
    read_scalars(G.h_res.data(), G.res_cur(), G.Nb, st);
    for (int i = 0; i < G.Nb; ++i) { CMBL_REQUIRE(G.h_res[i] == G.h_res[i], "conjugate_gradient: res is NaN"); if (res_host) res_host[i] = G.h_res[i]; }
    dev_copy(G.bestx.p, x, vb, st);
    G.begun = true;
}

template <class T> void cg_step(CgT<T>& G, double* res_host, cmblStream_t st) {
    CMBL_REQUIRE(G.begun, "cmbl_cg_step before cmbl_cg_begin");
    PlanT<T>& P = *G.P;
    C2<T>* p = (C2<T>*)G.p.p; C2<T>* Ap = (C2<T>*)G.Ap.p;
    const double scale = 1.0 / ((double)P.Ny * (double)P.Nx);
    cg_gradientf<T>(G, p, nullptr, true, Ap, st);                                        // Ap = A*p
    dot_partials<T>(P, CMBL_FOURIER, p, Ap, G.Npol, G.Nb, G.pAp_part(), st);             // dot(p, Ap)
    CgUpdate1Body<T> u1{G.nf() * G.Npol, G.nf(), G.Npol, P.Nyh, P.lam, scale, G.res_cur(), G.pAp_part(), (const T*)G.precond.p,
                        p, Ap, (C2<T>*)G.x.p, (C2<T>*)G.r.p, (C2<T>*)G.z.p, G.res_part()};
    launch(u1, G.Nb * RED_BLOCKS, sizeof(double) * u1.NT, st);
    CgUpdate2Body<T> u2{G.nf() * G.Npol, G.res_cur(), G.res_part(), G.res_next(), (const C2<T>*)G.z.p, p};
    launch(u2, G.Nb * RED_BLOCKS, 0, st);
    G.flip = 1 - G.flip;
    // res_host == NULL: α, β and res stay on the device and nothing synchronises — the caller polls when it wants to (every k-th
    // iteration, or only at the end of a fixed-length solve); the Nb doubles are read back only when asked for
    if (res_host) {
        read_scalars(G.h_res.data(), G.res_cur(), G.Nb, st);
        for (int i = 0; i < G.Nb; ++i) res_host[i] = G.h_res[i];
    }
}

#define INST(T)                                                                                             \
    template void cg_setup<T>(CgT<T>&, const cmbl_dataset_desc&, cmblStream_t);                             \
    template void cg_gradientf<T>(CgT<T>&, const C2<T>*, const C2<T>*, bool, C2<T>*, cmblStream_t);         \
    template void cg_begin<T>(CgT<T>&, const C2<T>*, bool, double*, cmblStream_t);                          \
    template void cg_step<T>(CgT<T>&, double*, cmblStream_t);
INST(float)
INST(double)

}  // namespace cmbl
