// Pullback through LenseFlow: the transpose flow negδvelocityᴴ (src/lenseflow.jl:176-214) integrated by RK4
// (src/numerical_algorithms.jl:11-24) on the joint state (f Map, δf Fourier, δϕ Fourier), as the reference's Zygote
// rule does (src/flowops.jl:40-68).
//
// One velocity evaluation at stage k (t = k/2n), written with the batched 2-D transforms of fft2d.cuh:
//     Łδf = irfft2(δf)                                   F = rfft2(f),  ∂ₓf = irfft2(iℓₓF),  ∂ᵧf = irfft2(iℓᵧF)
//     df/dt   = p₁∂ₓf + p₂∂ᵧf                             (velocity)
//     dδf/dt  = iℓₓ·rfft2(p₁Łδf) + iℓᵧ·rfft2(p₂Łδf)       (velocityᴴ)
//     w_i = Σ_pol Łδf·∂_i f          (src/proj_lambert.jl:423-430)
//     m = M⁻¹w   — reference: m₁ = M₁₁w₁+M₁₂w₂ ; m₂ = M₂₁·m₁+M₂₂w₂ (aliased in-place product, src/lenseflow.jl:198-200 with
//                  src/field_vectors.jl:48-49).  bug_compat = false uses the exact m₂ = M₂₁w₁+M₂₂w₂.
//     dδϕ/dt = iℓₓ·rfft2(m₁) + iℓᵧ·rfft2(m₂) + Σ_ij (−iℓ_i)(−iℓ_j)·rfft2(t·p_j·m_i)
//              (the (1,2) and (2,1) terms share their multiplier and are summed before the transform: 5 transforms, not 6)
// M⁻¹₁₂ ≡ M⁻¹₂₁ (the reference reads [2,1] twice, src/field_vectors.jl:87), so the cache holds 3 maps per time.
// This path is not on the headline roofline: it is built from the library's general transforms plus
// three fused pointwise kernels; the caches may be in the row-grouped layout of the fast stage kernels (index remap).
#include "flow.cuh"
#include "../../include/cmbl_b200.h"

namespace cmbl {

// spectra of the two derivatives of f: GX = iℓₓF, GY = iℓᵧF
template <class T> struct GradSpecBody {
    static constexpr int NT = 256;
    static const char* name() { return "grad_spec"; }
    int Nx, Nyh; const T* lx; const T* ly; const C2<T>* F; C2<T>* GX; C2<T>* GY; size_t total;     // total = C*Nx*Nyh
    DEV void operator()(int blk, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) {
            size_t e = (size_t)blk * NT + tid;
            if (e < total) {
                size_t r = e % ((size_t)Nx * Nyh);
                int kx = (int)(r / Nyh), ky = (int)(r - (size_t)kx * Nyh);
                C2<T> f = F[e];
                GX[e] = cmul(mk<T>((T)0, lx[kx]), f);
                GY[e] = cmul(mk<T>((T)0, ly[ky]), f);
            }
        }
    }
};

// pointwise part of negδvelocityᴴ; one thread per pixel and batch item
template <class T> struct DeltaPointBody {
    static constexpr int NT = 256;
    static const char* name() { return "delta_point"; }
    int Npol, Nb, Nbphi, Nx, Ny, G; bool bug_compat; T t;
    const T* ldf; const T* gx; const T* gy;          // [Nb][Npol] maps (reference layout)
    const T* pk; const T* mk3;                        // p[k]: [Nbphi][2] maps, M⁻¹[k]: [Nbphi][3] maps (layout G)
    T* dfdt; T* a1; T* a2;                            // [Nb][Npol] maps
    T* six;                                           // [Nb][5] maps: m1, m2, t p1 m1, t (p2 m1 + p1 m2), t p2 m2  (the two mixed terms share (−iℓx)(−iℓy))
    DEV void operator()(int blk, unsigned char*) const {
        const size_t nmap = (size_t)Nx * Ny;
        CMBL_FOR_THREADS(tid, NT) {
            size_t e = (size_t)blk * NT + tid;
            if (e < nmap * Nb) {
                const size_t b = e / nmap, r = e - b * nmap;
                size_t rc = r;                                                          // index inside the caches
                if (G > 0) { const int x = (int)(r / Ny), y = (int)(r - (size_t)x * Ny); rc = ((size_t)(y / G) * Nx + x) * G + (y % G); }
                const size_t bp = (Nbphi == 1) ? 0 : b;
                const T p1 = pk[(bp * 2 + 0) * nmap + rc], p2 = pk[(bp * 2 + 1) * nmap + rc];
                const T m11 = mk3[(bp * 3 + 0) * nmap + rc], m21 = mk3[(bp * 3 + 1) * nmap + rc], m22 = mk3[(bp * 3 + 2) * nmap + rc];
                T w1 = 0, w2 = 0;
                for (int pol = 0; pol < Npol; ++pol) {
                    const size_t i = (b * Npol + pol) * nmap + r;
                    const T l = ldf[i], dx = gx[i], dy = gy[i];
                    dfdt[i] = p1 * dx + p2 * dy;
                    a1[i] = p1 * l; a2[i] = p2 * l;
                    w1 += l * dx; w2 += l * dy;
                }
                const T m1 = m11 * w1 + m21 * w2;                                      // M₁₂ ≡ M₂₁
                const T m2 = bug_compat ? (m21 * m1 + m22 * w2) : (m21 * w1 + m22 * w2);
                T* s = six + b * 5 * nmap + r;
                s[0] = m1; s[nmap] = m2;
                s[2 * nmap] = t * p1 * m1; s[3 * nmap] = t * p2 * m1 + t * p1 * m2; s[4 * nmap] = t * p2 * m2;
            }
        }
    }
};

// spectral part: dδf/dt = iℓₓA1 + iℓᵧA2 ;  dδϕ/dt = iℓₓM1 + iℓᵧM2 + Σ_ij (−iℓ_i)(−iℓ_j) R_ij
template <class T> struct DeltaSpecBody {
    static constexpr int NT = 256;
    static const char* name() { return "delta_spec"; }
    int Nx, Nyh, C, Nb; const T* lx; const T* ly;
    const C2<T>* A1; const C2<T>* A2; const C2<T>* S6; C2<T>* ddf; C2<T>* ddphi;
    DEV void operator()(int blk, unsigned char*) const {
        const size_t nf = (size_t)Nx * Nyh;
        CMBL_FOR_THREADS(tid, NT) {
            size_t e = (size_t)blk * NT + tid;
            if (e < nf * (size_t)(C + Nb)) {
                const size_t pl = e / nf, r = e - pl * nf;
                const int kx = (int)(r / Nyh), ky = (int)(r - (size_t)kx * Nyh);
                const C2<T> d1 = mk<T>((T)0, lx[kx]), d2 = mk<T>((T)0, ly[ky]);
                if (pl < (size_t)C) ddf[e] = cmul(d1, A1[e]) + cmul(d2, A2[e]);
                else {
                    const size_t b = pl - C;
                    const C2<T>* s = S6 + b * 5 * nf + r;
                    const C2<T> n1 = mk<T>((T)0, -lx[kx]), n2 = mk<T>((T)0, -ly[ky]);
                    C2<T> v = cmul(d1, s[0]) + cmul(d2, s[nf]);
                    v = v + cmul(n1, cmul(n1, s[2 * nf]));       // i=1 (m1), j=1 (p1)
                    v = v + cmul(n1, cmul(n2, s[3 * nf]));       // (i,j) = (1,2) and (2,1): rfft2 is linear, one transform for both
                    v = v + cmul(n2, cmul(n2, s[4 * nf]));       // i=2, j=2
                    ddphi[b * nf + r] = v;
                }
            }
        }
    }
};

// out = (x ? x : 0) + s*k on real arrays (complex arrays are passed as reals of twice the length); out may alias x or k
template <class T> struct AxpyBody {
    static constexpr int NT = 256;
    static const char* name() { return "axpy"; }
    size_t n; const T* x; const T* k; T s; T* out;
    DEV void operator()(int blk, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) {
            size_t e = (size_t)blk * NT + tid;
            if (e < n) out[e] = (x ? x[e] : (T)0) + s * k[e];
        }
    }
};
template <class T> static void axpy(size_t n, const T* x, const T* k, T s, T* out, cmblStream_t st) {
    AxpyBody<T> b{n, x, k, s, out};
    launch(b, (int)((n + b.NT - 1) / b.NT), 0, st);
}

// One RK4 stage update of the joint state (three components stored as real arrays):
//   acc = (acc_in ? acc_in : y) + cb·k ;  u = y + ca·k  (u == nullptr: not needed — last stage, or the δϕ component, which
//   never feeds the velocity).  k is read once for both results.
template <class T> struct RkUpdateBody {
    static constexpr int NT = 256;
    static const char* name() { return "rk_update"; }
    struct Seg { size_t n; const T* y; const T* acc_in; const T* k; T* acc; T* u; };
    Seg seg[3]; size_t off1, off2, total; T ca, cb;
    DEV void operator()(int blk, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) {
            size_t e = (size_t)blk * NT + tid;
            if (e < total) {
                const int si = e < off1 ? 0 : (e < off2 ? 1 : 2);
                const Seg& g = seg[si];
                const size_t i = e - (si == 0 ? 0 : (si == 1 ? off1 : off2));
                const T kk = g.k[i], yy = g.y[i];
                g.acc[i] = (g.acc_in ? g.acc_in[i] : yy) + cb * kk;
                if (g.u) g.u[i] = yy + ca * kk;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Fused transpose-δ flow (transform lengths with fast stage kernels).  Three observations turn the 12 two-dimensional transforms of one
// velocity evaluation into the stage kernels the plain flows already run:
//   * the f leg IS the forward flow (its velocity p₁∂ₓf + p₂∂ᵧf is what flow_rows + flow_cols compute), and the δf leg IS the adjoint
//     flow, which is integrated in map space (flow.cuh): Ł(δf) at a stage is simply the δf leg's stage map;
//   * the column kernel of the f leg has both derivative maps ∂ₓf, ∂ᵧf in registers in its epilogue and exports them (DMODE);
//   * δϕ never feeds back into the velocity: dδϕ/dt = g(t, f, δf).  RK4 is then a quadrature for it, and because every term of g is a
//     fixed linear operator applied to a MAP — iℓₓ·rfft2(m₁) + iℓᵧ·rfft2(m₂) + Σ_ij (−iℓ_i)(−iℓ_j)·rfft2(t p_j m_i) — the five maps are
//     accumulated over the 4n stages with the RK weights and transformed ONCE at the end: 5·Nb transforms per pullback instead of
//     5·Nb per stage.
// Per stage: forward stage kernels (f leg, derivative export), one pointwise kernel (w = Σ_pol Łδf·∇f, m = M⁻¹w, the five
// accumulators), adjoint stage kernels (δf leg).  Same trajectory as the reference's joint RK4 (src/flowops.jl:40-68).
// ---------------------------------------------------------------------------------------------------------------
template <class T> struct DeltaAccBody {
    static constexpr int NT = 256;
    static const char* name() { return "delta_acc"; }
    int Npol, Nb, Nbphi; size_t nmap; bool bug_compat; T t, wgt;
    const T* ud; const T* gx; const T* gy;           // [Nb][Npol] maps: δf leg stage state, ∂ₓf, ∂ᵧf
    const T* pk; const T* mk3;                        // p[k]: [Nbphi][2], M⁻¹[k]: [Nbphi][3]  (every map of this kernel in the same layout)
    T* A;                                             // [Nb][5] accumulators: m1, m2, t p1 m1, t (p2 m1 + p1 m2), t p2 m2
    DEV void operator()(int blk, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) {
            size_t e = (size_t)blk * NT + tid;
            if (e < nmap * Nb) {
                const size_t b = e / nmap, r = e - b * nmap;
                const size_t bp = (Nbphi == 1) ? 0 : b;
                const T p1 = pk[(bp * 2 + 0) * nmap + r], p2 = pk[(bp * 2 + 1) * nmap + r];
                const T m11 = mk3[(bp * 3 + 0) * nmap + r], m21 = mk3[(bp * 3 + 1) * nmap + r], m22 = mk3[(bp * 3 + 2) * nmap + r];
                T w1 = 0, w2 = 0;
                for (int pol = 0; pol < Npol; ++pol) {
                    const size_t i = (b * Npol + pol) * nmap + r;
                    const T l = ud[i];
                    w1 += l * gx[i]; w2 += l * gy[i];
                }
                const T m1 = m11 * w1 + m21 * w2;                                      // M₁₂ ≡ M₂₁
                const T m2 = bug_compat ? (m21 * m1 + m22 * w2) : (m21 * w1 + m22 * w2);
                T* a = A + b * 5 * nmap + r;
                a[0] += wgt * m1; a[nmap] += wgt * m2;
                a[2 * nmap] += wgt * (t * p1 * m1); a[3 * nmap] += wgt * (t * p2 * m1 + t * p1 * m2); a[4 * nmap] += wgt * (t * p2 * m2);
            }
        }
    }
};

// δϕ = iℓₓ·S0 + iℓᵧ·S1 + (−iℓₓ)²·S2 + (−iℓₓ)(−iℓᵧ)·S3 + (−iℓᵧ)²·S4 with S = rfft2 of the five accumulators
template <class T> struct DeltaPhiSpecBody {
    static constexpr int NT = 256;
    static const char* name() { return "delta_phi_spec"; }
    int Nx, Nyh, Nb; const T* lx; const T* ly; const C2<T>* S; C2<T>* dphi;
    DEV void operator()(int blk, unsigned char*) const {
        const size_t nf = (size_t)Nx * Nyh;
        CMBL_FOR_THREADS(tid, NT) {
            size_t e = (size_t)blk * NT + tid;
            if (e < nf * (size_t)Nb) {
                const size_t b = e / nf, r = e - b * nf;
                const int kx = (int)(r / Nyh), ky = (int)(r - (size_t)kx * Nyh);
                const C2<T> d1 = mk<T>((T)0, lx[kx]), d2 = mk<T>((T)0, ly[ky]), n1 = mk<T>((T)0, -lx[kx]), n2 = mk<T>((T)0, -ly[ky]);
                const C2<T>* s = S + b * 5 * nf + r;
                C2<T> v = cmul(d1, s[0]) + cmul(d2, s[nf]);
                v = v + cmul(n1, cmul(n1, s[2 * nf]));
                v = v + cmul(n1, cmul(n2, s[3 * nf]));
                v = v + cmul(n2, cmul(n2, s[4 * nf]));
                dphi[e] = v;
            }
        }
    }
};

template <class T> static void flow_grad_fused(FlowT<T>& F, int op, const T* fout, const C2<T>* delta, C2<T>* dfield, C2<T>* dphi,
                                               bool bug_compat, cmblStream_t st) {
    PlanT<T>& P = *F.P;
    const size_t nmap = P.map_elems(), nf = P.four_elems();
    const int C = F.C, Nb = F.Nb, n = F.nsteps, G = flow_rg_rows(P);
    const size_t mC = nmap * C;
    CMBL_REQUIRE(G > 0 && G == F.pcache_G, "fused transpose-δ flow needs the row-grouped caches");
    flow_reserve(F, st);
    T* f3[3]; T* d3[3];                               // (y, acc, u) of the two legs, row-grouped
    for (int i = 0; i < 3; ++i) { f3[i] = (T*)F.gq_f[i].reserve(sizeof(T) * mC); d3[i] = (T*)F.gq_d[i].reserve(sizeof(T) * mC); }
    const int Gd = flow_rg_direct(F);               // the transforms address the row-grouped buffers directly where they can
    T* ref = Gd ? nullptr : (T*)F.gq_ref.reserve(sizeof(T) * mC);                     // a reference-layout map (C planes) for the conversions
    T* gx = (T*)F.gq_gx.reserve(sizeof(T) * mC); T* gy = (T*)F.gq_gy.reserve(sizeof(T) * mC);
    T* A = (T*)F.gq_A.reserve(sizeof(T) * 5 * nmap * Nb); T* Aref = Gd ? nullptr : (T*)F.gq_Aref.reserve(sizeof(T) * 5 * nmap * Nb);
    C2<T>* spec = (C2<T>*)F.gq_spec.reserve(sizeof(C2<T>) * 5 * nf * Nb);
    // initial state: f leg = the forward result, δf leg = irfft2(Δ) (+ what a map cannot carry, flow.cuh), δϕ integrand accumulators = 0
    convert_layout<T, true>(P, G, fout, f3[0], C, st);
    if (Gd) flow_adj_prepare<T>(F, delta, d3[0], st, Gd);
    else { flow_adj_prepare<T>(F, delta, ref, st); convert_layout<T, true>(P, G, ref, d3[0], C, st); }
    dev_zero(A, sizeof(T) * 5 * nmap * Nb, st);
    const int k0 = (op == CMBL_OP_L) ? 2 * n : 0, k1 = (op == CMBL_OP_L) ? 0 : 2 * n;
    const int sgn = k1 > k0 ? 1 : -1;
    const double h = (double)sgn / n;
    const T h2 = (T)(h / 2), h1 = (T)h, h6 = (T)(h / 6), h3 = (T)(h / 3);
    const T* mk3 = reinterpret_cast<const T*>(F.minv.p);
    int kk = k0;
    for (int step = 0; step < n; ++step) {
        for (int s = 0; s < 4; ++s) {
            const int kq = kk + (s == 0 ? 0 : (s < 3 ? sgn : 2 * sgn));
            const T ca = (s < 2) ? h2 : h1, cb = (s == 0 || s == 3) ? h6 : h3;
            auto stage = [&](T** L3, bool adj, T* dxo, T* dyo) {
                T* y = L3[0]; T* acc = L3[1]; T* ub = L3[2];
                const T* u = (s == 0) ? y : ub; const T* ybase = (s == 3) ? nullptr : y; const T* acc_in = (s == 0) ? nullptr : acc;
                T* acc_out = (s == 3) ? y : acc; T* u_out = (s == 3) ? nullptr : ub;
                if (adj) flow_stage<T, true>(F, 0, C, u, kq, cb, ybase, acc_in, acc_out, u_out, ca, cb, st);
                else flow_stage<T, false>(F, 0, C, u, kq, cb, ybase, acc_in, acc_out, u_out, ca, cb, st, dxo, dyo);
            };
            const T* ud = (s == 0) ? d3[0] : d3[2];                                   // δf leg's stage state, read BEFORE that leg advances
            stage(f3, false, gx, gy);
            {
                DeltaAccBody<T> b;
                b.Npol = F.Npol; b.Nb = Nb; b.Nbphi = F.Nbphi; b.nmap = nmap; b.bug_compat = bug_compat;
                b.t = (T)((double)kq / (double)(2 * n)); b.wgt = cb;
                b.ud = ud; b.gx = gx; b.gy = gy; b.pk = F.pk(kq); b.mk3 = mk3 + (size_t)kq * F.Nbphi * 3 * nmap; b.A = A;
                launch(b, (int)((nmap * Nb + b.NT - 1) / b.NT), 0, st);
            }
            stage(d3, true, nullptr, nullptr);
        }
        kk += 2 * sgn;
    }
    // δf: back to the reference layout, rfft2, restore the rows / Nyquist terms a map cannot carry
    if (Gd) flow_adj_finish<T>(F, d3[0], dfield, st, Gd);
    else { convert_layout<T, false>(P, G, d3[0], ref, C, st); flow_adj_finish<T>(F, ref, dfield, st); }
    // δϕ: the five accumulated maps, transformed once
    if (Gd) rfft2<T>(P, A, spec, 5 * Nb, st, Gd);
    else { convert_layout<T, false>(P, G, A, Aref, 5 * Nb, st); rfft2<T>(P, Aref, spec, 5 * Nb, st); }
    DeltaPhiSpecBody<T> b{P.Nx, P.Nyh, Nb, P.lx, P.ly, spec, dphi};
    launch(b, (int)((nf * (size_t)Nb + b.NT - 1) / b.NT), 0, st);
}

static bool grad_fused_enabled() { static const bool v = [] { const char* e = getenv("CMBL_GRAD_FUSED"); return !e || atoi(e) != 0; }(); return v; }

template <class T> static void flow_grad_general(FlowT<T>& F, int op, const T* fout, const C2<T>* delta, C2<T>* dfield, C2<T>* dphi,
                                                 bool bug_compat, cmblStream_t st);

template <class T> void flow_grad(FlowT<T>& F, int op, const T* fout, const C2<T>* delta, C2<T>* dfield, C2<T>* dphi,
                                  bool bug_compat, cmblStream_t st) {
    CMBL_REQUIRE(F.have_p && F.have_minv, "cmbl_lenseflow_grad needs cmbl_lenseflow_precompute(..., with_minv = 1)");
    if (grad_fused_enabled() && flow_rg_rows(*F.P) > 0 && F.pcache_G > 0) flow_grad_fused<T>(F, op, fout, delta, dfield, dphi, bug_compat, st);
    else flow_grad_general<T>(F, op, fout, delta, dfield, dphi, bug_compat, st);
}

// general path (any power-of-two size): one velocity evaluation = 12 batched 2-D transforms + three pointwise kernels
template <class T> static void flow_grad_general(FlowT<T>& F, int op, const T* fout, const C2<T>* delta, C2<T>* dfield, C2<T>* dphi,
                                                 bool bug_compat, cmblStream_t st) {
    PlanT<T>& P = *F.P;
    const size_t nmap = P.map_elems(), nf = P.four_elems();
    const int C = F.C, Nb = F.Nb, n = F.nsteps;
    const size_t mC = nmap * C, fC = nf * C, fB = nf * Nb;
    // scratch: state copies (y: f, δf, δϕ), stage inputs u, accumulators acc, velocities k, and the work arrays of one evaluation
    auto R = [&](DevBuf& b, size_t bytes) { return b.reserve(bytes); };
    T* yf = (T*)R(F.g_yf, sizeof(T) * mC);           C2<T>* yd = (C2<T>*)R(F.g_yd, sizeof(C2<T>) * fC);   C2<T>* yp = (C2<T>*)R(F.g_yp, sizeof(C2<T>) * fB);
    T* uf = (T*)R(F.g_uf, sizeof(T) * mC);           C2<T>* ud = (C2<T>*)R(F.g_ud, sizeof(C2<T>) * fC);
    T* af = (T*)R(F.g_af, sizeof(T) * mC);           C2<T>* ad = (C2<T>*)R(F.g_ad, sizeof(C2<T>) * fC);   C2<T>* ap = (C2<T>*)R(F.g_ap, sizeof(C2<T>) * fB);
    T* kf = (T*)R(F.g_kf, sizeof(T) * mC);           C2<T>* kd = (C2<T>*)R(F.g_kd, sizeof(C2<T>) * fC);   C2<T>* kp = (C2<T>*)R(F.g_kp, sizeof(C2<T>) * fB);
    T* ldf = (T*)R(F.g_ldf, sizeof(T) * mC);         T* gxy = (T*)R(F.g_gxy, sizeof(T) * 2 * mC);         T* a12 = (T*)R(F.g_a12, sizeof(T) * 2 * mC);
    T* six = (T*)R(F.g_six, sizeof(T) * 5 * nmap * Nb);
    C2<T>* spec = (C2<T>*)R(F.g_spec, sizeof(C2<T>) * (2 * fC > 5 * fB ? 2 * fC : 5 * fB));
    C2<T>* spec2 = (C2<T>*)R(F.g_spec2, sizeof(C2<T>) * 2 * fC);

    auto velocity = [&](int k, const T* f, const C2<T>* df, T* of, C2<T>* od, C2<T>* ophi) {
        const T t = (T)((double)k / (double)(2 * n));
        irfft2<T>(P, df, ldf, C, st);
        rfft2<T>(P, f, spec, C, st);
        {
            GradSpecBody<T> b{P.Nx, P.Nyh, P.lx, P.ly, spec, spec2, spec2 + fC, fC};
            launch(b, (int)((fC + b.NT - 1) / b.NT), 0, st);
        }
        irfft2<T>(P, spec2, gxy, 2 * C, st);
        {
            DeltaPointBody<T> b;
            b.Npol = F.Npol; b.Nb = Nb; b.Nbphi = F.Nbphi; b.Nx = P.Nx; b.Ny = P.Ny; b.G = F.pcache_G; b.bug_compat = bug_compat; b.t = t;
            b.ldf = ldf; b.gx = gxy; b.gy = gxy + mC;
            b.pk = F.pk(k); b.mk3 = reinterpret_cast<T*>(F.minv.p) + (size_t)k * F.Nbphi * 3 * nmap;
            b.dfdt = of; b.a1 = a12; b.a2 = a12 + mC; b.six = six;
            launch(b, (int)((nmap * Nb + b.NT - 1) / b.NT), 0, st);
        }
        rfft2<T>(P, a12, spec2, 2 * C, st);
        rfft2<T>(P, six, spec, 5 * Nb, st);
        {
            DeltaSpecBody<T> b{P.Nx, P.Nyh, C, Nb, P.lx, P.ly, spec2, spec2 + fC, spec, od, ophi};
            launch(b, (int)((nf * (size_t)(C + Nb) + b.NT - 1) / b.NT), 0, st);
        }
    };

    dev_copy(yf, fout, sizeof(T) * mC, st);
    dev_copy(yd, delta, sizeof(C2<T>) * fC, st);
    dev_zero(yp, sizeof(C2<T>) * fB, st);
    const int k0 = (op == CMBL_OP_L) ? 2 * n : 0, k1 = (op == CMBL_OP_L) ? 0 : 2 * n;
    const int sgn = k1 > k0 ? 1 : -1;
    const double h = (double)sgn / n;
    const T h2 = (T)(h / 2), h1 = (T)h, h6 = (T)(h / 6), h3 = (T)(h / 3);
    auto rr = [](C2<T>* p) { return reinterpret_cast<T*>(p); };
    int kk = k0;
    for (int step = 0; step < n; ++step) {
        for (int s = 0; s < 4; ++s) {
            const int kq = kk + (s == 0 ? 0 : (s < 3 ? sgn : 2 * sgn));
            const T* f_in = (s == 0) ? yf : uf; const C2<T>* d_in = (s == 0) ? yd : ud;
            velocity(kq, f_in, d_in, kf, kd, kp);
            const T cb = (s == 0 || s == 3) ? h6 : h3, ca = (s < 2) ? h2 : h1;
            {
                // stages 1-3: acc = (s == 0 ? y : acc) + cb k, u = y + ca k;  stage 4: y = acc + cb k  (written as acc := y_out)
                RkUpdateBody<T> b;
                const bool last = (s == 3);
                b.seg[0] = {mC, last ? af : yf, (s == 0 || last) ? nullptr : af, kf, last ? yf : af, last ? nullptr : uf};
                b.seg[1] = {2 * fC, last ? rr(ad) : rr(yd), (s == 0 || last) ? nullptr : rr(ad), rr(kd), last ? rr(yd) : rr(ad), last ? nullptr : rr(ud)};
                b.seg[2] = {2 * fB, last ? rr(ap) : rr(yp), (s == 0 || last) ? nullptr : rr(ap), rr(kp), last ? rr(yp) : rr(ap), nullptr};
                b.off1 = mC; b.off2 = mC + 2 * fC; b.total = mC + 2 * fC + 2 * fB; b.ca = ca; b.cb = cb;
                launch(b, (int)((b.total + b.NT - 1) / b.NT), 0, st);
            }
        }
        kk += 2 * sgn;
    }
    dev_copy(dfield, yd, sizeof(C2<T>) * fC, st);
    dev_copy(dphi, yp, sizeof(C2<T>) * fB, st);
}

template void flow_grad<float>(FlowT<float>&, int, const float*, const C2<float>*, C2<float>*, C2<float>*, bool, cmblStream_t);
template void flow_grad<double>(FlowT<double>&, int, const double*, const C2<double>*, C2<double>*, C2<double>*, bool, cmblStream_t);

}  // namespace cmbl
