// Elementwise / reduction kernels: DiagOp products (src/specialops.jl:9-10), QU<->EB rotation
// (src/proj_lambert.jl:253-271), dot (src/proj_lambert.jl:318-328) and the fused vector updates of
// conjugate_gradient (src/numerical_algorithms.jl:99-108).
#pragma once
#include "plan.cuh"

namespace cmbl {


template <class T> HD T nan2zero(T v) { return (v - v == (T)0) ? v : (T)0; }          // isfinite(v) ? v : 0
template <class T> HD C2<T> nan2zero(C2<T> v) {                                       // complex: both parts finite
    bool fin = (v.x - v.x == (T)0) && (v.y - v.y == (T)0);
    return fin ? v : mk<T>((T)0, (T)0);
}

// ---------------------------------------------------------------------------------------------------------------
// Cℓ_to_2D / Cℓ_to_Cov (src/proj_lambert.jl:173-175,361-364): out[kx][ky] = nan2zero(Cℓ(ℓmag)) / units with Cℓ the linear
// interpolation of the table (ell[n] ascending, cl[n]; NaN outside it, src/numerical_algorithms.jl:148-177) and ℓmag = √(ℓx²+ℓy²) in T.
// ---------------------------------------------------------------------------------------------------------------
template <class T> struct ClTo2DBody {
    static constexpr int NT = 256;
    static const char* name() { return "cl_to_2d"; }
    int Nx, Nyh, n; const T* lx; const T* ly; const double* ell; const double* cl; T units; T* out;
    DEV void operator()(int blk, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) {
            const size_t e = (size_t)blk * NT + tid;
            if (e < (size_t)Nx * Nyh) {
                const int kx = (int)(e / Nyh), ky = (int)(e - (size_t)kx * Nyh);
                const T lm = sqrt(lx[kx] * lx[kx] + ly[ky] * ly[ky]);
                const double x = (double)lm;
                double v = 0;                                                  // outside the table: NaN -> nan2zero -> 0
                if (x >= ell[0] && x <= ell[n - 1]) {
                    int lo = 0, hi = n - 1;                                    // last i with ell[i] <= x
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (ell[mid] <= x) lo = mid; else hi = mid; }
                    if (lo == n - 1 || ell[lo] == x) v = cl[lo];
                    else v = (cl[lo + 1] - cl[lo]) / (ell[lo + 1] - ell[lo]) * (x - ell[lo]) + cl[lo];
                    if (!(v - v == 0)) v = 0;
                }
                out[e] = (T)v / units;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// out = a ⊙ x + b ⊙ y on fields, a and b per-batch scalars (BatchedReal broadcasts of the reference: `@. x + α*Δ`, the RK / leap-frog
// axpys of the callers, src/batching.jl:9-45).  Arrays are addressed as reals (a complex field is 2·n reals); y may be NULL (b ignored).
// ---------------------------------------------------------------------------------------------------------------
constexpr int AXPBY_MAX_NB = 64;
template <class T> struct AxpbyBody {
    static constexpr int NT = 256;
    static const char* name() { return "field_axpby"; }
    size_t per_batch, total; const T* x; const T* y; T* out; int na, nb;
    double a[AXPBY_MAX_NB], b[AXPBY_MAX_NB];
    DEV void operator()(int blk, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) {
            const size_t e = (size_t)blk * NT + tid;
            if (e < total) {
                const size_t bi = e / per_batch;
                const T aa = (T)a[na == 1 ? 0 : bi];
                T v = aa * x[e];
                if (y) v += (T)b[nb == 1 ? 0 : bi] * y[e];
                out[e] = v;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// DiagOp * f , DiagOp \ f
// ---------------------------------------------------------------------------------------------------------------
template <class T, bool CPLX> struct DiagMulBody {
    static constexpr int NT = 256;
    static const char* name() { return "diag_mul"; }
    size_t plane, total; int Cd; bool ldiv;
    const T* diag; const void* in; void* out;
    DEV void operator()(int blk, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) {
            size_t e = (size_t)blk * NT + tid;
            if (e < total) {
                size_t c = e / plane, r = e - c * plane;
                T dg = diag[(c % Cd) * plane + r];
                if (CPLX) {
                    C2<T> v = reinterpret_cast<const C2<T>*>(in)[e];
                    v = ldiv ? nan2zero(mk<T>(v.x / dg, v.y / dg)) : mk<T>(dg * v.x, dg * v.y);
                    reinterpret_cast<C2<T>*>(out)[e] = v;
                } else {
                    T v = reinterpret_cast<const T*>(in)[e];
                    reinterpret_cast<T*>(out)[e] = ldiv ? nan2zero(v / dg) : dg * v;
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// fused Fourier-space chain on harmonic-basis fields (Npol = 1: no rotation; Npol = 2: planes (E,B) / (Q,U);
// Npol = 3: planes (I,E,B) / (I,Q,U), the rotation acts on planes 2:3, src/proj_lambert.jl:284,292):
//   v = in;  v *= din;  v = d − v (or −v when neg);  v *= pre;  v *= pre2;  v = Rot(v);  v *= post;  v −= sdiag·sub;  v = Rot2(v);  out = v
// Npol ≤ 2: every diagonal is REAL with Npol planes shared across the batch (NULL = skip).
// Npol = 3: every operator is a BlockDiagIEB (src/specialops.jl:61-82) of 4 REAL planes [ΣTE[1,1], ΣTE[2,1], ΣTE[2,2], ΣB]
//           applied to (I,E,B):  i′ = A11 i + A21 e,  e′ = A21 i + A22 e,  b′ = ΣB b   (the 2×2 block is symmetric).
// rot: 0 none, 1 EB→QU, 2 QU→EB  (src/proj_lambert.jl:253-271)
// ---------------------------------------------------------------------------------------------------------------
template <class T> struct FourierChainBody {
    static constexpr int NT = 256;
    static const char* name() { return "fourier_chain"; }
    int Npol, Nb; size_t nf; int rot; bool neg;
    const T *sin2phi, *cos2phi;
    const C2<T>* in; const T* din; const C2<T>* d; const T* pre; const T* post; const T* sdiag; const C2<T>* sub; C2<T>* out;
    const T* pre2 = nullptr;                          // second pre-rotation operator (Npol = 3 only; diagonals are fused instead)
    int rot2 = 0;                                     // second rotation applied after `post` (QU→EB, ×B, EB→QU in one pass)
    HD void block(const T* A, size_t r, C2<T>& i, C2<T>& e, C2<T>& b) const {
        const T a11 = A[r], a21 = A[nf + r], a22 = A[2 * nf + r], bb = A[3 * nf + r];
        const C2<T> ni = mk<T>(a11 * i.x + a21 * e.x, a11 * i.y + a21 * e.y);
        const C2<T> ne = mk<T>(a21 * i.x + a22 * e.x, a21 * i.y + a22 * e.y);
        i = ni; e = ne; b = cscale(b, bb);
    }
    HD void rotate(int dir, size_t r, C2<T>& a, C2<T>& c) const {
        const T s = sin2phi[r], co = cos2phi[r];
        C2<T> na, nc;
        if (dir == 1) {      // Q = −E c + B s ; U = −E s − B c
            na = mk<T>(-a.x * co + c.x * s, -a.y * co + c.y * s);
            nc = mk<T>(-a.x * s - c.x * co, -a.y * s - c.y * co);
        } else {             // E = −Q c − U s ; B = Q s − U c
            na = mk<T>(-a.x * co - c.x * s, -a.y * co - c.y * s);
            nc = mk<T>(a.x * s - c.x * co, a.y * s - c.y * co);
        }
        a = na; c = nc;
    }
    HD C2<T> head(C2<T> v, size_t r, int pol, size_t e) const {
        if (din) v = cscale(v, din[pol * nf + r]);
        if (d) v = d[e] - v; else if (neg) v = mk<T>(-v.x, -v.y);
        if (pre) v = cscale(v, pre[pol * nf + r]);
        return v;
    }
    HD C2<T> tail(C2<T> v, size_t r, int pol, size_t e) const {
        if (post) v = cscale(v, post[pol * nf + r]);
        if (sub) v = v - cscale(sub[e], sdiag[pol * nf + r]);
        return v;
    }
    DEV void operator()(int blk, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) {
            size_t t = (size_t)blk * NT + tid;
            if (Npol == 2) {
                if (t < nf * Nb) {
                    size_t b = t / nf, r = t - b * nf;
                    size_t e0 = (b * 2) * nf + r, e1 = e0 + nf;
                    C2<T> a = head(in[e0], r, 0, e0), c = head(in[e1], r, 1, e1);
                    if (rot) rotate(rot, r, a, c);
                    a = tail(a, r, 0, e0); c = tail(c, r, 1, e1);
                    if (rot2) rotate(rot2, r, a, c);
                    out[e0] = a; out[e1] = c;
                }
            } else if (Npol == 3) {
                if (t < nf * Nb) {
                    size_t b = t / nf, r = t - b * nf;
                    size_t e0 = (b * 3) * nf + r, e1 = e0 + nf, e2 = e1 + nf;
                    C2<T> i = in[e0], a = in[e1], c = in[e2];
                    if (din) block(din, r, i, a, c);
                    if (d) { i = d[e0] - i; a = d[e1] - a; c = d[e2] - c; }
                    else if (neg) { i = mk<T>(-i.x, -i.y); a = mk<T>(-a.x, -a.y); c = mk<T>(-c.x, -c.y); }
                    if (pre) block(pre, r, i, a, c);
                    if (pre2) block(pre2, r, i, a, c);
                    if (rot) rotate(rot, r, a, c);
                    if (post) block(post, r, i, a, c);
                    if (sub) {
                        C2<T> si = sub[e0], sa = sub[e1], sc = sub[e2];
                        block(sdiag, r, si, sa, sc);
                        i = i - si; a = a - sa; c = c - sc;
                    }
                    if (rot2) rotate(rot2, r, a, c);
                    out[e0] = i; out[e1] = a; out[e2] = c;
                }
            } else {
                if (t < nf * Nb * Npol) {
                    size_t cpl = t / nf, r = t - cpl * nf;
                    int pol = (int)(cpl % Npol);
                    out[t] = tail(head(in[t], r, pol, t), r, pol, t);
                }
            }
        }
    }
};

// BlockDiagIEB applied to an IEBFourier field (src/specialops.jl:77-82,87-88): mode 0  L*f, 1  L\f = pinv(L)*f, 2  sqrt(L)*f.
// block = 4 REAL half-planes [ΣTE[1,1], ΣTE[2,1], ΣTE[2,2], ΣB]; the 2×2 pinv / sqrt follow src/field_vectors.jl:62-78
// (the off-diagonal is read from [2,1] for both positions).
template <class T> struct BlockIebBody {
    static constexpr int NT = 256;
    static const char* name() { return "blockdiag_ieb"; }
    size_t nf; int Nb, mode; const T* A; const C2<T>* in; C2<T>* out;
    HD static T pinv(T v) { return v == (T)0 ? (T)0 : (T)1 / v; }
    DEV void operator()(int blk, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) {
            size_t t = (size_t)blk * NT + tid;
            if (t < nf * Nb) {
                size_t b = t / nf, r = t - b * nf;
                T a = A[r], c = A[nf + r], d = A[2 * nf + r], e = A[3 * nf + r];
                if (mode == 1) { const T id = pinv(a * d - c * c); const T na = d * id, nc = -(c * id), nd = a * id; a = na; c = nc; d = nd; e = pinv(e); }
                else if (mode == 2) {
                    const T s = sqrt(a * d - c * c), tt = pinv(sqrt(a + (d + 2 * s)));
                    const T na = tt * (a + s), nc = tt * c, nd = tt * (d + s); a = na; c = nc; d = nd; e = sqrt(e);
                }
                const size_t e0 = (b * 3) * nf + r, e1 = e0 + nf, e2 = e1 + nf;
                const C2<T> i = in[e0], ee = in[e1], bb = in[e2];
                out[e0] = mk<T>(a * i.x + c * ee.x, a * i.y + c * ee.y);
                out[e1] = mk<T>(c * i.x + d * ee.x, c * i.y + d * ee.y);
                out[e2] = cscale(bb, e);
            }
        }
    }
};

// standalone rotation with arbitrary plane stride (cmbl_qu_eb)
template <class T> struct QuEbBody {
    static constexpr int NT = 256;
    static const char* name() { return "qu_eb"; }
    int Nb, stride_planes, first_plane, dir; size_t nf;
    const T *sin2phi, *cos2phi; const C2<T>* in; C2<T>* out;
    DEV void operator()(int blk, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) {
            size_t t = (size_t)blk * NT + tid;
            if (t < nf * Nb) {
                size_t b = t / nf, r = t - b * nf;
                size_t e0 = (b * stride_planes + first_plane) * nf + r, e1 = e0 + nf;
                C2<T> a = in[e0], c = in[e1];
                T s = sin2phi[r], co = cos2phi[r];
                if (dir == 0) { out[e0] = mk<T>(-a.x * co + c.x * s, -a.y * co + c.y * s); out[e1] = mk<T>(-a.x * s - c.x * co, -a.y * s - c.y * co); }
                else { out[e0] = mk<T>(-a.x * co - c.x * s, -a.y * co - c.y * s); out[e1] = mk<T>(a.x * s - c.x * co, a.y * s - c.y * co); }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// dot: partial[b][j] (double), j < RED_BLOCKS.   Fourier: Σ Re(conj(a) b) λ[ky] / (Ny Nx);  Map: Σ a b
// ---------------------------------------------------------------------------------------------------------------
template <class T, bool CPLX> struct DotBody {
    static constexpr int NT = 256;
    static const char* name() { return "dot"; }
    size_t per_batch; int Nyh; const T* lam; double scale;
    const void* a; const void* b; double* partial;
    DEV void operator()(int blk, unsigned char* smem) const {
        double* sm = reinterpret_cast<double*>(smem);
        const int bi = blk / RED_BLOCKS, j = blk % RED_BLOCKS;
        const size_t base = (size_t)bi * per_batch;
        CMBL_FOR_THREADS(tid, NT) {
            double s = 0;
            for (size_t e = (size_t)j * NT + tid; e < per_batch; e += (size_t)RED_BLOCKS * NT) {
                if (CPLX) {
                    C2<T> x = reinterpret_cast<const C2<T>*>(a)[base + e], y = reinterpret_cast<const C2<T>*>(b)[base + e];
                    s += ((double)x.x * (double)y.x + (double)x.y * (double)y.y) * (double)lam[e % Nyh];
                } else {
                    s += (double)reinterpret_cast<const T*>(a)[base + e] * (double)reinterpret_cast<const T*>(b)[base + e];
                }
            }
            sm[tid] = s;
        }
        CMBL_SYNC();
        CMBL_FOR_THREADS(tid, NT) {
            if (tid == 0) { double s = 0; for (int i = 0; i < NT; ++i) s += sm[i]; partial[blk] = s * scale; }
        }
    }
};

HD double sum_partials(const double* p) { double s = 0; for (int i = 0; i < RED_BLOCKS; ++i) s += p[i]; return s; }

// ---------------------------------------------------------------------------------------------------------------
// CG update 1 (numerical_algorithms.jl:101-105):  α = res/Σ pAp;  x += α p;  r −= α Ap;  z = M \ r;  partial(res′ = r·z)
// ---------------------------------------------------------------------------------------------------------------
template <class T> struct CgUpdate1Body {
    static constexpr int NT = 256;
    static const char* name() { return "cg_update1"; }
    size_t per_batch, nf; int Npol, Nyh; const T* lam; double scale;
    const double* res; const double* pAp_part; const T* Mdiag;
    const C2<T>* p; const C2<T>* Ap; C2<T>* x; C2<T>* r; C2<T>* z; double* res_part;
    DEV void operator()(int blk, unsigned char* smem) const {
        double* sm = reinterpret_cast<double*>(smem);
        const int bi = blk / RED_BLOCKS, j = blk % RED_BLOCKS;
        const size_t base = (size_t)bi * per_batch;
        CMBL_FOR_THREADS(tid, NT) {
            const T alpha = (T)(res[bi] / sum_partials(pAp_part + (size_t)bi * RED_BLOCKS));
            double s = 0;
            if (Npol == 3) {
                // BlockDiagIEB preconditioner: z = pinv(M) * r (src/specialops.jl:78); Mdiag = the 4 planes of pinv(M)
                for (size_t e = (size_t)j * NT + tid; e < nf; e += (size_t)RED_BLOCKS * NT) {
                    C2<T> rv[3];
                    for (int c = 0; c < 3; ++c) {
                        const size_t g = base + c * nf + e;
                        C2<T> pv = p[g], av = Ap[g], xv = x[g], rr = r[g];
                        x[g] = mk<T>(xv.x + alpha * pv.x, xv.y + alpha * pv.y);
                        rv[c] = mk<T>(rr.x - alpha * av.x, rr.y - alpha * av.y);
                        r[g] = rv[c];
                    }
                    const T a11 = Mdiag[e], a21 = Mdiag[nf + e], a22 = Mdiag[2 * nf + e], bb = Mdiag[3 * nf + e];
                    C2<T> zv[3] = {mk<T>(a11 * rv[0].x + a21 * rv[1].x, a11 * rv[0].y + a21 * rv[1].y),
                                   mk<T>(a21 * rv[0].x + a22 * rv[1].x, a21 * rv[0].y + a22 * rv[1].y), cscale(rv[2], bb)};
                    double q = 0;
                    for (int c = 0; c < 3; ++c) { z[base + c * nf + e] = zv[c]; q += (double)rv[c].x * (double)zv[c].x + (double)rv[c].y * (double)zv[c].y; }
                    s += q * (double)lam[e % Nyh];
                }
            } else
            for (size_t e = (size_t)j * NT + tid; e < per_batch; e += (size_t)RED_BLOCKS * NT) {
                C2<T> pv = p[base + e], av = Ap[base + e], xv = x[base + e], rv = r[base + e];
                xv = mk<T>(xv.x + alpha * pv.x, xv.y + alpha * pv.y);
                rv = mk<T>(rv.x - alpha * av.x, rv.y - alpha * av.y);
                T m = Mdiag[e];                                   // Npol planes, shared across the batch
                C2<T> zv = nan2zero(mk<T>(rv.x / m, rv.y / m));
                x[base + e] = xv; r[base + e] = rv; z[base + e] = zv;
                s += ((double)rv.x * (double)zv.x + (double)rv.y * (double)zv.y) * (double)lam[e % Nyh];
            }
            sm[tid] = s;
        }
        CMBL_SYNC();
        CMBL_FOR_THREADS(tid, NT) {
            if (tid == 0) { double s = 0; for (int i = 0; i < NT; ++i) s += sm[i]; res_part[blk] = s * scale; }
        }
    }
};

// CG update 2 (:106-107):  res′ = Σ partial;  p = z + (res′/res) p;  res_out[b] = res′
template <class T> struct CgUpdate2Body {
    static constexpr int NT = 256;
    static const char* name() { return "cg_update2"; }
    size_t per_batch; const double* res; const double* res_part; double* res_out;
    const C2<T>* z; C2<T>* p;
    DEV void operator()(int blk, unsigned char*) const {
        const int bi = blk / RED_BLOCKS, j = blk % RED_BLOCKS;
        const size_t base = (size_t)bi * per_batch;
        CMBL_FOR_THREADS(tid, NT) {
            const double rn = sum_partials(res_part + (size_t)bi * RED_BLOCKS);
            const T beta = (T)(rn / res[bi]);
            for (size_t e = (size_t)j * NT + tid; e < per_batch; e += (size_t)RED_BLOCKS * NT) {
                C2<T> zv = z[base + e], pv = p[base + e];
                p[base + e] = mk<T>(zv.x + beta * pv.x, zv.y + beta * pv.y);
            }
            if (j == 0 && tid == 0) res_out[bi] = rn;
        }
    }
};

// r = b − Ax ; z = M \ r ; p = z ; partial(res = r·z)   (numerical_algorithms.jl:89-92)
template <class T> struct CgInitBody {
    static constexpr int NT = 256;
    static const char* name() { return "cg_init"; }
    size_t per_batch; int Nyh; const T* lam; double scale; const T* Mdiag;
    const C2<T>* b; const C2<T>* Ax; C2<T>* r; C2<T>* z; C2<T>* p; double* res_part;
    size_t nf_block = 0;                              // > 0: BlockDiagIEB preconditioner (Npol = 3), Mdiag = 4 planes of pinv(M)
    DEV void operator()(int blk, unsigned char* smem) const {
        double* sm = reinterpret_cast<double*>(smem);
        const int bi = blk / RED_BLOCKS, j = blk % RED_BLOCKS;
        const size_t base = (size_t)bi * per_batch;
        CMBL_FOR_THREADS(tid, NT) {
            double s = 0;
            if (nf_block) {
                const size_t nf = nf_block;
                for (size_t e = (size_t)j * NT + tid; e < nf; e += (size_t)RED_BLOCKS * NT) {
                    C2<T> rv[3];
                    for (int c = 0; c < 3; ++c) {
                        const size_t g = base + c * nf + e;
                        rv[c] = b[g];
                        if (Ax) rv[c] = rv[c] - Ax[g];
                        r[g] = rv[c];
                    }
                    const T a11 = Mdiag[e], a21 = Mdiag[nf + e], a22 = Mdiag[2 * nf + e], bb = Mdiag[3 * nf + e];
                    C2<T> zv[3] = {mk<T>(a11 * rv[0].x + a21 * rv[1].x, a11 * rv[0].y + a21 * rv[1].y),
                                   mk<T>(a21 * rv[0].x + a22 * rv[1].x, a21 * rv[0].y + a22 * rv[1].y), cscale(rv[2], bb)};
                    double q = 0;
                    for (int c = 0; c < 3; ++c) { z[base + c * nf + e] = zv[c]; p[base + c * nf + e] = zv[c]; q += (double)rv[c].x * (double)zv[c].x + (double)rv[c].y * (double)zv[c].y; }
                    s += q * (double)lam[e % Nyh];
                }
            } else
            for (size_t e = (size_t)j * NT + tid; e < per_batch; e += (size_t)RED_BLOCKS * NT) {
                C2<T> rv = b[base + e];
                if (Ax) rv = rv - Ax[base + e];
                T m = Mdiag[e];
                C2<T> zv = nan2zero(mk<T>(rv.x / m, rv.y / m));
                r[base + e] = rv; z[base + e] = zv; p[base + e] = zv;
                s += ((double)rv.x * (double)zv.x + (double)rv.y * (double)zv.y) * (double)lam[e % Nyh];
            }
            sm[tid] = s;
        }
        CMBL_SYNC();
        CMBL_FOR_THREADS(tid, NT) {
            if (tid == 0) { double s = 0; for (int i = 0; i < NT; ++i) s += sm[i]; res_part[blk] = s * scale; }
        }
    }
};

// res_out[b] = Σ partial[b][:]
struct SumPartialsBody {
    static constexpr int NT = 32;
    static const char* name() { return "sum_partials"; }
    int Nb; const double* part; double* out;
    DEV void operator()(int, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) { for (int b = tid; b < Nb; b += NT) out[b] = sum_partials(part + (size_t)b * RED_BLOCKS); }
    }
};

template <class T> void diag_mul(PlanT<T>& P, int basis, const T* diag, int Cd, const void* in, void* out, int C, bool ldiv, cmblStream_t st);
template <class T> void qu_eb(PlanT<T>& P, int dir, const C2<T>* in, C2<T>* out, int Nb, int stride_planes, int first_plane, cmblStream_t st);
template <class T> void blockdiag_ieb(PlanT<T>& P, int mode, const T* block, const C2<T>* in, C2<T>* out, int Nb, cmblStream_t st);
// per-batch dot into device partial sums; returns pointer to partial[Nb][RED_BLOCKS] (plan scratch)
template <class T> void dot_partials(PlanT<T>& P, int basis, const void* a, const void* b, int Npol, int Nb, double* partial, cmblStream_t st);

}  // namespace cmbl
