// LenseFlow on the device (src/lenseflow.jl, src/flowops.jl:11-14, RK4 src/numerical_algorithms.jl:11-24).
//
// The reference evaluates the velocity  v = p₁·∂ₓf + p₂·∂ᵧf  with three full 2-D FFTs per RK stage
// (src/lenseflow.jl:150-161).  ∂ᵧ only involves transforms along y and ∂ₓ only transforms along x, so here a stage is
// two 1-D spectral-derivative kernels and no 2-D transform at all:
//     row kernel    (FlowRowBody): tiles of rows, all x      → ∂ₓ (Hermitian part), Nyquist-in-x line N(y)
//     column kernel (FlowColBody): tiles of columns, all y   → ∂ᵧ, the Nyquist correction, velocity, RK4 update
// What makes this exact rather than approximate is the reference's treatment of the x-Nyquist mode: ℓ at Nyquist is
// −(N/2)Δℓ and is not zeroed (src/proj_lambert.jl:63-64), so iℓₓF is non-Hermitian there and the final C2R along y
// (which drops Im of the ky=0, Ny/2 rows) turns that mode into  (ℓ_N/Nx)(−1)^x · J[N](y),  N(y)=Σₓ(−1)^x f(y,x),
// J = Hilbert-type multiplier i·sign(ky) with DC/Nyquist removed.  The column kernel adds exactly this rank-one term.
// The adjoint flow (src/lenseflow.jl:163-174), whose state the reference keeps in Fourier space, is integrated in map
// space in divergence form; the non-Hermitian spectrum entries the reference accumulates on the ky∈{0,Ny/2} rows are
// carried as two 1-D accumulators and restored by AdjFixBody after the final rfft2.
#pragma once
#include "fft2d.cuh"

// launch shape of the two generic stage kernels: 256 threads per tile, two blocks per SM (128 registers) — at Nside=2048 18 % (fp64) /
// 11 % (fp32) faster than 128 threads x 3 blocks (profiles/r01_generic_flow_launch_shape.log)
#ifndef CMBL_FLOW_NT
#define CMBL_FLOW_NT 256
#endif
#ifndef CMBL_FLOW_MINB
#define CMBL_FLOW_MINB 2
#endif

namespace cmbl {

// p[k] pointer for plane c:  pcache layout [k][Nbphi][2][Nx][Ny]
template <class T> HD const T* p_plane(const T* pk, int c, int Npol, int Nbphi, int comp, size_t nmap) {
    int bphi = (Nbphi == 1) ? 0 : c / Npol;
    return pk + ((size_t)bphi * 2 + comp) * nmap;
}

// ---------------------------------------------------------------------------------------------------------------
// row kernel: tmp = ∂ₓ_herm(g),  g = u (forward) or p₁·u (adjoint);  nline = N(y) = Σₓ (−1)^x g.
// The block that finishes a plane last (ticket counter) also turns that plane's N(y) into jn(y) = cN·J[N](y).
// ---------------------------------------------------------------------------------------------------------------
template <class T, bool ADJ> struct RowMid {
    const T* mult; T* nline_c; T* nacc_c; T wgt; int y0;
    template <int R> HD void run(int l, int i0, C2<T>* v) const {
        if (i0 == 0) {
            C2<T> nq = v[R / 2];                                  // Nyquist coefficient sits at tile position R/2
            nline_c[y0 + 2 * l] = nq.x; nline_c[y0 + 2 * l + 1] = nq.y;
            if (ADJ) { nacc_c[y0 + 2 * l] += wgt * nq.x; nacc_c[y0 + 2 * l + 1] += wgt * nq.y; }
        }
#pragma unroll
        for (int q = 0; q < R; ++q) { T m = CMBL_LDG(&mult[i0 + q]); v[q] = mk<T>(-m * v[q].y, m * v[q].x); }
    }
};

template <class T> struct SignMid {                                // J: multiplier i·sign(k)/N (0 at DC / Nyquist)
    const T* mult;
    template <int R> HD void run(int, int i0, C2<T>* v) const {
#pragma unroll
        for (int q = 0; q < R; ++q) { T m = CMBL_LDG(&mult[i0 + q]); v[q] = mk<T>(-m * v[q].y, m * v[q].x); }
    }
};

template <class T, bool ADJ> struct FlowRowBody {
    static constexpr int NT = CMBL_FLOW_NT, MINB = CMBL_FLOW_MINB;
    static const char* name() { return "flow_rows"; }
    Fft1D<T> fx, fy; const T* mult; const T* mult_sign_y; T cN;
    int Ny, Nx, L, logL, tiles_per_plane, Npol, Nbphi, cbase;
    const T* u; const T* pk; T* tmp; T* nline; T* jn; T* nacc; T wgt; int* counter;
    static HD size_t smem_bytes(const Fft1D<T>& fx, const Fft1D<T>& fy, int L) {
        size_t a = Tile<T, false>::bytes(fx.N, L, fx.sk), b = Tile<T, false>::bytes(fy.N, 1, fy.sk);
        return (a > b ? a : b) + 16;
    }
    DEV void operator()(int blk, unsigned char* smem) const {
        const int c = cbase + blk / tiles_per_plane, y0 = (blk % tiles_per_plane) * 2 * L;
        const size_t nmap = (size_t)Ny * Nx;
        Tile<T, false> tv = line_tile<T>(smem, L, fx);
        int* flag = reinterpret_cast<int*>(smem + smem_bytes(fx, fy, L) - 16);
        const T* uc = u + (size_t)c * nmap + y0;
        const T* p1 = ADJ ? p_plane(pk, c, Npol, Nbphi, 0, nmap) + y0 : nullptr;
        CMBL_FOR_THREADS(tid, NT) {
            for (int e = tid; e < (Nx << logL); e += NT) {
                const int x = e >> logL, l = e & (L - 1);
                const size_t idx = (size_t)x * Ny + 2 * l;
                C2<T> v = *reinterpret_cast<const C2<T>*>(uc + idx);
                if (ADJ) { C2<T> pp = *reinterpret_cast<const C2<T>*>(p1 + idx); v.x *= pp.x; v.y *= pp.y; }
                tv.at(l, x) = v;
            }
        }
        CMBL_SYNC();
        RowMid<T, ADJ> mid{mult, nline + (size_t)c * Ny, ADJ ? nacc + (size_t)c * Ny : nullptr, wgt, y0};
        fft_spectral_op<T, false, NT>(tv, fx, mid);
        T* tc = tmp + (size_t)c * nmap + y0;
        CMBL_FOR_THREADS(tid, NT) {
            for (int e = tid; e < (Nx << logL); e += NT) {
                const int x = e >> logL, l = e & (L - 1);
                *reinterpret_cast<C2<T>*>(tc + (size_t)x * Ny + 2 * l) = tv.at(l, x);
            }
        }
        // ---- last block of this plane: jn = cN · J[N] ------------------------------------------------------------
        CMBL_FOR_THREADS(tid, NT) { mem_fence(); }                  // publish this block's nline entries device-wide
        CMBL_SYNC();
        CMBL_FOR_THREADS(tid, NT) {
            if (tid == 0) {
                mem_fence();
                *flag = (atomic_inc_int(counter + c) == tiles_per_plane - 1) ? 1 : 0;
            }
        }
        CMBL_SYNC();
        if (*flag) {
            Tile<T, false> t1 = line_tile<T>(smem, 1, fy);
            CMBL_FOR_THREADS(tid, NT) {
                if (tid == 0) mem_fence();
                for (int y = tid; y < Ny; y += NT) t1.at(0, y) = mk<T>(ld_cg(nline + (size_t)c * Ny + y), (T)0);
            }
            CMBL_SYNC();
            SignMid<T> smid{mult_sign_y};
            fft_spectral_op<T, false, NT>(t1, fy, smid);
            CMBL_FOR_THREADS(tid, NT) {
                for (int y = tid; y < Ny; y += NT) jn[(size_t)c * Ny + y] = cN * t1.at(0, y).x;
                if (tid == 0) counter[c] = 0;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// column kernel: ∂ᵧ, Nyquist correction, velocity k, RK4 update (128-bit global accesses)
//   forward: k = p₁·(tmp ± jn) + p₂·∂ᵧu          adjoint: k = tmp ± jn + ∂ᵧ(p₂·u)         (+ for even x, − for odd x)
//   acc_out = (acc_in ? acc_in : ybase) + cb·k ;  u_out = ybase + ca·k  (if u_out)
// ---------------------------------------------------------------------------------------------------------------
template <class T, bool ADJ> struct ColMid {
    const T* mult_d; T* macc_c; T wgt; int x0;
    template <int R> HD void run(int l, int i0, C2<T>* v) const {
        if (ADJ && i0 == 0) {
            C2<T> nq = v[R / 2];
            macc_c[x0 + 2 * l] += wgt * nq.x; macc_c[x0 + 2 * l + 1] += wgt * nq.y;
        }
#pragma unroll
        for (int q = 0; q < R; ++q) { T m = CMBL_LDG(&mult_d[i0 + q]); v[q] = mk<T>(-m * v[q].y, m * v[q].x); }
    }
};

template <class T, bool ADJ> struct FlowColBody {
    static constexpr int NT = CMBL_FLOW_NT, MINB = CMBL_FLOW_MINB;
    static const char* name() { return "flow_cols"; }
    Fft1D<T> fy; const T* mult_d;
    int Ny, Nx, L, logNyv, tiles_per_plane, Npol, Nbphi, cbase;
    const T* u; const T* pk; const T* tmp; const T* jn; T* macc; T wgt;
    const T* ybase; const T* acc_in; T* acc_out; T* u_out; T ca, cb;
    DEV void operator()(int blk, unsigned char* smem) const {
        constexpr int V = Vec<T>::N;
        const int c = cbase + blk / tiles_per_plane, x0 = (blk % tiles_per_plane) * 2 * L;
        const size_t nmap = (size_t)Ny * Nx, off = (size_t)c * nmap + (size_t)x0 * Ny;
        Tile<T, false> tv = line_tile<T>(smem, L, fy);
        const T* uc = u + off;
        const T* p1 = p_plane(pk, c, Npol, Nbphi, 0, nmap) + (size_t)x0 * Ny;
        const T* p2 = p_plane(pk, c, Npol, Nbphi, 1, nmap) + (size_t)x0 * Ny;
        const int nvec = L << logNyv;                                   // vectors per tile column-set
        CMBL_FOR_THREADS(tid, NT) {
            for (int e = tid; e < nvec; e += NT) {
                const int l = e >> logNyv, y = (e & ((1 << logNyv) - 1)) * V;
                const size_t ia = (size_t)(2 * l) * Ny + y, ib = ia + Ny;
                Vec<T> a = vload(uc + ia), b = vload(uc + ib);
                if (ADJ) {
                    Vec<T> pa = vload(p2 + ia), pb = vload(p2 + ib);
#pragma unroll
                    for (int k = 0; k < V; ++k) { a.v[k] *= pa.v[k]; b.v[k] *= pb.v[k]; }
                }
                C2<T>* dst = &tv.at(l, y);                               // V consecutive tile positions (V | 16)
#pragma unroll
                for (int k = 0; k < V; ++k) dst[k] = mk<T>(a.v[k], b.v[k]);
            }
        }
        CMBL_SYNC();
        ColMid<T, ADJ> mid{mult_d, ADJ ? macc + (size_t)c * Nx : nullptr, wgt, x0};
        fft_spectral_op<T, false, NT>(tv, fy, mid);
        const T* tc = tmp + off;
        const T* jc = jn + (size_t)c * Ny;
        CMBL_FOR_THREADS(tid, NT) {
            for (int e = tid; e < nvec; e += NT) {
                const int l = e >> logNyv, y = (e & ((1 << logNyv) - 1)) * V;
                const size_t ia = (size_t)(2 * l) * Ny + y, ib = ia + Ny;
                const C2<T>* src = &tv.at(l, y);
                Vec<T> j = vload(jc + y), ta = vload(tc + ia), tb = vload(tc + ib), ka, kb;
                if (ADJ) {
#pragma unroll
                    for (int k = 0; k < V; ++k) { C2<T> z = src[k]; ka.v[k] = ta.v[k] + j.v[k] + z.x; kb.v[k] = tb.v[k] - j.v[k] + z.y; }
                } else {
                    Vec<T> p1a = vload(p1 + ia), p1b = vload(p1 + ib), p2a = vload(p2 + ia), p2b = vload(p2 + ib);
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        C2<T> z = src[k];
                        ka.v[k] = p1a.v[k] * (ta.v[k] + j.v[k]) + p2a.v[k] * z.x;
                        kb.v[k] = p1b.v[k] * (tb.v[k] - j.v[k]) + p2b.v[k] * z.y;
                    }
                }
                Vec<T> ba, bb;
#pragma unroll
                for (int k = 0; k < V; ++k) { ba.v[k] = 0; bb.v[k] = 0; }
                if (ybase) { ba = vload(ybase + off + ia); bb = vload(ybase + off + ib); }
                Vec<T> a0 = ba, b0 = bb;
                if (acc_in) { a0 = vload(acc_in + off + ia); b0 = vload(acc_in + off + ib); }
#pragma unroll
                for (int k = 0; k < V; ++k) { a0.v[k] += cb * ka.v[k]; b0.v[k] += cb * kb.v[k]; }
                vstore(acc_out + off + ia, a0); vstore(acc_out + off + ib, b0);
                if (u_out) {
#pragma unroll
                    for (int k = 0; k < V; ++k) { ba.v[k] += ca * ka.v[k]; bb.v[k] += ca * kb.v[k]; }
                    vstore(u_out + off + ia, ba); vstore(u_out + off + ib, bb);
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// adjoint flow bookkeeping on the ky ∈ {0, Ny/2} rows of the Fourier state
// ---------------------------------------------------------------------------------------------------------------
// rows0[c][r][kx] = Y[c][kx][r ? Ny/2 : 0]
template <class T> struct AdjRowsSaveBody {
    static constexpr int NT = 256;
    static const char* name() { return "adj_rows_save"; }
    int Nx, Nyh; const C2<T>* Y; C2<T>* rows0;
    DEV void operator()(int blk, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) {
            for (int e = tid; e < 2 * Nx; e += NT) {
                int r = e / Nx, kx = e - r * Nx;
                rows0[((size_t)blk * 2 + r) * Nx + kx] = Y[((size_t)blk * Nx + kx) * Nyh + (r ? Nyh - 1 : 0)];
            }
        }
    }
};

// out = rfft2(y_final) + (I − P)·Y₀ + Nyquist accumulators   (one block per plane; see header comment)
template <class T> struct AdjFixBody {
    static constexpr int NT = 256;
    static const char* name() { return "adj_fix"; }
    Fft1D<T> fx; int Ny, Nx, Nyh; T lxN, lyN;
    const C2<T>* rows0; const T* nacc; const T* macc; C2<T>* out;
    DEV void operator()(int blk, unsigned char* smem) const {
        const int c = blk;
        C2<T>* line = reinterpret_cast<C2<T>*>(smem);
        T* red = reinterpret_cast<T*>(line + Tile<T, false>::pitch_for(Nx, fx.sk));   // [2][NT]
        Tile<T, false> tv = line_tile<T>(smem, 1, fx);
        CMBL_FOR_THREADS(tid, NT) {
            for (int x = tid; x < Nx; x += NT) tv.at(0, x) = mk<T>(macc[(size_t)c * Nx + x], (T)0);
            T s0 = 0, s1 = 0;
            for (int y = tid; y < Ny; y += NT) { T v = nacc[(size_t)c * Ny + y]; s0 += v; s1 += (y & 1) ? -v : v; }
            red[tid] = s0; red[NT + tid] = s1;
        }
        CMBL_SYNC();
        fft_forward_passes<T, false, NT>(tv, fx, 0, fx.npass);
        C2<T>* oc = out + (size_t)c * Nx * Nyh;
        const C2<T>* r0 = rows0 + (size_t)c * 2 * Nx;
        CMBL_FOR_THREADS(tid, NT) {
            for (int e = tid; e < 2 * Nx; e += NT) {
                int r = e / Nx, kx = e - r * Nx;
                C2<T> X = r0[(size_t)r * Nx + kx], Xm = r0[(size_t)r * Nx + ((Nx - kx) & (Nx - 1))];
                C2<T> add = mk<T>((X.x - Xm.x) * (T)0.5, (X.y + Xm.y) * (T)0.5);     // (X − conj X₋ₖ)/2
                if (r == 1) {                                                          // + iℓyN · FFTx(Macc)
                    C2<T> m = tv.at(0, CMBL_LDG(&fx.pos[kx]));
                    add.x += -lyN * m.y; add.y += lyN * m.x;
                }
                if (kx == Nx / 2) {                                                    // + iℓxN · Σ(±)Nacc
                    T s = 0;
                    for (int i = 0; i < NT; ++i) s += red[r * NT + i];
                    add.y += lxN * s;
                }
                C2<T>& o = oc[(size_t)kx * Nyh + (r ? Nyh - 1 : 0)];
                o = o + add;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// precompute! (src/lenseflow.jl:131-142; gradhess src/specialops.jl:184-188; pinv! src/field_vectors.jl:86-94)
// ---------------------------------------------------------------------------------------------------------------
// Φ[bphi] -> 5 spectra [bphi][5]: g1=iℓxΦ, g2=iℓyΦ, H11=iℓx g1, H21=iℓx g2, H22=iℓy g2
template <class T> struct GradHessSpecBody {
    static constexpr int NT = 256;
    static const char* name() { return "gradhess_spec"; }
    int Nx, Nyh; const T* lx; const T* ly; const C2<T>* phi; C2<T>* out; size_t total;   // total = Nbphi*Nx*Nyh
    DEV void operator()(int blk, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) {
            size_t e = (size_t)blk * NT + tid;
            if (e < total) {
                size_t nf = (size_t)Nx * Nyh;
                size_t b = e / nf, r = e - b * nf;
                int kx = (int)(r / Nyh), ky = (int)(r - (size_t)kx * Nyh);
                C2<T> d1 = mk<T>((T)0, lx[kx]), d2 = mk<T>((T)0, ly[ky]);
                C2<T> F = phi[e];
                C2<T> g1 = cmul(d1, F), g2 = cmul(d2, F);
                C2<T>* o = out + b * 5 * nf + r;
                o[0] = g1; o[nf] = g2; o[2 * nf] = cmul(d1, g1); o[3 * nf] = cmul(d1, g2); o[4 * nf] = cmul(d2, g2);
            }
        }
    }
};

// maps gh[bphi][5] -> pcache[k][bphi][2], minv[k][bphi][3] (m11, m21, m22) for k = 0..2n
template <class T> struct PCacheBody {
    static constexpr int NT = 256;
    static const char* name() { return "pcache"; }
    int nk, Nbphi; size_t nmap; const T* gh; T* pcache; T* minv;
    int G, Nx, Ny;                       // G > 0: write the caches in the row-grouped layout of flow_fast.cuh
    DEV void operator()(int blk, unsigned char*) const {
        CMBL_FOR_THREADS(tid, NT) {
            size_t e = (size_t)blk * NT + tid;
            if (e < nmap * Nbphi) {
                size_t b = e / nmap, r = e - b * nmap;
                const size_t rin = r;
                if (G > 0) { const int x = (int)(r / Ny), y = (int)(r - (size_t)x * Ny); r = ((size_t)(y / G) * Nx + x) * G + (y % G); }
                const T* g = gh + b * 5 * nmap + rin;
                T g1 = g[0], g2 = g[nmap], H11 = g[2 * nmap], H21 = g[3 * nmap], H22 = g[4 * nmap];
                for (int k = 0; k < nk; ++k) {
                    T t = (T)((double)k / (double)(nk - 1));
                    T a = (T)1 + t * H11, d = (T)1 + t * H22, bb = t * H21;
                    T det = a * d - bb * bb;
                    T idet = (det == (T)0) ? (T)0 : (T)1 / det;              // scalar pinv: 0 -> 0
                    T m11 = idet * d, m21 = -idet * bb, m22 = idet * a;      // m12 == m21 (field_vectors.jl:87 reads [2,1] twice)
                    size_t o = ((size_t)k * Nbphi + b) * 2 * nmap + r;
                    pcache[o] = m11 * g1 + m21 * g2;
                    pcache[o + nmap] = m21 * g1 + m22 * g2;
                    if (minv) {
                        size_t om = ((size_t)k * Nbphi + b) * 3 * nmap + r;
                        minv[om] = m11; minv[om + nmap] = m21; minv[om + 2 * nmap] = m22;
                    }
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// get_max_lensing_step (src/lenseflow.jl:242-256): per pixel the roots α of det(𝕀 + ∇∇(ϕ + α η)) = a α² + b α + c; per batch item the
// smallest positive root.  gh_phi / gh_eta: the 5 maps per batch item of GradHessSpecBody (only the Hessian entries are used; the
// reference reads ϕ₁₂ for both off-diagonal entries).  partial[b][RED_BLOCKS]: block minima (+inf when a block saw no positive root).
// ---------------------------------------------------------------------------------------------------------------
template <class T> struct MaxStepBody {
    static constexpr int NT = 256;
    static const char* name() { return "max_lensing_step"; }
    size_t nmap; const T* gh_phi; const T* gh_eta; double* partial;
    DEV void operator()(int blk, unsigned char* smem) const {
        double* sm = reinterpret_cast<double*>(smem);
        const int bi = blk / RED_BLOCKS, j = blk % RED_BLOCKS;
        const T* P = gh_phi + (size_t)bi * 5 * nmap; const T* E = gh_eta + (size_t)bi * 5 * nmap;
        CMBL_FOR_THREADS(tid, NT) {
            double m = HUGE_VAL;
            for (size_t e = (size_t)j * NT + tid; e < nmap; e += (size_t)RED_BLOCKS * NT) {
                const T p11 = P[2 * nmap + e], p12 = P[3 * nmap + e], p22 = P[4 * nmap + e];
                const T e11 = E[2 * nmap + e], e12 = E[3 * nmap + e], e22 = E[4 * nmap + e];
                const T a = e11 * e22 - e12 * e12;
                const T b = e11 * ((T)1 + p22) + e22 * ((T)1 + p11) - (T)2 * e12 * p12;
                const T c = ((T)1 + p11) * ((T)1 + p22) - p12 * p12;
                const T sq = sqrt(b * b - (T)4 * a * c);                  // NaN where there is no real root: fails every comparison below
                const T a1 = (-b + sq) / ((T)2 * a), a2 = (-b - sq) / ((T)2 * a);
                if (a1 > (T)0 && (double)a1 < m) m = (double)a1;
                if (a2 > (T)0 && (double)a2 < m) m = (double)a2;
            }
            sm[tid] = m;
        }
        CMBL_SYNC();
        CMBL_FOR_THREADS(tid, NT) {
            if (tid == 0) { double m = HUGE_VAL; for (int i = 0; i < NT; ++i) m = sm[i] < m ? sm[i] : m; partial[blk] = m; }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
struct FlowBase { PlanBase* plan = nullptr; virtual ~FlowBase() {} };

template <class T> struct FlowT : FlowBase {
    PlanT<T>* P = nullptr;
    int nsteps = 7, Npol = 1, Nb = 1, Nbphi = 1, C = 1;
    bool have_p = false, have_minv = false;
    bool integrated_once = false;       // every lazily allocated work buffer of the stage kernels exists (set by flow_integrate_range)
    int pcache_G = 0;                    // layout of pcache / minv: 0 = reference layout, else rows per row group
    DevBuf yrg;                          // row-grouped copy of the ODE state
    DevBuf g_yf, g_yd, g_yp, g_uf, g_ud, g_up, g_af, g_ad, g_ap, g_kf, g_kd, g_kp, g_ldf, g_gxy, g_a12, g_six, g_spec, g_spec2;   // δ-flow scratch (flow_grad.cu)
    DevBuf gq_f[3], gq_d[3], gq_ref, gq_gx, gq_gy, gq_A, gq_Aref, gq_spec;      // fused δ-flow (flow_grad.cu): (y, acc, u) of the f and δf legs, derivative maps, δϕ integrand accumulators
    DevBuf jnblk;                        // per-block private J[N] line of the fast column kernel (fallback when the shared line is not published)
    DevBuf jnflag; int jn_epoch = 0;     // per-plane publication flags of the shared J[N] lines (value = launch epoch)
    DevBuf pcache, minv, ybuf, acc, ubuf, tmp, nline, jn, counter, nacc, macc, rows0, spec, gh;
    size_t nmap() const { return P->map_elems(); }
    const T* pk(int k) const { return reinterpret_cast<T*>(pcache.p) + (size_t)k * Nbphi * 2 * nmap(); }
};

template <class T> void flow_precompute(FlowT<T>& F, const void* phi, int phi_basis, bool with_minv, cmblStream_t st);
// building blocks shared with the transpose-δ flow (flow_grad.cu)
template <class T> int flow_rg_rows(const PlanT<T>& P);                       // rows per group of the internal row-grouped layout (0: generic kernels)
template <class T> void flow_reserve(FlowT<T>& F, cmblStream_t st);
template <class T, bool TO_RG> void convert_layout(PlanT<T>& P, int G, const T* in, T* out, int C, cmblStream_t st);
// one RK stage (row kernel + column kernel) on planes [c0, c0+nC); dx_out/dy_out (fast forward kernels only): also export ∂ₓu, ∂ᵧu
template <class T, bool ADJ> void flow_stage(FlowT<T>& F, int c0, int nC, const T* u, int kq, T wgt, const T* ybase, const T* acc_in, T* acc_out, T* u_out,
                                             T ca, T cb, cmblStream_t st, T* dx_out = nullptr, T* dy_out = nullptr);
template <class T> void flow_adj_prepare(FlowT<T>& F, const C2<T>* Y0, T* y, cmblStream_t st, int G = 0);     // G > 0: y in the row-grouped layout (fft2d.cuh)
template <class T> void flow_adj_finish(FlowT<T>& F, const T* y, C2<T>* Yout, cmblStream_t st, int G = 0);
// rows per group when the transforms around a flow can hand over the integrator's row-grouped buffer directly (0 = convert layouts)
template <class T> int flow_rg_direct(FlowT<T>& F);
// integrate the map-space flow in place on y from stage index k0 to k1 (0 or 2n)
// rg_state: y already is the row-grouped state (all F.C planes) — no layout conversion on entry / exit
template <class T> void flow_integrate(FlowT<T>& F, bool adj, T* y, int k0, int k1, cmblStream_t st, bool rg_state = false);
template <class T> void flow_integrate_range(FlowT<T>& F, bool adj, T* y, int k0, int k1, int c0, int nC, cmblStream_t st, bool rg_state = false);
template <class T> void flow_apply(FlowT<T>& F, int op, const void* in, void* out, cmblStream_t st);
template <class T> int flow_kernel_path(FlowT<T>& F);
// smallest positive α per batch item with det(𝕀 + ∇∇(ϕ + α η)) = 0 somewhere (host doubles; +inf if none); synchronises
template <class T> void max_lensing_step(PlanT<T>& P, const void* phi, int phi_basis, const void* eta, int eta_basis, int Nb, double* out_host, cmblStream_t st);

}  // namespace cmbl
