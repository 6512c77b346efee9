// In-shared-memory, in-place, power-of-two complex FFT passes used by every transform kernel of the library.
//
// Forward transforms run decimation-in-frequency (natural order in, mixed-radix digit-reversed order out);
// inverse transforms run the exact mirror (decimation-in-time: digit-reversed in, natural out).  A spectral
// derivative (forward → multiply → inverse) therefore needs no reordering at all, and the last forward pass, the
// pointwise multiplier and the first inverse pass fuse into one register-resident "middle" step.
// Replaces the FFTW / CUFFT plans the reference reaches through src/util_fft.jl:19-44.
#pragma once
#include "common.cuh"

namespace cmbl {

constexpr int MAX_PASSES = 6;

// One 1-D transform length. `W[t] = exp(-2πi t/N)`; `pos[k]` = tile position that holds frequency k after the
// forward passes. Device pointers (host pointers in the emulator build).
template <class T> struct Fft1D {
    int N = 0, logN = 0, npass = 0, sk = 3;      // sk = log2(last radix): skew shift of line-major tiles
    int radix[MAX_PASSES] = {0, 0, 0, 0, 0, 0};
    const C2<T>* W = nullptr;
    const int* pos = nullptr;
};

// Shared-memory tile of L complex lines of length N.
//   LFAST = true : element (l, i) at (i + (i >> 4))*L + l   (lines interleaved; used when lines are strided in global memory).  One
//                  padding slot per 16 elements: in the twiddle-free radix-16 pass the lanes of a warp hold neighbouring butterflies
//                  whose elements are 16·L slots apart — a multiple of the 128-byte bank width for the usual L·sizeof(C2) = 64 B —
//                  and the skew spreads them over both halves (27-45 % of the row pass's shared-memory wavefronts were conflicts).
//   LFAST = false: element (l, i) at l*pitch + i + (i >> sk)   (line-major; sk = log2 of the LAST radix of the schedule,
//                  so that the stride-R accesses of the fused middle pass are bank-conflict free)
template <class T, bool LFAST> struct Tile {
    C2<T>* s; int L; int pitch; int sk;
    HD int phys(int i) const { return i + (i >> sk); }
    HD C2<T>& at(int l, int i) const { return LFAST ? s[(i + (i >> 4)) * L + l] : s[l * pitch + i + (i >> sk)]; }
    // pitch ≡ 2 (mod 16) elements: transposing tile loads (lines fastest across threads) stay conflict free
    static HD int pitch_for(int N, int sk) { int p = N + (N >> sk); return p + ((18 - (p & 15)) & 15); }
    static HD size_t bytes(int N, int L, int sk) { return sizeof(C2<T>) * (size_t)(LFAST ? (N + (N >> 4)) * L : pitch_for(N, sk) * L); }
};
template <class T> HD Tile<T, false> line_tile(unsigned char* smem, int L, const Fft1D<T>& f) {
    Tile<T, false> t; t.s = reinterpret_cast<C2<T>*>(smem); t.L = L; t.pitch = Tile<T, false>::pitch_for(f.N, f.sk); t.sk = f.sk; return t;
}

// ---------------------------------------------------------------------------------------------------------------
// register butterflies: v[q] <- sum_m v[m] * exp(∓2πi m q / R)   (INV: + sign, unnormalised)
// ---------------------------------------------------------------------------------------------------------------
template <class T, bool INV> HD C2<T> mul_mi(C2<T> d) { return INV ? mk<T>(-d.y, d.x) : mk<T>(d.y, -d.x); }   // (∓i)·d

template <class T, bool INV> HD void dft2(C2<T>& a, C2<T>& b) { C2<T> t = a; a = t + b; b = t - b; }

template <class T, bool INV> HD void dft4(C2<T>& x0, C2<T>& x1, C2<T>& x2, C2<T>& x3) {
    C2<T> a = x0 + x2, b = x0 - x2, c = x1 + x3, d = mul_mi<T, INV>(x1 - x3);
    x0 = a + c; x2 = a - c; x1 = b + d; x3 = b - d;
}

template <class T, bool INV> HD void dft8(C2<T>* v) {
    const T h = (T)0.70710678118654752440084436210485;
    dft4<T, INV>(v[0], v[2], v[4], v[6]);                 // E[k] in v[0],v[2],v[4],v[6]
    dft4<T, INV>(v[1], v[3], v[5], v[7]);                 // O[k] in v[1],v[3],v[5],v[7]
    // w8^k O[k]; forward w8 = (1 - i)/sqrt2
    C2<T> o0 = v[1];
    C2<T> o1 = INV ? mk<T>((v[3].x - v[3].y) * h, (v[3].x + v[3].y) * h) : mk<T>((v[3].x + v[3].y) * h, (v[3].y - v[3].x) * h);
    C2<T> o2 = mul_mi<T, INV>(v[5]);
    C2<T> o3 = INV ? mk<T>((-v[7].x - v[7].y) * h, (v[7].x - v[7].y) * h) : mk<T>((v[7].y - v[7].x) * h, (-v[7].x - v[7].y) * h);
    C2<T> e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    v[0] = e0 + o0; v[4] = e0 - o0;
    v[1] = e1 + o1; v[5] = e1 - o1;
    v[2] = e2 + o2; v[6] = e2 - o2;
    v[3] = e3 + o3; v[7] = e3 - o3;
}

template <class T, bool INV> HD void dft16(C2<T>* v) {
    // x[4a+b]: inner DFT4 over a for each b, twiddle w16^(b*k1), outer DFT4 over b -> X[k1 + 4 k2]
    const T c1 = (T)0.92387953251128675612818318939679, s1 = (T)0.38268343236508977172845998403040;
    const T h = (T)0.70710678118654752440084436210485;
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4<T, INV>(v[b], v[4 + b], v[8 + b], v[12 + b]);     // Y_b[k1] at v[4*k1 + b]
    // w16^m, forward = (cos, -sin)(2πm/16)
    const T wc[10] = {1, c1, h, s1, 0, -s1, -h, -c1, -1, -c1};
    const T ws[10] = {0, s1, h, c1, 1, c1, h, s1, 0, -s1};
#pragma unroll
    for (int k1 = 1; k1 < 4; ++k1)
#pragma unroll
        for (int b = 1; b < 4; ++b) {
            int m = b * k1;                                                             // 1..9
            C2<T> w = mk<T>(wc[m], INV ? ws[m] : -ws[m]);
            v[4 * k1 + b] = cmul(v[4 * k1 + b], w);
        }
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4<T, INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);   // X[k1+4k2] at v[4*k1+k2]
    // transpose to natural order: X[k] with k = k1 + 4 k2 currently at 4*k1 + k2
    C2<T> t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = v[i];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = t[4 * k1 + k2];
    }
}

template <class T, int R, bool INV> HD void dftR(C2<T>* v) {
    if (R == 2) dft2<T, INV>(v[0], v[1]);
    else if (R == 4) dft4<T, INV>(v[0], v[1], v[2], v[3]);
    else if (R == 8) dft8<T, INV>(v);
    else dft16<T, INV>(v);
}

// ---------------------------------------------------------------------------------------------------------------
// one block-wide pass over sub-transforms of length n (stride s = n/R) on every line of the tile.
//   forward: butterfly then twiddle;  inverse: conjugate twiddle then inverse butterfly.
// Called by ONE thread (tid) of NT; the caller provides the block-wide loop and the syncs.
// ---------------------------------------------------------------------------------------------------------------
template <class T, bool LFAST, int R, bool INV>
HD void fft_pass_thread(const Tile<T, LFAST>& tv, const Fft1D<T>& f, int logn, int tid, int NT) {
    const int logR = (R == 2 ? 1 : R == 4 ? 2 : R == 8 ? 3 : 4);
    const int logs = logn - logR, s = 1 << logs;
    const int nb = f.N >> logR;                 // butterflies per line
    const int tws = f.logN - logn;              // twiddle index shift: W_n^(j q) = W_N^(j q << tws)
    const int ntask = nb * tv.L;
    for (int task = tid; task < ntask; task += NT) {
        int l, j;
        if (LFAST) { l = task % tv.L; j = task / tv.L; } else { j = task % nb; l = task / nb; }
        const int jj = j & (s - 1);
        const int i0 = ((j >> logs) << logn) + jj;
        C2<T> v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = tv.at(l, i0 + (m << logs));
        if (!INV) {
            dftR<T, R, false>(v);
            if (logs > 0) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmul(v[q], CMBL_LDG(&f.W[(jj * q) << tws]));
            }
        } else {
            if (logs > 0) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], CMBL_LDG(&f.W[(jj * q) << tws]));
            }
            dftR<T, R, true>(v);
        }
#pragma unroll
        for (int m = 0; m < R; ++m) tv.at(l, i0 + (m << logs)) = v[m];
    }
}

// Line-major tiles: every thread owns a fixed butterfly index j for the whole pass, keeps its R-1 twiddles and R tile
// offsets in registers and sweeps the lines of the tile with them.
template <class T, int R, bool INV, int NT>
HD void fft_pass_lines_thread(const Tile<T, false>& tv, const Fft1D<T>& f, int logn, int tid) {
    constexpr int logR = (R == 2 ? 1 : R == 4 ? 2 : R == 8 ? 3 : 4);
    constexpr bool CACHE_TW = (R <= 8);
    const int logs = logn - logR, s = 1 << logs;
    const int lognb = f.logN - logR, nb = 1 << lognb;
    const int tws = f.logN - logn;
    int jstep, l0, lstep;
    if (nb >= NT) { jstep = NT; l0 = 0; lstep = 1; } else { jstep = nb; l0 = tid >> lognb; lstep = NT >> lognb; }
    for (int j = (nb >= NT) ? tid : (tid & (nb - 1)); j < nb; j += jstep) {
        const int jj = j & (s - 1);
        const int i0 = ((j >> logs) << logn) + jj;
        int off[R];
#pragma unroll
        for (int m = 0; m < R; ++m) off[m] = tv.phys(i0 + (m << logs));
        C2<T> w[R];
        if (CACHE_TW && logs > 0) {
#pragma unroll
            for (int q = 1; q < R; ++q) w[q] = CMBL_LDG(&f.W[(jj * q) << tws]);
        }
        for (int l = l0; l < tv.L; l += lstep) {
            C2<T>* line = tv.s + l * tv.pitch;
            C2<T> v[R];
#pragma unroll
            for (int m = 0; m < R; ++m) v[m] = line[off[m]];
            if (!INV) {
                dftR<T, R, false>(v);
                if (logs > 0) {
#pragma unroll
                    for (int q = 1; q < R; ++q) v[q] = cmul(v[q], CACHE_TW ? w[q] : CMBL_LDG(&f.W[(jj * q) << tws]));
                }
            } else {
                if (logs > 0) {
#pragma unroll
                    for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], CACHE_TW ? w[q] : CMBL_LDG(&f.W[(jj * q) << tws]));
                }
                dftR<T, R, true>(v);
            }
#pragma unroll
            for (int m = 0; m < R; ++m) line[off[m]] = v[m];
        }
        if (nb < NT) break;
    }
}

template <class T, bool INV, int NT>
HD void fft_pass_lines_dispatch(const Tile<T, false>& tv, const Fft1D<T>& f, int R, int logn, int tid) {
    switch (R) {
        case 2: fft_pass_lines_thread<T, 2, INV, NT>(tv, f, logn, tid); break;
        case 4: fft_pass_lines_thread<T, 4, INV, NT>(tv, f, logn, tid); break;
        case 8: fft_pass_lines_thread<T, 8, INV, NT>(tv, f, logn, tid); break;
        default: fft_pass_lines_thread<T, 16, INV, NT>(tv, f, logn, tid); break;
    }
}

template <class T, bool LFAST, bool INV>
HD void fft_pass_dispatch(const Tile<T, LFAST>& tv, const Fft1D<T>& f, int R, int logn, int tid, int NT) {
    switch (R) {
        case 2: fft_pass_thread<T, LFAST, 2, INV>(tv, f, logn, tid, NT); break;
        case 4: fft_pass_thread<T, LFAST, 4, INV>(tv, f, logn, tid, NT); break;
        case 8: fft_pass_thread<T, LFAST, 8, INV>(tv, f, logn, tid, NT); break;
        default: fft_pass_thread<T, LFAST, 16, INV>(tv, f, logn, tid, NT); break;
    }
}

template <class T, bool LFAST, bool INV, int NT>
HD void fft_pass_any(const Tile<T, LFAST>& tv, const Fft1D<T>& f, int R, int logn, int tid) {
    if constexpr (LFAST) fft_pass_dispatch<T, true, INV>(tv, f, R, logn, tid, NT);
    else fft_pass_lines_dispatch<T, INV, NT>(tv, f, R, logn, tid);
}

// Forward passes [p0, p1) of the plan (block-wide; includes the trailing sync of each pass).
template <class T, bool LFAST, int NT>
DEV void fft_forward_passes(const Tile<T, LFAST>& tv, const Fft1D<T>& f, int p0, int p1) {
    int logn = f.logN;
    for (int p = 0; p < p1; ++p) {
        if (p >= p0) {
            CMBL_FOR_THREADS(tid, NT) fft_pass_any<T, LFAST, false, NT>(tv, f, f.radix[p], logn, tid);
            CMBL_SYNC();
        }
        logn -= ilog2(f.radix[p]);
    }
}

// Inverse passes: mirrors forward passes p1-1 down to p0.
template <class T, bool LFAST, int NT>
DEV void fft_inverse_passes(const Tile<T, LFAST>& tv, const Fft1D<T>& f, int p0, int p1) {
    int lognp[MAX_PASSES + 1];
    lognp[0] = f.logN;
    for (int p = 0; p < f.npass; ++p) lognp[p + 1] = lognp[p] - ilog2(f.radix[p]);
    for (int p = p1 - 1; p >= p0; --p) {
        CMBL_FOR_THREADS(tid, NT) fft_pass_any<T, LFAST, true, NT>(tv, f, f.radix[p], lognp[p], tid);
        CMBL_SYNC();
    }
}

// Fused middle of a spectral operator: last forward pass (sub-length R, no twiddles) → mid.run<R>(l, i0, v) on the R
// spectrum values at tile positions i0..i0+R-1 → first inverse pass.  One thread's share.
template <class T, bool LFAST, int R, int NT, class Mid>
HD void fft_middle_thread(const Tile<T, LFAST>& tv, const Fft1D<T>& f, int tid, const Mid& mid) {
    constexpr int logR = (R == 2 ? 1 : R == 4 ? 2 : R == 8 ? 3 : 4);
    const int lognb = f.logN - logR, nb = 1 << lognb;
    if constexpr (LFAST) {
        const int ntask = nb * tv.L;
        for (int task = tid; task < ntask; task += NT) {
            const int l = task % tv.L, j = task / tv.L;
            const int i0 = j << logR;
            C2<T> v[R];
#pragma unroll
            for (int m = 0; m < R; ++m) v[m] = tv.at(l, i0 + m);
            dftR<T, R, false>(v);
            mid.template run<R>(l, i0, v);
            dftR<T, R, true>(v);
#pragma unroll
            for (int m = 0; m < R; ++m) tv.at(l, i0 + m) = v[m];
        }
    } else {
        int jstep, l0, lstep;
        if (nb >= NT) { jstep = NT; l0 = 0; lstep = 1; } else { jstep = nb; l0 = tid >> lognb; lstep = NT >> lognb; }
        for (int j = (nb >= NT) ? tid : (tid & (nb - 1)); j < nb; j += jstep) {
            const int i0 = j << logR;
            const int o0 = tv.phys(i0);                       // the R positions of one block are contiguous
            for (int l = l0; l < tv.L; l += lstep) {
                C2<T>* line = tv.s + l * tv.pitch + o0;
                C2<T> v[R];
#pragma unroll
                for (int m = 0; m < R; ++m) v[m] = line[m];
                dftR<T, R, false>(v);
                mid.template run<R>(l, i0, v);
                dftR<T, R, true>(v);
#pragma unroll
                for (int m = 0; m < R; ++m) line[m] = v[m];
            }
            if (nb < NT) break;
        }
    }
}

// Whole spectral operator on a tile: forward passes, fused middle, inverse passes (block-wide, ends with a sync).
template <class T, bool LFAST, int NT, class Mid>
DEV void fft_spectral_op(const Tile<T, LFAST>& tv, const Fft1D<T>& f, const Mid& mid) {
    fft_forward_passes<T, LFAST, NT>(tv, f, 0, f.npass - 1);
    const int R = f.radix[f.npass - 1];
    CMBL_FOR_THREADS(tid, NT) {
        switch (R) {
            case 2: fft_middle_thread<T, LFAST, 2, NT>(tv, f, tid, mid); break;
            case 4: fft_middle_thread<T, LFAST, 4, NT>(tv, f, tid, mid); break;
            case 8: fft_middle_thread<T, LFAST, 8, NT>(tv, f, tid, mid); break;
            default: fft_middle_thread<T, LFAST, 16, NT>(tv, f, tid, mid); break;
        }
    }
    CMBL_SYNC();
    fft_inverse_passes<T, LFAST, NT>(tv, f, 0, f.npass - 1);
}

}  // namespace cmbl
