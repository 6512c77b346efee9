// Column passes of the general 2-D transforms (rfft2 / irfft2, fft2d.cuh) on the tile machinery of the fast stage kernels
// (flow_fast.cuh) for Ny ∈ {256, 512, 1024, 2048}: persistent blocks, the next tile in flight while the current one is transformed,
// planar XOR-swizzled tiles (plane = column; two real columns are the real and imaginary part of one complex line), the twiddled
// sweeps of FastColBody.  The generic kernels of fft2d.cuh (one tile per block: load → passes → store, nothing overlapped) stay for
// every other length.  Replaces the column half of the FFTW / CUFFT plans behind src/util_fft.jl:26-27,44.
//
//   rfft2 columns : tile lands by cp.async → sweep 1 → sweep 2 → radix 16 in registers → the 16 values of every thread go to the
//                   NATURAL frequency order of the same tile buffer (all threads have read: one barrier) → the two real columns'
//                   spectra are separated (X_a[k] = (Z[k] + conj Z[N−k])/2, X_b[k] = (Z[k] − conj Z[N−k])/2i) and leave as
//                   contiguous half-spectra [x][ky].
//   irfft2 columns: the M half-spectra of a tile are one contiguous range of the input: ONE bulk copy (TMA engine) into a staging
//                   buffer, issued a tile ahead → Z[k] = A[k] + i·B[k], Z[N−k] = conj A[k] + i·conj B[k] into the natural order of
//                   the tile buffer (Im of the ky = 0 and ky = Ny/2 rows ignored, like FFTW / cuFFT c2r) → every thread gathers the
//                   16 values of one radix-16 butterfly (barrier) → inverse radix 16 → inverse sweeps 2 and 1 → the real planes
//                   leave scaled by 1/(Ny·Nx), optionally times a Map-basis diagonal (the pixel mask of the CG operator).
// Natural order k ↔ tile order: after the forward sweeps [R1, R2, 16] frequency k = q1 + R1·q2 + R1·R2·q3 sits at position
// q1·(N/R1) + q2·16 + q3 (plan.cu fft_positions).  With the chunk-level XOR swizzle of the tile both the scattered accesses (a warp:
// 8 values of q2 × 4 of q1 for one q3) and the contiguous ones are bank-conflict free.
#pragma once
#include "flow_fast.cuh"

#ifndef CMBL_FFT_FAST_MINB
#define CMBL_FFT_FAST_MINB 3
#endif

namespace cmbl {

template <class T, int LOGN> struct FastFftColCommon {
    typedef FastColBody<T, LOGN, false> FC;
    static constexpr int N = FC::N, V = FC::V, CH = FC::CH, NT = FC::NT, L = FC::L, M = FC::M, TILE = FC::TILE, TB = FC::TB, LGM = FC::LGM;
    static constexpr int R1 = FC::R1, R2 = FC::R2, NB16 = N / 16, CPB = 16 / V, NH = N / 2;
    static constexpr int TPT = L * NB16 / NT;                    // radix-16 butterflies per thread (1 in fp64, 2 in fp32)
    static_assert((L * NB16) % NT == 0 && TPT >= 1 && TPT <= 2, "one or two radix-16 butterflies per thread");
    static constexpr int MINB = (N > 1024) ? 1 : CMBL_FFT_FAST_MINB;
    typedef typename FC::template Tw<R1> Tw1;
    typedef typename FC::template Tw<R2> Tw2;

    FC fc;                                                       // twiddle tables and layout parameters (reference layout = one row group: G = Ny)
    // element k (natural order) of plane p
    DEV int nat(int k, int p) const { return fc.swzp(k / V, p) * V + (k % V); }
    // the thread's t-th radix-16 butterfly: line pair l, tile-order block j; frequency of its value q3 = kbase + NB16·q3
    static DEV void task_of(int tid, int t, int& l, int& j, int& kbase) {
        const int task = tid + t * NT;
        l = task / NB16; j = task % NB16;
        kbase = j / R2 + R1 * (j % R2);
    }
    DEV void load16_tile(const T* buf, int l, int j, C2<T>* v) const {           // 16 consecutive tile positions
        const T* re = buf + (2 * l) * N; const T* im = re + N;
        const int xa = fc.pxor(2 * l), xb = fc.pxor(2 * l + 1);
#pragma unroll
        for (int cc = 0; cc < CPB; ++cc) {
            const int o = swz8(j * CPB + cc);
            Vec<T> a = vload(re + (o ^ xa) * V), b = vload(im + (o ^ xb) * V);
#pragma unroll
            for (int e = 0; e < V; ++e) v[cc * V + e] = mk<T>(a.v[e], b.v[e]);
        }
    }
    DEV void store16_tile(T* buf, int l, int j, const C2<T>* v) const {
        T* re = buf + (2 * l) * N; T* im = re + N;
        const int xa = fc.pxor(2 * l), xb = fc.pxor(2 * l + 1);
#pragma unroll
        for (int cc = 0; cc < CPB; ++cc) {
            const int o = swz8(j * CPB + cc);
            Vec<T> a, b;
#pragma unroll
            for (int e = 0; e < V; ++e) { a.v[e] = v[cc * V + e].x; b.v[e] = v[cc * V + e].y; }
            vstore(re + (o ^ xa) * V, a); vstore(im + (o ^ xb) * V, b);
        }
    }
    DEV void load16_nat(const T* buf, int l, int kbase, C2<T>* v) const {         // the 16 frequencies kbase + NB16·q3
        const T* re = buf + (2 * l) * N; const T* im = re + N;
#pragma unroll
        for (int q = 0; q < 16; ++q) { const int k = kbase + NB16 * q; v[q] = mk<T>(re[nat(k, 2 * l)], im[nat(k, 2 * l + 1)]); }
    }
    DEV void store16_nat(T* buf, int l, int kbase, const C2<T>* v) const {
        T* re = buf + (2 * l) * N; T* im = re + N;
#pragma unroll
        for (int q = 0; q < 16; ++q) { const int k = kbase + NB16 * q; re[nat(k, 2 * l)] = v[q].x; im[nat(k, 2 * l + 1)] = v[q].y; }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// rfft2, column pass
// ---------------------------------------------------------------------------------------------------------------
template <class T, int LOGN> struct FastR2CColBody : FastFftColCommon<T, LOGN> {
    typedef FastFftColCommon<T, LOGN> Cm;
    using Cm::N; using Cm::V; using Cm::NT; using Cm::L; using Cm::M; using Cm::TILE; using Cm::TB; using Cm::TPT; using Cm::NH; using Cm::fc;
    static constexpr int MINB = Cm::MINB;
#ifdef CMBL_EMU
    static constexpr size_t SMEM = (size_t)TB * 3;               // + snapshot for the serial emulation of "read all — barrier — write all"
#else
    static constexpr size_t SMEM = (size_t)TB * 2;
#endif
    static const char* name() { return "rfft2_cols"; }
    int Nyh, tiles_per_plane, ntiles, nblocks;
    const T* in; C2<T>* out;

    DEV void operator()(int blk, unsigned char* smem) const {
        T* const sbase = reinterpret_cast<T*>(smem);
        const size_t nmap = (size_t)N * fc.Nx;
        typename Cm::Tw1 w1; typename Cm::Tw2 w2;
        auto src_of = [&](int t) { return in + (size_t)(t / tiles_per_plane) * nmap; };
        auto x0_of = [&](int t) { return (t % tiles_per_plane) * M; };
        int tile = blk, cur = 0;
        if (tile < ntiles) { CMBL_FOR_THREADS(tid, NT) { fc.issue_tile(src_of(tile), x0_of(tile), sbase, tid); cp_async_commit(); } }
        CMBL_FOR_THREADS(tid, NT) { CMBL_PRE_END(fc.load_tw1(w1, tid)); }
        for (; tile < ntiles; tile += nblocks, cur ^= 1) {
            T* const buf = sbase + cur * TILE;
            T* const nbuf = sbase + (cur ^ 1) * TILE;
            const int c = tile / tiles_per_plane, x0 = x0_of(tile), next = tile + nblocks;
            CMBL_FOR_THREADS(tid, NT) { cp_async_wait_all(); }
            CMBL_SYNC();
            if (next < ntiles) { CMBL_FOR_THREADS(tid, NT) { fc.issue_tile(src_of(next), x0_of(next), nbuf, tid); cp_async_commit(); } }
            CMBL_FOR_THREADS(tid, NT) { CMBL_PRE_START(fc.load_tw1(w1, tid)); fc.template pass1<false>(buf, nullptr, tid, w1); CMBL_PRE_END(fc.load_tw2(w2, tid)); }
            CMBL_SYNC();
            CMBL_FOR_THREADS(tid, NT) { CMBL_PRE_START(fc.load_tw2(w2, tid)); fc.template pass2<false>(buf, tid, w2); }
            CMBL_SYNC();
            // last forward sweep (radix 16, no twiddles) in registers; its results go to the natural frequency order of the same buffer
#ifdef CMBL_EMU
            T* const snap = sbase + 2 * TILE;
            memcpy(snap, buf, TB);
            CMBL_FOR_THREADS(tid, NT) {
                for (int t = 0; t < TPT; ++t) {
                    int l, j, kb; Cm::task_of(tid, t, l, j, kb);
                    C2<T> v[16]; this->load16_tile(snap, l, j, v); dft16<T, false>(v); this->store16_nat(buf, l, kb, v);
                }
            }
#else
            {
                C2<T> v[TPT][16];
#pragma unroll
                for (int t = 0; t < TPT; ++t) { int l, j, kb; Cm::task_of(threadIdx.x, t, l, j, kb); this->load16_tile(buf, l, j, v[t]); dft16<T, false>(v[t]); }
                __syncthreads();
#pragma unroll
                for (int t = 0; t < TPT; ++t) { int l, j, kb; Cm::task_of(threadIdx.x, t, l, j, kb); this->store16_nat(buf, l, kb, v[t]); }
            }
#endif
            CMBL_SYNC();
            // separate the two real columns of every line and store the half-spectra (ky = 0 .. Ny/2)
            C2<T>* const dst = out + ((size_t)c * fc.Nx + x0) * Nyh;
            CMBL_FOR_THREADS(tid, NT) {
                const T h = (T)0.5;
                for (int idx = tid; idx < L * NH + L; idx += NT) {
                    const int l = idx < L * NH ? idx / NH : idx - L * NH, k = idx < L * NH ? idx % NH : NH;
                    const int km = (N - k) & (N - 1);
                    const T* re = buf + (2 * l) * N; const T* im = re + N;
                    const C2<T> zk = mk<T>(re[this->nat(k, 2 * l)], im[this->nat(k, 2 * l + 1)]);
                    const C2<T> zm = mk<T>(re[this->nat(km, 2 * l)], im[this->nat(km, 2 * l + 1)]);
                    dst[(size_t)(2 * l) * Nyh + k] = mk<T>((zk.x + zm.x) * h, (zk.y - zm.y) * h);
                    dst[(size_t)(2 * l + 1) * Nyh + k] = mk<T>((zk.y + zm.y) * h, (zm.x - zk.x) * h);
                }
                CMBL_PRE_END(fc.load_tw1(w1, tid));
            }
            // the next iteration's first barrier orders these shared-memory reads before the buffer is refilled (two iterations later)
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// irfft2, column pass
// ---------------------------------------------------------------------------------------------------------------
template <class T, int LOGN> struct FastC2RColBody : FastFftColCommon<T, LOGN> {
    typedef FastFftColCommon<T, LOGN> Cm;
    using Cm::N; using Cm::V; using Cm::CH; using Cm::NT; using Cm::L; using Cm::M; using Cm::TILE; using Cm::TB; using Cm::TPT; using Cm::NH; using Cm::fc;
    static constexpr int MINB = Cm::MINB;
    static constexpr int NYH = NH + 1;
    static constexpr unsigned SB = (unsigned)(M * NYH * sizeof(C2<T>));          // the tile's half-spectra: one contiguous range of the input
    static_assert(SB % 16 == 0, "bulk copies move multiples of 16 bytes");
#ifdef CMBL_EMU
    static constexpr size_t SMEM = (size_t)TB * 2 + SB + 16;
#else
    static constexpr size_t SMEM = (size_t)TB + SB + 16;
#endif
    static const char* name() { return "irfft2_cols"; }
    int tiles_per_plane, ntiles, nblocks; T scale;
    const C2<T>* in; T* out;
    const T* post_diag; int post_planes;

    DEV void operator()(int blk, unsigned char* smem) const {
        T* const buf = reinterpret_cast<T*>(smem);
        const C2<T>* const stg = reinterpret_cast<const C2<T>*>(smem + TB);
        uint64_t* const bar = reinterpret_cast<uint64_t*>(smem + TB + SB);
        const size_t nmap = (size_t)N * fc.Nx;
        typename Cm::Tw1 w1; typename Cm::Tw2 w2;
        auto src_of = [&](int t) { return in + ((size_t)(t / tiles_per_plane) * fc.Nx + (size_t)(t % tiles_per_plane) * M) * NYH; };
        CMBL_FOR_THREADS(tid, NT) { if (tid == 0) { mbar_init(bar, 1); mbar_init_fence(); } }
        CMBL_SYNC();
        int tile = blk;
        if (tile < ntiles) { CMBL_FOR_THREADS(tid, NT) { if (tid == 0) { mbar_expect_tx(bar, SB); bulk_load(const_cast<C2<T>*>(stg), src_of(tile), SB, bar); } } }
        for (int it = 0; tile < ntiles; tile += nblocks, ++it) {
            const int c = tile / tiles_per_plane, x0 = (tile % tiles_per_plane) * M, next = tile + nblocks;
            CMBL_SYNC();                                            // the previous tile's planes have left the buffer
            // Z[k] = A[k] + i·B[k] and Z[N−k] = conj A[k] + i·conj B[k], natural order
            CMBL_FOR_THREADS(tid, NT) {
#ifndef CMBL_EMU
                mbar_wait(bar, (unsigned)(it & 1));
#endif
                for (int idx = tid; idx < L * NYH; idx += NT) {
                    const int l = idx / NYH, k = idx % NYH;
                    const bool edge = (k == 0 || k == NH);
                    C2<T> a = stg[(2 * l) * NYH + k], b = stg[(2 * l + 1) * NYH + k];
                    if (edge) { a.y = 0; b.y = 0; }                // c2r ignores Im of the DC and Nyquist rows
                    T* re = buf + (2 * l) * N; T* im = re + N;
                    re[this->nat(k, 2 * l)] = a.x - b.y; im[this->nat(k, 2 * l + 1)] = a.y + b.x;
                    if (!edge) { re[this->nat(N - k, 2 * l)] = a.x + b.y; im[this->nat(N - k, 2 * l + 1)] = b.x - a.y; }
                }
            }
            CMBL_SYNC();
            if (next < ntiles) { CMBL_FOR_THREADS(tid, NT) { if (tid == 0) { mbar_expect_tx(bar, SB); bulk_load(const_cast<C2<T>*>(stg), src_of(next), SB, bar); } } }
            // every thread gathers the 16 frequencies of its radix-16 butterflies; first inverse sweep in registers, results in tile order
#ifdef CMBL_EMU
            T* const snap = reinterpret_cast<T*>(smem + TB + SB + 16);
            memcpy(snap, buf, TB);
            CMBL_FOR_THREADS(tid, NT) {
                for (int t = 0; t < TPT; ++t) {
                    int l, j, kb; Cm::task_of(tid, t, l, j, kb);
                    C2<T> v[16]; this->load16_nat(snap, l, kb, v); dft16<T, true>(v); this->store16_tile(buf, l, j, v);
                }
            }
#else
            {
                C2<T> v[TPT][16];
#pragma unroll
                for (int t = 0; t < TPT; ++t) { int l, j, kb; Cm::task_of(threadIdx.x, t, l, j, kb); this->load16_nat(buf, l, kb, v[t]); }
                __syncthreads();
#pragma unroll
                for (int t = 0; t < TPT; ++t) { int l, j, kb; Cm::task_of(threadIdx.x, t, l, j, kb); dft16<T, true>(v[t]); this->store16_tile(buf, l, j, v[t]); }
                fc.load_tw2(w2, threadIdx.x);
            }
#endif
            CMBL_SYNC();
            CMBL_FOR_THREADS(tid, NT) { CMBL_PRE_START(fc.load_tw2(w2, tid)); fc.template pass2<true>(buf, tid, w2); CMBL_PRE_END(fc.load_tw1(w1, tid)); }
            CMBL_SYNC();
            CMBL_FOR_THREADS(tid, NT) { CMBL_PRE_START(fc.load_tw1(w1, tid)); fc.template pass1<true>(buf, nullptr, tid, w1); }
            CMBL_SYNC();
            // the real planes (plane = column) leave in memory order, scaled, optionally times a Map-basis diagonal
            T* const dst = out + (size_t)c * nmap;
            const T* const pd = post_diag ? post_diag + (size_t)(c % post_planes) * nmap : nullptr;
            CMBL_FOR_THREADS(tid, NT) {
#pragma unroll 4
                for (int k = tid; k < M * CH; k += NT) {
                    int p, ch; size_t goff; fc.chunk_of(k, x0, p, ch, goff);
                    Vec<T> z = vload(buf + p * N + fc.swzp(ch, p) * V);
#pragma unroll
                    for (int e = 0; e < V; ++e) z.v[e] *= scale;
                    if (pd) { const Vec<T> d = vload_stream(pd + goff);
#pragma unroll
                        for (int e = 0; e < V; ++e) z.v[e] = d.v[e] * z.v[e]; }       // same product order as DiagMulBody on the stored map
                    vstore(dst + goff, z);
                }
            }
        }
    }
};

}  // namespace cmbl
