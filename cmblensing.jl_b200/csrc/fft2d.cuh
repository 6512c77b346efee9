// Batched 2-D real FFT over planes laid out like the reference's arrays: Map (Ny,Nx,C) column-major = [c][x][y],
// Fourier (Ny/2+1,Nx,C) = [c][x][ky]  (src/proj_cartesian.jl:51-56; transforms src/util_fft.jl:26-27 m_rfft!/m_irfft!).
//   rfft2 : column pass (two real columns packed as one complex line, R2C along y) then row pass (C2C along x,
//           in place on the output).  Unnormalised.
//   irfft2: row pass (inverse C2C along x, into plan scratch so the input is never clobbered, util_fft.jl:44) then
//           column pass (C2R along y: imaginary parts of the ky=0 and ky=Ny/2 rows are ignored, like FFTW / cuFFT).
//           Normalised by 1/(Ny·Nx).
#pragma once
#include "plan.cuh"

// launch shape of the three general-transform kernels (threads per block, minimum resident blocks per SM → register cap).
// Uncapped, the fp64 kernels take 196-255 registers and run ONE 256-thread block per SM, so a block's load → passes → store
// phases overlap with nothing; capping at 128 registers (a few hundred bytes of spills) doubles the residency and is 26 %
// faster at Nside=1024 (profiles/r01_fft_launch_shape.log; 128 threads × 3-4 blocks and 256 × 3 measured, not better).
#ifndef CMBL_FFT_NT
#define CMBL_FFT_NT 256
#endif
#ifndef CMBL_FFT_MINB
#define CMBL_FFT_MINB 2
#endif
#ifndef CMBL_FFT_ROW_UNR
#define CMBL_FFT_ROW_UNR 16         // strided loads of the row pass kept in flight per thread (the tile's runs are L·sizeof(C2) = 64 bytes):
                                    // all 16 of a thread instead of 4 — fft2_rows 118 -> 111 us fp64, 81 -> 72 us fp32 (profiles/r02_fft_row_unroll.log)
#endif

namespace cmbl {

constexpr int FFT_ROW_UNR = CMBL_FFT_ROW_UNR;

inline int tile_budget_bytes() {
    static int v = [] { const char* e = getenv("CMBL_TILE_KB"); int kb = e ? atoi(e) : 70; if (kb < 8) kb = 8; if (kb > 200) kb = 200; return kb * 1024; }();
    return v;
}
inline int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }

// lines per line-major tile of transforms of length N (other dimension Nother: 2 real lines per complex line)
template <class T> int col_lines(const Fft1D<T>& f, int Nother, int extra = 0) {
    size_t per = sizeof(C2<T>) * (size_t)Tile<T, false>::pitch_for(f.N, f.sk);
    int L = (int)(tile_budget_bytes() / per) - extra;
    if (L < 1) L = 1;
    L = pow2_floor(L);
    if (L > Nother / 2) L = Nother / 2;
    if (L > 64) L = 64;
    return L;
}
// lines per tile for row kernels (interleaved tile)
template <class T> int row_lines(int N, int maxL) {
    size_t per = Tile<T, true>::bytes(N, 1, 0);
    int L = (int)(tile_budget_bytes() / per);
    if (L < 1) L = 1;
    L = pow2_floor(L);
    if (L > maxL) L = pow2_floor(maxL);
    if (L > 64) L = 64;
    return L;
}

// ---------------------------------------------------------------------------------------------------------------
// column pass of rfft2
// ---------------------------------------------------------------------------------------------------------------
template <class T> struct R2CColBody {
    static constexpr int NT = CMBL_FFT_NT, MINB = CMBL_FFT_MINB;
    static const char* name() { return "rfft2_cols"; }
    Fft1D<T> fy; int Ny, Nx, Nyh, L, tiles_per_plane;
    const T* in; C2<T>* out;
    DEV void operator()(int blk, unsigned char* smem) const {
        const int c = blk / tiles_per_plane, x0 = (blk % tiles_per_plane) * 2 * L;
        Tile<T, false> tv = line_tile<T>(smem, L, fy);
        const T* src = in + ((size_t)c * Nx + x0) * Ny;
        CMBL_FOR_THREADS(tid, NT) {
            for (int e = tid; e < L * Ny; e += NT) {                     // Ny is a power of two: shifts, no division
                const int l = e >> fy.logN, y = e & (Ny - 1);
                tv.at(l, y) = mk<T>(src[(size_t)(2 * l) * Ny + y], src[(size_t)(2 * l + 1) * Ny + y]);
            }
        }
        CMBL_SYNC();
        fft_forward_passes<T, false, NT>(tv, fy, 0, fy.npass);
        C2<T>* dst = out + ((size_t)c * Nx + x0) * Nyh;
        CMBL_FOR_THREADS(tid, NT) {
            for (int k = tid; k < Nyh; k += NT) {                        // Nyh = Ny/2+1 is odd: loop per line instead of dividing
                const int pk = CMBL_LDG(&fy.pos[k]), pm = CMBL_LDG(&fy.pos[(Ny - k) & (Ny - 1)]);
                const T h = (T)0.5;
                for (int l = 0; l < L; ++l) {
                    const C2<T> zk = tv.at(l, pk), zm = tv.at(l, pm);
                    dst[(size_t)(2 * l) * Nyh + k] = mk<T>((zk.x + zm.x) * h, (zk.y - zm.y) * h);
                    dst[(size_t)(2 * l + 1) * Nyh + k] = mk<T>((zk.y + zm.y) * h, (zm.x - zk.x) * h);
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// row pass (C2C along x) on a chunk of L consecutive ky, forward (in place allowed) or inverse (in -> out)
// ---------------------------------------------------------------------------------------------------------------
template <class T, bool INV> struct C2CRowBody {
    static constexpr int NT = CMBL_FFT_NT, MINB = CMBL_FFT_MINB;
    static const char* name() { return "fft2_rows"; }
    Fft1D<T> fx; int Nx, Nyh, L, tiles_per_plane;
    const C2<T>* in; C2<T>* out;
    DEV void operator()(int blk, unsigned char* smem) const {
        const int c = blk / tiles_per_plane, k0 = (blk % tiles_per_plane) * L;
        Tile<T, true> tv{reinterpret_cast<C2<T>*>(smem), L, 0, 0};
        const C2<T>* src = in + (size_t)c * Nx * Nyh;
        C2<T>* dst = out + (size_t)c * Nx * Nyh;
        const int logL = ilog2(L);                                              // L is a power of two
        CMBL_FOR_THREADS(tid, NT) {
#pragma unroll FFT_ROW_UNR
            for (int e = tid; e < L * Nx; e += NT) {
                const int x = e >> logL, l = e & (L - 1);
                C2<T> v = (k0 + l < Nyh) ? src[(size_t)x * Nyh + k0 + l] : mk<T>(0, 0);
                tv.at(l, INV ? CMBL_LDG(&fx.pos[x]) : x) = v;
            }
        }
        CMBL_SYNC();
        if (INV) fft_inverse_passes<T, true, NT>(tv, fx, 0, fx.npass);
        else fft_forward_passes<T, true, NT>(tv, fx, 0, fx.npass);
        CMBL_FOR_THREADS(tid, NT) {
            for (int e = tid; e < L * Nx; e += NT) {
                const int x = e >> logL, l = e & (L - 1);
                if (k0 + l < Nyh) dst[(size_t)x * Nyh + k0 + l] = tv.at(l, INV ? x : CMBL_LDG(&fx.pos[x]));
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// column pass of irfft2 (C2R along y)
// ---------------------------------------------------------------------------------------------------------------
template <class T> struct C2RColBody {
    static constexpr int NT = CMBL_FFT_NT, MINB = CMBL_FFT_MINB;
    static const char* name() { return "irfft2_cols"; }
    Fft1D<T> fy; int Ny, Nx, Nyh, L, tiles_per_plane; T scale;
    const C2<T>* in; T* out;
    const T* post_diag = nullptr; int post_planes = 1;          // optional Map-basis diagonal applied to the result (plane c uses plane c % post_planes)
    DEV void operator()(int blk, unsigned char* smem) const {
        const int c = blk / tiles_per_plane, x0 = (blk % tiles_per_plane) * 2 * L;
        Tile<T, false> tv = line_tile<T>(smem, L, fy);
        const C2<T>* src = in + ((size_t)c * Nx + x0) * Nyh;
        CMBL_FOR_THREADS(tid, NT) {
            for (int k = tid; k < Nyh; k += NT) {
                const bool edge = (k == 0 || 2 * k == Ny);
                const int pk = CMBL_LDG(&fy.pos[k]), pm = edge ? 0 : CMBL_LDG(&fy.pos[Ny - k]);
                for (int l = 0; l < L; ++l) {
                    C2<T> a = src[(size_t)(2 * l) * Nyh + k], b = src[(size_t)(2 * l + 1) * Nyh + k];
                    if (edge) { a.y = 0; b.y = 0; }                            // c2r ignores Im of DC / Nyquist rows
                    tv.at(l, pk) = mk<T>(a.x - b.y, a.y + b.x);
                    if (!edge) tv.at(l, pm) = mk<T>(a.x + b.y, b.x - a.y);
                }
            }
        }
        CMBL_SYNC();
        fft_inverse_passes<T, false, NT>(tv, fy, 0, fy.npass);
        T* dst = out + ((size_t)c * Nx + x0) * Ny;
        const T* pd = post_diag ? post_diag + ((size_t)(c % post_planes) * Nx + x0) * Ny : nullptr;
        CMBL_FOR_THREADS(tid, NT) {
            for (int e = tid; e < L * Ny; e += NT) {
                const int l = e >> fy.logN, y = e & (Ny - 1);
                C2<T> z = tv.at(l, y);
                const size_t ia = (size_t)(2 * l) * Ny + y, ib = ia + Ny;
                T a = z.x * scale, b = z.y * scale;
                if (pd) { a = pd[ia] * a; b = pd[ib] * b; }                    // same product order as DiagMulBody on the stored map
                dst[ia] = a; dst[ib] = b;
            }
        }
    }
};

// G > 0 (both directions): the MAP side is in the row-grouped layout of the fast stage kernels with G rows per group (flow_fast.cuh) instead
// of the reference layout — the transforms that open and close a flow then hand over the integrator's own buffer and the two layout
// conversions disappear.  Only where fft_rg_io_ok(P) (the persistent column kernels of fft2d_fast.cuh address either layout).
template <class T> bool fft_rg_io_ok(const PlanT<T>& P);
template <class T> void rfft2(PlanT<T>& P, const T* map, C2<T>* four, int C, cmblStream_t st, int G = 0);
// post_diag: optional REAL Map-basis diagonal (post_planes planes, broadcast over the batch) multiplied into the result — the
// pixel mask of M = Mfourier·Mpix rides on the transform's store instead of a separate pass over the maps
template <class T> void irfft2(PlanT<T>& P, const C2<T>* four, T* map, int C, cmblStream_t st, const T* post_diag = nullptr, int post_planes = 1, int G = 0);

}  // namespace cmbl
