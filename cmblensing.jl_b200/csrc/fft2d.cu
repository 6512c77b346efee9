#include "fft2d.cuh"
#include "fft2d_fast.cuh"
#include <algorithm>

namespace cmbl {

// column passes on the persistent tile kernels of fft2d_fast.cuh (Ny = 256 … 2048); CMBL_FFT_FAST=0 keeps the generic ones (A/B, tests)
static bool fft_fast_enabled() { static const bool v = [] { const char* e = getenv("CMBL_FFT_FAST"); return !e || atoi(e) != 0; }(); return v; }
template <class T> static bool fft_cols_fast_ok(const PlanT<T>& P) {
    if (!fft_fast_enabled() || !fast_len_ok(P.Ny) || !P.ay.ftw1) return false;
    return P.Nx % (fast_tile_bytes(P.Ny) / (int)sizeof(T) / P.Ny) == 0;
}
template <class T> bool fft_rg_io_ok(const PlanT<T>& P) { return fft_cols_fast_ok(P); }
template <class T, class B> static void fft_cols_fast_setup(PlanT<T>& P, B& b, int nC, int G) {
    b.fc.tw1 = P.ay.ftw1; b.fc.tw2 = P.ay.ftw2;
    b.fc.Nx = P.Nx; b.fc.G = G ? G : P.Ny; b.fc.lgGV = ilog2(b.fc.G / B::V);   // the reference layout is the row-grouped layout with a single group
    b.tiles_per_plane = P.Nx / B::M; b.ntiles = nC * b.tiles_per_plane;
    b.nblocks = std::min(b.ntiles, persistent_blocks<B>(B::SMEM));
}
template <class T, int LOGN> static void rfft2_cols_fast(PlanT<T>& P, const T* in, C2<T>* out, int nC, cmblStream_t st, int G) {
    typedef FastR2CColBody<T, LOGN> B;
    B b{};
    fft_cols_fast_setup<T, B>(P, b, nC, G);
    b.Nyh = P.Nyh; b.in = in; b.out = out;
    launch(b, b.nblocks, B::SMEM, st);
}
template <class T, int LOGN> static void irfft2_cols_fast(PlanT<T>& P, const C2<T>* in, T* out, int nC, const T* post_diag, int post_planes, cmblStream_t st, int G) {
    typedef FastC2RColBody<T, LOGN> B;
    B b{};
    fft_cols_fast_setup<T, B>(P, b, nC, G);
    b.scale = (T)1 / ((T)P.Ny * (T)P.Nx); b.in = in; b.out = out; b.post_diag = post_diag; b.post_planes = post_planes;
    launch(b, b.nblocks, B::SMEM, st);
}

// Planes per pass pair.  The two 1-D passes of a 2-D transform exchange a half-plane spectrum per plane; transforming a few
// planes at a time keeps that intermediate (and, for irfft2, a scratch buffer reused by every group) inside the 126 MB L2.
// Measured on B200 at Nside=1024, 16 planes (profiles/r01_fft_chunk.log): it does NOT pay — these kernels are bound by their
// load → transform → store structure, not by DRAM (16 planes at once 160 µs per pass; 2-plane groups 8 × 33 µs) — so the
// default is all planes in one pass pair.  CMBL_FFT_CHUNK_MB = spectrum bytes per group (0 = all planes).
template <class T> static int fft_chunk_planes(const PlanT<T>& P, int C) {
    static const int mb = [] { const char* e = getenv("CMBL_FFT_CHUNK_MB"); return e ? atoi(e) : 0; }();
    if (mb <= 0) return C;
    const size_t per = sizeof(C2<T>) * P.four_elems();
    int n = (int)(((size_t)mb << 20) / per);
    if (n < 1) n = 1;
    return n < C ? n : C;
}

template <class T> void rfft2(PlanT<T>& P, const T* map, C2<T>* four, int C, cmblStream_t st, int G) {
    if (C <= 0) return;
    CMBL_REQUIRE(!G || fft_cols_fast_ok(P), "row-grouped map input needs the persistent column kernels");
    const int chunk = fft_chunk_planes(P, C);
    for (int c0 = 0; c0 < C; c0 += chunk) {
        const int nC = (C - c0 < chunk) ? C - c0 : chunk;
        const T* in = map + (size_t)c0 * P.map_elems();
        C2<T>* out = four + (size_t)c0 * P.four_elems();
        if (fft_cols_fast_ok(P)) {
            switch (P.Ny) {
                case 256: rfft2_cols_fast<T, 8>(P, in, out, nC, st, G); break;
                case 512: rfft2_cols_fast<T, 9>(P, in, out, nC, st, G); break;
                case 1024: rfft2_cols_fast<T, 10>(P, in, out, nC, st, G); break;
                default: rfft2_cols_fast<T, 11>(P, in, out, nC, st, G); break;
            }
        } else {
            R2CColBody<T> b;
            b.fy = P.ay.fft; b.Ny = P.Ny; b.Nx = P.Nx; b.Nyh = P.Nyh;
            b.L = col_lines<T>(P.ay.fft, P.Nx); b.tiles_per_plane = P.Nx / (2 * b.L);
            b.in = in; b.out = out;
            launch(b, nC * b.tiles_per_plane, Tile<T, false>::bytes(P.Ny, b.L, P.ay.fft.sk), st);
        }
        {
            C2CRowBody<T, false> b;
            b.fx = P.ax.fft; b.Nx = P.Nx; b.Nyh = P.Nyh;
            b.L = row_lines<T>(P.Nx, P.Nyh); b.tiles_per_plane = (P.Nyh + b.L - 1) / b.L;
            b.in = out; b.out = out;
            launch(b, nC * b.tiles_per_plane, Tile<T, true>::bytes(P.Nx, b.L, 0), st);
        }
    }
}

template <class T> void irfft2(PlanT<T>& P, const C2<T>* four, T* map, int C, cmblStream_t st, const T* post_diag, int post_planes, int G) {
    if (C <= 0) return;
    CMBL_REQUIRE(!G || (fft_cols_fast_ok(P) && !post_diag), "row-grouped map output needs the persistent column kernels and no Map-basis diagonal");
    const int chunk = fft_chunk_planes(P, C);
    C2<T>* scratch = reinterpret_cast<C2<T>*>(P.scratch_four.reserve(sizeof(C2<T>) * P.four_elems() * (size_t)chunk));   // reused by every group
    for (int c0 = 0; c0 < C; c0 += chunk) {
        const int nC = (C - c0 < chunk) ? C - c0 : chunk;
        {
            C2CRowBody<T, true> b;
            b.fx = P.ax.fft; b.Nx = P.Nx; b.Nyh = P.Nyh;
            b.L = row_lines<T>(P.Nx, P.Nyh); b.tiles_per_plane = (P.Nyh + b.L - 1) / b.L;
            b.in = four + (size_t)c0 * P.four_elems(); b.out = scratch;
            launch(b, nC * b.tiles_per_plane, Tile<T, true>::bytes(P.Nx, b.L, 0), st);
        }
        CMBL_REQUIRE(!post_diag || (c0 % post_planes == 0), "plane groups must start on a diagonal period");
        if (fft_cols_fast_ok(P)) {
            T* out = map + (size_t)c0 * P.map_elems();
            switch (P.Ny) {
                case 256: irfft2_cols_fast<T, 8>(P, scratch, out, nC, post_diag, post_planes, st, G); break;
                case 512: irfft2_cols_fast<T, 9>(P, scratch, out, nC, post_diag, post_planes, st, G); break;
                case 1024: irfft2_cols_fast<T, 10>(P, scratch, out, nC, post_diag, post_planes, st, G); break;
                default: irfft2_cols_fast<T, 11>(P, scratch, out, nC, post_diag, post_planes, st, G); break;
            }
        } else {
            C2RColBody<T> b;
            b.fy = P.ay.fft; b.Ny = P.Ny; b.Nx = P.Nx; b.Nyh = P.Nyh;
            b.L = col_lines<T>(P.ay.fft, P.Nx); b.tiles_per_plane = P.Nx / (2 * b.L);
            b.scale = (T)1 / ((T)P.Ny * (T)P.Nx);
            b.in = scratch; b.out = map + (size_t)c0 * P.map_elems();
            b.post_diag = post_diag; b.post_planes = post_planes;
            launch(b, nC * b.tiles_per_plane, Tile<T, false>::bytes(P.Ny, b.L, P.ay.fft.sk), st);
        }
    }
}

template void rfft2<float>(PlanT<float>&, const float*, C2<float>*, int, cmblStream_t, int);
template void rfft2<double>(PlanT<double>&, const double*, C2<double>*, int, cmblStream_t, int);
template void irfft2<float>(PlanT<float>&, const C2<float>*, float*, int, cmblStream_t, const float*, int, int);
template void irfft2<double>(PlanT<double>&, const C2<double>*, double*, int, cmblStream_t, const double*, int, int);
template bool fft_rg_io_ok<float>(const PlanT<float>&);
template bool fft_rg_io_ok<double>(const PlanT<double>&);

}  // namespace cmbl
