#include "fft2d.cuh"

namespace cmbl {

template <class T> void rfft2(PlanT<T>& P, const T* map, C2<T>* four, int C, cmblStream_t st) {
    if (C <= 0) return;
    {
        R2CColBody<T> b;
        b.fy = P.ay.fft; b.Ny = P.Ny; b.Nx = P.Nx; b.Nyh = P.Nyh;
        b.L = col_lines<T>(P.ay.fft, P.Nx); b.tiles_per_plane = P.Nx / (2 * b.L);
        b.in = map; b.out = four;
        launch(b, C * b.tiles_per_plane, Tile<T, false>::bytes(P.Ny, b.L, P.ay.fft.sk), st);
    }
    {
        C2CRowBody<T, false> b;
        b.fx = P.ax.fft; b.Nx = P.Nx; b.Nyh = P.Nyh;
        b.L = row_lines<T>(P.Nx, P.Nyh); b.tiles_per_plane = (P.Nyh + b.L - 1) / b.L;
        b.in = four; b.out = four;
        launch(b, C * b.tiles_per_plane, Tile<T, true>::bytes(P.Nx, b.L, 0), st);
    }
}

template <class T> void irfft2(PlanT<T>& P, const C2<T>* four, T* map, int C, cmblStream_t st) {
    if (C <= 0) return;
    C2<T>* scratch = reinterpret_cast<C2<T>*>(P.scratch_four.reserve(sizeof(C2<T>) * P.four_elems() * (size_t)C));
    {
        C2CRowBody<T, true> b;
        b.fx = P.ax.fft; b.Nx = P.Nx; b.Nyh = P.Nyh;
        b.L = row_lines<T>(P.Nx, P.Nyh); b.tiles_per_plane = (P.Nyh + b.L - 1) / b.L;
        b.in = four; b.out = scratch;
        launch(b, C * b.tiles_per_plane, Tile<T, true>::bytes(P.Nx, b.L, 0), st);
    }
    {
        C2RColBody<T> b;
        b.fy = P.ay.fft; b.Ny = P.Ny; b.Nx = P.Nx; b.Nyh = P.Nyh;
        b.L = col_lines<T>(P.ay.fft, P.Nx); b.tiles_per_plane = P.Nx / (2 * b.L);
        b.scale = (T)1 / ((T)P.Ny * (T)P.Nx);
        b.in = scratch; b.out = map;
        launch(b, C * b.tiles_per_plane, Tile<T, false>::bytes(P.Ny, b.L, P.ay.fft.sk), st);
    }
}

template void rfft2<float>(PlanT<float>&, const float*, C2<float>*, int, cmblStream_t);
template void rfft2<double>(PlanT<double>&, const double*, C2<double>*, int, cmblStream_t);
template void irfft2<float>(PlanT<float>&, const C2<float>*, float*, int, cmblStream_t);
template void irfft2<double>(PlanT<double>&, const C2<double>*, double*, int, cmblStream_t);

}  // namespace cmbl
