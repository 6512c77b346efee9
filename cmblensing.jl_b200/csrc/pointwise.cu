#include "pointwise.cuh"
#include "../../include/cmbl_b200.h"

namespace cmbl {

template <class T> void diag_mul(PlanT<T>& P, int basis, const T* diag, int Cd, const void* in, void* out, int C, bool ldiv, cmblStream_t st) {
    CMBL_REQUIRE(Cd >= 1 && C % Cd == 0, "diagonal planes must divide field planes");
    if (basis == CMBL_FOURIER) {
        DiagMulBody<T, true> b{P.four_elems(), P.four_elems() * (size_t)C, Cd, ldiv, diag, in, out};
        launch(b, (int)((b.total + b.NT - 1) / b.NT), 0, st);
    } else {
        DiagMulBody<T, false> b{P.map_elems(), P.map_elems() * (size_t)C, Cd, ldiv, diag, in, out};
        launch(b, (int)((b.total + b.NT - 1) / b.NT), 0, st);
    }
}

template <class T> void qu_eb(PlanT<T>& P, int dir, const C2<T>* in, C2<T>* out, int Nb, int stride_planes, int first_plane, cmblStream_t st) {
    CMBL_REQUIRE(dir == 0 || dir == 1, "dir must be 0 (EB->QU) or 1 (QU->EB)");
    CMBL_REQUIRE(stride_planes >= 2 && first_plane >= 0 && first_plane + 2 <= stride_planes, "bad plane stride / first plane");
    QuEbBody<T> b{Nb, stride_planes, first_plane, dir, P.four_elems(), P.sin2phi, P.cos2phi, in, out};
    launch(b, (int)((b.nf * Nb + b.NT - 1) / b.NT), 0, st);
}

template <class T> void blockdiag_ieb(PlanT<T>& P, int mode, const T* block, const C2<T>* in, C2<T>* out, int Nb, cmblStream_t st) {
    CMBL_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (L*f), 1 (L\\f) or 2 (sqrt(L)*f)");
    BlockIebBody<T> b{P.four_elems(), Nb, mode, block, in, out};
    launch(b, (int)((b.nf * Nb + b.NT - 1) / b.NT), 0, st);
}

template <class T> void dot_partials(PlanT<T>& P, int basis, const void* a, const void* b, int Npol, int Nb, double* partial, cmblStream_t st) {
    if (basis == CMBL_FOURIER) {
        DotBody<T, true> k{P.four_elems() * (size_t)Npol, P.Nyh, P.lam, 1.0 / ((double)P.Ny * (double)P.Nx), a, b, partial};
        launch(k, Nb * RED_BLOCKS, sizeof(double) * k.NT, st);
    } else {
        DotBody<T, false> k{P.map_elems() * (size_t)Npol, P.Nyh, P.lam, 1.0, a, b, partial};
        launch(k, Nb * RED_BLOCKS, sizeof(double) * k.NT, st);
    }
}

#define INST(T)                                                                                                   \
    template void diag_mul<T>(PlanT<T>&, int, const T*, int, const void*, void*, int, bool, cmblStream_t);        \
    template void qu_eb<T>(PlanT<T>&, int, const C2<T>*, C2<T>*, int, int, int, cmblStream_t);                    \
    template void blockdiag_ieb<T>(PlanT<T>&, int, const T*, const C2<T>*, C2<T>*, int, cmblStream_t);            \
    template void dot_partials<T>(PlanT<T>&, int, const void*, const void*, int, int, double*, cmblStream_t);
INST(float)
INST(double)

}  // namespace cmbl
