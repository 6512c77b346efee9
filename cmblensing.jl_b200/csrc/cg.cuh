// CG Wiener filter on the device: gradientf_logpdf (src/dataset.jl:76-80), Hessian preconditioner (:129-132),
// argmaxf_logpdf (src/maximization.jl:17-42) and conjugate_gradient (src/numerical_algorithms.jl:73-134).
#pragma once
#include "flow.cuh"
#include "pointwise.cuh"
#include "../../include/cmbl_b200.h"

namespace cmbl {

struct CgBase { PlanBase* plan = nullptr; virtual ~CgBase() {} };

template <class T> struct CgT : CgBase {
    PlanT<T>* P = nullptr; FlowT<T>* F = nullptr;
    int Npol = 1, Nb = 1, C = 1;
    // dataset diagonals (device, caller-owned) and derived diagonals (owned)
    const T *Cf = nullptr, *Cn = nullptr, *Cnhat = nullptr, *B = nullptr, *Bhat = nullptr, *Mf = nullptr, *mask = nullptr;
    const C2<T>* d = nullptr;
    DevBuf inv_Cf, inv_Cn, inv_Cn_Mf, precond;          // pinv(Cf), pinv(Cn), Mf·pinv(Cn), Hessian preconditioner
    // CG vectors (harmonic basis, C half-planes each) and scratch
    DevBuf x, r, z, p, Ap, b, bestx, w1, w2, m1;
    DevBuf scal;                                          // doubles: res[2][Nb], pAp_part[Nb][RED], res_part[Nb][RED]
    int flip = 0; bool begun = false;
    std::vector<double> h_res;
    size_t nf() const { return P->four_elems(); }
    size_t nmap() const { return P->map_elems(); }
    double* res_cur() const { return reinterpret_cast<double*>(scal.p) + (size_t)flip * Nb; }
    double* res_next() const { return reinterpret_cast<double*>(scal.p) + (size_t)(1 - flip) * Nb; }
    double* pAp_part() const { return reinterpret_cast<double*>(scal.p) + 2 * (size_t)Nb; }
    double* res_part() const { return pAp_part() + (size_t)Nb * RED_BLOCKS; }
};

template <class T> void cg_setup(CgT<T>& G, const cmbl_dataset_desc& ds, cmblStream_t st);
// out = gradientf_logpdf(f, d)   (d == nullptr with d_zero: data ≡ 0)
template <class T> void cg_gradientf(CgT<T>& G, const C2<T>* f, const C2<T>* d, bool d_zero, C2<T>* out, cmblStream_t st);
template <class T> void cg_begin(CgT<T>& G, const C2<T>* fstart, bool offset, double* res_host, cmblStream_t st);
template <class T> void cg_step(CgT<T>& G, double* res_host, cmblStream_t st);

}  // namespace cmbl
