/* cmbl_b200.h — C ABI of libcmbl_b200.so, the B200 (sm_100a) implementation of CMBLensing.jl's flat-sky hot path.
 *
 * The reference (marius311/CMBLensing.jl @ 8e75a7c) has no FFI of its own: its GPU backend is a set of Julia methods
 * dispatched on the storage type `A<:CuArray` of `BaseField{B,M,T,A}` (src/base_fields.jl:14, ext/CMBLensingCUDAExt.jl:28).
 * Each entry point below names the reference method(s) it replaces; the Julia-side `ccall` shim a maintainer would add
 * is shown in INTEGRATION.md (julia/CMBLensingB200Ext.jl).
 *
 * Conventions
 *   - every function returns 0 on success, a negative code on failure; cmbl_last_error() returns the message of the
 *     last failure on the calling thread (the Julia shim turns it into an ErrorException, like the reference's error()).
 *   - all array pointers are DEVICE pointers to caller-owned, densely packed buffers in the reference's own layout:
 *       Map     array (Ny, Nx, Npol, Nb) column-major real T              (src/proj_cartesian.jl:51-56)
 *       Fourier array (Ny/2+1, Nx, Npol, Nb) column-major Complex{T} (interleaved re,im)
 *     C below always means the number of (Ny,Nx) planes = Npol*Nb.  The `_host` variants take HOST pointers and include
 *     the host<->device copies.
 *   - dtype: 0 = Float32, 1 = Float64.  basis: 0 = Map, 1 = Fourier.
 *   - `stream` is a cudaStream_t (pass CUDA.stream() from Julia, 0 for the default stream).  Calls are asynchronous on it
 *     unless they return host values.  Handles are not thread-safe; distinct handles are independent.
 */
#ifndef CMBL_B200_H
#define CMBL_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cmbl_plan cmbl_plan;
typedef struct cmbl_flow cmbl_flow;
typedef struct cmbl_cg cmbl_cg;
typedef struct cmbl_comm cmbl_comm;

#define CMBL_OK 0
#define CMBL_ERR_INVALID (-1)
#define CMBL_ERR_CUDA (-2)
#define CMBL_ERR_NUMERIC (-3)

#define CMBL_MAP 0
#define CMBL_FOURIER 1

/* LenseFlow operations, src/flowops.jl:11-14 */
#define CMBL_OP_L 0       /* Lϕ * f   : Map -> Map,         t: 0 -> 1 */
#define CMBL_OP_LH 1      /* Lϕ' * f  : Fourier -> Fourier, t: 1 -> 0 */
#define CMBL_OP_LINV 2    /* Lϕ \ f   : Map -> Map,         t: 1 -> 0 */
#define CMBL_OP_LHINV 3   /* Lϕ' \ f  : Fourier -> Fourier, t: 0 -> 1 */

const char* cmbl_last_error(void);
const char* cmbl_version(void);
/* number of kernels this library has launched in this process so far (bench.py reports the per-step delta) */
long long cmbl_launch_count(void);
/* per-kernel device timing (CUDA events around every launch of this library, on the launching stream) between
 * cmbl_profile_begin() and cmbl_profile_end(); the latter synchronises and returns lines "kernel count total_ms\n".
 * Diagnostic only (bench.py's roofline line): event pairs add launch gaps, so never enable it inside a timed region. */
int cmbl_profile_begin(void);
const char* cmbl_profile_end(void);

/* ---- plan: ProjLambert metadata + FFT plan (src/proj_lambert.jl:24-75, src/util_fft.jl:32-39) ------------------- */
int cmbl_plan_create(cmbl_plan** plan, int device, int Ny, int Nx, double theta_pix_arcmin, int dtype);
int cmbl_plan_destroy(cmbl_plan* plan);
/* copies grid metadata to HOST arrays (any pointer may be NULL): lx[Nx], ly[Ny/2+1], lam_rfft[Ny/2+1],
 * sin2phi/cos2phi [(Ny/2+1)*Nx] in the Fourier layout, all of the plan's dtype; scalars[5] (double) =
 * {Δx, Δℓx, Δℓy, Ωpix, nyquist} */
int cmbl_plan_grids(cmbl_plan* plan, void* lx, void* ly, void* lam_rfft, void* sin2phi, void* cos2phi, double* scalars);

/* Cℓ_to_Cov(:I, proj, Cℓ; units) (src/proj_lambert.jl:173-175,361-364): out = nan2zero(Cℓ(ℓmag)) / units as ONE real half-plane (Ny/2+1, Nx) of the
 * plan's dtype on the device; Cℓ = linear interpolation of the HOST table (ell ascending), NaN -> 0 outside it.  units = 0 means Ωpix.
 * (:P / :IP operators are stacks of such planes: EE, BB / TT, TE, EE, BB.)  Synchronises. */
int cmbl_cl_to_cov(cmbl_plan* plan, const double* ell_host, const double* cl_host, int n, double units, void* out, void* stream);

/* ---- batched 2-D real FFT: m_rfft! / m_irfft! (src/util_fft.jl:26-27; call sites src/proj_lambert.jl:245-300) ---- */
int cmbl_rfft2(cmbl_plan* plan, const void* map, void* four, int C, void* stream);        /* unnormalised            */
int cmbl_irfft2(cmbl_plan* plan, const void* four, void* map, int C, void* stream);       /* 1/(Ny Nx); input intact */

/* ---- diagonal / basis operators ------------------------------------------------------------------------------------ */
/* DiagOp*f and DiagOp\f (src/specialops.jl:9-10): out = diag .* in, or nan2zero(in ./ diag) when `ldiv` != 0.
 * basis selects the element type of in/out (Map: real, Fourier: complex); diag is REAL with Cd planes, Cd == C or
 * Cd == Npol_d (broadcast over the batch: plane c uses diag plane c % Cd). in may alias out. */
int cmbl_diag_mul(cmbl_plan* plan, int basis, const void* diag, int Cd, const void* in, void* out, int C, int ldiv, void* stream);
/* Field broadcasts with per-batch scalars (BatchedReal, src/batching.jl:9-45): out = a .* x .+ b .* y, e.g. `x + α*Δ`, the leap-frog and
 * line-search updates of the callers.  a_host[na], b_host[nb]: HOST doubles, length 1 or Nb; y may be NULL (out = a .* x).  x, y, out:
 * fields of Npol*Nb planes in `basis` (out may alias x or y).  Nb <= 64. */
int cmbl_field_axpby(cmbl_plan* plan, int basis, const double* a_host, int na, const void* x, const double* b_host, int nb, const void* y_or_null,
                     void* out, int Npol, int Nb, void* stream);
/* QU<->EB rotation in Fourier space (src/proj_lambert.jl:253-271). dir 0: EB->QU, 1: QU->EB.  The two planes of each of
 * the Nb pairs are consecutive; pair_stride_planes = Npol of the array (2 for QU, 3 for IQU with first_plane = 1). */
int cmbl_qu_eb(cmbl_plan* plan, int dir, const void* in, void* out, int Nb, int pair_stride_planes, int first_plane, void* stream);
/* BlockDiagIEB on an IEBFourier field of Nb batch items (src/specialops.jl:61-118; 2x2 sqrt/pinv src/field_vectors.jl:62-78).
 * block: 4 REAL half-planes [SigmaTE[1,1], SigmaTE[2,1], SigmaTE[2,2], SigmaB] (Cl_to_Cov(:IP) builds the 2x2 block symmetric,
 * src/proj_lambert.jl:368-371).  mode 0: L*f, 1: L\f = pinv(L)*f, 2: sqrt(L)*f (simulate).  in may alias out. */
int cmbl_blockdiag_ieb(cmbl_plan* plan, int mode, const void* block, const void* in, void* out, int Nb, void* stream);
/* dot(a,b) per batch item (src/proj_lambert.jl:318-328): Map: Σ a·b; Fourier: Σ Re(conj(a) b) λ_rfft / (Ny Nx).
 * out_host[Nb] (double, HOST); synchronises the stream. */
int cmbl_dot(cmbl_plan* plan, int basis, const void* a, const void* b, int Npol, int Nb, double* out_host, void* stream);

/* ---- LenseFlow (src/lenseflow.jl, src/flowops.jl) ----------------------------------------------------------------- */
/* CachedLenseFlow for fields with Npol planes per batch item, Nb_f batch items, and Nb_phi (1 or Nb_f) distinct ϕ.
 * nsteps = RK4 steps (reference default 7, src/lenseflow.jl:29). */
int cmbl_lenseflow_create(cmbl_flow** flow, cmbl_plan* plan, int nsteps, int Npol, int Nb_f, int Nb_phi);
int cmbl_lenseflow_destroy(cmbl_flow* flow);
/* precompute! (src/lenseflow.jl:131-142): ϕ (Nb_phi planes; phi_basis Map or Fourier) -> p[τ] (and M⁻¹[τ] when
 * with_minv != 0, needed only by cmbl_lenseflow_grad) at the 2*nsteps+1 times. */
int cmbl_lenseflow_precompute(cmbl_flow* flow, const void* phi, int phi_basis, int with_minv, void* stream);
/* *, \ on FlowOp / Adjoint (src/flowops.jl:11-14). ops 0,2: in/out Map; ops 1,3: in/out Fourier. in may alias out. */
int cmbl_lenseflow_apply(cmbl_flow* flow, int op, const void* in, void* out, void* stream);
/* same, HOST buffers, copies included (the end-to-end path bench.py times).  Returns when out_host is valid.  For the
 * map-space flows (ops 0, 2) the batch items are pipelined: H2D of the next item group, integration of the current one and
 * D2H of the finished one overlap on separate streams (pin the host buffers to get the overlap; CMBL_HOST_CHUNKS = number
 * of groups, default 3 = a quarter / half / quarter of the batch: the first and last group are the exposed transfers, the
 * middle one keeps the persistent stage kernels filled; 1 = copy-in / compute / copy-out back to back). */
int cmbl_lenseflow_apply_host(cmbl_flow* flow, int op, const void* in_host, void* out_host, void* stream);
/* the same for a caller that streams many fields through one operator: returns once the work is queued; successive calls overlap (H2D of
 * call i+1 and D2H of call i-1 run during the integration of call i: two staging slots, two copy streams), so a call costs
 * max(integration, one-way transfer) instead of their sum.  in_host must stay unchanged, and out_host is valid only, after
 * cmbl_lenseflow_host_sync() (which waits for every asynchronous call of the calling host thread). */
int cmbl_lenseflow_apply_host_async(cmbl_flow* flow, int op, const void* in_host, void* out_host, void* stream);
int cmbl_lenseflow_host_sync(void);
/* pullback through Lϕ*f (op 0) or Lϕ\f (op 2): negδvelocityᴴ transpose flow (src/lenseflow.jl:176-214, src/flowops.jl:40-68).
 * f_out_map = the forward result (Map), delta = cotangent (Fourier). Outputs: dfield (Fourier, C planes), dphi (Fourier,
 * Nb_f planes: one per batch item, also when a single ϕ is shared by the batch — sum them for the gradient w.r.t. the
 * shared ϕ).  Needs cmbl_lenseflow_precompute(..., with_minv = 1).  bug_compat != 0 reproduces the reference's aliased
 * 2x2 product (src/lenseflow.jl:198-200 with src/field_vectors.jl:48-49). */
int cmbl_lenseflow_grad(cmbl_flow* flow, int op, const void* f_out_map, const void* delta_four, void* dfield_four,
                        void* dphi_four, int bug_compat, void* stream);
/* get_max_lensing_step(ϕ, η) (src/lenseflow.jl:242-256): per batch item the smallest α > 0 at which 𝕀 + ∇∇(ϕ + α η) becomes singular in
 * some pixel (out_host[Nb], double, +inf if there is none) — ϕ + α η stays in the weak-lensing regime LenseFlow needs for α below it.
 * ϕ, η: Nb planes each, Map or Fourier.  Synchronises. */
int cmbl_max_lensing_step(cmbl_plan* plan, const void* phi, int phi_basis, const void* eta, int eta_basis, int Nb, double* out_host, void* stream);
/* which stage kernels this flow runs (diagnostic): bit 0 = fast persistent row kernel, bit 1 = fast persistent column kernel
 * (csrc/flow_fast.cuh; transform length 256/512/1024/2048), 0 = generic kernels of csrc/flow.cuh */
int cmbl_lenseflow_kernel_path(cmbl_flow* flow);
/* read back one cached p map set for tests: out_host = p[k] as (Ny,Nx,2,Nb_phi) of the plan's dtype */
int cmbl_lenseflow_get_p(cmbl_flow* flow, int k, void* out_host);

/* ---- CG Wiener filter: argmaxf_logpdf for a BaseDataSet with diagonal Cf, Cn, B, Mfourier and a pixel mask ---------
 * (src/maximization.jl:17-42, src/dataset.jl:76-80,129-132, src/numerical_algorithms.jl:73-134).
 * Npol = 1 | 2: all diagonals are REAL device arrays of `Npol` half-planes in the harmonic basis of the field (Fourier |
 * EBFourier), shared by every batch item.  Npol = 3 (pol = :IP, IEBFourier): every operator is a BlockDiagIEB given as 4 REAL
 * half-planes [SigmaTE[1,1], SigmaTE[2,1], SigmaTE[2,2], SigmaB] (see cmbl_blockdiag_ieb); the preconditioner is then the
 * BlockDiagIEB pinv(Cf) + Bhat'Mhat'pinv(Cnhat)Mhat Bhat (src/specialops.jl:99-102) and `M \ r` = pinv(M)*r (:78).
 * mask_pix is a REAL Map-basis diagonal of Npol planes or NULL. */
typedef struct cmbl_dataset_desc {
    int Npol, Nb;
    const void* Cf;        /* signal covariance          */
    const void* Cn;        /* noise covariance           */
    const void* Cnhat;     /* approximate noise cov (preconditioner) */
    const void* B;         /* beam                       */
    const void* Bhat;
    const void* Mf;        /* Fourier part of the mask M */
    const void* mask_pix;  /* pixel part of M (Map basis, Npol planes) or NULL */
    const void* d;         /* data, harmonic basis, Npol*Nb half-planes (complex) */
} cmbl_dataset_desc;

int cmbl_cg_create(cmbl_cg** cg, cmbl_flow* flow, const cmbl_dataset_desc* ds, void* stream);
int cmbl_cg_destroy(cmbl_cg* cg);
/* builds b (and a₀), x = fstart (or 0), r, z, p; returns res = dot(r,z) per batch item in res_host[Nb] (history entry i=1) */
int cmbl_cg_begin(cmbl_cg* cg, const void* fstart_or_null, int offset, double* res_host, void* stream);
/* one iteration of the loop body (numerical_algorithms.jl:99-121); returns the new res per batch item in res_host[Nb], or — with
 * res_host == NULL — leaves it on the device and does not synchronise (poll with a later call that passes a buffer) */
int cmbl_cg_step(cmbl_cg* cg, double* res_host, void* stream);
/* record the current x as bestx (the caller implements the lock-step `all(res<bestres)` rule, possibly across ranks) */
int cmbl_cg_mark_best(cmbl_cg* cg, void* stream);
/* copy bestx (which=0) or the current x (which=1) to f_out (harmonic basis, device) */
int cmbl_cg_result(cmbl_cg* cg, int which, void* f_out, void* stream);
/* whole solve on one device: conjugate_gradient's loop incl. bestx and `all(res<tol)` stop; res_hist_host[nsteps*Nb] */
int cmbl_wiener_cg(cmbl_cg* cg, const void* fstart_or_null, void* f_out, int nsteps, double tol, int offset,
                   int* iters_out, double* res_hist_host, void* stream);
/* gradientf_logpdf (src/dataset.jl:76-80) at f with data d (d_or_null = NULL: the dataset's d); all harmonic basis, device */
int cmbl_gradientf_logpdf(cmbl_cg* cg, const void* f, const void* d_or_null, int d_is_zero, void* out, void* stream);

/* ---- multi-GPU: the path shards over independent batch items / chains (src/batching.jl, pmap over chains src/sampling.jl:292-307) with no
 * data-path collective.  What couples the shards is a handful of scalars: conjugate_gradient's lock-step rules all(res<bestres) /
 * all(res<tol) over the batch (src/numerical_algorithms.jl:110-121) and MAP_joint's batch-summed line-search objective
 * (src/maximization.jl:197).  One process per GPU; rank 0 calls cmbl_comm_unique_id and hands the 128 bytes to the other ranks by whatever
 * the host program uses (MPI, Distributed.jl, torch.distributed); every rank then calls cmbl_comm_init on its device.  NCCL is loaded
 * with dlopen at that moment (CMBL_NCCL_LIB overrides the path); single-GPU users never need it. */
int cmbl_comm_unique_id(void* id128);
int cmbl_comm_init(cmbl_comm** comm, int nranks, int rank, const void* id128);
int cmbl_comm_destroy(cmbl_comm* comm);
/* in-place all-reduce of n (<= 64) HOST doubles over the ranks; op 0 = sum, 1 = min, 2 = max.  Synchronises the stream. */
int cmbl_comm_allreduce(cmbl_comm* comm, double* values_host, int n, int op, void* stream);
/* cmbl_wiener_cg with the batch sharded over the ranks of `comm` (NULL: this rank alone): every rank runs the same number of iterations and
 * keeps bestx by the rule evaluated over all batch items of all ranks — one 16-byte all-reduce per iteration. */
int cmbl_wiener_cg_sharded(cmbl_cg* cg, cmbl_comm* comm_or_null, const void* fstart_or_null, void* f_out, int nsteps, double tol, int offset,
                           int* iters_out, double* res_hist_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif
