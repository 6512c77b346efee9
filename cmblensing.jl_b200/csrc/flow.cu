#include "flow_fast.cuh"
#include "../../include/cmbl_b200.h"
#include <algorithm>

namespace cmbl {

template <class T> void flow_precompute(FlowT<T>& F, const void* phi, int phi_basis, bool with_minv, cmblStream_t st) {
    PlanT<T>& P = *F.P;
    const size_t nmap = P.map_elems(), nf = P.four_elems();
    const int nk = 2 * F.nsteps + 1;
    const C2<T>* phif;
    if (phi_basis == CMBL_FOURIER) phif = reinterpret_cast<const C2<T>*>(phi);
    else {
        C2<T>* s = reinterpret_cast<C2<T>*>(F.spec.reserve(sizeof(C2<T>) * nf * F.Nbphi));
        rfft2<T>(P, reinterpret_cast<const T*>(phi), s, F.Nbphi, st);
        phif = s;
    }
    // the 5 gradient/Hessian spectra live behind the (optional) ϕ spectrum in the same scratch buffer
    DevBuf& ghs = F.gh;
    C2<T>* spec5 = reinterpret_cast<C2<T>*>(ghs.reserve(sizeof(C2<T>) * nf * 5 * F.Nbphi + sizeof(T) * nmap * 5 * F.Nbphi));
    T* maps5 = reinterpret_cast<T*>(spec5 + nf * 5 * F.Nbphi);
    {
        GradHessSpecBody<T> b{P.Nx, P.Nyh, P.lx, P.ly, phif, spec5, nf * (size_t)F.Nbphi};
        launch(b, (int)((b.total + b.NT - 1) / b.NT), 0, st);
    }
    irfft2<T>(P, spec5, maps5, 5 * F.Nbphi, st);
    T* pc = reinterpret_cast<T*>(F.pcache.reserve(sizeof(T) * nmap * 2 * F.Nbphi * nk));
    T* mi = with_minv ? reinterpret_cast<T*>(F.minv.reserve(sizeof(T) * nmap * 3 * F.Nbphi * nk)) : nullptr;
    {
        F.pcache_G = flow_rg_rows(P);
        PCacheBody<T> b{nk, F.Nbphi, nmap, maps5, pc, mi, F.pcache_G, P.Nx, P.Ny};
        launch(b, (int)((nmap * F.Nbphi + b.NT - 1) / b.NT), 0, st);
    }
    F.have_p = true; F.have_minv = with_minv;
}

static bool fast_enabled() { static const bool v = [] { const char* e = getenv("CMBL_FLOW_FAST"); return !e || atoi(e) != 0; }(); return v; }
static unsigned fast_stagger_ns() { static const unsigned v = [] { const char* e = getenv("CMBL_FLOW_STAGGER_NS"); return e ? (unsigned)atoi(e) : 0u; }(); return v; }
static int device_sms() {
#ifdef CMBL_EMU
    return 2;
#else
    static thread_local int v = 0;
    if (!v) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); }
    return v;
#endif
}
// blocks that compute and publish each J[N] plane pair at launch start (0 = none: every block computes private copies — test knob)
static int fast_cols_jn_red() { static const int v = [] { const char* e = getenv("CMBL_COL_JN_RED"); return e ? atoi(e) : 3; }(); return v; }
// bounded patience of a flag look: polls of ~0.15 us each (default ~15 us in total; 0 = look once)
static int fast_cols_jn_polls() { static const int v = [] { const char* e = getenv("CMBL_COL_JN_POLLS"); return e ? atoi(e) : 100; }(); return v; }
// opt-in experiment (CMBL_COL_PAIR=1): the Q and U tile of the same columns run as a 2-block cluster whose epilogues start together (cluster
// barrier), hoping that the p maps both read come from DRAM once.  Measured (profiles/r02_col_cluster_pair.log): DRAM reads unchanged
// (683 vs 679 MB per launch) and +4 % time, so it is off by default.
static int fast_cols_pair() { static const int v = [] { const char* e = getenv("CMBL_COL_PAIR"); return e ? atoi(e) : 0; }(); return v; }
// Programmatic dependent launch between the stage kernels (the next kernel's blocks are scheduled and load their tables while the previous one
// drains; griddepcontrol.wait orders the data).  Measured by shape (profiles/r02_pdl_by_shape.log): fp64 gains everywhere (0.3-2.5 % on the
// machine-filling launches, 13-21 % on small ones, 15 % on the adjoint flow of Nside=1024 batch 1), fp32 gains only on launches of a few tiles
// (13-25 % at 256², but +2..+12 % on the larger ones) — so: on for fp64, on for fp32 launches with fewer tiles than half the SMs.  CMBL_PDL=0/1 overrides.
template <class T> static int fast_pdl_hint(int ntiles) { return (sizeof(T) == 8 || 2 * ntiles <= device_sms()) ? 1 : 0; }
static int fast_block_cap(int full) {     // experiment knob: cap the persistent grid at N blocks per SM
    static const int v = [] { const char* e = getenv("CMBL_FLOW_BLOCKS_PER_SM"); return e ? atoi(e) : 0; }();
    return v > 0 ? std::min(full, v * device_sms()) : full;
}
// L2 prefetch of the column kernel's epilogue operands: 0 off, 1-4 = issued before sweep 2 / the middle / sweep 4 / sweep 5.
// Default: before sweep 4 for fp64 and for the adjoint kernels of both precisions (measured at Nside=1024 batch 8, fp64: L*f
// 7.12 -> 7.04 ms, L'*f 8.11 -> 7.63 ms, Nside=512 -4 %; fp32: L'*f 4.47 -> 4.32 ms), off for the fp32 forward kernel (neutral to
// slightly negative).  CMBL_FLOW_PF overrides all.  (profiles/r01_prefetch_sweep.log)
static int fast_pf(size_t elem_bytes, bool adj) {
    static const int v = [] { const char* e = getenv("CMBL_FLOW_PF"); return e ? atoi(e) : -1; }();
    return v >= 0 ? v : ((elem_bytes == 8 || adj) ? 3 : 0);
}

template <class T, int LOGN, bool ADJ>
static void fast_rows(FlowT<T>& F, int c0, int nC, const T* u, int kq, T wgt, cmblStream_t st) {
    PlanT<T>& P = *F.P;
    static const bool use_tma = [] { const char* e = getenv("CMBL_ROW_TMA"); return !e || atoi(e) != 0; }();
    if (use_tma || fast_row_tile_bytes<T>(1 << LOGN) != 32768) {
        typedef TmaRowBody<T, LOGN, ADJ> B;
        B b;
        b.fx = P.ax.fft; b.mult = P.ax.fmult_deriv;
        b.Ny = P.Ny; b.tiles_per_plane = P.Ny / B::ROWS; b.ntiles = nC * b.tiles_per_plane;
        b.nblocks = std::min(b.ntiles, fast_block_cap(persistent_blocks<B>(B::SMEM)));
        b.Npol = F.Npol; b.Nbphi = F.Nbphi; b.cbase = c0;
        b.u = u; b.pk = F.pk(kq); b.tmp = reinterpret_cast<T*>(F.tmp.p); b.nline = reinterpret_cast<T*>(F.nline.p);
        b.nacc = reinterpret_cast<T*>(F.nacc.p); b.wgt = wgt;
        launch(b, b.nblocks, B::SMEM, st, 1, fast_pdl_hint<T>(b.ntiles));
        return;
    }
    if constexpr (fast_row_tile_bytes<T>(1 << LOGN) == 32768) {
    typedef FastRowBody<T, LOGN, ADJ> B;
    B b;
    b.fx = P.ax.fft; b.mult = P.ax.fmult_deriv;
    b.Ny = P.Ny; b.tiles_per_plane = P.Ny / B::ROWS; b.ntiles = nC * b.tiles_per_plane;
    b.nblocks = std::min(b.ntiles, fast_block_cap(persistent_blocks<B>(B::SMEM)));
    b.Npol = F.Npol; b.Nbphi = F.Nbphi; b.cbase = c0; b.sms = device_sms(); b.stagger_ns = fast_stagger_ns();
    b.u = u; b.pk = F.pk(kq); b.tmp = reinterpret_cast<T*>(F.tmp.p); b.nline = reinterpret_cast<T*>(F.nline.p);
    b.nacc = reinterpret_cast<T*>(F.nacc.p); b.wgt = wgt;
    launch(b, b.nblocks, B::SMEM, st);
    }
}
template <class T, class B>
static void fast_cols_launch(FlowT<T>& F, int c0, int nC, const T* u, int kq, T wgt, const T* ybase, const T* acc_in, T* acc_out, T* u_out, T ca, T cb,
                             int private_lines, bool adj, cmblStream_t st, T* dx_out = nullptr, T* dy_out = nullptr) {
    PlanT<T>& P = *F.P;
    B b;
    b.dx_out = dx_out; b.dy_out = dy_out;
    b.tw1 = P.ay.ftw1; b.tw2 = P.ay.ftw2; b.mult_d = P.ay.fmult_deriv; b.mult_sign = P.ay.fmult_sign; b.cN = P.ax.ell_nyq / (T)P.Nx;
    b.Nx = P.Nx; b.G = flow_rg_rows(P); b.lgGV = ilog2(b.G / B::V); b.tiles_per_plane = P.Nx / B::M; b.ntiles = nC * b.tiles_per_plane;
    // every block that can be resident; a launch with fewer tiles than that gets a few extra blocks that own no tile and only compute and
    // publish the J[N] lines, so that no tile-owning block of a one-wave launch is delayed by them
    b.jn_red = fast_cols_jn_red();
    // (experiment knob) forward kernel on QU fields: clusters of two blocks (Q tile, U tile of the same columns) with aligned epilogues
    const int cluster = (!adj && F.Npol == 2 && fast_cols_pair() && b.ntiles >= 4) ? 2 : 1;
    const int cap = cluster == 2 ? fast_block_cap(persistent_blocks_clustered<B>(B::SMEM, 2)) & ~1 : fast_block_cap(persistent_blocks<B>(B::SMEM));
    b.nblocks = std::min(b.ntiles + b.jn_red * ((nC + 1) / 2), cap);
    if (cluster == 2) { b.nblocks &= ~1; b.csync = 1; }
    b.Npol = F.Npol; b.Nbphi = F.Nbphi; b.cbase = c0; b.pf = fast_pf(sizeof(T), adj); b.sms = device_sms(); b.stagger_ns = fast_stagger_ns();
    b.u = u; b.pk = F.pk(kq); b.tmp = reinterpret_cast<T*>(F.tmp.p); b.macc = reinterpret_cast<T*>(F.macc.p); b.wgt = wgt;
    b.nline = reinterpret_cast<T*>(F.nline.p); b.jn = nullptr;
    b.jn_pub = reinterpret_cast<T*>(F.jn.p);
    b.jn_blk = reinterpret_cast<T*>(F.jnblk.reserve(sizeof(T) * (size_t)b.nblocks * private_lines * B::N)); b.jn_polls = fast_cols_jn_polls();
    if (F.jnflag.cap < sizeof(int) * (size_t)F.C) { F.jnflag.reserve(sizeof(int) * (size_t)F.C); dev_zero(F.jnflag.p, sizeof(int) * (size_t)F.C, st); }
    b.jn_flag = reinterpret_cast<int*>(F.jnflag.p); b.epoch = ++F.jn_epoch;
    b.ybase = ybase; b.acc_in = acc_in; b.acc_out = acc_out; b.u_out = u_out; b.ca = ca; b.cb = cb;
    launch(b, b.nblocks, B::SMEM, st, cluster, fast_pdl_hint<T>(b.ntiles));
}
template <class T, int LOGN, bool ADJ>
static void fast_cols(FlowT<T>& F, int c0, int nC, const T* u, int kq, T wgt, const T* ybase, const T* acc_in, T* acc_out, T* u_out, T ca, T cb, cmblStream_t st,
                      T* dx_out, T* dy_out) {
    if constexpr (!ADJ) {
        if (dx_out) { fast_cols_launch<T, FastColBody<T, LOGN, false, true>>(F, c0, nC, u, kq, wgt, ybase, acc_in, acc_out, u_out, ca, cb, 1, false, st, dx_out, dy_out); return; }
    }
    fast_cols_launch<T, FastColBody<T, LOGN, ADJ>>(F, c0, nC, u, kq, wgt, ybase, acc_in, acc_out, u_out, ca, cb, 1, ADJ, st);
}
template <class T> static bool fast_rows_ok(const PlanT<T>& P) {
    if (!fast_enabled() || !fast_len_ok(P.Nx) || !P.ax.ftw1) return false;
    const int rows = (fast_row_tile_bytes<T>(P.Nx) / (P.Nx * 16)) * (16 / (int)sizeof(T));
    return P.Ny % rows == 0;
}
template <class T> static bool fast_cols_ok(const PlanT<T>& P) {
    if (!fast_enabled() || !fast_len_ok(P.Ny) || !P.ay.ftw1) return false;
    const int cols = fast_tile_bytes(P.Ny) / (int)sizeof(T) / P.Ny;
    return P.Nx % cols == 0;
}

// rows per group of the row-grouped internal layout (flow_fast.cuh), 0 = the generic kernels on the reference layout
template <class T> int flow_rg_rows(const PlanT<T>& P) {
    if (!fast_rows_ok(P) || !fast_cols_ok(P) || P.Ny % 64 != 0 || P.Nx % 32 != 0) return 0;
    return (fast_row_tile_bytes<T>(P.Nx) / (P.Nx * 16)) * (16 / (int)sizeof(T));
}
template <class T, bool TO_RG> void convert_layout(PlanT<T>& P, int G, const T* in, T* out, int C, cmblStream_t st) {
    typedef LayoutBody<T, TO_RG> B;
    B b{P.Ny, P.Nx, G, in, out};
    launch(b, C * (P.Nx / B::TX) * (P.Ny / B::TY), B::SMEM, st);
}

template <class T, bool ADJ>
void flow_stage(FlowT<T>& F, int c0, int nC, const T* u, int kq, T wgt, const T* ybase, const T* acc_in, T* acc_out, T* u_out,
                T ca, T cb, cmblStream_t st, T* dx_out, T* dy_out) {
    PlanT<T>& P = *F.P;
    T* tmp = reinterpret_cast<T*>(F.tmp.p); T* nline = reinterpret_cast<T*>(F.nline.p); T* jn = reinterpret_cast<T*>(F.jn.p);
    const bool fast = flow_rg_rows(P) > 0;
    if (fast) {
        switch (P.Nx) {
            case 256: fast_rows<T, 8, ADJ>(F, c0, nC, u, kq, wgt, st); break;
            case 512: fast_rows<T, 9, ADJ>(F, c0, nC, u, kq, wgt, st); break;
            case 1024: fast_rows<T, 10, ADJ>(F, c0, nC, u, kq, wgt, st); break;
            default: fast_rows<T, 11, ADJ>(F, c0, nC, u, kq, wgt, st); break;
        }
    } else {
        FlowRowBody<T, ADJ> b;
        b.fx = P.ax.fft; b.fy = P.ay.fft; b.mult = P.ax.mult_deriv; b.mult_sign_y = P.ay.mult_sign; b.cN = P.ax.ell_nyq / (T)P.Nx;
        b.Ny = P.Ny; b.Nx = P.Nx; b.L = col_lines<T>(P.ax.fft, P.Ny); b.logL = ilog2(b.L); b.tiles_per_plane = P.Ny / (2 * b.L);
        b.Npol = F.Npol; b.Nbphi = F.Nbphi; b.cbase = c0;
        b.u = u; b.pk = F.pk(kq); b.tmp = tmp; b.nline = nline; b.jn = jn; b.nacc = reinterpret_cast<T*>(F.nacc.p); b.wgt = wgt;
        b.counter = reinterpret_cast<int*>(F.counter.p);
        launch(b, nC * b.tiles_per_plane, FlowRowBody<T, ADJ>::smem_bytes(b.fx, b.fy, b.L), st);
    }
    if (fast) {
        switch (P.Ny) {
            case 256: fast_cols<T, 8, ADJ>(F, c0, nC, u, kq, wgt, ybase, acc_in, acc_out, u_out, ca, cb, st, dx_out, dy_out); break;
            case 512: fast_cols<T, 9, ADJ>(F, c0, nC, u, kq, wgt, ybase, acc_in, acc_out, u_out, ca, cb, st, dx_out, dy_out); break;
            case 1024: fast_cols<T, 10, ADJ>(F, c0, nC, u, kq, wgt, ybase, acc_in, acc_out, u_out, ca, cb, st, dx_out, dy_out); break;
            default: fast_cols<T, 11, ADJ>(F, c0, nC, u, kq, wgt, ybase, acc_in, acc_out, u_out, ca, cb, st, dx_out, dy_out); break;
        }
    } else {
        CMBL_REQUIRE(!dx_out, "derivative export needs the fast stage kernels");
        FlowColBody<T, ADJ> b;
        b.fy = P.ay.fft; b.mult_d = P.ay.mult_deriv;
        b.Ny = P.Ny; b.Nx = P.Nx; b.L = col_lines<T>(P.ay.fft, P.Nx); b.logNyv = ilog2(P.Ny / Vec<T>::N); b.tiles_per_plane = P.Nx / (2 * b.L);
        b.Npol = F.Npol; b.Nbphi = F.Nbphi; b.cbase = c0;
        b.u = u; b.pk = F.pk(kq); b.tmp = tmp; b.jn = jn; b.macc = reinterpret_cast<T*>(F.macc.p); b.wgt = wgt;
        b.ybase = ybase; b.acc_in = acc_in; b.acc_out = acc_out; b.u_out = u_out; b.ca = ca; b.cb = cb;
        launch(b, nC * b.tiles_per_plane, Tile<T, false>::bytes(P.Ny, b.L, P.ay.fft.sk), st);
    }
}

// working buffers of the integrator, sized for all F.C planes
template <class T> void flow_reserve(FlowT<T>& F, cmblStream_t st) {
    PlanT<T>& P = *F.P;
    const size_t nmap = P.map_elems();
    if (flow_rg_rows(P)) F.yrg.reserve(sizeof(T) * nmap * F.C);
    F.acc.reserve(sizeof(T) * nmap * F.C);
    F.ubuf.reserve(sizeof(T) * nmap * F.C);
    F.tmp.reserve(sizeof(T) * nmap * F.C);
    F.nline.reserve(sizeof(T) * (size_t)P.Ny * F.C);
    F.jn.reserve(sizeof(T) * (size_t)P.Ny * F.C);
    if (F.counter.cap < sizeof(int) * (size_t)F.C) { F.counter.reserve(sizeof(int) * (size_t)F.C); dev_zero(F.counter.p, sizeof(int) * (size_t)F.C, st); }
}

// RK4 over stages k0 → k1 for the planes [c0, c0 + nC) of the caller's state `ycaller` (reference layout, F.C planes).
// Planes of different batch items are independent, so a caller may integrate plane ranges one after another (the pipelined
// host path below overlaps their transfers with the integration of their neighbours).
template <class T> void flow_integrate_range(FlowT<T>& F, bool adj, T* ycaller, int k0, int k1, int c0, int nC, cmblStream_t st, bool rg_state) {
    CMBL_REQUIRE(F.have_p, "LenseFlow used before cmbl_lenseflow_precompute");
    CMBL_REQUIRE(c0 >= 0 && nC >= 1 && c0 + nC <= F.C && c0 % F.Npol == 0 && nC % F.Npol == 0, "plane range must cover whole batch items");
    PlanT<T>& P = *F.P;
    const size_t nmap = P.map_elems();
    const int n = F.nsteps;
    const int G = flow_rg_rows(P);
    CMBL_REQUIRE(G == F.pcache_G, "p-cache layout does not match the kernel path");
    flow_reserve(F, st);
    CMBL_REQUIRE(!rg_state || G, "a row-grouped state needs the fast stage kernels");
    T* y = ycaller;
    if (G && !rg_state) {                           // integrate on a row-grouped copy of the state (flow_fast.cuh)
        y = reinterpret_cast<T*>(F.yrg.p);
        convert_layout<T, true>(P, G, ycaller + (size_t)c0 * nmap, y + (size_t)c0 * nmap, nC, st);
    }
    T* acc = reinterpret_cast<T*>(F.acc.p);
    T* ub = reinterpret_cast<T*>(F.ubuf.p);
    const int sgn = k1 > k0 ? 1 : -1;
    const double h = (double)sgn / n;
    const T h2 = (T)(h / 2), h1 = (T)h, h6 = (T)(h / 6), h3 = (T)(h / 3);
    int kk = k0;
    for (int step = 0; step < n; ++step) {
        for (int s = 0; s < 4; ++s) {
            const int kq = kk + (s == 0 ? 0 : (s < 3 ? sgn : 2 * sgn));
            const T* u = (s == 0) ? y : ub;
            const T* ybase = (s == 3) ? nullptr : y;
            const T* acc_in = (s == 0) ? nullptr : acc;
            T* acc_out = (s == 3) ? y : acc;
            T* u_out = (s == 3) ? nullptr : ub;
            const T ca = (s < 2) ? h2 : h1, cb = (s == 0 || s == 3) ? h6 : h3;
            if (adj) flow_stage<T, true>(F, c0, nC, u, kq, cb, ybase, acc_in, acc_out, u_out, ca, cb, st);
            else flow_stage<T, false>(F, c0, nC, u, kq, cb, ybase, acc_in, acc_out, u_out, ca, cb, st);
        }
        kk += 2 * sgn;
    }
    if (G && !rg_state) convert_layout<T, false>(P, G, y + (size_t)c0 * nmap, ycaller + (size_t)c0 * nmap, nC, st);
    F.integrated_once = true;
}

template <class T> int flow_rg_direct(FlowT<T>& F) {
    static const bool on = [] { const char* e = getenv("CMBL_RG_DIRECT"); return !e || atoi(e) != 0; }();
    const int G = flow_rg_rows(*F.P);
    return (on && G > 0 && fft_rg_io_ok(*F.P)) ? G : 0;
}
template <class T> void flow_integrate(FlowT<T>& F, bool adj, T* y, int k0, int k1, cmblStream_t st, bool rg_state) {
    static const int chunk_env = [] { const char* e = getenv("CMBL_FLOW_CHUNK"); return e ? atoi(e) : 0; }();
    int chunk = (chunk_env > 0 && chunk_env < F.C) ? chunk_env : F.C;             // planes integrated together (L2 residency)
    chunk = (chunk + F.Npol - 1) / F.Npol * F.Npol;
    for (int c0 = 0; c0 < F.C; c0 += chunk)
        flow_integrate_range<T>(F, adj, y, k0, k1, c0, (F.C - c0 < chunk) ? F.C - c0 : chunk, st, rg_state);
}

template <class T> void flow_apply(FlowT<T>& F, int op, const void* in, void* out, cmblStream_t st) {
    PlanT<T>& P = *F.P;
    const size_t nmap = P.map_elems(), nf = P.four_elems();
    const int n = F.nsteps;
    CMBL_REQUIRE(op >= 0 && op <= 3, "LenseFlow op must be 0..3");
    if (op == CMBL_OP_L || op == CMBL_OP_LINV) {
        T* y = reinterpret_cast<T*>(out);
        if (in != out) dev_copy(y, in, sizeof(T) * nmap * F.C, st);
        if (op == CMBL_OP_L) flow_integrate<T>(F, false, y, 0, 2 * n, st);
        else flow_integrate<T>(F, false, y, 2 * n, 0, st);
        return;
    }
    // adjoint flows: Fourier state integrated in map space (see flow.cuh)
    // the transforms around the flow read / write the integrator's row-grouped buffer directly where they can (no layout conversions)
    const int Gd = flow_rg_direct(F);
    T* y = reinterpret_cast<T*>(Gd ? F.yrg.reserve(sizeof(T) * nmap * F.C) : F.ybuf.reserve(sizeof(T) * nmap * F.C));
    flow_adj_prepare<T>(F, reinterpret_cast<const C2<T>*>(in), y, st, Gd);
    if (op == CMBL_OP_LH) flow_integrate<T>(F, true, y, 2 * n, 0, st, Gd > 0);
    else flow_integrate<T>(F, true, y, 0, 2 * n, st, Gd > 0);
    flow_adj_finish<T>(F, y, reinterpret_cast<C2<T>*>(out), st, Gd);
    (void)nf;
}

// Fourier state Y0 of an adjoint flow -> its map y = irfft2(Y0), with the ky ∈ {0, Ny/2} rows saved and the Nyquist accumulators cleared
template <class T> void flow_adj_prepare(FlowT<T>& F, const C2<T>* Y0, T* y, cmblStream_t st, int G) {
    PlanT<T>& P = *F.P;
    C2<T>* rows0 = reinterpret_cast<C2<T>*>(F.rows0.reserve(sizeof(C2<T>) * 2 * (size_t)P.Nx * F.C));
    T* nacc = reinterpret_cast<T*>(F.nacc.reserve(sizeof(T) * (size_t)P.Ny * F.C));
    T* macc = reinterpret_cast<T*>(F.macc.reserve(sizeof(T) * (size_t)P.Nx * F.C));
    {
        AdjRowsSaveBody<T> b{P.Nx, P.Nyh, Y0, rows0};
        launch(b, F.C, 0, st);
    }
    dev_zero(nacc, sizeof(T) * (size_t)P.Ny * F.C, st);
    dev_zero(macc, sizeof(T) * (size_t)P.Nx * F.C, st);
    irfft2<T>(P, Y0, y, F.C, st, nullptr, 1, G);
}
// integrated map y -> Fourier result: rfft2(y) plus what a map cannot carry (saved rows, Nyquist accumulators)
template <class T> void flow_adj_finish(FlowT<T>& F, const T* y, C2<T>* Yout, cmblStream_t st, int G) {
    PlanT<T>& P = *F.P;
    rfft2<T>(P, y, Yout, F.C, st, G);
    AdjFixBody<T> b;
    b.fx = P.ax.fft; b.Ny = P.Ny; b.Nx = P.Nx; b.Nyh = P.Nyh; b.lxN = P.ax.ell_nyq; b.lyN = P.ay.ell_nyq;
    b.rows0 = reinterpret_cast<C2<T>*>(F.rows0.p); b.nacc = reinterpret_cast<T*>(F.nacc.p); b.macc = reinterpret_cast<T*>(F.macc.p); b.out = Yout;
    size_t smem = sizeof(C2<T>) * Tile<T, false>::pitch_for(P.Nx, P.ax.fft.sk) + sizeof(T) * 2 * b.NT;
    launch(b, F.C, smem, st);
}

template <class T> void max_lensing_step(PlanT<T>& P, const void* phi, int phi_basis, const void* eta, int eta_basis, int Nb, double* out_host, cmblStream_t st) {
    const size_t nmap = P.map_elems(), nf = P.four_elems();
    DevBuf spec, ws;
    C2<T>* sp = reinterpret_cast<C2<T>*>(spec.reserve(sizeof(C2<T>) * nf * (size_t)Nb));
    // per field: 5 spectra + 5 maps per batch item; [ϕ | η]
    const size_t per = sizeof(C2<T>) * nf * 5 * Nb + sizeof(T) * nmap * 5 * Nb;
    char* w = reinterpret_cast<char*>(ws.reserve(2 * per + sizeof(double) * (size_t)Nb * RED_BLOCKS));
    const T* maps[2];
    for (int k = 0; k < 2; ++k) {
        const void* src = k ? eta : phi; const int basis = k ? eta_basis : phi_basis;
        const C2<T>* four = reinterpret_cast<const C2<T>*>(src);
        if (basis != CMBL_FOURIER) { rfft2<T>(P, reinterpret_cast<const T*>(src), sp, Nb, st); four = sp; }
        C2<T>* s5 = reinterpret_cast<C2<T>*>(w + k * per);
        T* m5 = reinterpret_cast<T*>(s5 + nf * 5 * Nb);
        GradHessSpecBody<T> b{P.Nx, P.Nyh, P.lx, P.ly, four, s5, nf * (size_t)Nb};
        launch(b, (int)((b.total + b.NT - 1) / b.NT), 0, st);
        irfft2<T>(P, s5, m5, 5 * Nb, st);
        maps[k] = m5;
    }
    double* part = reinterpret_cast<double*>(w + 2 * per);
    MaxStepBody<T> mb{nmap, maps[0], maps[1], part};
    launch(mb, Nb * RED_BLOCKS, sizeof(double) * mb.NT, st);
    std::vector<double> h((size_t)Nb * RED_BLOCKS);
    dev_download(h.data(), part, sizeof(double) * h.size(), st);
    for (int b = 0; b < Nb; ++b) {
        double m = HUGE_VAL;
        for (int i = 0; i < RED_BLOCKS; ++i) m = std::min(m, h[(size_t)b * RED_BLOCKS + i]);
        out_host[b] = m;
    }
#ifndef CMBL_EMU
    CMBL_CUDA(cudaStreamSynchronize(st));            // the scratch buffers are released on return
#endif
}

template <class T> int flow_kernel_path(FlowT<T>& F) { return flow_rg_rows(*F.P) > 0 ? 3 : 0; }

#define INST(T)                                                                                        \
    template void flow_precompute<T>(FlowT<T>&, const void*, int, bool, cmblStream_t);                 \
    template void flow_integrate<T>(FlowT<T>&, bool, T*, int, int, cmblStream_t, bool);                \
    template void flow_integrate_range<T>(FlowT<T>&, bool, T*, int, int, int, int, cmblStream_t, bool); \
    template int flow_rg_direct<T>(FlowT<T>&);                                                         \
    template void flow_apply<T>(FlowT<T>&, int, const void*, void*, cmblStream_t);                     \
    template int flow_kernel_path<T>(FlowT<T>&);                                                       \
    template int flow_rg_rows<T>(const PlanT<T>&);                                                     \
    template void flow_reserve<T>(FlowT<T>&, cmblStream_t);                                            \
    template void flow_adj_prepare<T>(FlowT<T>&, const C2<T>*, T*, cmblStream_t, int);                 \
    template void flow_adj_finish<T>(FlowT<T>&, const T*, C2<T>*, cmblStream_t, int);                  \
    template void convert_layout<T, true>(PlanT<T>&, int, const T*, T*, int, cmblStream_t);            \
    template void convert_layout<T, false>(PlanT<T>&, int, const T*, T*, int, cmblStream_t);           \
    template void flow_stage<T, false>(FlowT<T>&, int, int, const T*, int, T, const T*, const T*, T*, T*, T, T, cmblStream_t, T*, T*);  \
    template void flow_stage<T, true>(FlowT<T>&, int, int, const T*, int, T, const T*, const T*, T*, T*, T, T, cmblStream_t, T*, T*);   \
    template void max_lensing_step<T>(PlanT<T>&, const void*, int, const void*, int, int, double*, cmblStream_t);
INST(float)
INST(double)

}  // namespace cmbl
