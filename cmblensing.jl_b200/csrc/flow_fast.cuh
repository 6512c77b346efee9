// Fast path of the two LenseFlow stage kernels (flow.cuh) for transform lengths 256 / 512 / 1024 / 2048, written for B200:
//
//   * persistent blocks (grid = resident blocks of the whole GPU) that walk over tiles blk, blk+G, blk+2G, ...
//   * every global→shared transfer is asynchronous (cp.async, 16 B per request) into a double-buffered tile, so the
//     next tile is in flight during all five FFT sweeps of the current one; operands that are only consumed by the
//     column kernel's RK epilogue are pulled into L2 a few µs ahead with bulk L2 prefetches
//   * shared-memory layouts are built from the 16-byte chunk the copy engine delivers and XOR-swizzled at chunk
//     granularity so that all five sweeps (strides N/R1, 16, 1) are bank-conflict free with 128-bit accesses:
//       column kernel: planar tile, plane p = column x0+p, chunk = V consecutive y;  a thread owns V adjacent butterflies
//       row kernel   : tile [cl][x], chunk = V consecutive rows at one x (= V/2 complex lines); a thread owns one
//                      butterfly index and keeps its twiddles in registers for the whole launch
//   * schedule [R1, R2, 16] (plan.cu's schedule up to 1024, so the tile-order multiplier tables are shared; [8, 16, 16] with its own
//     tables at 2048, where tiles are 64 KB and blocks 256 threads): forward R1, forward R2, fused register-resident middle
//     (forward 16 · multiplier · inverse 16), inverse R2, inverse R1.
//   * the state of the ODE (y, u, tmp, acc) and the p-cache live in a library-internal ROW-GROUPED layout while the fast
//     kernels run:  element (x, y) of a plane at  ((y / G)·Nx + x)·G + y % G,  G = rows of one row-kernel tile
//     (G·sizeof(T) = 32 B at Nx = 1024).  A row tile (all x, G rows) is then ONE contiguous 32 KB run and a column tile
//     (M columns, all y) is Ny/G runs of M·G·sizeof(T) bytes (128 B fp64 / 256 B fp32) — in the reference's column-major
//     layout a row tile is 1024 pieces of 32 B at an 8 KB stride, which DRAM serves at < 20 % of its bandwidth
//     (measured: 1.2 TB/s).  LayoutBody converts caller buffers on entry / exit of flow_integrate; the transforms that open and close
//     a flow inside the library (fft2d_fast.cuh) address the row-grouped buffer directly.
// The arithmetic is the generic kernels' (same butterflies, twiddles and order of operations per element).
#pragma once
#include "flow.cuh"

namespace cmbl {

// ---------------------------------------------------------------------------------------------------------------
// asynchronous copy / cache-hint helpers (plain copies in the host emulator)
// ---------------------------------------------------------------------------------------------------------------
DEV void cp_async16(void* smem_dst, const void* gsrc) {
#ifdef __CUDA_ARCH__
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    // .ca, not .cg: the L1-bypassing form (LDGSTS.BYPASS) writes shared memory one 32-byte sector per wavefront — exactly 4x the
    // ideal wavefront count in the ncu source page, a third of the column kernel's shared-memory traffic; through L1 the landing
    // costs less (column kernel 188 -> 183 us fp64, 100 -> 97 us fp32; profiles/r01_cpasync_ca.log)
#ifdef CMBL_CPASYNC_CG
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
#endif
#else
    memcpy(smem_dst, gsrc, 16);
#endif
}
DEV void cp_async_commit() {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
DEV void cp_async_wait_all() {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}
// pull [p, p+bytes) into L2 (bytes multiple of 16)
DEV void l2_prefetch(const void* p, unsigned bytes) {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#else
    (void)p; (void)bytes;
#endif
}
DEV void l2_prefetch_line(const void* p) {                 // one 128-byte line
#ifdef __CUDA_ARCH__
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p) : "memory");
#else
    (void)p;
#endif
}
// streaming 128-bit global load that does not allocate in L1 (L1 is kept for the twiddle / multiplier tables)
template <class T> DEV Vec<T> vload_stream(const T* p);
template <> DEV Vec<float> vload_stream<float>(const float* p) {
    Vec<float> r;
#ifdef __CUDA_ARCH__
    asm("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
#else
    r = vload(p);
#endif
    return r;
}
template <> DEV Vec<double> vload_stream<double>(const double* p) {
    Vec<double> r;
#ifdef __CUDA_ARCH__
    asm("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(p));
#else
    r = vload(p);
#endif
    return r;
}
// two adjacent 16-byte chunks as one 256-bit streaming load (sm_100: ld.global.v8.f32 / v4.f64); p must be 32-byte aligned
template <class T> DEV void vload_stream2(const T* p, Vec<T>& a, Vec<T>& b);
template <> DEV void vload_stream2<float>(const float* p, Vec<float>& a, Vec<float>& b) {
#ifdef __CUDA_ARCH__
    asm("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.v[0]), "=f"(a.v[1]), "=f"(a.v[2]), "=f"(a.v[3]),
        "=f"(b.v[0]), "=f"(b.v[1]), "=f"(b.v[2]), "=f"(b.v[3]) : "l"(p));
#else
    a = vload(p); b = vload(p + 4);
#endif
}
template <> DEV void vload_stream2<double>(const double* p, Vec<double>& a, Vec<double>& b) {
#ifdef __CUDA_ARCH__
    asm("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.v[0]), "=d"(a.v[1]), "=d"(b.v[0]), "=d"(b.v[1]) : "l"(p));
#else
    a = vload(p); b = vload(p + 2);
#endif
}
// two adjacent 16-byte chunks as one 256-bit store (sm_100: st.global.v8.f32 / v4.f64); p must be 32-byte aligned
template <class T> DEV void vstore2(T* p, const Vec<T>& a, const Vec<T>& b);
template <> DEV void vstore2<float>(float* p, const Vec<float>& a, const Vec<float>& b) {
#ifdef __CUDA_ARCH__
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.v[0]), "f"(a.v[1]), "f"(a.v[2]), "f"(a.v[3]),
                 "f"(b.v[0]), "f"(b.v[1]), "f"(b.v[2]), "f"(b.v[3]) : "memory");
#else
    vstore(p, a); vstore(p + 4, b);
#endif
}
template <> DEV void vstore2<double>(double* p, const Vec<double>& a, const Vec<double>& b) {
#ifdef __CUDA_ARCH__
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a.v[0]), "d"(a.v[1]), "d"(b.v[0]), "d"(b.v[1]) : "memory");
#else
    vstore(p, a); vstore(p + 2, b);
#endif
}
// ticket with release semantics only (MEMBAR.ALL.GPU + ATOMG): unlike __threadfence() it does not invalidate L1
// start-up stagger of the co-resident persistent blocks of an SM (blocks b, b+#SM, b+2·#SM share an SM): identical blocks
// that start together stay phase-locked, so their shared-memory sweeps and their arithmetic collide instead of overlapping
DEV void stagger_start(int blk, int sms, unsigned ns) {
#ifdef __CUDA_ARCH__
    if (ns) { const unsigned k = (unsigned)(blk / sms); if (k) __nanosleep(k * ns); }
#else
    (void)blk; (void)sms; (void)ns;
#endif
}
// programmatic dependent launch: the stage kernels are launched back to back on one stream with programmatic stream
// serialization, so the next kernel's blocks are scheduled (and load their constant tables) while the previous one drains;
// pdl_wait() blocks until the previous kernel has completed and its writes are visible
DEV void pdl_wait() {
#ifdef __CUDA_ARCH__
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
DEV void pdl_launch_dependents() {
#ifdef __CUDA_ARCH__
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
// per-plane publication flag of a J[N] line inside one launch: the producer publishes with release semantics, a consumer LOOKS
// (acquire) and never waits — if the line is not there yet it computes its own copy
DEV void flag_publish(int* p, int v) {
#ifdef __CUDA_ARCH__
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
    __atomic_store_n(p, v, __ATOMIC_RELEASE);
#endif
}
// two flags behind ONE release fence (fence + relaxed stores is a release sequence; each st.release would pay its own MEMBAR)
DEV void flag_publish2(int* p, int* q, int v) {
#ifdef __CUDA_ARCH__
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(q), "r"(v) : "memory");       // q == p when there is only one flag
#else
    __atomic_store_n(p, v, __ATOMIC_RELEASE);
    __atomic_store_n(q, v, __ATOMIC_RELEASE);
#endif
}
DEV int flag_look(const int* p) {
#ifdef __CUDA_ARCH__
    int got; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(got) : "l"(p) : "memory"); return got;
#else
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#endif
}
// patient look: poll the flag for a BOUNDED time (a few microseconds more than a publisher needs after launch) before giving up.  The bound
// keeps the forward-progress guarantee — a block whose publishers are not resident stops waiting and computes its own copy — while a block
// that merely arrives a microsecond early does not pay for a private copy.
DEV int flag_look_patient(const int* p, int want, int polls) {
#ifdef __CUDA_ARCH__
    // polls are relaxed loads (served by L2, no side effect on this SM's L1); ONE acquire load — whose L1 invalidation is what makes the
    // published line visible to the plain loads that follow — once the flag has been seen or the patience is used up
    for (int i = 0; i < polls; ++i) {
        int got; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(got) : "l"(p) : "memory");
        if (got == want) break;
        __nanosleep(128);
    }
    return flag_look(p) == want;
#else
    (void)polls; return flag_look(p) == want;
#endif
}
// block-wide OR of a per-thread predicate (includes a barrier); the host emulator runs the threads of a block one after another in
// one host thread, so the value set by "thread 0" is already what every thread sees
DEV int block_or(int v) {
#ifdef __CUDA_ARCH__
    return __syncthreads_or(v);
#else
    return v;
#endif
}
DEV int block_and(int v) {
#ifdef __CUDA_ARCH__
    return __syncthreads_and(v);
#else
    return v;
#endif
}
DEV int ticket_release(int* p, int n) {
#ifdef __CUDA_ARCH__
    unsigned old;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"((unsigned)n) : "memory");
    return (int)old;
#else
    return __atomic_fetch_add(p, n, __ATOMIC_ACQ_REL);
#endif
}
template <class T> DEV Vec<T> vload_cg(const T* p) {         // 128-bit load that bypasses L1 (data published by another block)
#ifdef __CUDA_ARCH__
    Vec<T> r; const float4 q = __ldcg(reinterpret_cast<const float4*>(p)); memcpy(&r, &q, 16); return r;
#else
    return vload(p);
#endif
}
template <class T> DEV Vec<T> vload_ldg(const T* p) {        // cached read-only 128-bit load (tables)
#ifdef __CUDA_ARCH__
    Vec<T> r; const float4 q = __ldg(reinterpret_cast<const float4*>(p)); memcpy(&r, &q, 16); return r;
#else
    return vload(p);
#endif
}

template <int LOGN> struct FastSched {
    static constexpr int LOGR1 = (LOGN == 8) ? 2 : 3, LOGR2 = LOGN - 4 - LOGR1;
    static constexpr int R1 = 1 << LOGR1, R2 = 1 << LOGR2;
    static_assert(LOGN >= 8 && LOGN <= 11, "fast path covers N = 256, 512, 1024, 2048");      // 2048 = 8 · 16 · 16
};
inline bool fast_len_ok(int N) { return N == 256 || N == 512 || N == 1024 || N == 2048; }
// tile bytes and threads per block of the stage kernels for transform length N: 32 KB / 128 threads up to 1024, 64 KB / 256 threads at 2048
// (with 32 KB a column tile would be two columns wide and a row group two rows high: 32-byte runs in the row-grouped layout; the bigger tile
// keeps the runs at 128 B fp64 / 256 B fp32 and the bigger block keeps one butterfly task per thread and sweep)
constexpr int fast_tile_bytes(int N) { return N > 1024 ? 65536 : 32768; }
constexpr int fast_threads(int N) { return N > 1024 ? 256 : 128; }
// The ROW kernel's tile also fixes G, the rows per group of the row-grouped layout, and with it the run length of every access of the
// column kernel: M·G·sizeof(T) bytes.  At Nx = 1024 in fp64 a 32 KB row tile gives G = 4 and 128-byte runs — the one shape where the column
// kernel's DRAM streams are that short (fp32: 256 B; Nx = 512: 512 B) — so CMBL_ROW_TILE_KB_1024_F64 = 64 trades a row kernel at one
// 256-thread block per SM (the 2048 shape) for 256-byte runs in the column kernel.
#ifndef CMBL_ROW_TILE_KB_1024_F64
#define CMBL_ROW_TILE_KB_1024_F64 32
#endif
template <class T> constexpr int fast_row_tile_bytes(int N) { return (N == 1024 && sizeof(T) == 8) ? CMBL_ROW_TILE_KB_1024_F64 * 1024 : fast_tile_bytes(N); }
template <class T> constexpr int fast_row_threads(int N) { return fast_row_tile_bytes<T>(N) > 32768 ? 256 : 128; }

HD int swz8(int ch) { return ch ^ ((ch >> 3) & 7); }        // column kernel: chunk index within a plane
HD int swzx(int x) { return x ^ ((x >> 4) & 7); }           // row kernel: x index within a chunk-row

HD size_t rg_index(int x, int y, int Nx, int G) { return ((size_t)(y / G) * Nx + x) * G + (y % G); }

// standard (Ny,Nx) column-major plane  <->  row-grouped plane, staged through shared memory so that both sides move
// >= 512-byte runs.  One block = TX columns × TY rows of one plane.
template <class T, bool TO_RG> struct LayoutBody {
    static constexpr int NT = 256, TX = 32, TY = 64, V = 16 / (int)sizeof(T), PITCH = TY + V;
    static constexpr size_t SMEM = sizeof(T) * TX * PITCH;
    static const char* name() { return TO_RG ? "layout_to_rg" : "layout_from_rg"; }
    int Ny, Nx, G; const T* in; T* out;
    DEV void operator()(int blk, unsigned char* smem) const {
        T* t = reinterpret_cast<T*>(smem);
        const int txn = Nx / TX, tyn = Ny / TY;
        const int c = blk / (txn * tyn), r = blk % (txn * tyn), x0 = (r / tyn) * TX, y0 = (r % tyn) * TY;
        const size_t plane = (size_t)c * Ny * Nx;
        const int GV = G / V;
        CMBL_FOR_THREADS(tid, NT) {
            for (int e = tid; e < TX * TY / V; e += NT) {
                if (TO_RG) { const int x = e / (TY / V), yv = e % (TY / V); vstore(t + x * PITCH + yv * V, vload(in + plane + (size_t)(x0 + x) * Ny + y0 + yv * V)); }
                else { const int sub = e % GV, x = (e / GV) % TX, yb = e / (GV * TX);
                       vstore(t + x * PITCH + yb * G + sub * V, vload(in + plane + rg_index(x0 + x, y0 + yb * G + sub * V, Nx, G))); }
            }
        }
        CMBL_SYNC();
        CMBL_FOR_THREADS(tid, NT) {
            for (int e = tid; e < TX * TY / V; e += NT) {
                if (TO_RG) { const int sub = e % GV, x = (e / GV) % TX, yb = e / (GV * TX);
                             vstore(out + plane + rg_index(x0 + x, y0 + yb * G + sub * V, Nx, G), vload(t + x * PITCH + yb * G + sub * V)); }
                else { const int x = e / (TY / V), yv = e % (TY / V); vstore(out + plane + (size_t)(x0 + x) * Ny + y0 + yv * V, vload(t + x * PITCH + yv * V)); }
            }
        }
    }
};

// Values that a thread loads right BEFORE a barrier for use right after it (twiddles / multipliers of the next sweep) live
// in registers across the barrier on the device; the host emulator runs the threads of a block one after another, so
// there they are (re)loaded at the start of the consuming phase instead.
#ifdef CMBL_EMU
#define CMBL_PRE_END(stmt) ((void)0)
#define CMBL_PRE_START(stmt) stmt
#else
#define CMBL_PRE_END(stmt) stmt
#define CMBL_PRE_START(stmt) ((void)0)
#endif

// ---------------------------------------------------------------------------------------------------------------
// column kernel
// ---------------------------------------------------------------------------------------------------------------
#ifndef CMBL_COL_MINB
#define CMBL_COL_MINB 3
#endif
#ifndef CMBL_COL_UNR
#define CMBL_COL_UNR 2
#endif
#ifndef CMBL_COL_ABLATE
#define CMBL_COL_ABLATE 0        // experiments only (wrong results): 1 = no sweeps (landing + epilogue), 2 = no epilogue (landing + sweeps)
#endif
#ifndef CMBL_COL_UNR_ADJ
#define CMBL_COL_UNR_ADJ (sizeof(T) == 8 ? 4 : CMBL_COL_UNR)   // adjoint kernel (3 operands per unit, 2 blocks/SM): 4 units in flight in fp64
#endif                                                          // (L'*f 7.81 -> 7.71 ms; neutral in fp32; profiles/r02_col_epilogue_unroll.log)
// DMODE (forward kernel only): the epilogue also stores the two derivative maps it has in registers — ∂ₓu = tmp ± jn and ∂ᵧu — for
// the δϕ integrand of the transpose-δ flow (flow_grad.cu); a separate instantiation, so the plain kernel's register budget is untouched.
template <class T, int LOGN, bool ADJ, bool DMODE = false> struct FastColBody {
    static_assert(!(ADJ && DMODE), "derivative export exists for the forward kernel only");
    static constexpr int N = 1 << LOGN, V = 16 / (int)sizeof(T), CH = N / V;       // CH chunks per column
    static constexpr int TB = fast_tile_bytes(N), NT = fast_threads(N), MINB = (N > 1024) ? 1 : (ADJ || DMODE) ? 2 : CMBL_COL_MINB;
    static constexpr int R1 = FastSched<LOGN>::R1, R2 = FastSched<LOGN>::R2;
    static constexpr int S1 = N / R1, N2 = N / R1;                                   // pass-2 stride is 16
    static constexpr int TILE = TB / (int)sizeof(T);                                // reals per tile buffer
    static constexpr int L = TILE / (2 * N);                                         // complex lines (column pairs) per tile
    static constexpr int NB1 = S1 / V, NB2 = N / (R2 * V);                           // bundles per line in pass 1 / pass 2
    static constexpr int NBUF = ADJ ? 3 : 2;
    static constexpr size_t SMEM = (size_t)TB * NBUF;
    static constexpr bool PDL = true;
    static const char* name() { return "flow_cols"; }

    static constexpr int M = 2 * L, LGM = (M == 2 ? 1 : M == 4 ? 2 : M == 8 ? 3 : M == 16 ? 4 : M == 32 ? 5 : 6);
    static_assert((1 << LGM) == M, "columns per tile must be a power of two <= 64");

    const T* tw1; const T* tw2; const T* mult_d; const T* mult_sign; T cN;     // mult_sign, cN: the Nyquist line operator J (flow.cuh)
    const T* nline;                                  // N(y) per plane (row kernel of this stage)
    T* jn_pub; int* jn_flag; int epoch;              // launch-wide J[N] lines [plane][N]; jn_flag[plane] == epoch <=> this launch's line is published
    T* jn_blk; int jn_red, jn_polls;                 // per-block private J[N] line [block][N] (fallback); publishers per plane pair; bounded polls of a flag
    int csync = 0;                                   // launched in clusters of 2 blocks that synchronise before every epilogue
    int Nx, G, lgGV, tiles_per_plane, ntiles, nblocks, Npol, Nbphi, cbase, pf;     // G rows per row group, 2^lgGV = G / V
    int sms; unsigned stagger_ns;
    const T* u; const T* pk; const T* tmp; const T* jn; T* macc; T wgt;
    const T* ybase; const T* acc_in; T* acc_out; T* u_out; T ca, cb;
    T* dx_out = nullptr; T* dy_out = nullptr;        // DMODE: ∂ₓu, ∂ᵧu (row-grouped layout, like every other map of the launch)

    template <int R> struct Tw { Vec<T> r[R], i[R]; };

    // shared-memory chunk position: XOR swizzle within the plane plus a per-plane rotation (epilogue / copies touch
    // M planes at the same chunk index)
    DEV int pxor(int p) const { return (p << lgGV) & 7; }
    DEV int swzp(int ch, int p) const { return swz8(ch) ^ pxor(p); }
    // k-th chunk of a tile in MEMORY order (row-grouped layout): plane (column) p, chunk ch along y, element offset in the plane
    DEV void chunk_of(int k, int x0, int& p, int& ch, size_t& goff) const {
        const int sub = k & ((1 << lgGV) - 1), yb = k >> (lgGV + LGM);
        p = (k >> lgGV) & (M - 1); ch = (yb << lgGV) + sub;
        goff = ((size_t)yb * Nx + x0 + p) * G + sub * V;
    }
    // ---- global -> shared (asynchronous) ------------------------------------------------------------------
    DEV void issue_tile(const T* plane_base, int x0, T* buf, int tid) const {
#pragma unroll 4
        for (int k = tid; k < M * CH; k += NT) {
            int p, ch; size_t goff; chunk_of(k, x0, p, ch, goff);
            cp_async16(buf + p * N + swzp(ch, p) * V, plane_base + goff);
        }
    }
    template <int R> DEV void load_tw(Tw<R>& w, const T* tw, int tws, int toff) const {
#pragma unroll
        for (int q = 1; q < R; ++q) { w.r[q] = vload_ldg(tw + ((q - 1) * 2) * tws + toff); w.i[q] = vload_ldg(tw + ((q - 1) * 2 + 1) * tws + toff); }
    }
    DEV void load_tw1(Tw<R1>& w, int task) const { load_tw<R1>(w, tw1, S1, V * (task % NB1)); }
    static constexpr bool HALF2 = (R2 == 16);       // second sweep in half bundles (pass2_half): its twiddles are not kept across barriers
    DEV void load_tw2(Tw<R2>& w, int task) const { if constexpr (!HALF2) load_tw<R2>(w, tw2, 16, (V * (task % NB2)) & 15); }

    // ---- twiddled pass on V adjacent butterflies: chunk index of element m = c0 + m*cs --------------------------
    template <int R, bool INV> DEV void bundle_pass(T* re, T* im, int xr_, int xi_, int c0, int cs, const Tw<R>& w, const T* pre, const T* pim) const {
        Vec<T> xr[R], xi[R];
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int o = swz8(c0 + m * cs), oa = (o ^ xr_) * V, ob = (o ^ xi_) * V;
            xr[m] = vload(re + oa); xi[m] = vload(im + ob);
            if (pre) {
                Vec<T> a = vload(pre + oa), b = vload(pim + ob);
#pragma unroll
                for (int e = 0; e < V; ++e) { xr[m].v[e] *= a.v[e]; xi[m].v[e] *= b.v[e]; }
            }
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
            C2<T> v[R];
#pragma unroll
            for (int m = 0; m < R; ++m) v[m] = mk<T>(xr[m].v[e], xi[m].v[e]);
            if (!INV) {
                dftR<T, R, false>(v);
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmul(v[q], mk<T>(w.r[q].v[e], w.i[q].v[e]));
            } else {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], mk<T>(w.r[q].v[e], w.i[q].v[e]));
                dftR<T, R, true>(v);
            }
#pragma unroll
            for (int m = 0; m < R; ++m) { xr[m].v[e] = v[m].x; xi[m].v[e] = v[m].y; }
        }
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int o = swz8(c0 + m * cs);
            vstore(re + (o ^ xr_) * V, xr[m]); vstore(im + (o ^ xi_) * V, xi[m]);
        }
    }
    template <bool INV> DEV void pass1(T* buf, const T* pbuf, int tid, Tw<R1>& w, int nl = L) const {
        for (int task = tid; task < nl * NB1; task += NT) {
            const int l = task / NB1, b = task % NB1;
            if (task != tid) load_tw1(w, task);
            T* re = buf + (2 * l) * N; T* im = re + N;
            const T* pre = (ADJ && !INV && pbuf) ? pbuf + (2 * l) * N : nullptr;
            bundle_pass<R1, INV>(re, im, pxor(2 * l), pxor(2 * l + 1), b, S1 / V, w, pre, pre ? pre + N : nullptr);
        }
    }
    // R2 = 16 (N = 2048): a V-wide bundle would hold 2·16·V data values plus 2·15·V twiddles per thread.  Half bundles — V/2 adjacent
    // butterflies, 8-byte shared-memory accesses — stay in registers and give each of the 256 threads one task per line pair.
    struct alignas(8) HalfVec { T v[V / 2]; };
    template <bool INV> DEV void pass2_half(T* buf, int tid, int nl) const {
        constexpr int VB = V / 2, NBH = 2 * NB2, R = R2;
        for (int task = tid; task < nl * NBH; task += NT) {
            const int l = task / NBH, bh = task % NBH, b = bh >> 1, e0 = (bh & 1) * VB;
            const int j0 = V * b, jj0 = j0 & 15, a = j0 >> 4, c0 = (a * N2 + jj0) / V, cs = 16 / V;
            T* re = buf + (2 * l) * N; T* im = re + N;
            const int xr_ = pxor(2 * l), xi_ = pxor(2 * l + 1);
            HalfVec xr[R], xi[R];
#pragma unroll
            for (int m = 0; m < R; ++m) {
                const int o = swz8(c0 + m * cs);
                xr[m] = *reinterpret_cast<const HalfVec*>(re + (o ^ xr_) * V + e0); xi[m] = *reinterpret_cast<const HalfVec*>(im + (o ^ xi_) * V + e0);
            }
#pragma unroll
            for (int e = 0; e < VB; ++e) {
                C2<T> v[R];
#pragma unroll
                for (int m = 0; m < R; ++m) v[m] = mk<T>(xr[m].v[e], xi[m].v[e]);
                const T* tw = tw2 + jj0 + e0 + e;                         // ftw2[(q-1)][re|im][jj]
                if (!INV) {
                    dftR<T, R, false>(v);
#pragma unroll
                    for (int q = 1; q < R; ++q) v[q] = cmul(v[q], mk<T>(CMBL_LDG(tw + ((q - 1) * 2) * 16), CMBL_LDG(tw + ((q - 1) * 2 + 1) * 16)));
                } else {
#pragma unroll
                    for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], mk<T>(CMBL_LDG(tw + ((q - 1) * 2) * 16), CMBL_LDG(tw + ((q - 1) * 2 + 1) * 16)));
                    dftR<T, R, true>(v);
                }
#pragma unroll
                for (int m = 0; m < R; ++m) { xr[m].v[e] = v[m].x; xi[m].v[e] = v[m].y; }
            }
#pragma unroll
            for (int m = 0; m < R; ++m) {
                const int o = swz8(c0 + m * cs);
                *reinterpret_cast<HalfVec*>(re + (o ^ xr_) * V + e0) = xr[m]; *reinterpret_cast<HalfVec*>(im + (o ^ xi_) * V + e0) = xi[m];
            }
        }
    }
    template <bool INV> DEV void pass2(T* buf, int tid, Tw<R2>& w, int nl = L) const {
        if constexpr (HALF2) { pass2_half<INV>(buf, tid, nl); return; }
        for (int task = tid; task < nl * NB2; task += NT) {
            const int l = task / NB2, b = task % NB2;
            if (task != tid) load_tw2(w, task);
            const int j0 = V * b, jj0 = j0 & 15, a = j0 >> 4;
            T* re = buf + (2 * l) * N; T* im = re + N;
            bundle_pass<R2, INV>(re, im, pxor(2 * l), pxor(2 * l + 1), (a * N2 + jj0) / V, 16 / V, w, nullptr, nullptr);
        }
    }
    // fused middle: forward radix 16, multiplier iℓ/N (tile order), Nyquist bookkeeping of the adjoint flow, inverse radix 16
    DEV void middle(T* buf, int tid, T* macc_c, int x0, const T* mtab, int nl = L) const {
        constexpr int NB = N / 16, CPB = 16 / V;                    // butterflies per line, chunks per butterfly
        for (int task = tid; task < nl * NB; task += NT) {
            const int l = task / NB, j = task % NB;
            T* re = buf + (2 * l) * N; T* im = re + N;
            T mlt[16];
#pragma unroll
            for (int cc = 0; cc < CPB; ++cc) { Vec<T> m4 = vload_ldg(mtab + 16 * j + cc * V);
#pragma unroll
                for (int e = 0; e < V; ++e) mlt[cc * V + e] = m4.v[e]; }
            C2<T> v[16];
            const int xa = pxor(2 * l), xb = pxor(2 * l + 1);
#pragma unroll
            for (int cc = 0; cc < CPB; ++cc) {
                const int o = swz8(j * CPB + cc);
                Vec<T> a = vload(re + (o ^ xa) * V), b = vload(im + (o ^ xb) * V);
#pragma unroll
                for (int e = 0; e < V; ++e) v[cc * V + e] = mk<T>(a.v[e], b.v[e]);
            }
            dft16<T, false>(v);
            if (ADJ && macc_c && j == 0) { macc_c[x0 + 2 * l] += wgt * v[8].x; macc_c[x0 + 2 * l + 1] += wgt * v[8].y; }   // Nyquist coefficient: position 8
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = mk<T>(-mlt[q] * v[q].y, mlt[q] * v[q].x);
            dft16<T, true>(v);
#pragma unroll
            for (int cc = 0; cc < CPB; ++cc) {
                const int o = swz8(j * CPB + cc);
                Vec<T> a, b;
#pragma unroll
                for (int e = 0; e < V; ++e) { a.v[e] = v[cc * V + e].x; b.v[e] = v[cc * V + e].y; }
                vstore(re + (o ^ xa) * V, a); vstore(im + (o ^ xb) * V, b);
            }
        }
    }

    // ---- velocity + RK4 update (src/lenseflow.jl:150-174, src/numerical_algorithms.jl:11-24) --------------------------
    //   forward: k = p₁·(tmp ± jn) + p₂·∂ᵧu          adjoint: k = tmp ± jn + ∂ᵧ(p₂·u)         (+ for even x, − for odd x)
    //   acc_out = (acc_in ? acc_in : ybase) + cb·k ;  u_out = ybase + ca·k  (if u_out)
    // KIND 0: first stage (no acc_in), 1: middle stages, 2: last stage (no ybase, no u_out).  One 32-byte unit (two adjacent
    // chunks of one column) per thread and iteration, units taken in memory order (256-bit accesses); the global loads of
    // UNR iterations are issued before the first use.
    DEV void unit_of(int k, int x0, int& p, int& ch, size_t& goff) const {          // k-th 32-byte unit of the tile
        const int lgU = lgGV - 1, h = k & ((1 << lgU) - 1), yb = k >> (lgU + LGM);
        p = (k >> lgU) & (M - 1); ch = (yb << lgGV) + 2 * h;
        goff = ((size_t)yb * Nx + x0 + p) * G + 2 * h * V;
    }
    template <int KIND> DEV void epilogue(const T* buf, int tid, size_t pbase, int x0, const T* jc, const T* p1, const T* p2) const {
        constexpr bool YB = KIND != 2, AI = KIND != 0, UO = KIND != 2;
        constexpr int ITER = M * CH / 2 / NT, UNRW = ADJ ? CMBL_COL_UNR_ADJ : CMBL_COL_UNR, UNR = (ITER % UNRW == 0) ? UNRW : 2;
        static_assert(ITER % UNR == 0, "epilogue unroll");
        const T* tc = tmp + pbase;
        const T* yb = YB ? ybase + pbase : nullptr;
        const T* ai = AI ? acc_in + pbase : nullptr;
#pragma unroll 1
        for (int it = 0; it < ITER; it += UNR) {
            Vec<T> ta[UNR][2], p1a[UNR][2], p2a[UNR][2], ya[UNR][2], aa[UNR][2], jv[UNR][2];
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
                int p, ch; size_t g; unit_of(tid + (it + k) * NT, x0, p, ch, g);
                vload_stream2(tc + g, ta[k][0], ta[k][1]);
                if (!ADJ) { vload_stream2(p1 + g, p1a[k][0], p1a[k][1]); vload_stream2(p2 + g, p2a[k][0], p2a[k][1]); }
                if (YB) vload_stream2(yb + g, ya[k][0], ya[k][1]);
                if (AI) vload_stream2(ai + g, aa[k][0], aa[k][1]);
                jv[k][0] = vload(jc + ch * V); jv[k][1] = vload(jc + (ch + 1) * V);   // this block's own lines (L1 / L2)
            }
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
                int p, ch; size_t g; unit_of(tid + (it + k) * NT, x0, p, ch, g);
                const T sgn = (p & 1) ? (T)-1 : (T)1;                  // x0 is even: + for even x, − for odd x
                Vec<T> a0[2], u0[2], dxv[2], dyv[2];
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    const Vec<T> z = vload(buf + p * N + swzp(ch + s2, p) * V);
#pragma unroll
                    for (int q = 0; q < V; ++q) {
                        const T gxq = ta[k][s2].v[q] + sgn * jv[k][s2].v[q];
                        const T kk = ADJ ? gxq + z.v[q] : p1a[k][s2].v[q] * gxq + p2a[k][s2].v[q] * z.v[q];
                        const T y0 = YB ? ya[k][s2].v[q] : (T)0;
                        a0[s2].v[q] = (AI ? aa[k][s2].v[q] : y0) + cb * kk;
                        u0[s2].v[q] = y0 + ca * kk;
                        if (DMODE) { dxv[s2].v[q] = gxq; dyv[s2].v[q] = z.v[q]; }
                    }
                }
                vstore2(acc_out + pbase + g, a0[0], a0[1]);
                if (UO) vstore2(u_out + pbase + g, u0[0], u0[1]);
                if (DMODE) { vstore2(dx_out + pbase + g, dxv[0], dxv[1]); vstore2(dy_out + pbase + g, dyv[0], dyv[1]); }
            }
        }
    }
    // pull the epilogue operands of this tile into L2 (one bulk prefetch per contiguous run of M columns x G rows)
    DEV void prefetch_epilogue(int tid, size_t pbase, int x0, const T* p1, const T* p2) const {
        constexpr int LINE = 128 / (int)sizeof(T);                 // elements per 128-byte line
        const int lines_per_run = M * G / LINE;                   // a run = M columns x G rows, contiguous
        for (int i = tid; i < (N / G) * lines_per_run; i += NT) {
            const int yb = i / lines_per_run, ln = i % lines_per_run;
            const size_t g = ((size_t)yb * Nx + x0) * G + (size_t)ln * LINE;
            l2_prefetch_line(tmp + pbase + g);
            if (!ADJ) { l2_prefetch_line(p1 + g); l2_prefetch_line(p2 + g); }
            if (ybase) l2_prefetch_line(ybase + pbase + g);
            if (acc_in) l2_prefetch_line(acc_in + pbase + g);
        }
    }

    // jn = cN · J[N] for up to two planes at once: one more spectral operator on ONE complex line — J maps real lines to real lines,
    // so plane ca's N(y) rides in the real part and plane cb's (cb < 0: none) in the imaginary part — done in a tile buffer that is
    // not in use.  Results go to dst_a / dst_b (global memory).  Ends with a barrier.
    DEV void jn_pair(T* ws, int ca, int cb, T* dst_a, T* dst_b, Tw<R1>& w1, Tw<R2>& w2, int goff = 0, int bar = 0) const {
        CMBL_FOR_GROUP(tid, NT, goff) {
            for (int i = tid; i < 2 * CH; i += NT) {
                const int pl = i / CH, ch = i % CH, c = pl ? cb : ca;
                Vec<T> z; for (int e = 0; e < V; ++e) z.v[e] = 0;
                if (c >= 0) z = vload(nline + (size_t)c * N + ch * V);
                vstore(ws + pl * N + swzp(ch, pl) * V, z);
            }
            load_tw1(w1, tid);
        }
        group_sync(bar, NT);
        CMBL_FOR_GROUP(tid, NT, goff) { CMBL_PRE_START(load_tw1(w1, tid)); pass1<false>(ws, nullptr, tid, w1, 1); CMBL_PRE_END(load_tw2(w2, tid)); }
        group_sync(bar, NT);
        CMBL_FOR_GROUP(tid, NT, goff) { CMBL_PRE_START(load_tw2(w2, tid)); pass2<false>(ws, tid, w2, 1); }
        group_sync(bar, NT);
        CMBL_FOR_GROUP(tid, NT, goff) { middle(ws, tid, nullptr, 0, mult_sign, 1); CMBL_PRE_END(load_tw2(w2, tid)); }
        group_sync(bar, NT);
        CMBL_FOR_GROUP(tid, NT, goff) { CMBL_PRE_START(load_tw2(w2, tid)); pass2<true>(ws, tid, w2, 1); CMBL_PRE_END(load_tw1(w1, tid)); }
        group_sync(bar, NT);
        CMBL_FOR_GROUP(tid, NT, goff) { CMBL_PRE_START(load_tw1(w1, tid)); pass1<true>(ws, nullptr, tid, w1, 1); }
        group_sync(bar, NT);
        CMBL_FOR_GROUP(tid, NT, goff) {
            for (int i = tid; i < 2 * CH; i += NT) {
                const int pl = i / CH, ch = i % CH;
                T* dst = pl ? dst_b : dst_a;
                if (dst) {
                    Vec<T> z = vload(ws + pl * N + swzp(ch, pl) * V);
                    for (int e = 0; e < V; ++e) z.v[e] *= cN;
                    vstore(dst + ch * V, z);
                }
            }
        }
        group_sync(bar, NT);
    }

    // publisher blocks: compute the lines of one plane pair into jn_pub and raise the flags (the block's first tile is in flight)
    DEV void jn_publish(int blk, int nC, T* ws, Tw<R1>& w1, Tw<R2>& w2, int goff = 0, int bar = 0) const {
        const int npairs = (nC + 1) / 2;
        int b0, np;
        if (nblocks > ntiles) { b0 = ntiles; np = nblocks - ntiles; }  // small launch: the grid carries extra blocks that own no tile and only publish
        else {
            np = jn_red * npairs;
            b0 = ntiles % nblocks;                                     // round-robin: blocks b0.. own one tile less than blocks 0..b0-1
            if (b0 + np > nblocks) b0 = 0;
        }
        if (blk < b0 || blk >= b0 + np) return;
        const int pr = (blk - b0) % npairs, ca = cbase + 2 * pr, cb = (2 * pr + 1 < nC) ? ca + 1 : -1;
        jn_pair(ws, ca, cb, jn_pub + (size_t)ca * N, cb >= 0 ? jn_pub + (size_t)cb * N : nullptr, w1, w2, goff, bar);
        CMBL_FOR_GROUP(tid, NT, goff) { if (tid == 0) flag_publish2(jn_flag + ca, jn_flag + (cb >= 0 ? cb : ca), epoch); }
    }
    // RARE path of the column kernel (the publishers of plane c's J[N] line are not resident): a private copy computed from the same plane
    // pair (identical bits).  The work space is the tile buffer `ws` in which the next tile is landing: let it arrive, use the buffer, request
    // the tile again.  Deliberately NOT inlined: its registers and its code stay out of the hot loop.
    DEV_NOINLINE const T* jn_private(int c, int blk, T* ws, const T* next_src, int next_x0) const {
        Tw<R1> w1; Tw<R2> w2;
        CMBL_FOR_THREADS(tid, NT) { cp_async_wait_all(); }
        CMBL_SYNC();
        T* const mine = jn_blk + (size_t)blk * N;
        const int nC = ntiles / tiles_per_plane;
        const int pr = (c - cbase) / 2, ca = cbase + 2 * pr, cb = (2 * pr + 1 < nC) ? ca + 1 : -1;
        jn_pair(ws, ca, cb, c == ca ? mine : nullptr, c == cb ? mine : nullptr, w1, w2);
        if (next_src) { CMBL_FOR_THREADS(tid, NT) { issue_tile(next_src, next_x0, ws, tid); cp_async_commit(); } }
        return mine;
    }

    DEV void operator()(int blk, unsigned char* smem) const {
        T* const sbase = reinterpret_cast<T*>(smem);                 // plain arithmetic on the shared base keeps LDS/STS (no generic LD/ST)
        T* const pbuf = sbase + 2 * TILE;
        const size_t nmap = (size_t)N * Nx;
        const int kind = !acc_in ? 0 : (u_out ? 1 : 2);
        Tw<R1> w1; Tw<R2> w2;
        pdl_launch_dependents();
        pdl_wait();
        // Tile order: batch item, column tile, polarisation (fastest); tiles are taken round-robin t = b, b + nblocks, ... so that at any
        // moment the launch works on a window of neighbouring column tiles (their 128/256-byte runs share DRAM pages and the window's
        // pages fit the TLB reach; contiguous per-block ranges spread the blocks over the whole 0.7 GB working set and were measured
        // 25 % slower) and the Q and U (I, Q, U) tiles of the same columns are in flight together and share the p maps.
        // J[N] lines: NO block ever waits for another block.  At launch start a few blocks (jn_red per plane pair; blocks of the first
        // wave that own one tile less than the others) compute the lines of one plane pair each into the launch-wide array jn[plane][N]
        // and publish a per-plane flag — while their first tile is in flight.  Before a block's first epilogue on a plane it LOOKS at
        // the flag (thread 0 during the last sweep; the answer rides on the barrier that ends it): published (the normal case — the lines are ready a few
        // microseconds into the launch) -> it reads the shared line; not published (the producers are not resident: another stream
        // owns the SMs, MPS, ...) -> it computes a private copy of the same plane pair (identical bits) and goes on.  Correct and
        // deterministic under any residency.
        auto item_of = [&](int t) { return (t / Npol) / tiles_per_plane; };
        auto x0_of = [&](int t) { return ((t / Npol) % tiles_per_plane) * M; };
        int tile = blk, cur = 0, cj = -1;
        const T* jline = nullptr;
        if (tile < ntiles) {
            const int c = cbase + item_of(tile) * Npol + tile % Npol, x0 = x0_of(tile);
            CMBL_FOR_THREADS(tid, NT) {
                issue_tile(u + (size_t)c * nmap, x0, sbase, tid);
                if (ADJ) issue_tile(p_plane(pk, c, Npol, Nbphi, 1, nmap), x0, pbuf, tid);
                cp_async_commit();
            }
        }
        jn_publish(blk, ntiles / tiles_per_plane, sbase + TILE, w1, w2);          // (publisher blocks only) first tile in flight
        stagger_start(blk, sms, stagger_ns);                                       // (experiment knob, default 0) de-phase the blocks of an SM
        CMBL_FOR_THREADS(tid, NT) { CMBL_PRE_END(load_tw1(w1, tid)); }
        for (; tile < ntiles; tile += nblocks, cur ^= 1) {
            T* const buf = sbase + cur * TILE;
            T* const nbuf = sbase + (cur ^ 1) * TILE;
            const int c = cbase + item_of(tile) * Npol + tile % Npol, x0 = x0_of(tile), next = tile + nblocks;
            const int cn = cbase + item_of(next) * Npol + next % Npol, x0n = x0_of(next);
            const T* const p1 = p_plane(pk, c, Npol, Nbphi, 0, nmap);
            const T* const p2 = p_plane(pk, c, Npol, Nbphi, 1, nmap);
            CMBL_FOR_THREADS(tid, NT) { cp_async_wait_all(); }
            CMBL_SYNC();
            if (next < ntiles) {                                       // the next tile lands while this one is transformed (DRAM is idle then)
                CMBL_FOR_THREADS(tid, NT) { issue_tile(u + (size_t)cn * nmap, x0n, nbuf, tid); cp_async_commit(); }
            }
            CMBL_FOR_THREADS(tid, NT) { CMBL_PRE_START(load_tw1(w1, tid)); if (CMBL_COL_ABLATE != 1) pass1<false>(buf, ADJ ? pbuf : nullptr, tid, w1); CMBL_PRE_END(load_tw2(w2, tid)); }
            CMBL_SYNC();
            if (ADJ && next < ntiles) {
                CMBL_FOR_THREADS(tid, NT) { issue_tile(p_plane(pk, cn, Npol, Nbphi, 1, nmap), x0n, pbuf, tid); cp_async_commit(); }
            }
            CMBL_FOR_THREADS(tid, NT) { if (pf == 1) prefetch_epilogue(tid, (size_t)c * nmap, x0, p1, p2); CMBL_PRE_START(load_tw2(w2, tid)); if (CMBL_COL_ABLATE != 1) pass2<false>(buf, tid, w2); }
            CMBL_SYNC();
            CMBL_FOR_THREADS(tid, NT) { if (pf == 2) prefetch_epilogue(tid, (size_t)c * nmap, x0, p1, p2); if (CMBL_COL_ABLATE != 1) middle(buf, tid, ADJ ? macc + (size_t)c * Nx : nullptr, x0, mult_d); CMBL_PRE_END(load_tw2(w2, tid)); }
            CMBL_SYNC();
            CMBL_FOR_THREADS(tid, NT) { if (pf == 3) prefetch_epilogue(tid, (size_t)c * nmap, x0, p1, p2); CMBL_PRE_START(load_tw2(w2, tid)); if (CMBL_COL_ABLATE != 1) pass2<true>(buf, tid, w2); CMBL_PRE_END(load_tw1(w1, tid)); }
            CMBL_SYNC();
            int fl = 1;
            CMBL_FOR_THREADS(tid, NT) {
                if (pf == 4) prefetch_epilogue(tid, (size_t)c * nmap, x0, p1, p2); CMBL_PRE_START(load_tw1(w1, tid)); if (CMBL_COL_ABLATE != 1) pass1<true>(buf, pbuf, tid, w1);
                if (tid == 0 && c != cj) fl = flag_look_patient(jn_flag + c, epoch, jn_polls);   // is this plane's J[N] line published?  (the answer rides on the barrier)
            }
            fl = block_and(fl);
            if (c != cj) {
                jline = fl ? jn_pub + (size_t)c * N : jn_private(c, blk, nbuf, next < ntiles ? u + (size_t)cn * nmap : nullptr, x0n);
                cj = c;
            }
            // (opt-in experiment, CMBL_COL_PAIR=1) blocks 2i and 2i+1 form a cluster and hold the Q and the U tile of the same columns; they
            // start their epilogues together.  A cluster barrier cannot deadlock: the hardware co-schedules the blocks of a cluster.
            if (csync) cluster_sync();
            CMBL_FOR_THREADS(tid, NT) {
                if (CMBL_COL_ABLATE == 2) { }
                else if (kind == 0) epilogue<0>(buf, tid, (size_t)c * nmap, x0, jline, p1, p2);
                else if (kind == 1) epilogue<1>(buf, tid, (size_t)c * nmap, x0, jline, p1, p2);
                else epilogue<2>(buf, tid, (size_t)c * nmap, x0, jline, p1, p2);
                CMBL_PRE_END(load_tw1(w1, tid));                      // twiddles of the next tile's first sweep
            }
            // the next iteration's first barrier orders these shared-memory reads before the tile buffer is refilled
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// row kernel
// ---------------------------------------------------------------------------------------------------------------
#ifndef CMBL_ABLATE
#define CMBL_ABLATE 0            // experiments only: 1 = row kernel without butterfly arithmetic, 2 = without inter-sweep barriers
#endif
#if CMBL_ABLATE == 2
#define ROW_SWEEP_SYNC() ((void)0)
#else
#define ROW_SWEEP_SYNC() CMBL_SYNC()
#endif
#ifndef CMBL_ROW_F64_MINB
#define CMBL_ROW_F64_MINB 2          // fp64 row kernel: 2 blocks/SM at 255 registers (no spills) beat 3 blocks/SM at 168 (measured 128 vs 140 us)
#endif
#ifndef CMBL_ROW_PAIR
#define CMBL_ROW_PAIR 1          // 1: two chunk-rows per butterfly step in every sweep; 0: only in the last (256-bit stores)
#endif
template <class T, int LOGN, bool ADJ> struct FastRowBody {
    static constexpr int NT = 128, MINB = ADJ ? 2 : (sizeof(T) == 8 ? CMBL_ROW_F64_MINB : 3);
    static constexpr int N = 1 << LOGN, V = 16 / (int)sizeof(T), H = V / 2;          // H complex lines per chunk
    static constexpr int FAST_TILE_BYTES = 32768;
    static_assert(LOGN <= 10, "the cp.async row kernel is kept for lengths up to 1024 (CMBL_ROW_TMA=0); 2048 runs on TmaRowBody only");
    static constexpr int R1 = FastSched<LOGN>::R1, R2 = FastSched<LOGN>::R2;
    static constexpr int S1 = N / R1, N2 = N / R1, NB2 = N / R2, NBM = N / 16;
    static constexpr int CPX = FAST_TILE_BYTES / (N * 16);                            // 16-byte chunks per x (even)
    static constexpr int ROWS = CPX * V, L = ROWS / 2;                                // rows / complex lines per tile
    static constexpr int TILE = FAST_TILE_BYTES / (int)sizeof(T);
    static constexpr size_t SMEM = (size_t)FAST_TILE_BYTES * (ADJ ? 3 : 2);
    static_assert(NT % S1 == 0 && NT % NB2 == 0 && NT % NBM == 0 && CPX % 2 == 0, "thread/butterfly mapping");
    static_assert((CPX / 2) % (NT / S1) == 0 && (CPX / 2) % (NT / NB2) == 0 && CPX % (NT / NBM) == 0, "chunk-row mapping");
    static constexpr bool PDL = true;
    static const char* name() { return "flow_rows"; }

    Fft1D<T> fx; const T* mult;
    int Ny, tiles_per_plane, ntiles, nblocks, Npol, Nbphi, cbase;
    int sms; unsigned stagger_ns;
    const T* u; const T* pk; T* tmp; T* nline; T* nacc; T wgt;

    struct Chunk { C2<T> c[H]; };
    static DEV Chunk ld(const T* p) { Vec<T> v = vload(p); Chunk r; memcpy(&r, &v, 16); return r; }
    static DEV void st(T* p, const Chunk& c) { Vec<T> v; memcpy(&v, &c, 16); vstore(p, v); }
    static DEV Vec<T> asvec(const Chunk& c) { Vec<T> v; memcpy(&v, &c, 16); return v; }

    DEV void issue_tile(const T* src /*the tile: one contiguous run in the row-grouped layout*/, T* buf, int tid) const {
        if (CMBL_ABLATE == 6) return;
#pragma unroll 4
        for (int k = tid; k < N * CPX; k += NT) {
            const int x = k / CPX, cl = k % CPX;
            cp_async16(buf + (cl * N + swzx(x)) * V, src + (size_t)k * V);
        }
    }
    DEV void load_w1(int tid, C2<T>* w) const {
        const int j = tid % S1;
#pragma unroll
        for (int q = 1; q < R1; ++q) w[q] = CMBL_LDG(&fx.W[j * q]);
    }
    DEV void load_w2(int tid, C2<T>* w) const {
        const int jj = (tid % NB2) & 15;
#pragma unroll
        for (int q = 1; q < R2; ++q) w[q] = CMBL_LDG(&fx.W[jj * q * R1]);
    }
    DEV void load_mult(int tid, T* m) const {
        const int j = tid % NBM;
#pragma unroll
        for (int cc = 0; cc < 16 / V; ++cc) { Vec<T> m4 = vload_ldg(mult + 16 * j + cc * V);
#pragma unroll
            for (int e = 0; e < V; ++e) m[cc * V + e] = m4.v[e]; }
    }
    // one twiddled radix-R sweep of one thread: butterfly index fixed (element m at x0 + m*xs), chunk-rows taken in adjacent
    // pairs (cl, cl+1) so that the two butterflies overlap and, for the last sweep, leave as one 256-bit store
    template <int R, bool INV, bool STORE_GLOBAL> DEV void pass(T* buf, const T* pbuf, int g0, int gs, int x0, int xs, const C2<T>* w, T* gdst) const {
        constexpr int PW = (STORE_GLOBAL || CMBL_ROW_PAIR) ? 2 : 1;      // chunk-rows handled together
        for (int g = g0; g < CPX / 2; g += gs) {
#pragma unroll
            for (int s0 = 0; s0 < 2; s0 += PW) {
                Chunk x[PW][R];
#pragma unroll
                for (int m = 0; m < R; ++m) {
#pragma unroll
                    for (int s = 0; s < PW; ++s) {
                        const int o = ((2 * g + s0 + s) * N + swzx(x0 + m * xs)) * V;
                        x[s][m] = ld(buf + o);
                        if (pbuf) {
                            Vec<T> p = vload(pbuf + o); Vec<T> a = asvec(x[s][m]);
#pragma unroll
                            for (int e = 0; e < V; ++e) a.v[e] *= p.v[e];
                            memcpy(&x[s][m], &a, 16);
                        }
                    }
                }
#pragma unroll
                for (int s = 0; s < PW; ++s) {
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        C2<T> v[R];
#pragma unroll
                        for (int m = 0; m < R; ++m) v[m] = x[s][m].c[h];
#if CMBL_ABLATE != 1
                        if (!INV) {
                            dftR<T, R, false>(v);
#pragma unroll
                            for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
                        } else {
#pragma unroll
                            for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], w[q]);
                            dftR<T, R, true>(v);
                        }
#else
                        v[0].x += w[1].x;          // ablation: data movement only
#endif
#pragma unroll
                        for (int m = 0; m < R; ++m) x[s][m].c[h] = v[m];
                    }
                }
#pragma unroll
                for (int m = 0; m < R; ++m) {
                    if (STORE_GLOBAL && CMBL_ABLATE == 5) { if (x[0][m].c[0].x == (T)1.2345e-30) vstore2(gdst + ((size_t)(x0 + m * xs) * CPX + 2 * g) * V, asvec(x[0][m]), asvec(x[PW - 1][m])); }
                    else if (STORE_GLOBAL) vstore2(gdst + ((size_t)(x0 + m * xs) * CPX + 2 * g) * V, asvec(x[0][m]), asvec(x[PW - 1][m]));
                    else {
#pragma unroll
                        for (int s = 0; s < PW; ++s) st(buf + ((2 * g + s0 + s) * N + swzx(x0 + m * xs)) * V, x[s][m]);
                    }
                }
            }
        }
    }
    // fused middle: forward radix 16 · (Nyquist line N(y), src/proj_lambert.jl:63-64) · multiplier iℓx/N · inverse radix 16
    static constexpr bool PRELOAD_MULT = true;
    DEV void middle(T* buf, int tid, const T* mlt_pre, T* nline_c, T* nacc_c, int y0) const {
        const int j = tid % NBM;
        for (int cl = tid / NBM; cl < CPX; cl += NT / NBM) {
            Chunk x[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) x[q] = ld(buf + (cl * N + swzx(16 * j + q)) * V);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                C2<T> v[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = x[q].c[h];
#if CMBL_ABLATE != 1
                dft16<T, false>(v);
#endif
                if (j == 0) {                                          // Nyquist coefficient sits at tile position 8
                    const int y = y0 + 2 * (cl * H + h);
                    nline_c[y] = v[8].x; nline_c[y + 1] = v[8].y;
                    if (ADJ) { nacc_c[y] += wgt * v[8].x; nacc_c[y + 1] += wgt * v[8].y; }
                }
                if (PRELOAD_MULT) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = mk<T>(-mlt_pre[q] * v[q].y, mlt_pre[q] * v[q].x);
                } else {
                    T mlt[16];
                    load_mult(tid, mlt);
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = mk<T>(-mlt[q] * v[q].y, mlt[q] * v[q].x);
                }
#if CMBL_ABLATE != 1
                dft16<T, true>(v);
#endif
#pragma unroll
                for (int q = 0; q < 16; ++q) x[q].c[h] = v[q];
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) st(buf + (cl * N + swzx(16 * j + q)) * V, x[q]);
        }
    }

    DEV void operator()(int blk, unsigned char* smem) const {
        T* const sbase = reinterpret_cast<T*>(smem);
        T* const pbuf = sbase + 2 * TILE;
        const size_t nmap = (size_t)N * Ny;
        // the thread's butterfly indices never change: its twiddles and multipliers stay in registers for the whole launch
        C2<T> w1[R1], w2[R2]; T mlt[16];
        pdl_launch_dependents();
        CMBL_FOR_THREADS(tid, NT) { CMBL_PRE_END(load_w1(tid, w1)); CMBL_PRE_END(load_w2(tid, w2)); CMBL_PRE_END(load_mult(tid, mlt)); }
        pdl_wait();
        // block b owns the contiguous tile range [tbeg, tend): at most a few planes per block, so the per-plane completion
        // ticket (a release atomic that has to wait for the block's stores) is paid once per plane, not once per tile
        const int tbeg = (int)((long long)blk * ntiles / nblocks), tend = (int)((long long)(blk + 1) * ntiles / nblocks);
        int tile = tbeg, cur = 0;
        if (tile < tend) {
            const int c = cbase + tile / tiles_per_plane;
            const size_t toff = (size_t)(tile % tiles_per_plane) * ROWS * N;
            CMBL_FOR_THREADS(tid, NT) {
                issue_tile(u + (size_t)c * nmap + toff, sbase, tid);
                if (ADJ) issue_tile(p_plane(pk, c, Npol, Nbphi, 0, nmap) + toff, pbuf, tid);
                cp_async_commit();
            }
        }
        for (; tile < tend; ++tile, cur ^= 1) {
            T* const buf = sbase + cur * TILE;
            T* const nbuf = sbase + (cur ^ 1) * TILE;
            const int c = cbase + tile / tiles_per_plane, y0 = (tile % tiles_per_plane) * ROWS, next = tile + 1;
            const int cn = cbase + next / tiles_per_plane;
            const size_t toff = (size_t)(tile % tiles_per_plane) * ROWS * N, toffn = (size_t)(next % tiles_per_plane) * ROWS * N;
            CMBL_FOR_THREADS(tid, NT) { cp_async_wait_all(); }
            CMBL_SYNC();
            if (next < tend) {
                CMBL_FOR_THREADS(tid, NT) { issue_tile(u + (size_t)cn * nmap + toffn, nbuf, tid); cp_async_commit(); }
            }
            CMBL_FOR_THREADS(tid, NT) {
                CMBL_PRE_START(load_w1(tid, w1));
                pass<R1, false, false>(buf, ADJ ? pbuf : nullptr, tid / S1, NT / S1, tid % S1, S1, w1, nullptr);
            }
            ROW_SWEEP_SYNC();
            if (ADJ && next < tend) {
                CMBL_FOR_THREADS(tid, NT) { issue_tile(p_plane(pk, cn, Npol, Nbphi, 0, nmap) + toffn, pbuf, tid); cp_async_commit(); }
            }
#if CMBL_ABLATE < 3 || CMBL_ABLATE == 7
            CMBL_FOR_THREADS(tid, NT) {
                CMBL_PRE_START(load_w2(tid, w2));
                const int j = tid % NB2;
                pass<R2, false, false>(buf, nullptr, tid / NB2, NT / NB2, (j >> 4) * N2 + (j & 15), 16, w2, nullptr);
            }
            ROW_SWEEP_SYNC();
            CMBL_FOR_THREADS(tid, NT) {
                CMBL_PRE_START(load_mult(tid, mlt));
                middle(buf, tid, mlt, nline + (size_t)c * Ny, ADJ ? nacc + (size_t)c * Ny : nullptr, y0);
            }
            ROW_SWEEP_SYNC();
            CMBL_FOR_THREADS(tid, NT) {
                CMBL_PRE_START(load_w2(tid, w2));
                const int j = tid % NB2;
                pass<R2, true, false>(buf, nullptr, tid / NB2, NT / NB2, (j >> 4) * N2 + (j & 15), 16, w2, nullptr);
            }
            ROW_SWEEP_SYNC();
#endif
            CMBL_FOR_THREADS(tid, NT) {
                CMBL_PRE_START(load_w1(tid, w1));
                pass<R1, true, true>(buf, nullptr, tid / S1, NT / S1, tid % S1, S1, w1, tmp + (size_t)c * nmap + toff);
            }
        }
    }
};


// ---------------------------------------------------------------------------------------------------------------
// bulk asynchronous copies (the TMA engine; no tensor map: one contiguous run) with mbarrier completion.  They move the row
// kernel's tiles without occupying the LSU / L1TEX data path, which the five FFT sweeps need (measured: scattered
// cp.async landing and 256-bit STG cost as much L1TEX time as three sweeps).  Plain memcpy in the host emulator.
// ---------------------------------------------------------------------------------------------------------------
DEV void mbar_init(uint64_t* b, int count) {
#ifdef __CUDA_ARCH__
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count) : "memory");
#else
    (void)b; (void)count;
#endif
}
DEV void mbar_init_fence() {
#ifdef __CUDA_ARCH__
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
DEV void mbar_expect_tx(uint64_t* b, unsigned bytes) {
#ifdef __CUDA_ARCH__
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
#else
    (void)b; (void)bytes;
#endif
}
DEV void mbar_wait(uint64_t* b, unsigned parity) {
#ifdef __CUDA_ARCH__
    asm volatile("{\n.reg .pred p;\nMBW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra MBD_%=;\nbra MBW_%=;\nMBD_%=:\n}"
                 ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(parity) : "memory");
#else
    (void)b; (void)parity;
#endif
}
DEV void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* b) {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
#else
    (void)b; memcpy(smem_dst, gsrc, bytes);
#endif
}
DEV void bulk_store(void* gdst, const void* smem_src, unsigned bytes) {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#else
    memcpy(gdst, smem_src, bytes);
#endif
}
DEV void bulk_wait_read_all() {                 // all bulk stores of this thread have finished READING shared memory
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}
DEV void bulk_wait_all() {                      // ... and have completed
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
}
DEV void fence_proxy_async() {                  // make this thread's shared-memory writes visible to the bulk-copy engine
#ifdef __CUDA_ARCH__
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// row kernel, bulk-copy version: the tile arrives and leaves as ONE contiguous 32 KB bulk copy each way (row-grouped
// layout = chunk order [x][cl]); the first sweep reads that linear order and writes the swizzled [cl][x] work layout in
// place, the last sweep does the reverse (a block holds a whole tile in registers during a sweep, so "read all — barrier —
// write all" makes the in-place permutation safe).  Two tile buffers: while tile t is transformed in one, tile t+1 lands
// in the other, which is also where tile t-1's result is read from by its outgoing copy.
// ---------------------------------------------------------------------------------------------------------------
template <class T, int LOGN, bool ADJ> struct TmaRowBody {
    static constexpr int N = 1 << LOGN, V = 16 / (int)sizeof(T), H = V / 2;
    static constexpr int FAST_TILE_BYTES = fast_row_tile_bytes<T>(N), NT = fast_row_threads<T>(N);
    static constexpr int MINB = (FAST_TILE_BYTES > 32768) ? 1 : ADJ ? 2 : (sizeof(T) == 8 ? CMBL_ROW_F64_MINB : 3);
    static constexpr int R1 = FastSched<LOGN>::R1, R2 = FastSched<LOGN>::R2;
    static constexpr int S1 = N / R1, N2 = N / R1, NB2 = N / R2, NBM = N / 16;
    static constexpr int CPX = FAST_TILE_BYTES / (N * 16);
    static constexpr int ROWS = CPX * V;
    static constexpr int TILE = FAST_TILE_BYTES / (int)sizeof(T);
    // chunk rows a thread takes together in a sweep (PW: 2 = adjacent pairs, always in the first / last sweep whose global side moves 256 bits;
    // 1 where the pairs would not give every thread a task: the radix-16 second sweep at N = 2048) and tasks per thread and sweep
    static constexpr int PW2 = ((CPX / 2) % (NT / NB2) == 0) ? 2 : 1;
    static constexpr int NP1 = (CPX / 2) / (NT / S1), NP2 = (CPX / PW2) / (NT / NB2);
    static_assert(NT % S1 == 0 && NT % NB2 == 0 && NT % NBM == 0 && CPX % 2 == 0 && NP1 >= 1 && NP2 >= 1, "thread/butterfly mapping");
    static_assert((CPX / 2) % (NT / S1) == 0 && (CPX / PW2) % (NT / NB2) == 0 && CPX % (NT / NBM) == 0, "chunk-row mapping");
#ifdef CMBL_EMU
    static constexpr size_t SMEM = (size_t)FAST_TILE_BYTES * (ADJ ? 3 : 2) + 64 + FAST_TILE_BYTES;   // + snapshot for the serial emulation
#else
    static constexpr size_t SMEM = (size_t)FAST_TILE_BYTES * (ADJ ? 3 : 2) + 64;
#endif
    static constexpr bool PDL = true;
    static const char* name() { return "flow_rows"; }

    Fft1D<T> fx; const T* mult;
    int Ny, tiles_per_plane, ntiles, nblocks, Npol, Nbphi, cbase;
    const T* u; const T* pk; T* tmp; T* nline; T* nacc; T wgt;

    struct Chunk { C2<T> c[H]; };
    static DEV Chunk ld(const T* p) { Vec<T> v = vload(p); Chunk r; memcpy(&r, &v, 16); return r; }
    static DEV void st(T* p, const Chunk& c) { Vec<T> v; memcpy(&v, &c, 16); vstore(p, v); }
    template <bool LIN> static DEV int off(int cl, int x) { return LIN ? (x * CPX + cl) * V : (cl * N + swzx(x)) * V; }

    DEV void load_w1(int tid, C2<T>* w) const {
        const int j = tid % S1;
#pragma unroll
        for (int q = 1; q < R1; ++q) w[q] = CMBL_LDG(&fx.W[j * q]);
    }
    DEV void load_w2(int tid, C2<T>* w) const {
        const int jj = (tid % NB2) & 15;
#pragma unroll
        for (int q = 1; q < R2; ++q) w[q] = CMBL_LDG(&fx.W[jj * q * R1]);
    }
    DEV void load_mult(int tid, T* m) const {
        const int j = tid % NBM;
#pragma unroll
        for (int cc = 0; cc < 16 / V; ++cc) { Vec<T> m4 = vload_ldg(mult + 16 * j + cc * V);
#pragma unroll
            for (int e = 0; e < V; ++e) m[cc * V + e] = m4.v[e]; }
    }
    // the three parts of one twiddled radix-R sweep of one thread (butterfly index fixed: element m at x0 + m*xs; chunk-row
    // pairs g0, g0+gs, ...)
    // `par` XOR-permutes which chunk row a thread keeps in which slot.  The rows are independent lines, so any consistent choice
    // is valid; in the sweeps that touch the LINEAR landing order (chunk (x, row) at (x·CPX + row)·16 B: a stride of CPX·16 bytes
    // across the threads of a quarter-warp, i.e. a CPX-way bank conflict for a fixed row) letting the row depend on the lane makes
    // the eight 128-bit accesses of a quarter-warp hit eight different 16-byte bank groups.  (The work layout is unaffected: a row
    // is a multiple of 4 KB there.)
    static DEV int lane_par(int tid) { return CPX == 2 ? ((tid >> 2) & 1) : CPX == 4 ? ((tid >> 1) & 3) : (tid & 7); }
    template <int R, int NP, bool LIN, int PW = 2> DEV void sw_load(const T* buf, const T* pbuf, int g0, int gs, int x0, int xs, Chunk (&x)[NP][PW][R], int par = 0) const {
#pragma unroll
        for (int i = 0; i < NP; ++i)
#pragma unroll
            for (int m = 0; m < R; ++m)
#pragma unroll
                for (int s = 0; s < PW; ++s) {
                    const int o = off<LIN>((PW * (g0 + i * gs) + s) ^ par, x0 + m * xs);
                    x[i][s][m] = ld(buf + o);
                    if (pbuf) {
                        Vec<T> p = vload(pbuf + o), a; memcpy(&a, &x[i][s][m], 16);
#pragma unroll
                        for (int e = 0; e < V; ++e) a.v[e] *= p.v[e];
                        memcpy(&x[i][s][m], &a, 16);
                    }
                }
    }
    template <int R, int NP, bool INV, int PW = 2> DEV void sw_compute(Chunk (&x)[NP][PW][R], const C2<T>* w) const {
#pragma unroll
        for (int i = 0; i < NP; ++i)
#pragma unroll
            for (int s = 0; s < PW; ++s)
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    C2<T> v[R];
#pragma unroll
                    for (int m = 0; m < R; ++m) v[m] = x[i][s][m].c[h];
                    if (!INV) {
                        dftR<T, R, false>(v);
#pragma unroll
                        for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
                    } else {
#pragma unroll
                        for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], w[q]);
                        dftR<T, R, true>(v);
                    }
#pragma unroll
                    for (int m = 0; m < R; ++m) x[i][s][m].c[h] = v[m];
                }
    }
    template <int R, int NP, bool LIN, int PW = 2> DEV void sw_store(T* buf, int g0, int gs, int x0, int xs, const Chunk (&x)[NP][PW][R], int par = 0) const {
#pragma unroll
        for (int i = 0; i < NP; ++i)
#pragma unroll
            for (int m = 0; m < R; ++m)
#pragma unroll
                for (int s = 0; s < PW; ++s) st(buf + off<LIN>((PW * (g0 + i * gs) + s) ^ par, x0 + m * xs), x[i][s][m]);
    }
    DEV void middle(T* buf, int tid, const T* mlt, T* nline_c, T* nacc_c, int y0) const {
        const int j = tid % NBM;
        for (int cl = tid / NBM; cl < CPX; cl += NT / NBM) {
            Chunk x[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) x[q] = ld(buf + off<false>(cl, 16 * j + q));
#pragma unroll
            for (int h = 0; h < H; ++h) {
                C2<T> v[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = x[q].c[h];
                dft16<T, false>(v);
                if (j == 0) {                                          // Nyquist coefficient sits at tile position 8
                    const int y = y0 + 2 * (cl * H + h);
                    nline_c[y] = v[8].x; nline_c[y + 1] = v[8].y;
                    if (ADJ) { nacc_c[y] += wgt * v[8].x; nacc_c[y + 1] += wgt * v[8].y; }
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = mk<T>(-mlt[q] * v[q].y, mlt[q] * v[q].x);
                dft16<T, true>(v);
#pragma unroll
                for (int q = 0; q < 16; ++q) x[q].c[h] = v[q];
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) st(buf + off<false>(cl, 16 * j + q), x[q]);
        }
    }

    DEV void operator()(int blk, unsigned char* smem) const {
        T* const sbase = reinterpret_cast<T*>(smem);
        T* const pbuf = sbase + 2 * TILE;
        uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + (size_t)FAST_TILE_BYTES * (ADJ ? 3 : 2));   // [0],[1]: tile buffers, [2]: p tile
#ifdef CMBL_EMU
        T* const snap = reinterpret_cast<T*>(smem + (size_t)FAST_TILE_BYTES * (ADJ ? 3 : 2) + 64);
#endif
        const size_t nmap = (size_t)N * Ny;
        constexpr unsigned TB = FAST_TILE_BYTES;
        C2<T> w1[R1], w2[R2]; T mlt[16];
        pdl_launch_dependents();
        CMBL_FOR_THREADS(tid, NT) {
            CMBL_PRE_END(load_w1(tid, w1)); CMBL_PRE_END(load_w2(tid, w2)); CMBL_PRE_END(load_mult(tid, mlt));
            if (tid == 0) { mbar_init(bars + 0, 1); mbar_init(bars + 1, 1); mbar_init(bars + 2, 1); mbar_init_fence(); }
        }
        CMBL_SYNC();
        pdl_wait();
        const int tbeg = (int)((long long)blk * ntiles / nblocks), tend = (int)((long long)(blk + 1) * ntiles / nblocks);
        auto tile_ptr = [&](const T* base, int t) { return base + (size_t)(cbase + t / tiles_per_plane) * nmap + (size_t)(t % tiles_per_plane) * ROWS * N; };
        auto p_ptr = [&](int t) { return p_plane(pk, cbase + t / tiles_per_plane, Npol, Nbphi, 0, nmap) + (size_t)(t % tiles_per_plane) * ROWS * N; };
        if (tbeg < tend) {
            CMBL_FOR_THREADS(tid, NT) {
                if (tid == 0) {
                    mbar_expect_tx(bars + 0, TB); bulk_load(sbase, tile_ptr(u, tbeg), TB, bars + 0);
                    if (ADJ) { mbar_expect_tx(bars + 2, TB); bulk_load(pbuf, p_ptr(tbeg), TB, bars + 2); }
                }
            }
        }
        int it = 0;
        for (int tile = tbeg; tile < tend; ++tile, ++it) {
            const int cur = it & 1;
            T* const buf = sbase + cur * TILE;
            T* const nbuf = sbase + (cur ^ 1) * TILE;
            const int c = cbase + tile / tiles_per_plane, y0 = (tile % tiles_per_plane) * ROWS, next = tile + 1;
            // ---- first sweep: linear landing order -> swizzled work layout, in place ------------------------------------
#ifdef CMBL_EMU
            memcpy(snap, buf, TB);
            CMBL_FOR_THREADS(tid, NT) {
                load_w1(tid, w1);
                Chunk x[NP1][2][R1];
                sw_load<R1, NP1, true>(snap, ADJ ? pbuf : nullptr, tid / S1, NT / S1, tid % S1, S1, x, lane_par(tid));
                sw_compute<R1, NP1, false>(x, w1);
                sw_store<R1, NP1, false>(buf, tid / S1, NT / S1, tid % S1, S1, x, lane_par(tid));
            }
#else
            {
                mbar_wait(bars + cur, (unsigned)((it >> 1) & 1));
                if (ADJ) mbar_wait(bars + 2, (unsigned)(it & 1));
                Chunk x[NP1][2][R1];
                sw_load<R1, NP1, true>(buf, ADJ ? pbuf : nullptr, threadIdx.x / S1, NT / S1, threadIdx.x % S1, S1, x, lane_par(threadIdx.x));
                __syncthreads();
                sw_compute<R1, NP1, false>(x, w1);
                sw_store<R1, NP1, false>(buf, threadIdx.x / S1, NT / S1, threadIdx.x % S1, S1, x, lane_par(threadIdx.x));
            }
#endif
            CMBL_SYNC();
            if (next < tend) {                                         // the other buffer: its outgoing copy (tile-1) must have read it
                CMBL_FOR_THREADS(tid, NT) {
                    if (tid == 0) {
                        bulk_wait_read_all();
                        mbar_expect_tx(bars + (cur ^ 1), TB); bulk_load(nbuf, tile_ptr(u, next), TB, bars + (cur ^ 1));
                        if (ADJ) { mbar_expect_tx(bars + 2, TB); bulk_load(pbuf, p_ptr(next), TB, bars + 2); }
                    }
                }
            }
            CMBL_FOR_THREADS(tid, NT) {
                CMBL_PRE_START(load_w2(tid, w2));
                const int j = tid % NB2;
                Chunk x[NP2][PW2][R2];
                sw_load<R2, NP2, false, PW2>(buf, nullptr, tid / NB2, NT / NB2, (j >> 4) * N2 + (j & 15), 16, x);
                sw_compute<R2, NP2, false, PW2>(x, w2);
                sw_store<R2, NP2, false, PW2>(buf, tid / NB2, NT / NB2, (j >> 4) * N2 + (j & 15), 16, x);
            }
            CMBL_SYNC();
            CMBL_FOR_THREADS(tid, NT) {
                CMBL_PRE_START(load_mult(tid, mlt));
                middle(buf, tid, mlt, nline + (size_t)c * Ny, ADJ ? nacc + (size_t)c * Ny : nullptr, y0);
            }
            CMBL_SYNC();
            CMBL_FOR_THREADS(tid, NT) {
                CMBL_PRE_START(load_w2(tid, w2));
                const int j = tid % NB2;
                Chunk x[NP2][PW2][R2];
                sw_load<R2, NP2, false, PW2>(buf, nullptr, tid / NB2, NT / NB2, (j >> 4) * N2 + (j & 15), 16, x);
                sw_compute<R2, NP2, true, PW2>(x, w2);
                sw_store<R2, NP2, false, PW2>(buf, tid / NB2, NT / NB2, (j >> 4) * N2 + (j & 15), 16, x);
            }
            CMBL_SYNC();
            // ---- last sweep: swizzled work layout -> linear order, in place; then one outgoing bulk copy ----------------
#ifdef CMBL_EMU
            memcpy(snap, buf, TB);
            CMBL_FOR_THREADS(tid, NT) {
                load_w1(tid, w1);
                Chunk x[NP1][2][R1];
                sw_load<R1, NP1, false>(snap, nullptr, tid / S1, NT / S1, tid % S1, S1, x, lane_par(tid));
                sw_compute<R1, NP1, true>(x, w1);
                sw_store<R1, NP1, true>(buf, tid / S1, NT / S1, tid % S1, S1, x, lane_par(tid));
            }
#else
            {
                Chunk x[NP1][2][R1];
                sw_load<R1, NP1, false>(buf, nullptr, threadIdx.x / S1, NT / S1, threadIdx.x % S1, S1, x, lane_par(threadIdx.x));
                __syncthreads();
                sw_compute<R1, NP1, true>(x, w1);
                sw_store<R1, NP1, true>(buf, threadIdx.x / S1, NT / S1, threadIdx.x % S1, S1, x, lane_par(threadIdx.x));
                fence_proxy_async();
            }
#endif
            CMBL_SYNC();
            CMBL_FOR_THREADS(tid, NT) { if (tid == 0) bulk_store(const_cast<T*>(tile_ptr(tmp, tile)), buf, TB); }
        }
        CMBL_FOR_THREADS(tid, NT) { if (tid == 0) bulk_wait_all(); }
    }
};

}  // namespace cmbl
