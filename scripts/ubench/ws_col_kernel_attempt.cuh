// RECORD OF A MEASURED, REJECTED EXPERIMENT (round 2) — not compiled into the library.
// Warp-specialised forward column kernel: FFT group + epilogue group per block, three tile buffers, named-barrier hand-over, register
// file split with setmaxnreg (176 / 80).  Bit-identical results (all GPU parity tests passed with it), but 282 us per launch at
// Nside=1024 QU batch 8 fp64 against 178-188 us for FastColBody: with 80 registers an epilogue thread can keep only ONE 160-byte unit
// (5 operands x 32 B) in flight, i.e. 20 KB per block / 40 KB per SM, and the epilogue group becomes latency-bound at ~3 TB/s
// (profiles/r02_ws_col_kernel.log).  The software-pipelined single-role variant that preceded it (epilogue slices of tile i-1 carried
// in registers across the sweeps of tile i, 2 blocks/SM) spilled 1.3-2.9 KB per thread: the twiddled radix-8 sweeps alone take every
// register ptxas is given.  What the experiment says: overlapping the epilogue's DRAM traffic with the sweeps needs the operands
// staged in SHARED memory by the bulk-copy engine (no register cost per byte in flight), which needs the shared memory of a third
// resident block.
// Drop-in for flow_fast.cuh (derives from FastColBody<T, LOGN, false>); launched with persistent_blocks / fast_cols_launch.
// ---------------------------------------------------------------------------------------------------------------
// column kernel, warp-specialised (forward flow).  FastColBody spends half of its time line in the RK epilogue waiting for p₁, p₂,
// tmp, y, acc: 87 % of the kernel's DRAM bytes are requested in a phase during which nothing is transformed, and nothing is requested
// while the five sweeps run.  Here a block has two roles that run CONCURRENTLY on different tiles:
//     FFT group  (threads   0..127): lands tile i+1 (cp.async), runs the five sweeps on tile i            — shared-memory / FP64 pipes
//     EPI group  (threads 128..255): velocity + RK4 update of tile i−1 (256-bit global loads and stores)   — the memory system
// Three tile buffers rotate between "landing", "being transformed" and "being consumed"; the hand-over is two named barriers per
// buffer (full: FFT arrives / EPI waits; empty: EPI arrives / FFT waits) — no block-wide __syncthreads in the steady state.  The
// register file is split with setmaxnreg: the sweeps need ~170 registers, the epilogue ~80, and (176 + 80)·128 threads·2 blocks is
// exactly the 64 K registers of an SM.  112 KB of shared memory per block (3 tiles + the work space of the J[N] fallback): two
// blocks per SM.  Same arithmetic, element by element, as FastColBody.
// ---------------------------------------------------------------------------------------------------------------
DEV void reg_inc_fft() {
#ifdef __CUDA_ARCH__
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;" ::: "memory");
#endif
}
DEV void reg_dec_epi() {
#ifdef __CUDA_ARCH__
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;" ::: "memory");
#endif
}
DEV void fence_block() {
#ifdef __CUDA_ARCH__
    __threadfence_block();
#endif
}

template <class T, int LOGN> struct WsColBody : FastColBody<T, LOGN, false> {
    typedef FastColBody<T, LOGN, false> Base;
    static constexpr int NG = 128, NT = 256, MINB = 2;
    static constexpr int N = Base::N, V = Base::V, M = Base::M, CH = Base::CH, TILE = Base::TILE, R1 = Base::R1, R2 = Base::R2;
    static constexpr size_t SMEM = (size_t)FAST_TILE_BYTES * 3 + sizeof(T) * 2 * N;
    static constexpr bool PDL = true;
    static const char* name() { return "flow_cols"; }
    static constexpr int ITER = M * CH / 2 / NG;                       // 32-byte units per EPI thread and tile
    enum { BAR_FFT = 1, BAR_EPI = 2, BAR_FULL = 3, BAR_EMPTY = 6 };    // named barriers (FULL / EMPTY: + buffer slot 0..2)
    template <int R> using Tw = typename Base::template Tw<R>;

    struct Ctx {                                                       // per-block constants of one launch
        T* sbase; T* jws; T* jmine; size_t nmap; int tile0, tstep, ntl, nC;
    };
    DEV int tile_of(const Ctx& x, int i) const { return x.tile0 + i * x.tstep; }
    DEV int plane_of(int t) const { return this->cbase + ((t / this->Npol) / this->tiles_per_plane) * this->Npol + t % this->Npol; }
    DEV int x0_of(int t) const { return ((t / this->Npol) % this->tiles_per_plane) * M; }

    // ---- FFT group: tile i --------------------------------------------------------------------------------------------------
    DEV void fft_step(const Ctx& x, int i, Tw<R1>& w1, Tw<R2>& w2) const {
        const int tile = tile_of(x, i), slot = i % 3, nslot = (i + 1) % 3;
        T* const buf = x.sbase + slot * TILE;
        CMBL_FOR_GROUP(tid, NG, 0) { cp_async_wait_all(); }
        group_sync(BAR_FFT, NG);
        if (i + 1 < x.ntl) {
            if (i + 1 >= 3) group_sync(BAR_EMPTY + nslot, NT);           // the epilogue of tile i−2 has let go of that buffer
            const int next = tile_of(x, i + 1);
            CMBL_FOR_GROUP(tid, NG, 0) { this->issue_tile(this->u + (size_t)plane_of(next) * x.nmap, x0_of(next), x.sbase + nslot * TILE, tid); cp_async_commit(); }
        }
        CMBL_FOR_GROUP(tid, NG, 0) { this->load_tw1(w1, tid); this->template pass1<false>(buf, nullptr, tid, w1); }
        group_sync(BAR_FFT, NG);
        CMBL_FOR_GROUP(tid, NG, 0) { this->load_tw2(w2, tid); this->template pass2<false>(buf, tid, w2); }
        group_sync(BAR_FFT, NG);
        CMBL_FOR_GROUP(tid, NG, 0) { this->middle(buf, tid, nullptr, x0_of(tile), this->mult_d); }
        group_sync(BAR_FFT, NG);
        CMBL_FOR_GROUP(tid, NG, 0) { this->load_tw2(w2, tid); this->template pass2<true>(buf, tid, w2); }
        group_sync(BAR_FFT, NG);
        CMBL_FOR_GROUP(tid, NG, 0) { this->load_tw1(w1, tid); this->template pass1<true>(buf, nullptr, tid, w1); }
        fence_block();
        group_arrive(BAR_FULL + slot, NT);                               // ∂ᵧu of tile i is in the buffer
    }

    // ---- EPI group: velocity + RK4 update of one tile (src/lenseflow.jl:150-161, src/numerical_algorithms.jl:11-24) ----------
    template <int KIND> DEV void epi_tile(const T* buf, int tid, size_t pbase, int x0, const T* jc, const T* p1, const T* p2) const {
        constexpr bool YB = KIND != 2, AI = KIND != 0, UO = KIND != 2;
        const T* tc = this->tmp + pbase;
        const T* yb = YB ? this->ybase + pbase : nullptr;
        const T* ai = AI ? this->acc_in + pbase : nullptr;
        T* const ao = this->acc_out + pbase;
        T* const uo = UO ? this->u_out + pbase : nullptr;
        const T ca = this->ca, cb = this->cb;
#pragma unroll 1
        for (int it = 0; it < ITER; ++it) {
            int p, ch; size_t g; this->unit_of(tid + it * NG, x0, p, ch, g);
            Vec<T> ta[2], p1a[2], p2a[2], ya[2], aa[2];
            vload_stream2(tc + g, ta[0], ta[1]);
            vload_stream2(p1 + g, p1a[0], p1a[1]);
            vload_stream2(p2 + g, p2a[0], p2a[1]);
            if (YB) vload_stream2(yb + g, ya[0], ya[1]);
            if (AI) vload_stream2(ai + g, aa[0], aa[1]);
            const T sgn = (p & 1) ? (T)-1 : (T)1;                       // x0 is even: + for even x, − for odd x
            Vec<T> a0[2], u0[2];
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                const Vec<T> z = vload(buf + p * N + this->swzp(ch + s2, p) * V);
                const Vec<T> jv = vload(jc + (ch + s2) * V);
#pragma unroll
                for (int q = 0; q < V; ++q) {
                    const T kk = p1a[s2].v[q] * (ta[s2].v[q] + sgn * jv.v[q]) + p2a[s2].v[q] * z.v[q];
                    const T y0 = YB ? ya[s2].v[q] : (T)0;
                    a0[s2].v[q] = (AI ? aa[s2].v[q] : y0) + cb * kk;
                    u0[s2].v[q] = y0 + ca * kk;
                }
            }
            vstore2(ao + g, a0[0], a0[1]);
            if (UO) vstore2(uo + g, u0[0], u0[1]);
        }
    }
    template <int KIND> DEV void epi_step(const Ctx& x, int i, int& cj, const T*& jline, Tw<R1>& w1, Tw<R2>& w2) const {
        const int tile = tile_of(x, i), slot = i % 3, c = plane_of(tile), x0 = x0_of(tile);
        int fl = 0;
        CMBL_FOR_GROUP(tid, NG, NG) { if (tid == 0 && c != cj) fl = (flag_look(this->jn_flag + c) == this->epoch); }   // J[N] line of this plane published?
        group_sync(BAR_FULL + slot, NT);
        if (c != cj) {
            const int ok = group_or(BAR_EPI, NG, fl);
            jline = this->jn_resolve(ok, c, x.nC, x.jws, x.jmine, w1, w2, NG, BAR_EPI);
            cj = c;
        }
        const T* const p1 = p_plane(this->pk, c, this->Npol, this->Nbphi, 0, x.nmap);
        const T* const p2 = p_plane(this->pk, c, this->Npol, this->Nbphi, 1, x.nmap);
        CMBL_FOR_GROUP(tid, NG, NG) { epi_tile<KIND>(x.sbase + slot * TILE, tid, (size_t)c * x.nmap, x0, jline, p1, p2); }
        if (i + 3 < x.ntl) group_arrive(BAR_EMPTY + slot, NT);           // (reads of the buffer are complete: they fed the stores above)
    }

    DEV void operator()(int blk, unsigned char* smem) const {
        const int kind = !this->acc_in ? 0 : (this->u_out ? 1 : 2);
        if (kind == 0) run<0>(blk, smem); else if (kind == 1) run<1>(blk, smem); else run<2>(blk, smem);
    }
    template <int KIND> DEV void run(int blk, unsigned char* smem) const {
        Ctx x;
        x.sbase = reinterpret_cast<T*>(smem); x.jws = x.sbase + 3 * TILE; x.jmine = this->jn_blk + (size_t)blk * N;
        x.nmap = (size_t)N * this->Nx; x.nC = this->ntiles / this->tiles_per_plane;
        const int nblocks = this->nblocks, ntiles = this->ntiles;
        x.tstep = this->contig ? 1 : nblocks;
        x.tile0 = this->contig ? (int)((long long)blk * ntiles / nblocks) : blk;
        const int tend = this->contig ? (int)((long long)(blk + 1) * ntiles / nblocks) : ntiles;
        x.ntl = x.tile0 < tend ? (tend - x.tile0 + x.tstep - 1) / x.tstep : 0;
        Tw<R1> w1; Tw<R2> w2;
        pdl_launch_dependents();
        pdl_wait();
#ifdef __CUDA_ARCH__
        if (threadIdx.x < NG) {
            reg_inc_fft();
            if (x.ntl > 0) { this->issue_tile(this->u + (size_t)plane_of(x.tile0) * x.nmap, x0_of(x.tile0), x.sbase, (int)threadIdx.x); cp_async_commit(); }
            this->jn_publish(blk, x.nC, x.jws, w1, w2, 0, BAR_FFT);      // (publisher blocks only) while the first tile is in flight
            for (int i = 0; i < x.ntl; ++i) fft_step(x, i, w1, w2);
        } else {
            reg_dec_epi();
            int cj = -1; const T* jline = nullptr;
            for (int i = 0; i < x.ntl; ++i) epi_step<KIND>(x, i, cj, jline, w1, w2);
        }
#else
        // host emulator: the two roles of a block run one after the other, tile by tile
        if (x.ntl > 0) { CMBL_FOR_GROUP(tid, NG, 0) { this->issue_tile(this->u + (size_t)plane_of(x.tile0) * x.nmap, x0_of(x.tile0), x.sbase, tid); } }
        this->jn_publish(blk, x.nC, x.jws, w1, w2, 0, BAR_FFT);
        int cj = -1; const T* jline = nullptr;
        for (int i = 0; i < x.ntl; ++i) { fft_step(x, i, w1, w2); epi_step<KIND>(x, i, cj, jline, w1, w2); }
#endif
    }
};

