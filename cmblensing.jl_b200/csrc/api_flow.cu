// extern "C": LenseFlow entry points.
#include "api_common.cuh"
#include "flow.cuh"

CMBL_FLOW_STRUCT;

namespace cmbl {
template <class T> void flow_grad(FlowT<T>& F, int op, const T* fout, const C2<T>* delta, C2<T>* dfield, C2<T>* dphi,
                                  bool bug_compat, cmblStream_t st);
}

extern "C" {

int cmbl_lenseflow_create(cmbl_flow** flow, cmbl_plan* plan, int nsteps, int Npol, int Nb_f, int Nb_phi) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(flow && plan && plan->p, "NULL argument");
    CMBL_REQUIRE(nsteps >= 1 && nsteps <= 64, "nsteps must be in 1..64");
    CMBL_REQUIRE(Npol >= 1 && Npol <= 3 && Nb_f >= 1, "Npol must be 1..3 and Nb_f >= 1");
    CMBL_REQUIRE(Nb_phi == 1 || Nb_phi == Nb_f, "Nb_phi must be 1 or Nb_f (batch sizes must broadcast, src/batching.jl)");
    auto h = std::make_unique<cmbl_flow>();
    h->plan = plan;
    CMBL_DISPATCH(plan->p.get(), {
        auto F = std::make_unique<cmbl::FlowT<T>>();
        F->plan = &P; F->P = &P; F->nsteps = nsteps; F->Npol = Npol; F->Nb = Nb_f; F->Nbphi = Nb_phi; F->C = Npol * Nb_f;
        h->f = std::move(F);
    });
    *flow = h.release();
    CMBL_API_END
}

int cmbl_lenseflow_destroy(cmbl_flow* flow) {
    CMBL_API_BEGIN
    delete flow;
    CMBL_API_END
}

int cmbl_lenseflow_precompute(cmbl_flow* flow, const void* phi, int phi_basis, int with_minv, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(flow && flow->f && phi, "NULL argument");
    CMBL_REQUIRE(phi_basis == CMBL_MAP || phi_basis == CMBL_FOURIER, "phi_basis must be Map or Fourier");
    CMBL_DISPATCH(flow->f->plan, cmbl::flow_precompute<T>(*static_cast<cmbl::FlowT<T>*>(flow->f.get()), phi, phi_basis,
                                                          with_minv != 0, as_stream(stream)));
    CMBL_API_END
}

int cmbl_lenseflow_apply(cmbl_flow* flow, int op, const void* in, void* out, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(flow && flow->f && in && out, "NULL argument");
    CMBL_DISPATCH(flow->f->plan, cmbl::flow_apply<T>(*static_cast<cmbl::FlowT<T>*>(flow->f.get()), op, in, out, as_stream(stream)));
    CMBL_API_END
}

#ifndef CMBL_EMU
namespace {
// copy streams / events of the pipelined host path (per host thread, created on first use)
struct HostPipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_start = nullptr, ev_fin = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_done;
    // asynchronous calls: two staging slots used alternately; slot_free[s] = the D2H of the call that used slot s last has finished
    cmbl::DevBuf slot[2]; cudaEvent_t slot_free[2] = {nullptr, nullptr}, a_in[2] = {nullptr, nullptr}, a_done[2] = {nullptr, nullptr};
    bool slot_used[2] = {false, false}; unsigned ncalls = 0;
    void ensure(int n) {
        if (!s_in) {
            CMBL_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
            CMBL_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
            CMBL_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
            CMBL_CUDA(cudaEventCreateWithFlags(&ev_fin, cudaEventDisableTiming));
            for (int i = 0; i < 2; ++i) {
                CMBL_CUDA(cudaEventCreateWithFlags(&slot_free[i], cudaEventDisableTiming));
                CMBL_CUDA(cudaEventCreateWithFlags(&a_in[i], cudaEventDisableTiming));
                CMBL_CUDA(cudaEventCreateWithFlags(&a_done[i], cudaEventDisableTiming));
            }
        }
        while ((int)ev_in.size() < n) {
            cudaEvent_t a, b;
            CMBL_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
            CMBL_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
            ev_in.push_back(a); ev_done.push_back(b);
        }
    }
};
HostPipe& host_pipe() { static thread_local HostPipe hp; return hp; }
int host_chunks() { static const int v = [] { const char* e = getenv("CMBL_HOST_CHUNKS"); return e ? atoi(e) : 3; }(); return v; }
}  // namespace
#endif

int cmbl_lenseflow_apply_host(cmbl_flow* flow, int op, const void* in_host, void* out_host, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(flow && flow->f && in_host && out_host, "NULL argument");
    CMBL_REQUIRE(op >= 0 && op <= 3, "LenseFlow op must be 0..3");
    CMBL_DISPATCH(flow->f->plan, {
        auto& F = *static_cast<cmbl::FlowT<T>*>(flow->f.get());
        const bool four = (op == CMBL_OP_LH || op == CMBL_OP_LHINV);
        const size_t bytes = (four ? sizeof(cmbl::C2<T>) * P.four_elems() : sizeof(T) * P.map_elems()) * (size_t)F.C;
        static thread_local cmbl::DevBuf io;
        void* d = io.reserve(bytes);
#ifdef CMBL_EMU
        memcpy(d, in_host, bytes);
        cmbl::flow_apply<T>(F, op, d, d, as_stream(stream));
        cmbl::dev_download(out_host, d, bytes, as_stream(stream));
#else
        cudaStream_t st = as_stream(stream);
        int nch = host_chunks();
        if (nch > F.Nb) nch = F.Nb;
        if (four || nch <= 1 || !F.integrated_once) {          // (first use of a handle: the serial path also allocates the work buffers)
            CMBL_CUDA(cudaMemcpyAsync(d, in_host, bytes, cudaMemcpyHostToDevice, st));
            cmbl::flow_apply<T>(F, op, d, d, st);
            cmbl::dev_download(out_host, d, bytes, st);
        } else {
            // Map-space flows (L*f, L\f): batch items are independent, so the batch moves through a three-stage pipeline —
            // H2D of items i+1.. on one copy stream, the integration of item group i on the caller's stream, D2H of finished
            // groups on a second copy stream (PCIe is full duplex) — instead of copy-in, compute, copy-out back to back.
            HostPipe& hp = host_pipe();
            // group sizes: the first and the last group are the exposed transfers, so they are small (a quarter of the batch);
            // the middle of the batch moves in larger groups that keep the persistent stage kernels filled
            std::vector<int> sizes;
            if (nch == 2 || F.Nb < 4) { const int per = (F.Nb + nch - 1) / nch; for (int b0 = 0; b0 < F.Nb; b0 += per) sizes.push_back(F.Nb - b0 < per ? F.Nb - b0 : per); }
            else {
                const int q = F.Nb / 4 > 0 ? F.Nb / 4 : 1, mid = F.Nb - 2 * q, nm = nch - 2, per = (mid + nm - 1) / nm;
                sizes.push_back(q);
                for (int b0 = 0; b0 < mid; b0 += per) sizes.push_back(mid - b0 < per ? mid - b0 : per);
                sizes.push_back(q);
            }
            const int ng = (int)sizes.size();
            hp.ensure(ng);
            const int n = F.nsteps;
            const size_t plane_b = sizeof(T) * P.map_elems();
            const char* hin = static_cast<const char*>(in_host); char* hout = static_cast<char*>(out_host); char* dd = static_cast<char*>(d);
            CMBL_CUDA(cudaEventRecord(hp.ev_start, st));                       // the staging buffer is free once earlier work on `st` is done
            CMBL_CUDA(cudaStreamWaitEvent(hp.s_in, hp.ev_start, 0));
            // All groups integrate on the caller's stream, one after another: they share the handle's work buffers.
            for (int i = 0, b0 = 0; i < ng; b0 += sizes[i], ++i) {
                const size_t off = plane_b * (size_t)b0 * F.Npol, cb = plane_b * (size_t)sizes[i] * F.Npol;
                CMBL_CUDA(cudaMemcpyAsync(dd + off, hin + off, cb, cudaMemcpyHostToDevice, hp.s_in));
                CMBL_CUDA(cudaEventRecord(hp.ev_in[i], hp.s_in));
            }
            for (int i = 0, b0 = 0; i < ng; b0 += sizes[i], ++i) {
                const size_t off = plane_b * (size_t)b0 * F.Npol, cb = plane_b * (size_t)sizes[i] * F.Npol;
                cudaStream_t sc = st;
                CMBL_CUDA(cudaStreamWaitEvent(sc, hp.ev_in[i], 0));
                cmbl::flow_integrate_range<T>(F, false, reinterpret_cast<T*>(d), op == CMBL_OP_L ? 0 : 2 * n, op == CMBL_OP_L ? 2 * n : 0,
                                              b0 * F.Npol, sizes[i] * F.Npol, sc);
                CMBL_CUDA(cudaEventRecord(hp.ev_done[i], sc));
                CMBL_CUDA(cudaStreamWaitEvent(hp.s_out, hp.ev_done[i], 0));
                CMBL_CUDA(cudaMemcpyAsync(hout + off, dd + off, cb, cudaMemcpyDeviceToHost, hp.s_out));
            }
            CMBL_CUDA(cudaEventRecord(hp.ev_fin, hp.s_out));
            CMBL_CUDA(cudaStreamWaitEvent(st, hp.ev_fin, 0));
            CMBL_CUDA(cudaStreamSynchronize(st));                              // out_host is valid on return, as before
        }
#endif
    });
    CMBL_API_END
}

// Asynchronous variant for a caller that streams many fields through one operator: returns as soon as the work is queued.  Successive
// calls overlap — the H2D copy of call i+1 runs during the integration of call i, whose D2H copy runs during the integration of call
// i+1 (two device staging slots, two copy streams) — so the steady-state cost of a call is max(integration, one-way transfer) instead of
// their sum.  in_host must stay unchanged and out_host is valid only after cmbl_lenseflow_host_sync().
int cmbl_lenseflow_apply_host_async(cmbl_flow* flow, int op, const void* in_host, void* out_host, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(flow && flow->f && in_host && out_host, "NULL argument");
    CMBL_REQUIRE(op >= 0 && op <= 3, "LenseFlow op must be 0..3");
    CMBL_DISPATCH(flow->f->plan, {
        auto& F = *static_cast<cmbl::FlowT<T>*>(flow->f.get());
        const bool four = (op == CMBL_OP_LH || op == CMBL_OP_LHINV);
        const size_t bytes = (four ? sizeof(cmbl::C2<T>) * P.four_elems() : sizeof(T) * P.map_elems()) * (size_t)F.C;
#ifdef CMBL_EMU
        static thread_local cmbl::DevBuf io;
        void* d = io.reserve(bytes);
        memcpy(d, in_host, bytes);
        cmbl::flow_apply<T>(F, op, d, d, as_stream(stream));
        cmbl::dev_download(out_host, d, bytes, as_stream(stream));
#else
        cudaStream_t st = as_stream(stream);
        HostPipe& hp = host_pipe();
        hp.ensure(1);
        const int s = (int)(hp.ncalls++ & 1);
        if (hp.slot[s].cap < bytes) {                                          // (re)allocation: drain everything that may still use the slot
            CMBL_CUDA(cudaStreamSynchronize(hp.s_in)); CMBL_CUDA(cudaStreamSynchronize(hp.s_out)); CMBL_CUDA(cudaStreamSynchronize(st));
            hp.slot_used[s] = false;
        }
        void* d = hp.slot[s].reserve(bytes);
        if (hp.slot_used[s]) CMBL_CUDA(cudaStreamWaitEvent(hp.s_in, hp.slot_free[s], 0));         // the slot's previous result has left the device
        CMBL_CUDA(cudaMemcpyAsync(d, in_host, bytes, cudaMemcpyHostToDevice, hp.s_in));
        CMBL_CUDA(cudaEventRecord(hp.a_in[s], hp.s_in));
        CMBL_CUDA(cudaStreamWaitEvent(st, hp.a_in[s], 0));
        cmbl::flow_apply<T>(F, op, d, d, st);
        CMBL_CUDA(cudaEventRecord(hp.a_done[s], st));
        CMBL_CUDA(cudaStreamWaitEvent(hp.s_out, hp.a_done[s], 0));
        CMBL_CUDA(cudaMemcpyAsync(out_host, d, bytes, cudaMemcpyDeviceToHost, hp.s_out));
        CMBL_CUDA(cudaEventRecord(hp.slot_free[s], hp.s_out));
        hp.slot_used[s] = true;
#endif
    });
    CMBL_API_END
}

// waits until every cmbl_lenseflow_apply_host_async call of this host thread has delivered its out_host
int cmbl_lenseflow_host_sync(void) {
    CMBL_API_BEGIN
#ifndef CMBL_EMU
    HostPipe& hp = host_pipe();
    if (hp.s_out) { CMBL_CUDA(cudaStreamSynchronize(hp.s_in)); CMBL_CUDA(cudaStreamSynchronize(hp.s_out)); }
#endif
    CMBL_API_END
}

int cmbl_lenseflow_grad(cmbl_flow* flow, int op, const void* f_out_map, const void* delta_four, void* dfield_four,
                        void* dphi_four, int bug_compat, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(flow && flow->f && f_out_map && delta_four && dfield_four && dphi_four, "NULL argument");
    CMBL_REQUIRE(op == CMBL_OP_L || op == CMBL_OP_LINV, "gradient is defined for op 0 (L*f) and 2 (L\\f)");
    CMBL_DISPATCH(flow->f->plan, cmbl::flow_grad<T>(*static_cast<cmbl::FlowT<T>*>(flow->f.get()), op, (const T*)f_out_map,
                  (const cmbl::C2<T>*)delta_four, (cmbl::C2<T>*)dfield_four, (cmbl::C2<T>*)dphi_four, bug_compat != 0, as_stream(stream)));
    CMBL_API_END
}

int cmbl_max_lensing_step(cmbl_plan* plan, const void* phi, int phi_basis, const void* eta, int eta_basis, int Nb, double* out_host, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(plan && plan->p && phi && eta && out_host, "NULL argument");
    CMBL_REQUIRE(Nb >= 1, "Nb must be >= 1");
    CMBL_REQUIRE((phi_basis == CMBL_MAP || phi_basis == CMBL_FOURIER) && (eta_basis == CMBL_MAP || eta_basis == CMBL_FOURIER), "basis must be Map or Fourier");
    CMBL_DISPATCH(plan->p.get(), cmbl::max_lensing_step<T>(P, phi, phi_basis, eta, eta_basis, Nb, out_host, as_stream(stream)));
    CMBL_API_END
}

int cmbl_lenseflow_kernel_path(cmbl_flow* flow) {
    if (!flow || !flow->f) return 0;
    int r = 0;
    try { CMBL_DISPATCH(flow->f->plan, r = cmbl::flow_kernel_path<T>(*static_cast<cmbl::FlowT<T>*>(flow->f.get()))); } catch (...) { r = 0; }
    return r;
}

int cmbl_lenseflow_get_p(cmbl_flow* flow, int k, void* out_host) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(flow && flow->f && out_host, "NULL argument");
    CMBL_DISPATCH(flow->f->plan, {
        auto& F = *static_cast<cmbl::FlowT<T>*>(flow->f.get());
        CMBL_REQUIRE(F.have_p, "precompute first");
        CMBL_REQUIRE(k >= 0 && k <= 2 * F.nsteps, "k out of range");
        const size_t n = F.nmap() * 2 * F.Nbphi;
        if (F.pcache_G == 0) cmbl::dev_download(out_host, F.pk(k), sizeof(T) * n, 0);
        else {                                           // undo the row-grouped layout of the fast path
            std::vector<T> h(n);
            cmbl::dev_download(h.data(), F.pk(k), sizeof(T) * n, 0);
            T* o = reinterpret_cast<T*>(out_host);
            const int G = F.pcache_G, Nx = P.Nx, Ny = P.Ny;
            for (size_t pl = 0; pl < (size_t)2 * F.Nbphi; ++pl)
                for (int x = 0; x < Nx; ++x)
                    for (int y = 0; y < Ny; ++y)
                        o[pl * F.nmap() + (size_t)x * Ny + y] = h[pl * F.nmap() + ((size_t)(y / G) * Nx + x) * G + (y % G)];
        }
    });
    CMBL_API_END
}

}  // extern "C"
