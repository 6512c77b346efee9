// extern "C": the one collective the path has when a batch is sharded over GPUs — conjugate_gradient's lock-step bookkeeping
// (all(res < bestres), all(res < tol) over EVERY batch item, src/numerical_algorithms.jl:110-121) and MAP_joint's batch-summed line-search
// objective (src/maximization.jl:197): a handful of scalars per iteration.  NCCL is resolved at run time (dlopen) the first time a
// communicator is created, so the library itself links against cudart only and a single-GPU user never loads it.
#include "api_common.cuh"
#include "cg.cuh"
#ifndef CMBL_EMU
#include <dlfcn.h>
#endif

struct cmbl_cg { std::unique_ptr<cmbl::CgBase> g; cmbl_flow* flow; };

namespace {
#ifndef CMBL_EMU
typedef struct ncclComm* ncclComm_t;
struct NcclId { char internal[128]; };
struct Nccl {
    void* so = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int /*dtype*/, int /*op*/, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
Nccl& nccl() {
    static Nccl n;
    if (!n.so) {
        const char* env = getenv("CMBL_NCCL_LIB");
        // a copy that is already mapped into the process (e.g. the one PyTorch or CUDA.jl/NCCL.jl brought) wins over a second one
        const char* names[] = {env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) { n.so = dlopen(nm, RTLD_NOW | RTLD_NOLOAD); if (n.so) break; }
        for (const char* nm : names) { if (n.so) break; n.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); }
        CMBL_REQUIRE(n.so != nullptr, "cmbl_comm: libnccl.so.2 not found (set CMBL_NCCL_LIB to its path)");
        auto sym = [&](const char* s) { void* p = dlsym(n.so, s); CMBL_REQUIRE(p != nullptr, std::string("cmbl_comm: NCCL symbol missing: ") + s); return p; };
        n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
        n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
        n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
    }
    return n;
}
#define CMBL_NCCL(call) do { int _r = (call); if (_r != 0) throw ::cmbl::Error(std::string("cuda/nccl " #call ": ") + nccl().GetErrorString(_r)); } while (0)
#endif
}  // namespace

struct cmbl_comm {
    int nranks = 1, rank = 0;
#ifndef CMBL_EMU
    ncclComm_t comm = nullptr;
    double* dbuf = nullptr;            // device staging for the scalars (64 doubles)
#endif
};

namespace cmbl {
// in-place all-reduce of a few host doubles; op 0 sum, 1 min, 2 max
void comm_allreduce(cmbl_comm* c, double* v, int n, int op, cmblStream_t st) {
    if (!c || c->nranks == 1) return;
#ifdef CMBL_EMU
    (void)v; (void)n; (void)op; (void)st;
    throw Error("the host emulator has no collectives");
#else
    CMBL_REQUIRE(n >= 1 && n <= 64, "cmbl_comm_allreduce carries 1..64 scalars");
    CMBL_CUDA(cudaMemcpyAsync(c->dbuf, v, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    const int ncclFloat64 = 8, ops[3] = {0 /*sum*/, 3 /*min*/, 2 /*max*/};
    CMBL_NCCL(nccl().AllReduce(c->dbuf, c->dbuf, (size_t)n, ncclFloat64, ops[op], c->comm, st));
    CMBL_CUDA(cudaMemcpyAsync(v, c->dbuf, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CMBL_CUDA(cudaStreamSynchronize(st));
#endif
}

// conjugate_gradient's loop (numerical_algorithms.jl:99-121) with the two lock-step rules taken over the batch items of every rank
template <class T>
void wiener_cg_loop(CgT<T>& G, cmbl_comm* comm, const C2<T>* fstart, void* f_out, int nsteps, double tol, bool offset, int* iters_out,
                    double* res_hist_host, cmblStream_t st) {
    const int Nb = G.Nb;
    const size_t vb = sizeof(C2<T>) * G.nf() * G.C;
    std::vector<double> res(Nb), best(Nb);
    cg_begin<T>(G, fstart, offset, res.data(), st);
    best = res;
    if (res_hist_host) for (int b = 0; b < Nb; ++b) res_hist_host[b] = res[b];
    int i = 1;
    for (i = 2; i <= nsteps; ++i) {
        cg_step<T>(G, res.data(), st);
        double flags[2] = {1.0, 1.0};                                         // {all(res < bestres), all(res < tol)} on this rank
        for (int b = 0; b < Nb; ++b) { if (!(res[b] < best[b])) flags[0] = 0.0; if (!(res[b] < tol)) flags[1] = 0.0; }
        comm_allreduce(comm, flags, 2, 1 /*min = logical and*/, st);
        if (flags[0] != 0.0) { best = res; dev_copy(G.bestx.p, G.x.p, vb, st); }
        if (res_hist_host) for (int b = 0; b < Nb; ++b) res_hist_host[(size_t)(i - 1) * Nb + b] = res[b];
        if (flags[1] != 0.0) break;
    }
    if (iters_out) *iters_out = (i > nsteps) ? nsteps : i;
    dev_copy(f_out, G.bestx.p, vb, st);
}
template void wiener_cg_loop<float>(CgT<float>&, cmbl_comm*, const C2<float>*, void*, int, double, bool, int*, double*, cmblStream_t);
template void wiener_cg_loop<double>(CgT<double>&, cmbl_comm*, const C2<double>*, void*, int, double, bool, int*, double*, cmblStream_t);
}  // namespace cmbl

extern "C" {

int cmbl_comm_unique_id(void* id128) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(id128 != nullptr, "NULL argument");
#ifdef CMBL_EMU
    memset(id128, 0, 128);
#else
    NcclId id; CMBL_NCCL(nccl().GetUniqueId(&id)); memcpy(id128, &id, 128);
#endif
    CMBL_API_END
}

int cmbl_comm_init(cmbl_comm** comm, int nranks, int rank, const void* id128) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(comm && id128, "NULL argument");
    CMBL_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "need 0 <= rank < nranks");
    auto c = std::make_unique<cmbl_comm>();
    c->nranks = nranks; c->rank = rank;
#ifdef CMBL_EMU
    CMBL_REQUIRE(nranks == 1, "the host emulator has no collectives (nranks must be 1)");
#else
    NcclId id; memcpy(&id, id128, 128);
    CMBL_NCCL(nccl().CommInitRank(&c->comm, nranks, id, rank));             // on the calling thread's current device
    CMBL_CUDA(cudaMalloc(&c->dbuf, sizeof(double) * 64));
#endif
    *comm = c.release();
    CMBL_API_END
}

int cmbl_comm_destroy(cmbl_comm* comm) {
    CMBL_API_BEGIN
#ifndef CMBL_EMU
    if (comm) { if (comm->comm) nccl().CommDestroy(comm->comm); if (comm->dbuf) cudaFree(comm->dbuf); }
#endif
    delete comm;
    CMBL_API_END
}

int cmbl_comm_allreduce(cmbl_comm* comm, double* values_host, int n, int op, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(comm && values_host, "NULL argument");
    CMBL_REQUIRE(op >= 0 && op <= 2, "op must be 0 (sum), 1 (min) or 2 (max)");
    cmbl::comm_allreduce(comm, values_host, n, op, as_stream(stream));
    CMBL_API_END
}

int cmbl_wiener_cg_sharded(cmbl_cg* cg, cmbl_comm* comm_or_null, const void* fstart_or_null, void* f_out, int nsteps, double tol, int offset,
                           int* iters_out, double* res_hist_host, void* stream) {
    CMBL_API_BEGIN
    CMBL_REQUIRE(cg && cg->g && f_out, "NULL argument");
    CMBL_REQUIRE(nsteps >= 1, "nsteps must be >= 1");
    CMBL_DISPATCH(cg->g->plan, cmbl::wiener_cg_loop<T>(*static_cast<cmbl::CgT<T>*>(cg->g.get()), comm_or_null, (const cmbl::C2<T>*)fstart_or_null, f_out,
                                                       nsteps, tol, offset != 0, iters_out, res_hist_host, as_stream(stream)));
    CMBL_API_END
}

}  // extern "C"
