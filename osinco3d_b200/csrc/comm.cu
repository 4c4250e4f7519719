// comm.cu -- z-slab domain decomposition plumbing: ghost-plane exchange with grouped
// ncclSend/ncclRecv over NVLink and small all-reduces (SOR residual, scalar-clipping sums,
// statistics).  The reference has no parallelism at all (README.md:61); this is new.
//
// NCCL is loaded lazily with dlopen so that libo3d_b200.so itself has no link-time dependency
// on it (single-GPU users and the CPU-side "does it load" check never touch NCCL).  Inside a
// torch process the already-loaded bundled libnccl.so.2 is reused.
#include <dlfcn.h>

#include <cstring>
#include <vector>

#include "session.h"

namespace o3d {

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess_ = 0 };
// ncclDataType_t / ncclRedOp_t values (nccl.h, stable since 2.x)
enum { ncclUint8_ = 1, ncclUint64_ = 5, ncclFloat64_ = 8 };
enum { ncclSum_ = 0, ncclMax_ = 2, ncclMin_ = 3 };

struct Api {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

Api g_api;

bool load_api() {
    if (g_api.h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        const char* env = getenv("O3D_NCCL_LIB");
        if (env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!h) {
        set_error("cannot dlopen libnccl.so.2 (set O3D_NCCL_LIB): %s", dlerror());
        return false;
    }
#define O3D_SYM(field, name)                                         \
    *(void**)(&g_api.field) = dlsym(h, name);                        \
    if (!g_api.field) {                                              \
        set_error("libnccl: missing symbol %s", name);               \
        return false;                                                \
    }
    O3D_SYM(GetUniqueId, "ncclGetUniqueId")
    O3D_SYM(CommInitRank, "ncclCommInitRank")
    O3D_SYM(CommDestroy, "ncclCommDestroy")
    O3D_SYM(Send, "ncclSend")
    O3D_SYM(Recv, "ncclRecv")
    O3D_SYM(AllReduce, "ncclAllReduce")
    O3D_SYM(Broadcast, "ncclBroadcast")
    O3D_SYM(AllGather, "ncclAllGather")
    O3D_SYM(GroupStart, "ncclGroupStart")
    O3D_SYM(GroupEnd, "ncclGroupEnd")
    O3D_SYM(GetErrorString, "ncclGetErrorString")
#undef O3D_SYM
    g_api.h = h;
    return true;
}

}  // namespace

struct Comm {
    ncclComm_t nccl;
    int rank, nranks;
};

#define O3D_NCCL_CHECK(call)                                                             \
    do {                                                                                 \
        ncclResult_t _r = (call);                                                        \
        if (_r != ncclSuccess_) {                                                        \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                      \
                      g_api.GetErrorString ? g_api.GetErrorString(_r) : "nccl error");   \
            return O3D_ERR_COMM;                                                         \
        }                                                                                \
    } while (0)

int nccl_unique_id(unsigned char* out128) {
    if (!load_api()) return O3D_ERR_COMM;
    ncclUniqueId id;
    O3D_NCCL_CHECK(g_api.GetUniqueId(&id));
    memcpy(out128, id.internal, 128);
    return O3D_OK;
}

int comm_create(o3d_session* s) {
    s->comm = nullptr;
    if (s->cfg.nranks <= 1) return O3D_OK;
    if (!load_api()) return O3D_ERR_COMM;
    Comm* c = new Comm();
    c->rank = s->cfg.rank;
    c->nranks = s->cfg.nranks;
    ncclUniqueId id;
    memcpy(id.internal, s->cfg.nccl_id, 128);
    ncclResult_t r = g_api.CommInitRank(&c->nccl, c->nranks, id, c->rank);
    if (r != ncclSuccess_) {
        set_error("ncclCommInitRank failed: %s", g_api.GetErrorString(r));
        delete c;
        return O3D_ERR_COMM;
    }
    s->comm = c;
    // Establish the point-to-point connections to BOTH ring neighbours now, wrap link included.
    // NCCL sets a connection up lazily at the first send/recv between a pair, on the host, and
    // that can take seconds at 8 GPUs.  The stencil halos of a run whose nbcz is free-slip never
    // use the wrap link, but poisson_solver_0000 / _0011 wrap z whatever nbcz is
    // (src/initialization.f90:283-301): its first use would then fall inside sor_solve, while the
    // other ranks' persistent SOR kernels already spin on this rank's residual (observed at 8
    // GPUs: their bounded spin expired).  One 8-byte ring exchange here takes the set-up out of
    // the time loop.
    if (c->nranks > 1) {
        double* w = nullptr;
        if (cudaMalloc(&w, 4 * sizeof(double)) == cudaSuccess) {
            cudaMemsetAsync(w, 0, 4 * sizeof(double), s->st);
            const int up = (c->rank + 1) % c->nranks, dn = (c->rank + c->nranks - 1) % c->nranks;
            g_api.GroupStart();
            g_api.Send(w, 1, ncclFloat64_, up, c->nccl, s->st);
            g_api.Recv(w + 1, 1, ncclFloat64_, dn, c->nccl, s->st);
            g_api.Send(w + 2, 1, ncclFloat64_, dn, c->nccl, s->st);
            g_api.Recv(w + 3, 1, ncclFloat64_, up, c->nccl, s->st);
            g_api.GroupEnd();
            cudaStreamSynchronize(s->st);
            cudaFree(w);
        }
    }
    return O3D_OK;
}

// ---- peer memory (CUDA IPC over NVLink) for the persistent SOR kernel ------------------------
// Every rank exports three allocations: its PeerBlock and its two physical pp buffers.  The
// handles travel in one ncclAllGather; a rank maps all blocks (the residual maxima go all-to-all)
// and the pp buffers of its z neighbours (ghost planes are stored straight into them).
struct PeerState {
    int ok = 0;
    PeerBlock* mine = nullptr;
    PeerBlock* all[16] = {};
    double* nb_alloc[16][3] = {};  // [rank][physical pp buffer 0 / 1, rhs]: mapped allocation bases
    int nb_nz[16] = {};
    void* opened[48] = {};
    int nopened = 0;
};

namespace {
struct PeerExport {
    cudaIpcMemHandle_t block, p[3];  // p[0], p[1]: physical pp buffers, p[2]: O3D_F_RHS
    int nzl, ok, pad[2];
};
}  // namespace

int comm_peer_setup(o3d_session* s) {
    Comm* c = s->comm;
    if (!c || c->nranks > 16) return 1;
    if (s->peers) return s->peers->ok ? 0 : 1;
    if (s->peers_tried) return 1;
    s->peers_tried = 1;
    PeerState* ps = new PeerState();
    s->peers = ps;
    const int P = c->nranks, me = c->rank;
    PeerExport mine;
    memset(&mine, 0, sizeof(mine));
    mine.nzl = s->nzl;
    mine.ok = 1;
    // physical buffer 0 is the allocation that was O3D_F_PP at session start
    double* phys[2] = {s->base[s->pp_phys ? O3D_F_PP2 : O3D_F_PP],
                       s->base[s->pp_phys ? O3D_F_PP : O3D_F_PP2]};
    double* rhs_alloc = s->base[O3D_F_RHS];
    if (!phys[0] || !phys[1] || !rhs_alloc) mine.ok = 0;
    if (mine.ok && (cudaMalloc(&ps->mine, sizeof(PeerBlock)) != cudaSuccess ||
                    cudaMemset(ps->mine, 0, sizeof(PeerBlock)) != cudaSuccess))
        mine.ok = 0;
    if (mine.ok && (cudaIpcGetMemHandle(&mine.block, ps->mine) != cudaSuccess ||
                    cudaIpcGetMemHandle(&mine.p[0], phys[0]) != cudaSuccess ||
                    cudaIpcGetMemHandle(&mine.p[1], phys[1]) != cudaSuccess ||
                    cudaIpcGetMemHandle(&mine.p[2], rhs_alloc) != cudaSuccess)) {
        (void)cudaGetLastError();
        mine.ok = 0;
    }
    // all-gather the exports (every rank takes part, also one that failed so far)
    PeerExport* dev = nullptr;
    std::vector<PeerExport> allx(P);
    if (cudaMalloc(&dev, sizeof(PeerExport) * (size_t)(P + 1)) != cudaSuccess) return 1;
    cudaMemcpyAsync(dev + P, &mine, sizeof(mine), cudaMemcpyHostToDevice, s->st);
    ncclResult_t r = g_api.AllGather(dev + P, dev, sizeof(PeerExport), ncclUint8_, c->nccl, s->st);
    cudaMemcpyAsync(allx.data(), dev, sizeof(PeerExport) * (size_t)P, cudaMemcpyDeviceToHost, s->st);
    const cudaError_t ce = cudaStreamSynchronize(s->st);
    cudaFree(dev);
    if (r != ncclSuccess_ || ce != cudaSuccess) return 1;
    bool ok = true;
    for (int q = 0; q < P; ++q) ok = ok && allx[q].ok;
    // neighbours in the slab ring (the wrap link is mapped whether or not this solve uses it)
    const int up = (me + 1) % P, dn = (me + P - 1) % P;
    for (int q = 0; q < P && ok; ++q) {
        ps->nb_nz[q] = allx[q].nzl;
        if (q == me) {
            ps->all[q] = ps->mine;
            continue;
        }
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, allx[q].block, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            ok = false;
            break;
        }
        ps->opened[ps->nopened++] = ptr;
        ps->all[q] = static_cast<PeerBlock*>(ptr);
        if (q == up || q == dn) {
            for (int b = 0; b < 3 && ok; ++b) {
                if (cudaIpcOpenMemHandle(&ptr, allx[q].p[b], cudaIpcMemLazyEnablePeerAccess) !=
                    cudaSuccess) {
                    ok = false;
                    break;
                }
                ps->opened[ps->nopened++] = ptr;
                ps->nb_alloc[q][b] = static_cast<double*>(ptr);
            }
        }
    }
    if (!ok) (void)cudaGetLastError();
    // the decision must be the same on every rank: agree on it (min over ranks)
    double* flag = s->scal_d + 100;
    const double mineok = ok ? 1.0 : 0.0;
    double allok = 0.0;
    cudaMemcpyAsync(flag, &mineok, sizeof(double), cudaMemcpyHostToDevice, s->st);
    r = g_api.AllReduce(flag, flag, 1, ncclFloat64_, ncclMin_, c->nccl, s->st);
    cudaMemcpyAsync(&allok, flag, sizeof(double), cudaMemcpyDeviceToHost, s->st);
    if (cudaStreamSynchronize(s->st) != cudaSuccess || r != ncclSuccess_) return 1;
    ps->ok = allok > 0.5;
    return ps->ok ? 0 : 1;
}

int comm_peer_args(o3d_session* s, PeerSync* out) {
    PeerState* ps = s->peers;
    Comm* c = s->comm;
    if (!ps || !ps->ok || !c) return 1;
    memset(out, 0, sizeof(*out));
    const int P = c->nranks, me = c->rank;
    const int wrap = (s->sor_variant != 2);  // _0000 / _0011 wrap z, _111111 mirrors it
    const int up = (me + 1 < P) ? me + 1 : (wrap ? 0 : -1);
    const int dn = (me > 0) ? me - 1 : (wrap ? P - 1 : -1);
    out->nranks = P, out->rank = me;
    out->has_lo = dn >= 0, out->has_hi = up >= 0;
    out->iter_base = s->peer_iter_base;
    out->solve_base = s->peer_solves;
    out->push_init = 1;
    out->mine = ps->mine;
    for (int q = 0; q < P; ++q) out->all[q] = ps->all[q];
    const long long ioff = interior_offset(s->g);
    // role order of the neighbour's buffers = ours: all ranks swap in lockstep (same iteration
    // counts), so its current O3D_F_PP is its physical buffer pp_phys
    if (dn >= 0) {
        out->lo = ps->all[dn];
        out->lo_nz = ps->nb_nz[dn];
        out->lo_p[0] = ps->nb_alloc[dn][s->pp_phys] + ioff;
        out->lo_p[1] = ps->nb_alloc[dn][s->pp_phys ^ 1] + ioff;
        out->lo_rhs = ps->nb_alloc[dn][2] + ioff;
    }
    if (up >= 0) {
        out->hi = ps->all[up];
        out->hi_p[0] = ps->nb_alloc[up][s->pp_phys] + ioff;
        out->hi_p[1] = ps->nb_alloc[up][s->pp_phys ^ 1] + ioff;
        out->hi_rhs = ps->nb_alloc[up][2] + ioff;
    }
    return 0;
}

static void peer_destroy(o3d_session* s) {
    PeerState* ps = s->peers;
    if (!ps) return;
    for (int q = 0; q < ps->nopened; ++q) cudaIpcCloseMemHandle(ps->opened[q]);
    if (ps->ok && s->comm) {
        // an exporter must not free memory that an importer still maps (undefined behaviour per
        // the CUDA IPC contract): every rank has closed its mappings once this barrier is passed.
        // o3d_session_destroy of a multi-rank session is collective.
        double* flag = s->scal_d + 100;
        if (g_api.AllReduce(flag, flag, 1, ncclFloat64_, ncclMax_, s->comm->nccl, s->st) ==
            ncclSuccess_)
            cudaStreamSynchronize(s->st);
    }
    if (ps->mine) cudaFree(ps->mine);
    delete ps;
    s->peers = nullptr;
}

void comm_destroy(o3d_session* s) {
    peer_destroy(s);
    if (!s->comm) return;
    g_api.CommDestroy(s->comm->nccl);
    delete s->comm;
    s->comm = nullptr;
}

// Ghost-plane exchange of widths[f] planes per side of field f.  Planes are whole padded planes
// (px * py doubles, contiguous); `bases` are field allocation starts.  wrap != 0: the slab ring
// is periodic in z.  One grouped ncclSend/ncclRecv call on the COMMUNICATION stream, ordered
// after everything enqueued on the session stream so far.
int comm_exchange_async(o3d_session* s, double* const* bases, const int* widths, int nf, int wrap) {
    Comm* c = s->comm;
    if (!c) return O3D_OK;
    const int up = (c->rank + 1 < c->nranks) ? c->rank + 1 : (wrap ? 0 : -1);
    const int dn = (c->rank > 0) ? c->rank - 1 : (wrap ? c->nranks - 1 : -1);
    const long long sz = s->g.sz;
    O3D_CUDA_CHECK(cudaEventRecord(s->ev_ready, s->st));
    O3D_CUDA_CHECK(cudaStreamWaitEvent(s->st_comm, s->ev_ready, 0));
    trace_mark(s, 0, "exchange issued");
    trace_mark(s, 1, "exchange begin");
    O3D_NCCL_CHECK(g_api.GroupStart());
    for (int f = 0; f < nf; ++f) {
        const int width = widths[f];
        const size_t cnt = (size_t)width * (size_t)sz;
        double* b = bases[f];
        double* top_owned = b + sz * (long long)(GH + s->nzl - width);  // last `width` planes
        double* bot_owned = b + sz * (long long)GH;                     // first `width` planes
        double* ghost_lo = b + sz * (long long)(GH - width);
        double* ghost_hi = b + sz * (long long)(GH + s->nzl);
        if (up >= 0)
            O3D_NCCL_CHECK(g_api.Send(top_owned, cnt, ncclFloat64_, up, c->nccl, s->st_comm));
        if (dn >= 0)
            O3D_NCCL_CHECK(g_api.Recv(ghost_lo, cnt, ncclFloat64_, dn, c->nccl, s->st_comm));
        if (dn >= 0)
            O3D_NCCL_CHECK(g_api.Send(bot_owned, cnt, ncclFloat64_, dn, c->nccl, s->st_comm));
        if (up >= 0)
            O3D_NCCL_CHECK(g_api.Recv(ghost_hi, cnt, ncclFloat64_, up, c->nccl, s->st_comm));
    }
    O3D_NCCL_CHECK(g_api.GroupEnd());
    trace_mark(s, 1, "exchange end");
    O3D_CUDA_CHECK(cudaEventRecord(s->ev_halo, s->st_comm));
    s->halo_pending = 1;
    s->t_cnt[ST_HALO] += 1;
    return O3D_OK;
}

int comm_wait(o3d_session* s) {
    if (!s->comm || !s->halo_pending) return O3D_OK;
    O3D_CUDA_CHECK(cudaStreamWaitEvent(s->st, s->ev_halo, 0));
    s->halo_pending = 0;
    return O3D_OK;
}

// blocking form (in stream order): exchange, then the session stream waits for the ghosts
int comm_exchange(o3d_session* s, double* const* bases, int nf, int width, int wrap) {
    if (!s->comm) return O3D_OK;
    int widths[8];
    if (nf > 8) return O3D_ERR_INVALID;
    for (int f = 0; f < nf; ++f) widths[f] = width;
    span_begin(s, ST_HALO);
    int rc = comm_exchange_async(s, bases, widths, nf, wrap);
    if (!rc) rc = comm_wait(s);
    span_end(s, ST_HALO, 0);
    return rc;
}

// Replicate a buffer whose consecutive chunks were produced by different ranks: rank r owns
// elements [first[r], first[r] + count[r]); after the call every rank holds all of them.  One
// grouped set of in-place broadcasts on the session stream (multigrid: coarse right-hand sides).
int comm_allgather_chunks(o3d_session* s, double* buf, const long long* first,
                          const long long* count) {
    Comm* c = s->comm;
    if (!c) return O3D_OK;
    O3D_NCCL_CHECK(g_api.GroupStart());
    for (int r = 0; r < c->nranks; ++r) {
        if (count[r] <= 0) continue;
        O3D_NCCL_CHECK(g_api.Broadcast(buf + first[r], buf + first[r], (size_t)count[r],
                                       ncclFloat64_, r, c->nccl, s->st));
    }
    O3D_NCCL_CHECK(g_api.GroupEnd());
    return O3D_OK;
}

int split_edge(const o3d_session* s) {
    const int EDGE = 8;
    if (!s->comm || s->nzl < 4 * EDGE) return 0;
    return EDGE;
}

int comm_allreduce(o3d_session* s, double* dev, int n, int op) {
    Comm* c = s->comm;
    if (!c) return O3D_OK;
    // RED_MAXBITS: max of non-negative doubles through their bit patterns (uint64 max), which is
    // what the SOR residual accumulator holds
    const int dt = (op == RED_MAXBITS) ? ncclUint64_ : ncclFloat64_;
    const int ro = (op == RED_MAXBITS || op == RED_MAX || op == RED_ABSMAX)
                       ? ncclMax_
                       : (op == RED_MIN ? ncclMin_ : ncclSum_);
    trace_mark(s, 0, "allreduce begin");
    O3D_NCCL_CHECK(g_api.AllReduce(dev, dev, (size_t)n, dt, ro, c->nccl, s->st));
    trace_mark(s, 0, "allreduce end");
    return O3D_OK;
}

}  // namespace o3d
