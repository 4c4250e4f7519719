// comm.cu -- z-slab domain decomposition plumbing: ghost-plane exchange with grouped
// ncclSend/ncclRecv over NVLink and small all-reduces (SOR residual, scalar-clipping sums,
// statistics).  The reference has no parallelism at all (README.md:61); this is new.
//
// NCCL is loaded lazily with dlopen so that libo3d_b200.so itself has no link-time dependency
// on it (single-GPU users and the CPU-side "does it load" check never touch NCCL).  Inside a
// torch process the already-loaded bundled libnccl.so.2 is reused.
#include <dlfcn.h>

#include <cstring>

#include "session.h"

namespace o3d {

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess_ = 0 };
// ncclDataType_t / ncclRedOp_t values (nccl.h, stable since 2.x)
enum { ncclUint64_ = 5, ncclFloat64_ = 8 };
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
    return O3D_OK;
}

void comm_destroy(o3d_session* s) {
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
