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

// Ghost-plane exchange of `width` planes per side.  Planes are whole padded planes (px * py
// doubles, contiguous); `bases` are field allocation starts.  wrap != 0: the slab ring is
// periodic in z.
int comm_exchange(o3d_session* s, double* const* bases, int nf, int width, int wrap) {
    Comm* c = s->comm;
    if (!c) return O3D_OK;
    const int up = (c->rank + 1 < c->nranks) ? c->rank + 1 : (wrap ? 0 : -1);
    const int dn = (c->rank > 0) ? c->rank - 1 : (wrap ? c->nranks - 1 : -1);
    const long long sz = s->g.sz;
    const size_t cnt = (size_t)width * (size_t)sz;
    span_begin(s, ST_HALO);
    O3D_NCCL_CHECK(g_api.GroupStart());
    for (int f = 0; f < nf; ++f) {
        double* b = bases[f];
        double* top_owned = b + sz * (long long)(GH + s->nzl - width);  // last `width` planes
        double* bot_owned = b + sz * (long long)GH;                     // first `width` planes
        double* ghost_lo = b + sz * (long long)(GH - width);
        double* ghost_hi = b + sz * (long long)(GH + s->nzl);
        if (up >= 0) O3D_NCCL_CHECK(g_api.Send(top_owned, cnt, ncclFloat64_, up, c->nccl, s->st));
        if (dn >= 0) O3D_NCCL_CHECK(g_api.Recv(ghost_lo, cnt, ncclFloat64_, dn, c->nccl, s->st));
        if (dn >= 0) O3D_NCCL_CHECK(g_api.Send(bot_owned, cnt, ncclFloat64_, dn, c->nccl, s->st));
        if (up >= 0) O3D_NCCL_CHECK(g_api.Recv(ghost_hi, cnt, ncclFloat64_, up, c->nccl, s->st));
    }
    O3D_NCCL_CHECK(g_api.GroupEnd());
    span_end(s, ST_HALO, 1);
    return O3D_OK;
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
    O3D_NCCL_CHECK(g_api.AllReduce(dev, dev, (size_t)n, dt, ro, c->nccl, s->st));
    return O3D_OK;
}

}  // namespace o3d
