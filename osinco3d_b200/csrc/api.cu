// api.cu -- the C ABI of libo3d_b200.so (include/o3d_b200.h): session management, the
// device-resident time-step stages and the stateless host-pointer module procedures.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>

#include "march.cuh"
#include "session.h"

// ------------------------------------------------------------------------------------------
// process-wide state
// ------------------------------------------------------------------------------------------
namespace o3d {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

Schemes g_schemes;
int g_sor_order = O3D_SOR_RED_BLACK;
static int g_device = 0;
static int g_device_checked = 0;

int ensure_device() {
    if (g_device_checked == 1) return O3D_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1) {
        (void)cudaGetLastError();
        set_error("no CUDA device available (%s); libo3d_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return O3D_ERR_NO_DEVICE;
    }
    O3D_CUDA_CHECK(cudaSetDevice(g_device));
    g_device_checked = 1;
    return O3D_OK;
}

int poisson_variant_of(int bx1, int bxn, int by1, int byn) {
    // src/initialization.f90:283-301: only x and y flags are looked at
    if (bx1 == 0 && bxn == 0 && by1 == 0 && byn == 0) return 0;
    if (bx1 == 0 && bxn == 0 && by1 == 1 && byn == 1) return 1;
    if (bx1 == 1 && bxn == 1 && by1 == 1 && byn == 1) return 2;
    return -1;
}

int axis_bc(int b1, int bn, int* out) {
    // src/initialization.f90:228-242
    if (b1 == O3D_PERIODIC && bn == O3D_PERIODIC) {
        *out = O3D_PERIODIC;
        return O3D_OK;
    }
    if (b1 == O3D_FREE_SLIP && bn == O3D_FREE_SLIP) {
        *out = O3D_FREE_SLIP;
        return O3D_OK;
    }
    set_error("Unrecognized boundary layer types: %d %d", b1, bn);
    return O3D_ERR_BC;
}

// ---- session helpers --------------------------------------------------------------------
double* field(o3d_session* s, int id) {
    if (id < 0 || id >= O3D_F_COUNT) return nullptr;
    if (!s->base[id]) {
        const size_t n = (size_t)s->felems;
        double* p = nullptr;
        if (cudaMalloc(&p, n * sizeof(double)) != cudaSuccess) {
            (void)cudaGetLastError();
            set_error("cudaMalloc of field %d (%zu bytes) failed", id, n * sizeof(double));
            return nullptr;
        }
        cudaMemsetAsync(p, 0, n * sizeof(double), s->st);
        if (make_field_tmap(&s->tmap[id], p, s->g.px, s->g.py, s->g.nz + 2 * GH, MBX, MBY) ||
            make_field_tmap(&s->tmap_st[id], p, s->g.px, s->g.py, s->g.nz + 2 * GH, MTX, MTY)) {
            cudaFree(p);
            return nullptr;
        }
        s->base[id] = p;
        s->tmap_sor_ok[id] = 0;
        // an all-zero field has valid ghosts for any closure
        s->gaxes[id] = 0x1Fu;
        s->gpar[id] = natural_parity(id);
    }
    return s->base[id] + interior_offset(s->g);
}

FieldRef fref(o3d_session* s, int id) {
    FieldRef r;
    r.p = field(s, id);
    r.tm = &s->tmap[id];
    r.tms = &s->tmap_st[id];
    return r;
}

unsigned natural_parity(int id) {
    switch (id) {
        case O3D_F_UX: case O3D_F_UX_PRED: return 0x1u;
        case O3D_F_UY: case O3D_F_UY_PRED: return 0x2u;
        case O3D_F_UZ: case O3D_F_UZ_PRED: return 0x4u;
        default: return 0u;
    }
}

void touch(o3d_session* s, int id) {
    if (id >= 0 && id < O3D_F_COUNT) s->gaxes[id] = 0u;
}

int ensure_ghosts(o3d_session* s, const int* ids, int n, const unsigned* par, unsigned axes,
                  bool defer) {
    GhostArgs a;
    a.njobs = 0;
    double* xbase[6];
    int nx = 0;
    const bool zhalo = (s->g.bz_lo == BM_HALO || s->g.bz_hi == BM_HALO) && (axes & 0x4u);
    for (int q = 0; q < n; ++q) {
        const int id = ids[q];
        double* p = field(s, id);
        if (!p) return O3D_ERR_CUDA;
        // axes that are missing, or present with the wrong parity
        unsigned need = 0;
        for (int ax = 0; ax < 3; ++ax) {
            const unsigned bit = 1u << ax;
            if (!(axes & bit)) continue;
            const bool have = (s->gaxes[id] & bit) && ((s->gpar[id] & bit) == (par[q] & bit));
            if (!have) need |= bit;
        }
        if (!need) continue;
        unsigned fill = need;
        if (zhalo && (need & 0x4u)) {
            xbase[nx++] = s->base[id];
            // z ghosts on the wall sides were already written by the producer (bit 0x8), or this
            // rank has no wall side at all: nothing to fill locally along z
            const bool wall_ok = (s->gaxes[id] & 0x8u) && ((s->gpar[id] & 0x4u) == (par[q] & 0x4u));
            if (wall_ok || (s->g.bz_lo == BM_HALO && s->g.bz_hi == BM_HALO)) fill &= ~0x4u;
        }
        s->gaxes[id] |= need | ((need & 0x4u) ? 0x8u : 0u);
        s->gpar[id] = (s->gpar[id] & ~need) | (par[q] & need);
        if (fill) {
            GhostJob& jb = a.job[a.njobs++];
            jb.p = p, jb.par = par[q], jb.axes = fill;
        }
        if (a.njobs == 6) {
            if (launch_fill_ghosts(s->st, s->g, a)) {
                set_error("ghost fill launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                return O3D_ERR_CUDA;
            }
            a.njobs = 0;
        }
    }
    if (a.njobs && launch_fill_ghosts(s->st, s->g, a)) return O3D_ERR_CUDA;
    if (nx) {
        const int wrap = s->cfg.nbcz1 == O3D_PERIODIC;
        if (defer) {  // the caller overlaps the exchange with its interior launch
            const int widths[6] = {R, R, R, R, R, R};
            return comm_exchange_async(s, xbase, widths, nx, wrap);
        }
        const int rc = comm_exchange(s, xbase, nx, R, wrap);
        if (rc) return rc;
    }
    return O3D_OK;
}

// three fields, field q only along axis q (divergence operands): one fused launch
int ensure_ghosts_own_axis(o3d_session* s, const int* ids, const unsigned* par, bool defer) {
    GhostArgs a;
    a.njobs = 0;
    double* xbase[1];
    int nxch = 0;
    for (int q = 0; q < 3; ++q) {
        const int id = ids[q];
        double* p = field(s, id);
        if (!p) return O3D_ERR_CUDA;
        const unsigned bit = 1u << q;
        if ((s->gaxes[id] & bit) && ((s->gpar[id] & bit) == (par[q] & bit))) continue;
        bool fill = true;
        if (q == 2 && (s->g.bz_lo == BM_HALO || s->g.bz_hi == BM_HALO)) {
            xbase[nxch++] = s->base[id];
            const bool wall_ok = (s->gaxes[id] & 0x8u) && ((s->gpar[id] & bit) == (par[q] & bit));
            if (wall_ok || (s->g.bz_lo == BM_HALO && s->g.bz_hi == BM_HALO)) fill = false;
        }
        if (fill) {
            GhostJob& jb = a.job[a.njobs++];
            jb.p = p, jb.par = par[q], jb.axes = bit;
        }
        s->gaxes[id] |= bit | (q == 2 ? 0x8u : 0u);
        s->gpar[id] = (s->gpar[id] & ~bit) | (par[q] & bit);
    }
    if (a.njobs && launch_fill_ghosts(s->st, s->g, a)) return O3D_ERR_CUDA;
    if (nxch) {
        const int wrap = s->cfg.nbcz1 == O3D_PERIODIC;
        if (defer) {
            const int widths[1] = {R};
            return comm_exchange_async(s, xbase, widths, nxch, wrap);
        }
        return comm_exchange(s, xbase, nxch, R, wrap);
    }
    return O3D_OK;
}

int ensure_local_ghosts(o3d_session* s, int id, unsigned par, bool edges) {
    double* p = field(s, id);
    if (!p) return O3D_ERR_CUDA;
    const unsigned want = 0x1u | 0x2u | 0x8u | (edges ? 0x10u : 0u);
    const bool par_ok = ((s->gpar[id] ^ par) & 0x7u) == 0;
    if ((s->gaxes[id] & want) == want && par_ok) return O3D_OK;
    if (launch_fill_ghosts_full(s->st, s->g, p, par)) return O3D_ERR_CUDA;
    const bool zhalo = (s->g.bz_lo == BM_HALO || s->g.bz_hi == BM_HALO);
    s->gaxes[id] = 0x1u | 0x2u | 0x8u | 0x10u | (zhalo ? 0u : 0x4u);
    s->gpar[id] = par & 0x7u;
    return O3D_OK;
}

const CUtensorMap* sor_tmap(o3d_session* s, int id) {
    if (!field(s, id)) return nullptr;
    if (!s->tmap_sor_ok[id]) {
        if (make_field_tmap(&s->tmap_sor[id], s->base[id], s->g.px, s->g.py, s->g.nz + 2 * GH,
                            sor_tma_box_x(), sor_tma_box_y()))
            return nullptr;
        s->tmap_sor_ok[id] = 1;
    }
    return &s->tmap_sor[id];
}

void swap_pp(o3d_session* s) {
    std::swap(s->base[O3D_F_PP], s->base[O3D_F_PP2]);
    std::swap(s->tmap[O3D_F_PP], s->tmap[O3D_F_PP2]);
    std::swap(s->tmap_sor[O3D_F_PP], s->tmap_sor[O3D_F_PP2]);
    std::swap(s->tmap_st[O3D_F_PP], s->tmap_st[O3D_F_PP2]);
    std::swap(s->tmap_sor_ok[O3D_F_PP], s->tmap_sor_ok[O3D_F_PP2]);
    s->pp_phys ^= 1;
}

int ensure_ghosts1(o3d_session* s, int id, unsigned par, unsigned axes, bool defer) {
    return ensure_ghosts(s, &id, 1, &par, axes, defer);
}

// Launch a z-marching kernel so that a pending halo exchange overlaps it: the interior planes go
// to the session stream at once; the two boundary chunks are queued on the COMMUNICATION stream
// right behind the NCCL exchange, so they start the moment the ghosts have landed and run
// concurrently with the interior launch; the session stream then joins.  Everything the
// boundary chunks read must have been enqueued before the exchange was issued.
// `launch(stream, zmode, zedge)`.
int launch_overlapped(o3d_session* s, const std::function<int(cudaStream_t, int, int)>& launch) {
    const int edge = split_edge(s);
    if (!s->halo_pending || edge == 0) {
        int rc = comm_wait(s);
        if (rc) return rc;
        return launch(s->st, ZFULL, 0) ? O3D_ERR_CUDA : O3D_OK;
    }
    trace_mark(s, 0, "interior begin");
    if (launch(s->st, ZINTERIOR, edge)) return O3D_ERR_CUDA;
    trace_mark(s, 0, "interior end");
    if (launch(s->st_comm, ZBOUNDARY, edge)) return O3D_ERR_CUDA;
    trace_mark(s, 1, "boundary end");
    O3D_CUDA_CHECK(cudaEventRecord(s->ev_halo, s->st_comm));
    const int rc = comm_wait(s);
    trace_mark(s, 0, "joined");
    return rc;
}

static const int HIST_BASE[4] = {O3D_F_FUX1, O3D_F_FUY1, O3D_F_FUZ1, O3D_F_FPHI1};

int hist_id(const o3d_session* s, int c, int level) { return HIST_BASE[c] + s->lv[c][level - 1]; }

// translate a public field id (logical history levels) to a physical one
int phys_id(const o3d_session* s, int id) {
    for (int c = 0; c < 4; ++c)
        if (id >= HIST_BASE[c] && id < HIST_BASE[c] + 3) return hist_id(s, c, id - HIST_BASE[c] + 1);
    return id;
}

int ensure_partial(o3d_session* s, long long n) {
    if (n <= s->partial_n) return O3D_OK;
    if (s->partial) cudaFree(s->partial);
    s->partial = nullptr;
    s->partial_n = 0;
    O3D_CUDA_CHECK(cudaMalloc(&s->partial, (size_t)n * sizeof(double)));
    s->partial_n = n;
    return O3D_OK;
}

static cudaEvent_t get_event(o3d_session* s) {
    if (!s->free_events.empty()) {
        cudaEvent_t e = s->free_events.back();
        s->free_events.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void trace_mark(o3d_session* s, int comm, const char* name) {
    if (!s->trace_on) return;
    o3d_session::Mark m;
    cudaEventCreate(&m.e);
    m.name = name, m.comm = comm;
    cudaEventRecord(m.e, comm ? s->st_comm : s->st);
    s->marks.push_back(m);
}

static void trace_dump(o3d_session* s) {
    cudaStreamSynchronize(s->st);
    cudaStreamSynchronize(s->st_comm);
    fprintf(stderr, "[o3d trace] rank %d step %d (us since first mark; C = communication stream)\n",
            s->cfg.rank, s->step_count);
    for (auto& m : s->marks) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s->marks[0].e, m.e);
        fprintf(stderr, "[o3d trace] r%d %9.1f %s %s\n", s->cfg.rank, 1e3 * ms, m.comm ? "C" : " ",
                m.name);
    }
    for (auto& m : s->marks) cudaEventDestroy(m.e);
    s->marks.clear();
}

void span_begin(o3d_session* s, int stage) {
    if (!s->timers_on) return;
    o3d_session::Span sp;
    sp.a = get_event(s);
    sp.b = nullptr;
    sp.stage = stage;
    cudaEventRecord(sp.a, s->st);
    s->pending.push_back(sp);
}

void span_end(o3d_session* s, int stage, long long count) {
    s->t_cnt[stage] += count;
    if (!s->timers_on) return;
    for (size_t q = s->pending.size(); q-- > 0;) {
        if (s->pending[q].stage == stage && !s->pending[q].b) {
            s->pending[q].b = get_event(s);
            cudaEventRecord(s->pending[q].b, s->st);
            return;
        }
    }
}

static void resolve_spans(o3d_session* s) {
    if (s->pending.empty()) return;
    cudaStreamSynchronize(s->st);
    for (auto& sp : s->pending) {
        if (sp.b) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) s->t_ms[sp.stage] += ms;
            s->free_events.push_back(sp.b);
        }
        s->free_events.push_back(sp.a);
    }
    s->pending.clear();
}

void poll_flag(o3d_session* s) {
    if (!s->flag_pending) return;
    s->flag_pending = 0;
    if (*s->flag_h) s->diverged = 1;
}

void fill_geom(o3d_session* s) {
    const o3d_config& c = s->cfg;
    Geom& g = s->g;
    g.nx = c.nx, g.ny = c.ny, g.nz = s->nzl;
    g.px = pitch_for(c.nx);
    g.py = c.ny + 2 * GH;
    g.sy = g.px;
    g.sz = (long long)g.px * g.py;
    g.bx = (c.nbcx1 == O3D_PERIODIC) ? BM_WRAP : BM_MIRROR;
    g.by = (c.nbcy1 == O3D_PERIODIC) ? BM_WRAP : BM_MIRROR;
    const int mz = (c.nbcz1 == O3D_PERIODIC) ? BM_WRAP : BM_MIRROR;
    const int nr = c.nranks > 1 ? c.nranks : 1;
    g.bz_lo = (nr > 1 && (c.rank > 0 || mz == BM_WRAP)) ? BM_HALO : mz;
    g.bz_hi = (nr > 1 && (c.rank < nr - 1 || mz == BM_WRAP)) ? BM_HALO : mz;
    g.sim2d = c.sim2d;
    g.gz0 = s->z0;
    g.gnz = c.nz;
    g.zr_lo = g.zr_hi = 0;
    s->felems = field_elems(g);
    s->cx = make_coef(c.dx);
    s->cy = make_coef(c.dy);
    s->cz = make_coef(c.dz);
}

}  // namespace o3d

using namespace o3d;

// ------------------------------------------------------------------------------------------
// misc entry points
// ------------------------------------------------------------------------------------------
extern "C" {

const char* o3d_last_error(void) { return g_err; }
int o3d_abi_version(void) { return O3D_ABI_VERSION; }
int o3d_config_size(void) { return (int)sizeof(o3d_config); }
long long o3d_kernel_launches(void) { return g_launches.load(); }

int o3d_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int o3d_set_device(int device) {
    g_device = device;
    g_device_checked = 0;
    return ensure_device();
}

int o3d_host_alloc(void** ptr, unsigned long long bytes) {
    if (!ptr) return O3D_ERR_INVALID;
    int rc = ensure_device();
    if (rc) return rc;
    O3D_CUDA_CHECK(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault));
    return O3D_OK;
}
int o3d_host_free(void* ptr) {
    O3D_CUDA_CHECK(cudaFreeHost(ptr));
    return O3D_OK;
}
int o3d_host_register(void* ptr, unsigned long long bytes) {
    int rc = ensure_device();
    if (rc) return rc;
    O3D_CUDA_CHECK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
    return O3D_OK;
}
int o3d_host_unregister(void* ptr) {
    O3D_CUDA_CHECK(cudaHostUnregister(ptr));
    return O3D_OK;
}

int o3d_nccl_unique_id(unsigned char* out128) {
    if (!out128) return O3D_ERR_INVALID;
    return nccl_unique_id(out128);
}

int o3d_slab_partition(int nz, int nranks, int rank, int* z0, int* nz_local) {
    if (nz < 1 || nranks < 1 || rank < 0 || rank >= nranks) return O3D_ERR_INVALID;
    const int q = nz / nranks, r = nz % nranks;
    if (nz_local) *nz_local = q + (rank < r ? 1 : 0);
    if (z0) *z0 = rank * q + (rank < r ? rank : r);
    return O3D_OK;
}

int o3d_set_sor_order(int order) {
    if (order != O3D_SOR_RED_BLACK && order != O3D_SOR_LEXI_WAVEFRONT) return O3D_ERR_INVALID;
    g_sor_order = order;
    return O3D_OK;
}

// ------------------------------------------------------------------------------------------
// B. session
// ------------------------------------------------------------------------------------------
int o3d_session_create(const o3d_config* cfg, o3d_session** out) {
    if (!cfg || !out) return O3D_ERR_INVALID;
    *out = nullptr;
    int rc = ensure_device();
    if (rc) return rc;
    if (cfg->nx < 7 || cfg->ny < 7 || cfg->nz < 7) {
        set_error("grid extents must be >= 7 (stencil radius 3): %d %d %d", cfg->nx, cfg->ny,
                  cfg->nz);
        return O3D_ERR_INVALID;
    }
    int b;
    if ((rc = axis_bc(cfg->nbcx1, cfg->nbcxn, &b))) return rc;
    if ((rc = axis_bc(cfg->nbcy1, cfg->nbcyn, &b))) return rc;
    if (cfg->sim2d == 0 && (rc = axis_bc(cfg->nbcz1, cfg->nbczn, &b))) return rc;
    o3d_session* s = new (std::nothrow) o3d_session();
    if (!s) return O3D_ERR_INVALID;
    s->cfg = *cfg;
    const int nr = cfg->nranks > 1 ? cfg->nranks : 1;
    s->cfg.nranks = nr;
    if (nr == 1) s->cfg.rank = 0;
    if (cfg->rank < 0 || s->cfg.rank >= nr) {
        delete s;
        return O3D_ERR_INVALID;
    }
    // contiguous z slabs, remainder planes to the low ranks
    o3d_slab_partition(cfg->nz, nr, s->cfg.rank, &s->z0, &s->nzl);
    if (nr > 1 && s->nzl < 2 * R) {
        set_error("z slab of %d planes is thinner than two stencil halos", s->nzl);
        delete s;
        return O3D_ERR_INVALID;
    }
    s->nloc = (long long)cfg->nx * cfg->ny * s->nzl;
    for (int f = 0; f < O3D_F_COUNT; ++f)
        s->base[f] = nullptr, s->gaxes[f] = 0, s->gpar[f] = 0, s->tmap_sor_ok[f] = 0;
    s->stage_d = nullptr;
    for (int c = 0; c < 4; ++c)
        for (int l = 0; l < 3; ++l) s->lv[c][l] = l;
    s->sor_variant = poisson_variant_of(cfg->nbcx1, cfg->nbcxn, cfg->nbcy1, cfg->nbcyn);
    s->last_iters = 0;
    s->omega = cfg->omega;
    s->partial = nullptr;
    s->partial_n = 0;
    s->comm = nullptr;
    s->mg = nullptr;
    s->io = nullptr;
    s->pipe = nullptr;
    s->timers_on = 0;
    s->use_src = 0;
    for (int q2 = 0; q2 < 6; ++q2) s->t_ms[q2] = 0.0, s->t_cnt[q2] = 0;
    s->st = nullptr, s->st_comm = nullptr;
    s->ev_ready = nullptr, s->ev_halo = nullptr;
    s->halo_pending = 0;
    s->flag_pending = 0, s->diverged = 0;
    s->spec_arm = 0, s->spec_state = 0;
    s->trace_on = 0, s->step_count = 0;
    s->trace_step = getenv("O3D_TRACE") ? atoi(getenv("O3D_TRACE")) : -1;
    s->sw_a = nullptr, s->sw_b = nullptr;
    s->ctrl_d = nullptr, s->ctrl_h = nullptr, s->flag_d = nullptr, s->flag_h = nullptr;
    s->seam_sync_d = nullptr;
    s->persist_sync_d = nullptr, s->ev_ctrl = nullptr;
    s->pp_phys = 0, s->peers = nullptr, s->peers_tried = 0, s->peer_iter_base = 0ull, s->peer_solves = 0ull;
    s->last_sor_path = 0;
    s->scal_d = nullptr, s->scal_h = nullptr;
    fill_geom(s);
    cudaError_t e = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
        // highest priority: the exchange and the boundary chunks behind it are dispatched ahead
        // of the remaining interior CTAs
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        e = cudaStreamCreateWithPriority(&s->st_comm, cudaStreamNonBlocking, hi);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_halo, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&s->ctrl_d, sizeof(SorCtrl));
    if (e == cudaSuccess) e = cudaHostAlloc(&s->ctrl_h, sizeof(SorCtrl), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc(&s->seam_sync_d, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(s->seam_sync_d, 0, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&s->persist_sync_d, 32 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_ctrl, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&s->flag_d, sizeof(int));
    // the NaN / >1000 flag is sticky (only cleared when reported): it must start clean
    if (e == cudaSuccess) e = cudaMemset(s->flag_d, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaHostAlloc(&s->flag_h, sizeof(int), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc(&s->scal_d, 128 * sizeof(double));
    if (e == cudaSuccess) e = cudaHostAlloc(&s->scal_h, 128 * sizeof(double), cudaHostAllocDefault);
    if (e != cudaSuccess) {
        set_error("session allocation failed: %s", cudaGetErrorString(e));
        o3d_session_destroy(s);
        return O3D_ERR_CUDA;
    }
    *s->flag_h = 0;
    if ((rc = comm_create(s))) {
        o3d_session_destroy(s);
        return rc;
    }
    *out = s;
    return O3D_OK;
}

int o3d_session_destroy(o3d_session* s) {
    if (!s) return O3D_OK;
    if (s->st_comm) cudaStreamSynchronize(s->st_comm);
    if (s->st) cudaStreamSynchronize(s->st);
    io_destroy(s);
    comm_destroy(s);
    mg_destroy(s);
    pipe_destroy(s);
    for (int f = 0; f < O3D_F_COUNT; ++f)
        if (s->base[f]) cudaFree(s->base[f]);
    if (s->partial) cudaFree(s->partial);
    if (s->stage_d) cudaFree(s->stage_d);
    if (s->ctrl_d) cudaFree(s->ctrl_d);
    if (s->seam_sync_d) cudaFree(s->seam_sync_d);
    if (s->persist_sync_d) cudaFree(s->persist_sync_d);
    if (s->ev_ctrl) cudaEventDestroy(s->ev_ctrl);
    if (s->ctrl_h) cudaFreeHost(s->ctrl_h);
    if (s->flag_d) cudaFree(s->flag_d);
    if (s->flag_h) cudaFreeHost(s->flag_h);
    if (s->scal_d) cudaFree(s->scal_d);
    if (s->scal_h) cudaFreeHost(s->scal_h);
    for (auto& sp : s->pending) {
        cudaEventDestroy(sp.a);
        if (sp.b) cudaEventDestroy(sp.b);
    }
    for (auto e : s->free_events) cudaEventDestroy(e);
    if (s->sw_a) cudaEventDestroy(s->sw_a);
    if (s->sw_b) cudaEventDestroy(s->sw_b);
    if (s->ev_ready) cudaEventDestroy(s->ev_ready);
    if (s->ev_halo) cudaEventDestroy(s->ev_halo);
    if (s->st_comm) cudaStreamDestroy(s->st_comm);
    if (s->st) cudaStreamDestroy(s->st);
    delete s;
    return O3D_OK;
}

int o3d_session_slab(const o3d_session* s, int* z0, int* nz_local) {
    if (!s) return O3D_ERR_INVALID;
    if (z0) *z0 = s->z0;
    if (nz_local) *nz_local = s->nzl;
    return O3D_OK;
}

// staging buffer of up to STAGE_MAX doubles (512 MB): slabs larger than that move in z chunks, so
// that a 1024^3 session (18 fields x 8.8 GB) still fits one B200
static const long long STAGE_MAX = 64ll << 20;
static int stage_planes(const o3d_session* s) {
    const long long plane = (long long)s->cfg.nx * s->cfg.ny;
    long long n = STAGE_MAX / plane;
    if (n < 1) n = 1;
    if (n > s->nzl) n = s->nzl;
    return (int)n;
}
static int ensure_stage(o3d_session* s) {
    if (s->stage_d) return O3D_OK;
    const size_t elems = (size_t)stage_planes(s) * s->cfg.nx * s->cfg.ny;
    O3D_CUDA_CHECK(cudaMalloc(&s->stage_d, elems * sizeof(double)));
    return O3D_OK;
}

// planes [k0, k0 + nk) of a padded field <- / -> a contiguous (nx, ny, nk) host block
static int copy_planes(o3d_session* s, double* d, double* host, int k0, int nk, bool up) {
    int rc = ensure_stage(s);
    if (rc) return rc;
    const int step = stage_planes(s);
    const size_t plane = (size_t)s->cfg.nx * s->cfg.ny;
    for (int k = 0; k < nk; k += step) {
        const int m = (nk - k < step) ? nk - k : step;
        double* h = host + plane * (size_t)k;
        if (up) {
            O3D_CUDA_CHECK(cudaMemcpyAsync(s->stage_d, h, plane * m * sizeof(double),
                                           cudaMemcpyHostToDevice, s->st));
            if (launch_pack_planes(s->st, s->g, s->stage_d, d, k0 + k, m)) return O3D_ERR_CUDA;
        } else {
            if (launch_unpack_planes(s->st, s->g, d, s->stage_d, k0 + k, m)) return O3D_ERR_CUDA;
            O3D_CUDA_CHECK(cudaMemcpyAsync(h, s->stage_d, plane * m * sizeof(double),
                                           cudaMemcpyDeviceToHost, s->st));
        }
    }
    return O3D_OK;
}

// logical history level 1 must not alias level 2 / 3 when it is written from outside (see
// o3d_upload in include/o3d_b200.h)
static void unalias_level1(o3d_session* s, int fid) {
    for (int c = 0; c < 4; ++c) {
        const int hb = c < 3 ? O3D_F_FUX1 + 3 * c : O3D_F_FPHI1;
        if (fid == hb && (s->lv[c][0] == s->lv[c][1] || s->lv[c][0] == s->lv[c][2])) {
            for (int p = 0; p < 3; ++p)
                if (p != s->lv[c][1] && p != s->lv[c][2]) s->lv[c][0] = p;
        }
    }
}

// Host arrays are the reference's contiguous (nx,ny,nz) allocatables; device fields are padded.
// A copy goes through one contiguous staging buffer so that the PCIe transfer is a single
// full-speed DMA, followed / preceded by a pack kernel (HBM traffic, negligible next to PCIe).
int o3d_upload(o3d_session* s, int fid, const double* host) {
    if (!s || !host) return O3D_ERR_INVALID;
    return o3d_upload_planes(s, fid, host, 0, s->nzl);
}

int o3d_upload_planes(o3d_session* s, int fid, const double* host, int k0, int nk) {
    if (!s || !host || k0 < 0 || nk < 1 || k0 + nk > s->nzl) return O3D_ERR_INVALID;
    // History ids are LOGICAL levels.  After the first rotation levels 1 and 2 share a physical
    // buffer (level 1 is always rewritten by the next predictor before it is read, so the
    // reference's copy fu(:,:,:,2) = fu(:,:,:,1) is a pointer assignment here): an upload to
    // level 1 must not clobber level 2 -> give level 1 the free buffer first.
    unalias_level1(s, fid);
    const int id = phys_id(s, fid);
    double* d = field(s, id);
    if (!d) return O3D_ERR_CUDA;
    const int rc = copy_planes(s, d, const_cast<double*>(host), k0, nk, true);
    if (rc) return rc;
    touch(s, id);
    O3D_CUDA_CHECK(cudaStreamSynchronize(s->st));
    return O3D_OK;
}

int o3d_download_planes(o3d_session* s, int fid, double* host, int k0, int nk) {
    if (!s || !host || k0 < 0 || nk < 1 || k0 + nk > s->nzl) return O3D_ERR_INVALID;
    double* d = field(s, phys_id(s, fid));
    if (!d) return O3D_ERR_CUDA;
    const int rc = copy_planes(s, d, host, k0, nk, false);
    if (rc) return rc;
    O3D_CUDA_CHECK(cudaStreamSynchronize(s->st));
    return O3D_OK;
}

int o3d_download(o3d_session* s, int fid, double* host) {
    if (!s || !host) return O3D_ERR_INVALID;
    double* d = field(s, phys_id(s, fid));
    if (!d) return O3D_ERR_CUDA;
    int rc = copy_planes(s, d, host, 0, s->nzl, false);
    if (rc) return rc;
    O3D_CUDA_CHECK(cudaStreamSynchronize(s->st));
    // o3d_step leaves the NaN / >1000 guard of its correction in flight: a caller that steps and
    // then downloads without o3d_sync must not get a diverged state silently (the data is still
    // delivered, as the stateless o3d_correct_velocity does)
    poll_flag(s);
    if (s->diverged) {
        s->diverged = 0;
        cudaMemsetAsync(s->flag_d, 0, sizeof(int), s->st);
        set_error("velocity diverged: NaN or max(u) > 1000 (src/integration.f90:309-325)");
        return O3D_ERR_DIVERGED;
    }
    return O3D_OK;
}

int o3d_device_ptr(o3d_session* s, int fid, double** dptr) {
    if (!s || !dptr) return O3D_ERR_INVALID;
    *dptr = field(s, phys_id(s, fid));
    if (!*dptr) return O3D_ERR_CUDA;
    // make the lazy zero-fill visible to other streams
    O3D_CUDA_CHECK(cudaStreamSynchronize(s->st));
    return O3D_OK;
}

int o3d_session_layout(const o3d_session* s, long long* stride_j, long long* stride_k) {
    if (!s) return O3D_ERR_INVALID;
    if (stride_j) *stride_j = s->g.sy;
    if (stride_k) *stride_k = s->g.sz;
    return O3D_OK;
}

int o3d_mark_modified(o3d_session* s, int fid) {
    if (!s) return O3D_ERR_INVALID;
    touch(s, phys_id(s, fid));
    return O3D_OK;
}

int o3d_sync(o3d_session* s) {
    if (!s) return O3D_ERR_INVALID;
    O3D_CUDA_CHECK(cudaStreamSynchronize(s->st));
    poll_flag(s);
    if (s->diverged) {
        s->diverged = 0;
        cudaMemsetAsync(s->flag_d, 0, sizeof(int), s->st);
        set_error("velocity diverged: NaN or max(u) > 1000 (src/integration.f90:309-325)");
        return O3D_ERR_DIVERGED;
    }
    return O3D_OK;
}

int o3d_s_sor_path(const o3d_session* s, int* persistent, int* peer) {
    if (!s) return O3D_ERR_INVALID;
    if (persistent) *persistent = s->last_sor_path & 1;
    if (peer) *peer = (s->last_sor_path >> 1) & 1;
    return O3D_OK;
}

int o3d_get_omega(const o3d_session* s, double* omega) {
    if (!s || !omega) return O3D_ERR_INVALID;
    *omega = s->omega;
    return O3D_OK;
}
int o3d_set_omega(o3d_session* s, double omega) {
    if (!s) return O3D_ERR_INVALID;
    s->omega = omega;
    return O3D_OK;
}

int o3d_session_set_poisson(o3d_session* s, double eps, int kmax, int idyn, int multigrid) {
    if (!s) return O3D_ERR_INVALID;
    s->cfg.eps = eps, s->cfg.kmax = kmax, s->cfg.idyn = idyn, s->cfg.multigrid = multigrid;
    return O3D_OK;
}

int o3d_s_enable_timers(o3d_session* s, int on) {
    if (!s) return O3D_ERR_INVALID;
    resolve_spans(s);
    s->timers_on = on;
    return O3D_OK;
}

int o3d_s_timers(o3d_session* s, double* ms6, long long* counts6, int reset) {
    if (!s) return O3D_ERR_INVALID;
    resolve_spans(s);
    for (int q = 0; q < 6; ++q) {
        if (ms6) ms6[q] = s->t_ms[q];
        if (counts6) counts6[q] = s->t_cnt[q];
        if (reset) s->t_ms[q] = 0.0, s->t_cnt[q] = 0;
    }
    return O3D_OK;
}

int o3d_s_stopwatch_start(o3d_session* s) {
    if (!s) return O3D_ERR_INVALID;
    if (!s->sw_a) O3D_CUDA_CHECK(cudaEventCreate(&s->sw_a));
    if (!s->sw_b) O3D_CUDA_CHECK(cudaEventCreate(&s->sw_b));
    O3D_CUDA_CHECK(cudaEventRecord(s->sw_a, s->st));
    return O3D_OK;
}

int o3d_s_stopwatch_stop(o3d_session* s, double* ms) {
    if (!s || !ms || !s->sw_a) return O3D_ERR_INVALID;
    O3D_CUDA_CHECK(cudaEventRecord(s->sw_b, s->st));
    O3D_CUDA_CHECK(cudaEventSynchronize(s->sw_b));
    float t = 0.f;
    O3D_CUDA_CHECK(cudaEventElapsedTime(&t, s->sw_a, s->sw_b));
    *ms = t;
    return O3D_OK;
}

// ---- stages ------------------------------------------------------------------------------
static int ab_select(const o3d_config& c, int itime, double* adu, double* bdu, double* cdu) {
    // src/integration.f90:84-105
    if (c.itscheme == 1 || itime == 1) {
        *adu = c.adt[0], *bdu = c.bdt[0], *cdu = c.cdt[0];
    } else if (c.itscheme == 2 || itime == 2) {
        *adu = c.adt[1], *bdu = c.bdt[1], *cdu = c.cdt[1];
    } else if (c.itscheme == 3) {
        *adu = c.adt[2], *bdu = c.bdt[2], *cdu = c.cdt[2];
    } else {
        set_error("itscheme: %d unrecognized", c.itscheme);
        return O3D_ERR_ITSCHEME;
    }
    return O3D_OK;
}

// choose the physical buffer that receives the new f and rotate the logical levels the way the
// reference copies them (src/integration.f90:176-188, :453-459)
static int hist_target(o3d_session* s, int c) {
    const int itscheme = s->cfg.itscheme;
    if (itscheme == 3) return s->lv[c][2];
    for (int p = 0; p < 3; ++p)
        if (p != s->lv[c][1] && p != s->lv[c][2]) return p;
    return 0;
}
static void hist_rotate(o3d_session* s, int c, int target) {
    const int itscheme = s->cfg.itscheme;
    if (itscheme == 3) {
        const int old2 = s->lv[c][1];
        s->lv[c][0] = target, s->lv[c][1] = target, s->lv[c][2] = old2;
    } else if (itscheme == 2) {
        s->lv[c][0] = target, s->lv[c][1] = target;
    } else {
        s->lv[c][0] = target;
    }
}

static const int VEL_IDS[3] = {O3D_F_UX, O3D_F_UY, O3D_F_UZ};
static const int PRED_IDS[3] = {O3D_F_UX_PRED, O3D_F_UY_PRED, O3D_F_UZ_PRED};
static const unsigned NAT3[3] = {0x1u, 0x2u, 0x4u};
static const unsigned EVEN3[3] = {0u, 0u, 0u};

extern "C++" {
namespace o3d {
// arguments of the fused RHS + predictor launch for time step `itime` (tgt = physical history
// buffer that receives the new f of each component) ...
int rhs_prepare(o3d_session* s, int itime, RhsArgs& a, int* tgt) {
    const o3d_config& c = s->cfg;
    double adu, bdu, cdu;
    int rc = ab_select(c, itime, &adu, &bdu, &cdu);
    if (rc) return rc;
    for (int k = 0; k < 3; ++k) {
        a.u[k] = fref(s, VEL_IDS[k]);
        tgt[k] = hist_target(s, k);
        a.f2[k] = field(s, HIST_BASE[k] + s->lv[k][1]);
        a.f3[k] = field(s, HIST_BASE[k] + s->lv[k][2]);
        a.f1[k] = field(s, HIST_BASE[k] + tgt[k]);
        a.up[k] = field(s, PRED_IDS[k]);
        if (!a.u[k].p || !a.f2[k] || !a.f3[k] || !a.f1[k] || !a.up[k]) return O3D_ERR_CUDA;
    }
    // nu_t is written only when iles == 1 (src/integration.f90:108-112): a DNS session does not
    // pay a field for it (a 1024^3 DNS state is 18 fields on one B200)
    a.nu_t = (c.iles == 1) ? field(s, O3D_F_NU_T) : nullptr;
    if (c.iles == 1 && !a.nu_t) return O3D_ERR_CUDA;
    a.cx = s->cx, a.cy = s->cy, a.cz = s->cz;
    a.onere = 1.0 / c.re;  // src/integration.f90:106
    a.adu = adu, a.bdu = bdu, a.cdu = cdu;
    const double csd = c.cs * c.delta;
    a.csd2 = csd * csd;  // (cs*delta)**2, src/les_turbulence.f90:87
    a.iles = (c.iles == 1);
    return O3D_OK;
}

// ... and the bookkeeping after it: history rotation, ghost state of u* and nu_t
void rhs_finish(o3d_session* s, const int* tgt, bool iles) {
    const bool zhalo = (s->g.bz_lo == BM_HALO || s->g.bz_hi == BM_HALO);
    for (int k = 0; k < 3; ++k) {
        hist_rotate(s, k, tgt[k]);
        // the RHS kernel wrote the own-axis ghost images of u* (odd closure) with the interior
        // (k == 2: z images on the wall sides are in place, bit 0x8; a rank boundary still needs
        // the exchange)
        s->gaxes[PRED_IDS[k]] = (k == 2) ? (zhalo ? 0x8u : 0xCu) : (1u << k);
        s->gpar[PRED_IDS[k]] = NAT3[k];
    }
    if (iles) touch(s, O3D_F_NU_T);
}
}  // namespace o3d
}  // extern "C++"

int o3d_s_predict_velocity(o3d_session* s, int itime) {
    if (!s) return O3D_ERR_INVALID;
    RhsArgs a;
    int tgt[3];
    int rc = rhs_prepare(s, itime, a, tgt);
    if (rc) return rc;
    // parity table of src/integration.f90:118-165 = natural-parity ghosts of ux, uy, uz
    if ((rc = ensure_ghosts(s, VEL_IDS, 3, NAT3, 0x7u, true))) return rc;
    span_begin(s, ST_RHS);
    rc = launch_overlapped(
        s, [&](cudaStream_t q, int zm, int ze) { return launch_rhs(q, s->g, a, zm, ze); });
    if (rc) {
        set_error("rhs kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return rc;
    }
    span_end(s, ST_RHS, 1);
    rhs_finish(s, tgt, a.iles != 0);
    return O3D_OK;
}

int o3d_s_correct_pression(o3d_session* s, int* iters, double* dmax) {
    if (!s) return O3D_ERR_INVALID;
    const o3d_config& c = s->cfg;
    if (c.multigrid != 1 && s->sor_variant < 0) {
        set_error("poisson_solver pointer is null for these boundary flags "
                  "(src/initialization.f90:283-301)");
        return O3D_ERR_BC;
    }
    FieldRef up[3] = {fref(s, O3D_F_UX_PRED), fref(s, O3D_F_UY_PRED), fref(s, O3D_F_UZ_PRED)};
    double* rhs = field(s, O3D_F_RHS);
    double* pp = field(s, O3D_F_PP);
    if (!up[0].p || !up[1].p || !up[2].p || !rhs || !pp) return O3D_ERR_CUDA;
    int rc;
    // divergence(..., odd = 1): derxi(ux*), deryi(uy*), derzi(uz*); each field only needs the
    // ghosts of its own axis (src/differential_operators.f90:30-32)
    if ((rc = ensure_ghosts_own_axis(s, PRED_IDS, NAT3, true))) return rc;
    span_begin(s, ST_DIV);
    rc = launch_overlapped(s, [&](cudaStream_t q, int zm, int ze) {
        return launch_div(q, s->g, up, s->cx, s->cy, s->cz, 1, c.dt, rhs, zm, ze);
    });
    if (rc) return rc;
    span_end(s, ST_DIV, 1);
    {   // the divergence kernel wrote the even ghost images of rhs with the interior
        const bool zhalo = (s->g.bz_lo == BM_HALO || s->g.bz_hi == BM_HALO);
        s->gaxes[O3D_F_RHS] = zhalo ? 0xBu : 0xFu;
        s->gpar[O3D_F_RHS] = 0u;
    }
    if (c.multigrid == 1) {
        int cycles = 0;
        rc = mg_solve(s, pp, rhs, c.kmax, 5, 4, c.eps, &cycles, dmax);  // src/integration.f90:244
        if (iters) *iters = cycles;
    } else {
        rc = sor_solve(s, pp, rhs, iters, dmax);
    }
    return rc;  // the solvers record the ghost state of pp themselves
}

static int correct_velocity_impl(o3d_session* s, bool defer) {
    if (!s) return O3D_ERR_INVALID;
    const o3d_config& c = s->cfg;
    FieldRef up[3] = {fref(s, O3D_F_UX_PRED), fref(s, O3D_F_UY_PRED), fref(s, O3D_F_UZ_PRED)};
    double* u[3] = {field(s, O3D_F_UX), field(s, O3D_F_UY), field(s, O3D_F_UZ)};
    FieldRef pp = fref(s, O3D_F_PP);
    for (int k = 0; k < 3; ++k)
        if (!up[k].p || !u[k]) return O3D_ERR_CUDA;
    if (!pp.p) return O3D_ERR_CUDA;
    int rc;
    // (before the exchange is issued: the boundary chunks run on the communication stream)
    O3D_CUDA_CHECK(cudaMemsetAsync(s->flag_d, 0, sizeof(int), s->st));
    if ((rc = ensure_ghosts1(s, O3D_F_PP, 0u, 0x7u, true))) return rc;  // derxp/deryp/derzp
    span_begin(s, ST_CORR);
    rc = launch_overlapped(s, [&](cudaStream_t q, int zm, int ze) {
        return launch_corr(q, s->g, pp, up, u, s->cx, s->cy, s->cz, c.dt, s->flag_d, zm, ze);
    });
    if (rc) return rc;
    span_end(s, ST_CORR, 1);
    {   // the correction kernel wrote the natural-parity ghost images of u with the interior
        const bool zhalo = (s->g.bz_lo == BM_HALO || s->g.bz_hi == BM_HALO);
        for (int k = 0; k < 3; ++k) {
            s->gaxes[VEL_IDS[k]] = zhalo ? 0xBu : 0xFu;
            s->gpar[VEL_IDS[k]] = NAT3[k];
        }
    }
    O3D_CUDA_CHECK(
        cudaMemcpyAsync(s->flag_h, s->flag_d, sizeof(int), cudaMemcpyDeviceToHost, s->st));
    if (defer) {
        s->flag_pending = 1;
        return O3D_OK;
    }
    O3D_CUDA_CHECK(cudaStreamSynchronize(s->st));
    s->flag_pending = 0;
    if (*s->flag_h) {
        set_error("velocity diverged: NaN or max(u) > 1000 (src/integration.f90:309-325)");
        return O3D_ERR_DIVERGED;
    }
    return O3D_OK;
}

int o3d_s_correct_velocity(o3d_session* s) { return correct_velocity_impl(s, false); }

static void corr_wrote_ghosts(o3d_session* s) {
    // the correction kernel wrote the natural-parity ghost images of u with the interior
    const bool zhalo = (s->g.bz_lo == BM_HALO || s->g.bz_hi == BM_HALO);
    for (int k = 0; k < 3; ++k) {
        s->gaxes[VEL_IDS[k]] = zhalo ? 0xBu : 0xFu;
        s->gpar[VEL_IDS[k]] = NAT3[k];
    }
}

extern "C++" {
namespace o3d {
// Called by sor_solve with the first batch of passes queued and before its host poll: the same
// correction launch as correct_velocity_impl, but gated on the solver's control block and reading
// whichever ping-pong buffer holds the final iterate.  The NaN / >1000 flag is NOT cleared here
// (it is sticky until reported), so the copy below also carries the previous step's verdict.
int spec_correct_launch(o3d_session* s) {
    const o3d_config& c = s->cfg;
    FieldRef up[3] = {fref(s, O3D_F_UX_PRED), fref(s, O3D_F_UY_PRED), fref(s, O3D_F_UZ_PRED)};
    double* u[3] = {field(s, O3D_F_UX), field(s, O3D_F_UY), field(s, O3D_F_UZ)};
    FieldRef pp = fref(s, O3D_F_PP), alt = fref(s, O3D_F_PP2);
    for (int k = 0; k < 3; ++k)
        if (!up[k].p || !u[k]) return O3D_ERR_CUDA;
    if (!pp.p || !alt.p) return O3D_ERR_CUDA;
    if (c.nranks > 1 && !(s->last_sor_path & 2)) {
        // z slabs: the correction differentiates pp across the rank boundaries (3 ghost planes).
        // A peer-memory solve has stored them into the neighbours' ghost planes already (both
        // ping-pong buffers, every iteration); otherwise they travel now -- which buffer holds the
        // final iterate is only known on the device, so the planes of BOTH go (one grouped
        // exchange, 6 planes per side) and the gated kernel picks.
        const long long ioff = interior_offset(s->g);
        double* bases[2] = {pp.p - ioff, alt.p - ioff};
        const int widths[2] = {R, R};
        const int rc = comm_exchange_async(s, bases, widths, 2, c.nbcz1 == O3D_PERIODIC);
        if (rc) return rc;
    }
    span_begin(s, ST_CORR);
    {
        const int rc = launch_overlapped(s, [&](cudaStream_t q, int zm, int ze) {
            return launch_corr(q, s->g, pp, up, u, s->cx, s->cy, s->cz, c.dt, s->flag_d, zm, ze,
                               &alt, s->ctrl_d);
        });
        if (rc) return rc;
    }
    span_end(s, ST_CORR, 0);
    O3D_CUDA_CHECK(
        cudaMemcpyAsync(s->flag_h, s->flag_d, sizeof(int), cudaMemcpyDeviceToHost, s->st));
    s->flag_pending = 1;
    return O3D_OK;
}
}  // namespace o3d
}  // extern "C++"

int o3d_s_transeq(o3d_session* s, int itime) {
    if (!s) return O3D_ERR_INVALID;
    const o3d_config& c = s->cfg;
    double adu, bdu, cdu;
    int rc = ab_select(c, itime, &adu, &bdu, &cdu);
    if (rc) return rc;
    TranseqArgs a;
    double* phi = field(s, O3D_F_PHI);
    double* phi_new = field(s, O3D_F_SCRATCH0);
    const int tgt = hist_target(s, 3);
    a.phi = fref(s, O3D_F_PHI);
    a.phi_new = phi_new;
    for (int k = 0; k < 3; ++k) a.u[k] = field(s, O3D_F_UX + k);
    a.nu_t = field(s, O3D_F_NU_T);
    // src is never assigned in the reference (src/initialization.f90:159): NULL == 0
    a.src = s->use_src ? field(s, O3D_F_SCRATCH1) : nullptr;
    a.f2 = field(s, HIST_BASE[3] + s->lv[3][1]);
    a.f3 = field(s, HIST_BASE[3] + s->lv[3][2]);
    a.f1 = field(s, HIST_BASE[3] + tgt);
    if (!phi || !phi_new || !a.u[0] || !a.u[1] || !a.u[2] || !a.nu_t || !a.f2 || !a.f3 || !a.f1)
        return O3D_ERR_CUDA;
    const int nb = transeq_blocks(s->g);
    if ((rc = ensure_partial(s, 3ll * nb))) return rc;
    a.partial = s->partial;
    a.cx = s->cx, a.cy = s->cy, a.cz = s->cz;
    a.resc = c.re * c.sc;
    a.sc = c.sc;
    a.adu = adu, a.bdu = bdu, a.cdu = cdu;
    a.iles = (c.iles == 1);
    if ((rc = ensure_ghosts1(s, O3D_F_PHI, 0u, 0x7u))) return rc;  // all "p" closures, :412-419
    span_begin(s, ST_TRANSEQ);
    if (launch_transeq_rhs(s->st, s->g, a)) return O3D_ERR_CUDA;
    if (launch_sum_partials(s->st, s->partial, nb, 3, s->scal_d)) return O3D_ERR_CUDA;
    if (c.nranks > 1 && (rc = comm_allreduce(s, s->scal_d, 3, RED_SUM))) return rc;
    const double count = (double)((long long)c.nx * c.ny * c.nz);
    if (launch_transeq_clip(s->st, s->g, phi_new, phi, s->scal_d, count)) return O3D_ERR_CUDA;
    span_end(s, ST_TRANSEQ, 1);
    hist_rotate(s, 3, tgt);
    touch(s, O3D_F_PHI);
    touch(s, O3D_F_SCRATCH0);
    return O3D_OK;
}

int o3d_step(o3d_session* s, int itime, int* iters, double* dmax) {
    // src/osinco3d_main.f90:105-115
    int rc;
    s->trace_on = (s->step_count == s->trace_step);
    trace_mark(s, 0, "step begin");
    if ((rc = o3d_s_predict_velocity(s, itime))) return rc;
    trace_mark(s, 0, "predict done");
    {
        const char* e = getenv("O3D_SPEC");
        s->spec_arm = !(e && e[0] == '0');
    }
    rc = o3d_s_correct_pression(s, iters, dmax);
    s->spec_arm = 0;
    if (rc) return rc;
    trace_mark(s, 0, "pression done");
    // the guard of the PREVIOUS step's correction (and of this step's, if it was queued behind the
    // SOR passes) was examined at the Poisson solver's host poll
    if (s->diverged) {
        s->diverged = 0;
        cudaMemsetAsync(s->flag_d, 0, sizeof(int), s->st);
        set_error("velocity diverged: NaN or max(u) > 1000 (src/integration.f90:309-325)");
        return O3D_ERR_DIVERGED;
    }
    if (s->spec_state == 2) {
        // the correction already ran on the device, gated on the solver's exit
        s->t_cnt[ST_CORR] += 1;
        corr_wrote_ghosts(s);
    } else if ((rc = correct_velocity_impl(s, true))) {
        return rc;
    }
    s->spec_state = 0;
    trace_mark(s, 0, "correct done");
    if (s->cfg.nscr == 1 && (rc = o3d_s_transeq(s, itime))) return rc;
    if (s->trace_on) trace_dump(s);
    s->trace_on = 0;
    s->step_count++;
    return O3D_OK;
}

// ---- diagnostics -------------------------------------------------------------------------
int o3d_s_divergence(o3d_session* s, int fx, int fy, int fz, int dst, int odd) {
    if (!s) return O3D_ERR_INVALID;
    const int ids[3] = {phys_id(s, fx), phys_id(s, fy), phys_id(s, fz)};
    const int out_id = phys_id(s, dst);
    FieldRef f[3] = {fref(s, ids[0]), fref(s, ids[1]), fref(s, ids[2])};
    double* out = field(s, out_id);
    if (!f[0].p || !f[1].p || !f[2].p || !out) return O3D_ERR_CUDA;
    int rc;
    // src/differential_operators.f90:25-33: odd -> derxi/deryi/derzi, else derxp/deryp/derzp
    const unsigned dpar[3] = {odd ? 0x1u : 0u, odd ? 0x2u : 0u, odd ? 0x4u : 0u};
    if ((rc = ensure_ghosts_own_axis(s, ids, dpar))) return rc;
    if (launch_div(s->st, s->g, f, s->cx, s->cy, s->cz, 0, 1.0, out)) return O3D_ERR_CUDA;
    {
        const bool zhalo = (s->g.bz_lo == BM_HALO || s->g.bz_hi == BM_HALO);
        s->gaxes[out_id] = zhalo ? 0xBu : 0xFu;
        s->gpar[out_id] = 0u;
    }
    return O3D_OK;
}

int o3d_s_reduce(o3d_session* s, int fid, int op, double* out) {
    if (!s || !out || op < 0 || op > 3) return O3D_ERR_INVALID;
    double* f = field(s, phys_id(s, fid));
    if (!f) return O3D_ERR_CUDA;
    int rc;
    if ((rc = ensure_partial(s, reduce_blocks(s->g)))) return rc;
    if (launch_reduce(s->st, s->g, f, op, s->partial, s->scal_d + 8)) return O3D_ERR_CUDA;
    if (s->cfg.nranks > 1 && (rc = comm_allreduce(s, s->scal_d + 8, 1, op))) return rc;
    O3D_CUDA_CHECK(cudaMemcpyAsync(s->scal_h + 8, s->scal_d + 8, sizeof(double),
                                   cudaMemcpyDeviceToHost, s->st));
    O3D_CUDA_CHECK(cudaStreamSynchronize(s->st));
    *out = s->scal_h[8];
    return O3D_OK;
}

int o3d_s_function_stats(o3d_session* s, int fid, double* stats6) {
    if (!s || !stats6) return O3D_ERR_INVALID;
    if (s->cfg.nranks > 1) {
        set_error("function_stats with argmax is single-rank; use o3d_s_reduce per rank");
        return O3D_ERR_UNSUPPORTED;
    }
    double* f = field(s, phys_id(s, fid));
    if (!f) return O3D_ERR_CUDA;
    int rc;
    if ((rc = ensure_partial(s, 4096))) return rc;
    if (launch_function_stats(s->st, s->g, f, s->partial, s->scal_d + 16)) return O3D_ERR_CUDA;
    O3D_CUDA_CHECK(cudaMemcpyAsync(s->scal_h + 16, s->scal_d + 16, 6 * sizeof(double),
                                   cudaMemcpyDeviceToHost, s->st));
    O3D_CUDA_CHECK(cudaStreamSynchronize(s->st));
    for (int q = 0; q < 6; ++q) stats6[q] = s->scal_h[16 + q];
    return O3D_OK;
}

int o3d_s_statistics(o3d_session* s, double t, double* out17) {
    if (!s || !out17) return O3D_ERR_INVALID;
    FieldRef u[3] = {fref(s, O3D_F_UX), fref(s, O3D_F_UY), fref(s, O3D_F_UZ)};
    if (!u[0].p || !u[1].p || !u[2].p) return O3D_ERR_CUDA;
    const int nb = stats_blocks(s->g);
    int rc;
    if ((rc = ensure_partial(s, 16ll * nb))) return rc;
    // same parity table as calculate_nu_t (src/utils.f90:283-291,339-347)
    if ((rc = ensure_ghosts(s, VEL_IDS, 3, NAT3, 0x7u))) return rc;
    if (launch_stats(s->st, s->g, u, s->cx, s->cy, s->cz, 1.0 / s->cfg.re, s->partial))
        return O3D_ERR_CUDA;
    if (launch_sum_partials(s->st, s->partial, nb, 16, s->scal_d + 32)) return O3D_ERR_CUDA;
    if (s->cfg.nranks > 1 && (rc = comm_allreduce(s, s->scal_d + 32, 16, RED_SUM))) return rc;
    O3D_CUDA_CHECK(cudaMemcpyAsync(s->scal_h + 32, s->scal_d + 32, 16 * sizeof(double),
                                   cudaMemcpyDeviceToHost, s->st));
    O3D_CUDA_CHECK(cudaStreamSynchronize(s->st));
    const double cnt = (double)((long long)s->cfg.nx * s->cfg.ny * s->cfg.nz);
    const double* a = s->scal_h + 32;
    // stats.dat columns, src/IOfunctions.f90:504-552: t, e_k, eps, eps2, dzeta, u2,v2,w2, 9 x d1^2
    out17[0] = t;
    out17[1] = a[0] / cnt;
    out17[2] = a[1] / cnt;
    out17[3] = a[2] / cnt;
    out17[4] = a[3] / cnt;
    for (int q = 0; q < 12; ++q) out17[5 + q] = a[4 + q] / cnt;
    return O3D_OK;
}

int o3d_s_step_diagnostics(o3d_session* s, double* out23) {
    if (!s || !out23) return O3D_ERR_INVALID;
    const o3d_config& c = s->cfg;
    int rc;
    const int nb = diag_blocks(s->g);
    if ((rc = ensure_partial(s, 13ll * nb))) return rc;
    // divergence(..., 1) of u* and of u: own-axis odd ghosts (src/differential_operators.f90:30-32)
    FieldRef up[3] = {fref(s, O3D_F_UX_PRED), fref(s, O3D_F_UY_PRED), fref(s, O3D_F_UZ_PRED)};
    FieldRef u[3] = {fref(s, O3D_F_UX), fref(s, O3D_F_UY), fref(s, O3D_F_UZ)};
    for (int q = 0; q < 3; ++q)
        if (!up[q].p || !u[q].p) return O3D_ERR_CUDA;
    double* d0 = s->scal_d + 64;  // 13 values of u*, 13 of u, then phi min / max
    if ((rc = ensure_ghosts_own_axis(s, PRED_IDS, NAT3))) return rc;
    if (launch_diag(s->st, s->g, up, s->cx, s->cy, s->cz, s->partial, d0)) return O3D_ERR_CUDA;
    if ((rc = ensure_ghosts_own_axis(s, VEL_IDS, NAT3))) return rc;
    if (launch_diag(s->st, s->g, u, s->cx, s->cy, s->cz, s->partial, d0 + 13)) return O3D_ERR_CUDA;
    if (c.nscr == 1) {
        double* phi = field(s, O3D_F_PHI);
        if (!phi) return O3D_ERR_CUDA;
        if ((rc = ensure_partial(s, reduce_blocks(s->g)))) return rc;
        if (launch_reduce(s->st, s->g, phi, RED_MIN, s->partial, d0 + 26)) return O3D_ERR_CUDA;
        if (launch_reduce(s->st, s->g, phi, RED_MAX, s->partial, d0 + 27)) return O3D_ERR_CUDA;
    }
    double* h = s->scal_h + 64;
    auto fetch = [&]() -> int {
        O3D_CUDA_CHECK(cudaMemcpyAsync(h, d0, 28 * sizeof(double), cudaMemcpyDeviceToHost, s->st));
        O3D_CUDA_CHECK(cudaStreamSynchronize(s->st));
        poll_flag(s);
        return O3D_OK;
    };
    if ((rc = fetch())) return rc;
    if (c.nranks > 1) {
        // regroup by reduction type, reduce across the slabs, withdraw the arg-max position of a
        // rank whose maximum is below the global one (the smallest remaining position wins)
        double mn[10], mx[14], sm[2], loc_max[2] = {h[1], h[14]};
        for (int t = 0; t < 2; ++t) {
            const double* x = h + 13 * t;
            mn[4 * t] = x[0];
            for (int q = 0; q < 3; ++q) mn[4 * t + 1 + q] = x[4 + q];
            mx[7 * t] = x[1];
            for (int q = 0; q < 3; ++q) mx[7 * t + 1 + q] = x[7 + q], mx[7 * t + 4 + q] = x[10 + q];
            sm[t] = x[2];
        }
        mn[8] = h[26];
        double* w = s->scal_d + 96;  // 10 min | 14 max | 2 sum   (mn[8] = phi min, mn[9] spare)
        double stage[28];
        for (int q = 0; q < 10; ++q) stage[q] = (q < 9) ? mn[q] : 0.0;
        for (int q = 0; q < 14; ++q) stage[10 + q] = mx[q];
        stage[24] = h[27];  // phi max rides with the maxima
        stage[25] = sm[0], stage[26] = sm[1];
        O3D_CUDA_CHECK(cudaMemcpyAsync(w, stage, 27 * sizeof(double), cudaMemcpyHostToDevice, s->st));
        if ((rc = comm_allreduce(s, w, 10, RED_MIN))) return rc;
        if ((rc = comm_allreduce(s, w + 10, 15, RED_MAX))) return rc;
        if ((rc = comm_allreduce(s, w + 25, 2, RED_SUM))) return rc;
        double r[27];
        O3D_CUDA_CHECK(cudaMemcpyAsync(s->scal_h + 96, w, 27 * sizeof(double),
                                       cudaMemcpyDeviceToHost, s->st));
        O3D_CUDA_CHECK(cudaStreamSynchronize(s->st));
        for (int q = 0; q < 27; ++q) r[q] = s->scal_h[96 + q];
        double pos[2] = {h[3], h[16]};
        for (int t = 0; t < 2; ++t)
            if (loc_max[t] < r[10 + 7 * t]) pos[t] = 9.0e18;
        O3D_CUDA_CHECK(cudaMemcpyAsync(w, pos, 2 * sizeof(double), cudaMemcpyHostToDevice, s->st));
        if ((rc = comm_allreduce(s, w, 2, RED_MIN))) return rc;
        O3D_CUDA_CHECK(cudaMemcpyAsync(s->scal_h + 96, w, 2 * sizeof(double),
                                       cudaMemcpyDeviceToHost, s->st));
        O3D_CUDA_CHECK(cudaStreamSynchronize(s->st));
        for (int t = 0; t < 2; ++t) {
            double* x = h + 13 * t;
            x[0] = r[4 * t], x[1] = r[10 + 7 * t], x[2] = r[25 + t], x[3] = s->scal_h[96 + t];
            for (int q = 0; q < 3; ++q)
                x[4 + q] = r[4 * t + 1 + q], x[7 + q] = r[10 + 7 * t + 1 + q],
                x[10 + q] = r[10 + 7 * t + 4 + q];
        }
        h[26] = r[8], h[27] = r[24];
    }
    const double cnt = (double)((long long)c.nx * c.ny * c.nz);
    for (int t = 0; t < 2; ++t) {  // function_stats layout, src/functions.f90:27-63
        const double* x = h + 13 * t;
        double* o = out23 + 6 * t;
        o[0] = x[0], o[1] = x[1], o[2] = x[2] / cnt;
        long long m = (x[3] > 8.9e18) ? 0 : (long long)x[3];
        o[3] = (double)(m % c.nx + 1);
        o[4] = (double)((m / c.nx) % c.ny + 1);
        o[5] = (double)(m / ((long long)c.nx * c.ny) + 1);
    }
    const double* xu = h + 13;
    for (int q = 0; q < 3; ++q) out23[12 + q] = xu[4 + q], out23[15 + q] = xu[7 + q];
    // src/utils.f90:199-201
    out23[18] = xu[10] * c.dt / c.dx;
    out23[19] = xu[11] * c.dt / c.dy;
    out23[20] = xu[12] * c.dt / c.dz;
    out23[21] = (c.nscr == 1) ? h[26] : 0.0;
    out23[22] = (c.nscr == 1) ? h[27] : 0.0;
    return O3D_OK;
}

int o3d_s_old_values(o3d_session* s) {
    if (!s) return O3D_ERR_INVALID;
    // src/utils.f90:165-176; whole padded allocations (ghost cells included), device to device
    for (int c = 0; c < 3; ++c) {
        double* src = field(s, O3D_F_UX + c);
        double* dst = field(s, O3D_F_OLD_UX + c);
        if (!src || !dst) return O3D_ERR_CUDA;
        const long long off = interior_offset(s->g);
        O3D_CUDA_CHECK(cudaMemcpyAsync(dst - off, src - off, (size_t)s->felems * sizeof(double),
                                       cudaMemcpyDeviceToDevice, s->st));
        s->gaxes[O3D_F_OLD_UX + c] = s->gaxes[O3D_F_UX + c];
        s->gpar[O3D_F_OLD_UX + c] = s->gpar[O3D_F_UX + c];
    }
    return O3D_OK;
}

int o3d_s_calculate_residuals(o3d_session* s, double dt, double t_ref, double u_ref,
                              double* out15) {
    if (!s || !out15) return O3D_ERR_INVALID;
    const o3d_config& c = s->cfg;
    const double* un[3];
    const double* uo[3];
    for (int q = 0; q < 3; ++q) {
        un[q] = field(s, O3D_F_UX + q);
        uo[q] = field(s, O3D_F_OLD_UX + q);
        if (!un[q] || !uo[q]) return O3D_ERR_CUDA;
    }
    int rc;
    if ((rc = ensure_partial(s, 9ll * 296))) return rc;
    double* out9 = s->scal_d + 48;
    if (launch_residuals(s->st, s->g, un, uo, 2.0 * dt, s->partial, out9)) return O3D_ERR_CUDA;
    double* h = s->scal_h + 48;
    auto fetch = [&]() -> int {
        O3D_CUDA_CHECK(
            cudaMemcpyAsync(h, out9, 9 * sizeof(double), cudaMemcpyDeviceToHost, s->st));
        O3D_CUDA_CHECK(cudaStreamSynchronize(s->st));
        return O3D_OK;
    };
    if (c.nranks > 1) {
        // sums add up, maxima are global; a rank whose maximum is below the global one withdraws
        // its index, the largest remaining (= last in array order) index wins
        if ((rc = fetch())) return rc;
        const double local_mx[3] = {h[3], h[4], h[5]};
        if ((rc = comm_allreduce(s, out9, 3, RED_SUM))) return rc;
        if ((rc = comm_allreduce(s, out9 + 3, 3, RED_MAX))) return rc;
        if ((rc = fetch())) return rc;
        for (int q = 0; q < 3; ++q)
            if (local_mx[q] < h[3 + q]) h[6 + q] = -1.0;
        O3D_CUDA_CHECK(cudaMemcpyAsync(out9 + 6, h + 6, 3 * sizeof(double),
                                       cudaMemcpyHostToDevice, s->st));
        if ((rc = comm_allreduce(s, out9 + 6, 3, RED_MAX))) return rc;
    }
    if ((rc = fetch())) return rc;
    // src/utils.f90:147-152; real(nx*ny*nz) is a default (single precision) real
    const double cnt = (double)(float)(c.nx * c.ny * c.nz);
    for (int q = 0; q < 3; ++q) {
        out15[q] = (t_ref / u_ref) * sqrt((1.0 / cnt) * h[q]);
        out15[3 + q] = (t_ref / u_ref) * h[3 + q];
        const long long m = (long long)h[6 + q];
        out15[6 + 3 * q] = (m < 0) ? 0.0 : (double)(m % c.nx + 1);
        out15[7 + 3 * q] = (m < 0) ? 0.0 : (double)((m / c.nx) % c.ny + 1);
        out15[8 + 3 * q] = (m < 0) ? 0.0 : (double)(m / ((long long)c.nx * c.ny) + 1);
    }
    return O3D_OK;
}

int o3d_s_rotational(o3d_session* s, int rotx, int roty, int rotz) {
    if (!s) return O3D_ERR_INVALID;
    FieldRef u[3] = {fref(s, O3D_F_UX), fref(s, O3D_F_UY), fref(s, O3D_F_UZ)};
    const int rid[3] = {phys_id(s, rotx), phys_id(s, roty), phys_id(s, rotz)};
    double* r[3] = {field(s, rid[0]), field(s, rid[1]), field(s, rid[2])};
    for (int k = 0; k < 3; ++k)
        if (!u[k].p || !r[k]) return O3D_ERR_CUDA;
    int rc;
    // every curl term uses the even closure (src/differential_operators.f90:64-74): refill the
    // velocity ghosts with even parity for this launch; the next consumer restores its own
    if ((rc = ensure_ghosts(s, VEL_IDS, 3, EVEN3, 0x7u))) return rc;
    if (launch_rot(s->st, s->g, u, s->cx, s->cy, s->cz, r[0], r[1], r[2])) return O3D_ERR_CUDA;
    for (int k = 0; k < 3; ++k) touch(s, rid[k]);
    return O3D_OK;
}

int o3d_s_vorticity_magnitude(o3d_session* s, int dst) {
    if (!s) return O3D_ERR_INVALID;
    FieldRef u[3] = {fref(s, O3D_F_UX), fref(s, O3D_F_UY), fref(s, O3D_F_UZ)};
    const int vid = phys_id(s, dst);
    double* v = field(s, vid);
    if (!u[0].p || !u[1].p || !u[2].p || !v) return O3D_ERR_CUDA;
    int rc;
    // the curl's even closures (src/differential_operators.f90:64-74), as o3d_s_rotational
    if ((rc = ensure_ghosts(s, VEL_IDS, 3, EVEN3, 0x7u))) return rc;
    if (launch_vort(s->st, s->g, u, s->cx, s->cy, s->cz, v)) return O3D_ERR_CUDA;
    touch(s, vid);
    return O3D_OK;
}

int o3d_s_q_criterion(o3d_session* s, int dst) {
    if (!s) return O3D_ERR_INVALID;
    FieldRef u[3] = {fref(s, O3D_F_UX), fref(s, O3D_F_UY), fref(s, O3D_F_UZ)};
    const int qid = phys_id(s, dst);
    double* q = field(s, qid);
    if (!u[0].p || !u[1].p || !u[2].p || !q) return O3D_ERR_CUDA;
    int rc;
    // derxi/deryi/derzi on the diagonal, p closures elsewhere (:90-100) = natural parity
    if ((rc = ensure_ghosts(s, VEL_IDS, 3, NAT3, 0x7u))) return rc;
    if (launch_qcrit(s->st, s->g, u, s->cx, s->cy, s->cz, q)) return O3D_ERR_CUDA;
    touch(s, qid);
    return O3D_OK;
}

}  // extern "C"
