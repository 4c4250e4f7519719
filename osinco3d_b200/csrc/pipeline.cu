// pipeline.cu -- z-chunk pipelining of the stateless HOST-pointer procedures (modules.cu).
//
// o3d_predict_velocity / o3d_correct_velocity move 22 / 7 fields over PCIe per call against
// < 0.5 ms of kernel time, so their cost is the copy time.  The plain path runs upload ->
// kernel -> download strictly one after the other: only one PCIe direction is ever busy.  Both
// kernels are z-marching stencils whose output planes [za, zb) depend on the input planes
// [za-3, zb+3) alone, so here the host arrays are cut into C z chunks and three kinds of work
// overlap:
//     upload streams   : H2D of chunk j+1 into a staging slot + pack into the padded field
//     session stream   : ghost fill + kernel on chunk j (plane-range launch, Geom::zr_lo/zr_hi)
//     download streams : unpack of chunk j-1 + D2H
// i.e. both PCIe directions are busy at the same time.  A call costs about max(bytes up, bytes
// down) instead of their sum; it still returns only when every output array is complete, so the
// procedures stay drop-ins for the reference's (src/integration.f90:14, :257).
//
// The kernels, their arguments and the closure data are those of the plain path -- a plane-range
// launch computes every point exactly as the whole-slab launch does (the same property the
// interior / boundary split of the multi-GPU path relies on) -- so the results are bitwise equal
// (tests/test_gpu_pipeline.py).
//
// Order constraints: chunk c needs upload chunk c+1 (3 planes beyond its end); with a periodic z
// axis chunk 0 also needs the LAST planes (wrap ghosts), so it runs last.  Mirror z ghosts are
// filled one side at a time, as soon as that side's source planes have landed.
#include <cstdlib>
#include <vector>

#include "host_copier.h"
#include "session.h"

namespace o3d {

constexpr int MAX_CHUNKS = 64;

struct Pipe {
    cudaStream_t up[2], dn[2];
    double* slot_up[2];
    double* slot_dn[2];
    long long slot_elems;
    cudaEvent_t ev_start;
    std::vector<cudaEvent_t> ev;
    // host-side history shift (host_copier.h): worker threads, the events they wait on (created
    // with cudaEventBlockingSync: a waiting worker sleeps instead of spinning on a core the
    // caller may need), one "level 3 <- level 2 done" flag per chunk and component
    HostCopier* hc;
    std::vector<cudaEvent_t> hev;
    std::atomic<int> a_done[3 * MAX_CHUNKS];
};

namespace {

int g_pipe_setting = -1;  // -1: take O3D_PIPELINE from the environment at first use
// Host-side history shift of the pipelined o3d_predict_velocity (host_copier.h): -1 = take
// O3D_HOSTSHIFT from the environment at first use.
// Default: on where the worker pool reaches its 8 threads (>= 10 hardware threads).  Measured end
// to end on the 256^3 TGV step (B200, PCIe Gen5, 16-core host, pinned arrays;
// profiles/r2y_e2e_hostshift_ab.jsonl): 63.9 ms per step with everything downloaded, 59.2 - 60.0 ms
// with 8 or 16 workers, 69.8 ms with 3 (the memcpy then trails the transfers it replaces).
int g_hostshift = -1;
constexpr unsigned HOSTSHIFT_MIN_HW_THREADS = 10;
// Default number of chunks.  Measured end to end on the 256^3 TGV step (B200, PCIe Gen5, pinned
// arrays; profiles/r1p_e2e_pipeline.jsonl, r1q_e2e_pipeline.jsonl): 86.8 ms unpipelined, 71.3 /
// 65.8 / 65.1 ms with 4 / 8 / 16 chunks; flat (63.4 - 63.9 ms on a second box) from 12 to 32.
constexpr int DEFAULT_CHUNKS = 16;

constexpr int MIN_PLANES = 8;  // per chunk: > stencil radius + mirror source planes

const unsigned NAT3[3] = {0x1u, 0x2u, 0x4u};

// The schedule of one pipelined call (pure host logic; o3d_pipeline_plan exposes it to the CPU
// tests).  Chunks are uploaded in ascending order; after upload j every chunk c that has not been
// issued yet and has need[c] <= j is issued, in ascending order.
struct Plan {
    int C;                  // number of chunks (0: too thin, plain path)
    int z[MAX_CHUNKS + 1];  // chunk c = planes [z[c], z[c+1])
    int need[MAX_CHUNKS];   // last upload chunk that the kernel on chunk c reads
    int zfill[MAX_CHUNKS];  // z ghost sides filled just before chunk c runs: 1 = low, 2 = high
};

void make_plan(int nz, int chunks, bool wrapz, Plan& p) {
    int C = chunks;
    if (C > MAX_CHUNKS) C = MAX_CHUNKS;
    while (C >= 2 && nz / C < MIN_PLANES) --C;
    p.C = (C >= 2) ? C : 0;
    if (!p.C) return;
    for (int c = 0; c <= C; ++c) p.z[c] = (int)((long long)nz * c / C);
    for (int c = 0; c < C; ++c) {
        // three planes beyond its end = the next chunk; a periodic z axis makes chunk 0 read the
        // LAST planes (wrap ghosts), so it goes behind the last upload
        p.need[c] = (c + 1 < C) ? c + 1 : C - 1;
        p.zfill[c] = 0;
    }
    if (wrapz) {
        // both sides with chunk 0: their sources are planes nz-3 .. nz-1 and 0 .. 2, and chunk 0
        // is issued before the last chunk (ascending order among the chunks upload C-1 releases)
        p.need[0] = C - 1;
        p.zfill[0] = 3;
    } else {
        // mirror: the low ghosts copy planes 1 .. 3 (chunk 0), the high ghosts planes nz-4 ..
        // nz-2 (last chunk): each side as soon as its own chunk has landed
        p.zfill[0] = 1;
        p.zfill[C - 1] = 2;
    }
}

int hostshift_threads() {
    const char* e = getenv("O3D_HOSTSHIFT_THREADS");
    int n = e ? atoi(e) : 0;
    if (n < 1) {
        // leave two cores to the caller and the CUDA driver threads; 8 workers saturate what the
        // PCIe link frees up (6 fields of memcpy against 6 fields of D2H)
        const unsigned hw = std::thread::hardware_concurrency();
        n = hw ? (int)hw - 2 : 4;
        if (n > 8) n = 8;
    }
    return n < 1 ? 1 : (n > 64 ? 64 : n);
}

int wait_event(void* ev) { return (int)cudaEventSynchronize((cudaEvent_t)ev); }
void enter_device(int dev) { cudaSetDevice(dev); }

int ensure_pipe(o3d_session* s, long long slot_elems, int nev, int nhev = 0) {
    Pipe* p = s->pipe;
    if (!p) {
        p = new (std::nothrow) Pipe();
        if (!p) return O3D_ERR_INVALID;
        for (int l = 0; l < 2; ++l) p->up[l] = p->dn[l] = nullptr, p->slot_up[l] = p->slot_dn[l] = nullptr;
        p->slot_elems = 0;
        p->ev_start = nullptr;
        p->hc = nullptr;
        s->pipe = p;
        for (int l = 0; l < 2; ++l) {
            O3D_CUDA_CHECK(cudaStreamCreateWithFlags(&p->up[l], cudaStreamNonBlocking));
            O3D_CUDA_CHECK(cudaStreamCreateWithFlags(&p->dn[l], cudaStreamNonBlocking));
        }
        O3D_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming));
    }
    if (slot_elems > p->slot_elems) {
        for (int l = 0; l < 2; ++l) {
            if (p->slot_up[l]) cudaFree(p->slot_up[l]);
            if (p->slot_dn[l]) cudaFree(p->slot_dn[l]);
            p->slot_up[l] = p->slot_dn[l] = nullptr;
        }
        p->slot_elems = 0;
        for (int l = 0; l < 2; ++l) {
            O3D_CUDA_CHECK(cudaMalloc(&p->slot_up[l], (size_t)slot_elems * sizeof(double)));
            O3D_CUDA_CHECK(cudaMalloc(&p->slot_dn[l], (size_t)slot_elems * sizeof(double)));
        }
        p->slot_elems = slot_elems;
    }
    while ((int)p->ev.size() < nev) {
        cudaEvent_t e;
        O3D_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        p->ev.push_back(e);
    }
    while ((int)p->hev.size() < nhev) {
        cudaEvent_t e;
        O3D_CUDA_CHECK(
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));
        p->hev.push_back(e);
    }
    if (nhev && !p->hc) {
        int dev = 0;
        O3D_CUDA_CHECK(cudaGetDevice(&dev));
        // a host that cannot start the workers (thread limit) keeps the download path
        try {
            p->hc = new HostCopier(hostshift_threads(), wait_event, enter_device, dev);
        } catch (...) {
            p->hc = nullptr;
        }
    }
    return O3D_OK;
}

// One pipelined call: the chunk partition, the copy lanes and the order constraints.
struct Run {
    o3d_session* s;
    Pipe* p;
    Plan pl;
    int C;
    const int* z;           // = pl.z
    long long plane;        // nx * ny
    int n_up, n_dn;         // transfer counters: lane = counter & 1
    int dn_lane;            // lane of the last download()
    bool hostshift;

    int init(o3d_session* ses, int chunks, bool hs = false) {
        s = ses;
        make_plan(s->g.nz, chunks, s->g.bz_lo == BM_WRAP, pl);
        C = pl.C;
        z = pl.z;
        if (C < 2) return O3D_ERR_INVALID;
        int maxnk = 0;
        for (int c = 0; c < C; ++c)
            if (z[c + 1] - z[c] > maxnk) maxnk = z[c + 1] - z[c];
        plane = (long long)s->g.nx * s->g.ny;
        n_up = n_dn = 0;
        dn_lane = 0;
        hostshift = hs;
        int rc = ensure_pipe(s, plane * maxnk, 3 * C, hs ? 5 * C : 0);
        if (rc) return rc;
        p = s->pipe;
        hostshift = hs && p->hc != nullptr;
        if (hostshift)
            for (int i = 0; i < 3 * C; ++i) p->a_done[i].store(0, std::memory_order_relaxed);
        // the lanes start behind everything queued on the session stream so far (the lazy
        // zero fill of freshly allocated fields, the memsets of the caller)
        O3D_CUDA_CHECK(cudaEventRecord(p->ev_start, s->st));
        for (int l = 0; l < 2; ++l) {
            O3D_CUDA_CHECK(cudaStreamWaitEvent(p->up[l], p->ev_start, 0));
            O3D_CUDA_CHECK(cudaStreamWaitEvent(p->dn[l], p->ev_start, 0));
        }
        return O3D_OK;
    }
    cudaEvent_t ev_up(int c, int lane) const { return p->ev[2 * c + lane]; }
    cudaEvent_t ev_cmp(int c) const { return p->ev[2 * C + c]; }
    // events the host workers wait on: upload chunk c complete on lane l; level 1 of component k,
    // chunk c, has landed in the host array
    cudaEvent_t hev_up(int c, int lane) const { return p->hev[2 * c + lane]; }
    cudaEvent_t hev_l1(int c, int k) const { return p->hev[2 * C + 3 * c + k]; }
    long long chunk_off(int c) const { return (long long)z[c] * plane; }
    size_t chunk_bytes(int c) const { return (size_t)(plane * (z[c + 1] - z[c])) * sizeof(double); }
    // last upload chunk that chunk c reads
    int need(int c) const { return pl.need[c]; }
    // planes of chunk c: host array -> staging slot -> padded field `d` (interior origin)
    int upload(double* d, const double* host, int c) {
        const int lane = (n_up++) & 1;
        const int k0 = z[c], nk = z[c + 1] - z[c];
        O3D_CUDA_CHECK(cudaMemcpyAsync(p->slot_up[lane], host + (long long)k0 * plane,
                                       (size_t)(plane * nk) * sizeof(double),
                                       cudaMemcpyHostToDevice, p->up[lane]));
        if (launch_pack_planes(p->up[lane], s->g, p->slot_up[lane], d, k0, nk)) {
            set_error("pack launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return O3D_ERR_CUDA;
        }
        return O3D_OK;
    }
    int uploaded(int c) {
        for (int l = 0; l < 2; ++l) O3D_CUDA_CHECK(cudaEventRecord(ev_up(c, l), p->up[l]));
        return O3D_OK;
    }
    // the session stream may read everything up to upload chunk j
    int wait_uploads(int j) {
        for (int l = 0; l < 2; ++l) O3D_CUDA_CHECK(cudaStreamWaitEvent(s->st, ev_up(j, l), 0));
        return O3D_OK;
    }
    // chunk c has been computed on the session stream: its planes may be downloaded
    int computed(int c) {
        O3D_CUDA_CHECK(cudaEventRecord(ev_cmp(c), s->st));
        for (int l = 0; l < 2; ++l) O3D_CUDA_CHECK(cudaStreamWaitEvent(p->dn[l], ev_cmp(c), 0));
        return O3D_OK;
    }
    int download(const double* d, double* host, int c) {
        const int lane = dn_lane = (n_dn++) & 1;
        const int k0 = z[c], nk = z[c + 1] - z[c];
        if (launch_unpack_planes(p->dn[lane], s->g, d, p->slot_dn[lane], k0, nk)) {
            set_error("unpack launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return O3D_ERR_CUDA;
        }
        O3D_CUDA_CHECK(cudaMemcpyAsync(host + (long long)k0 * plane, p->slot_dn[lane],
                                       (size_t)(plane * nk) * sizeof(double),
                                       cudaMemcpyDeviceToHost, p->dn[lane]));
        return O3D_OK;
    }
    // ghost cells a kernel on chunk c reads, for `nf` fields with parity par[q]: the x / y faces of
    // the chunk's own planes, and the z faces the plan assigns to this chunk (Plan::zfill).
    // A z stencil only reads the thread's own column: no x / y ghosts in other planes are needed.
    int fill_ghosts(double* const* d, const unsigned* par, int nf, int c) {
        Geom gg = s->g;
        gg.zr_lo = z[c], gg.zr_hi = z[c + 1];
        GhostArgs ga;
        ga.njobs = 0;
        for (int q = 0; q < nf; ++q) {
            GhostJob& jb = ga.job[ga.njobs++];
            jb.p = d[q], jb.par = par[q], jb.axes = 0x3u;
            if (pl.zfill[c] == 1) jb.axes |= 0x4u | GHOST_Z_LO_ONLY;
            if (pl.zfill[c] == 2) jb.axes |= 0x4u | GHOST_Z_HI_ONLY;
            if (pl.zfill[c] == 3) jb.axes |= 0x4u;
        }
        if (launch_fill_ghosts(s->st, gg, ga)) {
            set_error("ghost fill launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return O3D_ERR_CUDA;
        }
        return O3D_OK;
    }
    // every stream idle: all host output arrays are complete (also called on the error paths --
    // no copy may still be in flight when the procedure returns)
    int drain() {
        cudaError_t e = cudaSuccess, q;
        for (int l = 0; l < 2; ++l) {
            if ((q = cudaStreamSynchronize(p->up[l])) != cudaSuccess) e = q;
            if ((q = cudaStreamSynchronize(p->dn[l])) != cudaSuccess) e = q;
        }
        if ((q = cudaStreamSynchronize(s->st)) != cudaSuccess) e = q;
        // ... and every host-made output is in place
        if (hostshift && p->hc) {
            const int he = p->hc->drain();
            if (he && e == cudaSuccess) e = (cudaError_t)he;
        }
        if (e != cudaSuccess) {
            set_error("pipelined procedure failed: %s", cudaGetErrorString(e));
            return O3D_ERR_CUDA;
        }
        return O3D_OK;
    }
};

}  // namespace

int pipe_setting() {
    if (g_pipe_setting < 0) {
        const char* e = getenv("O3D_PIPELINE");
        g_pipe_setting = e ? (atoi(e) > 0 ? atoi(e) : 0) : DEFAULT_CHUNKS;
    }
    return g_pipe_setting;
}

int pipe_chunks(int nz) {
    Plan p;
    make_plan(nz, pipe_setting(), false, p);
    return p.C;
}

int hostshift_setting() {
    if (g_hostshift < 0) {
        const char* e = getenv("O3D_HOSTSHIFT");
        g_hostshift = e ? (atoi(e) != 0)
                        : (std::thread::hardware_concurrency() >= HOSTSHIFT_MIN_HW_THREADS);
    }
    return g_hostshift;
}

void pipe_destroy(o3d_session* s) {
    Pipe* p = s->pipe;
    if (!p) return;
    delete p->hc;  // joins the workers (the queue is empty: every call drains it)
    p->hc = nullptr;
    for (auto e : p->hev) cudaEventDestroy(e);
    for (int l = 0; l < 2; ++l) {
        if (p->up[l]) cudaStreamSynchronize(p->up[l]);
        if (p->dn[l]) cudaStreamSynchronize(p->dn[l]);
    }
    for (int l = 0; l < 2; ++l) {
        if (p->slot_up[l]) cudaFree(p->slot_up[l]);
        if (p->slot_dn[l]) cudaFree(p->slot_dn[l]);
        if (p->up[l]) cudaStreamDestroy(p->up[l]);
        if (p->dn[l]) cudaStreamDestroy(p->dn[l]);
    }
    for (auto e : p->ev) cudaEventDestroy(e);
    if (p->ev_start) cudaEventDestroy(p->ev_start);
    delete p;
    s->pipe = nullptr;
}

// predict_velocity (src/integration.f90:14-197) on host arrays: u_h[3] in, f_h[3] = fux/fuy/fuz
// (nx,ny,nz,3) inout, up_h[3] and nu_t_h out.  The caller (o3d_predict_velocity) has set the
// configuration, reset the history mapping and zeroed nu_t when iles /= 1.
int pipe_predict_velocity(o3d_session* s, int itime, double* const* up_h,
                          const double* const* u_h, double* const* f_h, double* nu_t_h) {
    const int C = pipe_chunks(s->g.nz);
    if (C < 2 || s->cfg.nranks > 1) return O3D_ERR_INVALID;
    RhsArgs a;
    int tgt[3];
    int rc = rhs_prepare(s, itime, a, tgt);
    if (rc) return rc;
    const long long N = s->nloc;
    // device buffers of the inputs (history levels 2 and 3 as they are BEFORE the rotation) ...
    double* ud[3];
    double* f2d[3];
    double* f3d[3];
    for (int k = 0; k < 3; ++k) {
        ud[k] = a.u[k].p;
        f2d[k] = field(s, hist_id(s, k, 2));
        f3d[k] = field(s, hist_id(s, k, 3));
        if (!f2d[k] || !f3d[k]) return O3D_ERR_CUDA;
    }
    // ... and of the outputs: the three levels as the reference leaves them AFTER its shifting
    // copies (src/integration.f90:176-188).  rhs_finish is pure bookkeeping (level -> buffer map,
    // ghost state of u*), so it can run before the kernels are queued.
    rhs_finish(s, tgt, a.iles != 0);
    const double* fout[3][3];
    for (int k = 0; k < 3; ++k)
        for (int l = 0; l < 3; ++l) {
            fout[k][l] = field(s, hist_id(s, k, l + 1));
            if (!fout[k][l]) return O3D_ERR_CUDA;
        }
    // Host-side history shift (host_copier.h).  After the call the reference leaves
    //     itscheme = 3:  level 3 = old level 2,  level 2 = level 1 = new f
    //     itscheme = 2:  level 2 = level 1 = new f,  level 3 untouched
    //     otherwise   :  level 1 = new f,  levels 2 and 3 untouched     (src/integration.f90:176-188)
    // and the host holds the old levels: only level 1 has to cross PCIe.  The copies are memcpy
    // jobs on worker threads, ordered against the DMA transfers of the same chunk by events; the
    // result is bit for bit what the downloads would have delivered (memcpy and PCIe both move
    // bits).  DNS: nu_t = 0.d0 (src/integration.f90:112) is a host memset.
    const int itscheme = s->cfg.itscheme;
    Run r;
    if ((rc = r.init(s, C, hostshift_setting() != 0))) return rc;
    const bool hs = r.hostshift;
    auto body = [&]() -> int {
        int rc2;
        bool issued[MAX_CHUNKS] = {false};
        if (hs && !a.nu_t)
            for (int c = 0; c < C; ++c)
                r.p->hc->push(zero_job(nu_t_h, r.chunk_off(c), r.chunk_bytes(c)));
        for (int j = 0; j < C; ++j) {
            for (int k = 0; k < 3; ++k)
                if ((rc2 = r.upload(ud[k], u_h[k], j))) return rc2;
            for (int k = 0; k < 3; ++k) {
                // level 1 is overwritten before it is read (src/integration.f90:129)
                if ((rc2 = r.upload(f2d[k], f_h[k] + N, j))) return rc2;
                if ((rc2 = r.upload(f3d[k], f_h[k] + 2 * N, j))) return rc2;
            }
            if ((rc2 = r.uploaded(j))) return rc2;
            if (hs && itscheme == 3) {
                // job A: level 3 <- old level 2 of chunk j, once the DMA engine has read both
                for (int l = 0; l < 2; ++l)
                    O3D_CUDA_CHECK(cudaEventRecord(r.hev_up(j, l), r.p->up[l]));
                for (int k = 0; k < 3; ++k)
                    r.p->hc->push(shift_job_a(f_h[k], N, r.chunk_off(j), r.chunk_bytes(j),
                                              r.hev_up(j, 0), r.hev_up(j, 1),
                                              &r.p->a_done[3 * j + k]));
            }
            for (int c = 0; c < C; ++c) {
                if (issued[c] || r.need(c) > j) continue;
                issued[c] = true;
                if ((rc2 = r.wait_uploads(j))) return rc2;
                // parity table of src/integration.f90:118-165 = natural-parity ghosts of u
                if ((rc2 = r.fill_ghosts(ud, NAT3, 3, c))) return rc2;
                Geom gg = s->g;
                gg.zr_lo = r.z[c], gg.zr_hi = r.z[c + 1];
                if (launch_rhs(s->st, gg, a)) {
                    set_error("rhs kernel launch failed: %s",
                              cudaGetErrorString(cudaGetLastError()));
                    return O3D_ERR_CUDA;
                }
                if ((rc2 = r.computed(c))) return rc2;
                if (hs) {
                    // level 1 first: its arrival releases job B
                    for (int k = 0; k < 3; ++k) {
                        if ((rc2 = r.download(fout[k][0], f_h[k], c))) return rc2;
                        if (itscheme != 2 && itscheme != 3) continue;
                        O3D_CUDA_CHECK(cudaEventRecord(r.hev_l1(c, k), r.p->dn[r.dn_lane]));
                        // job B: level 2 <- new level 1 of chunk c.  The upload of the old
                        // level 2 of this chunk is complete (the kernel on chunk c waited for
                        // upload need(c) >= c); job A (pushed with upload c <= j) has read it
                        // once a_done is set.
                        r.p->hc->push(shift_job_b(
                            f_h[k], N, r.chunk_off(c), r.chunk_bytes(c), r.hev_l1(c, k),
                            (itscheme == 3) ? &r.p->a_done[3 * c + k] : nullptr));
                    }
                }
                for (int k = 0; k < 3; ++k)
                    if ((rc2 = r.download(a.up[k], up_h[k], c))) return rc2;
                // (DNS: the field the caller zeroed -- nu_t = 0.d0, src/integration.f90:112 --
                // since the kernel arguments carry no nu_t then)
                if (!hs || a.nu_t)
                    if ((rc2 = r.download(a.nu_t ? a.nu_t : field(s, O3D_F_NU_T), nu_t_h, c)))
                        return rc2;
                if (!hs)
                    for (int k = 0; k < 3; ++k)
                        for (int l = 0; l < 3; ++l)
                            if ((rc2 = r.download(fout[k][l], f_h[k] + (long long)l * N, c)))
                                return rc2;
            }
        }
        return O3D_OK;
    };
    rc = body();
    const int rcd = r.drain();
    // the interiors of the inputs were rewritten: whatever ghost state was recorded is stale
    for (int k = 0; k < 3; ++k) {
        touch(s, O3D_F_UX + k);
        touch(s, hist_id(s, k, 3));  // level 3 holds the uploaded level 2
    }
    return rc ? rc : rcd;
}

// correct_velocity (src/integration.f90:257-330) on host arrays: pp_h, up_h[3] in, u_h[3] out;
// O3D_ERR_DIVERGED after the outputs are complete if the NaN / > 1000 guard fired.
int pipe_correct_velocity(o3d_session* s, double* const* u_h, const double* const* up_h,
                          const double* pp_h) {
    const int C = pipe_chunks(s->g.nz);
    if (C < 2 || s->cfg.nranks > 1) return O3D_ERR_INVALID;
    const o3d_config& c = s->cfg;
    FieldRef up[3] = {fref(s, O3D_F_UX_PRED), fref(s, O3D_F_UY_PRED), fref(s, O3D_F_UZ_PRED)};
    double* u[3] = {field(s, O3D_F_UX), field(s, O3D_F_UY), field(s, O3D_F_UZ)};
    FieldRef pp = fref(s, O3D_F_PP);
    for (int k = 0; k < 3; ++k)
        if (!up[k].p || !u[k]) return O3D_ERR_CUDA;
    if (!pp.p) return O3D_ERR_CUDA;
    O3D_CUDA_CHECK(cudaMemsetAsync(s->flag_d, 0, sizeof(int), s->st));
    Run r;
    int rc = r.init(s, C);
    if (rc) return rc;
    auto body = [&]() -> int {
        int rc2;
        bool issued[MAX_CHUNKS] = {false};
        const unsigned even = 0u;
        for (int j = 0; j < C; ++j) {
            if ((rc2 = r.upload(pp.p, pp_h, j))) return rc2;
            for (int k = 0; k < 3; ++k)
                if ((rc2 = r.upload(up[k].p, up_h[k], j))) return rc2;
            if ((rc2 = r.uploaded(j))) return rc2;
            for (int cc = 0; cc < C; ++cc) {
                if (issued[cc] || r.need(cc) > j) continue;
                issued[cc] = true;
                if ((rc2 = r.wait_uploads(j))) return rc2;
                // derxp / deryp / derzp of pp (src/integration.f90:298-300): even ghosts
                if ((rc2 = r.fill_ghosts(&pp.p, &even, 1, cc))) return rc2;
                Geom gg = s->g;
                gg.zr_lo = r.z[cc], gg.zr_hi = r.z[cc + 1];
                if (launch_corr(s->st, gg, pp, up, u, s->cx, s->cy, s->cz, c.dt, s->flag_d)) {
                    set_error("correction kernel launch failed: %s",
                              cudaGetErrorString(cudaGetLastError()));
                    return O3D_ERR_CUDA;
                }
                if ((rc2 = r.computed(cc))) return rc2;
                for (int k = 0; k < 3; ++k)
                    if ((rc2 = r.download(u[k], u_h[k], cc))) return rc2;
            }
        }
        O3D_CUDA_CHECK(
            cudaMemcpyAsync(s->flag_h, s->flag_d, sizeof(int), cudaMemcpyDeviceToHost, s->st));
        return O3D_OK;
    };
    rc = body();
    const int rcd = r.drain();
    touch(s, O3D_F_PP);
    for (int k = 0; k < 3; ++k) {
        touch(s, O3D_F_UX_PRED + k);
        // the kernel wrote the natural-parity ghost images of u with the interior, as in the
        // whole-slab launch
        s->gaxes[O3D_F_UX + k] = 0xFu;
        s->gpar[O3D_F_UX + k] = NAT3[k];
    }
    s->flag_pending = 0;
    if (rc) return rc;
    if (rcd) return rcd;
    if (*s->flag_h) {
        set_error("velocity diverged: NaN or max(u) > 1000 (src/integration.f90:309-325)");
        return O3D_ERR_DIVERGED;
    }
    return O3D_OK;
}

}  // namespace o3d

extern "C" int o3d_set_pipeline(int chunks) {
    if (chunks < 0) return O3D_ERR_INVALID;
    o3d::g_pipe_setting = chunks;
    return O3D_OK;
}

extern "C" int o3d_get_pipeline(void) { return o3d::pipe_setting(); }

extern "C" int o3d_set_hostshift(int on) {
    if (on < 0 || on > 1) return O3D_ERR_INVALID;
    o3d::g_hostshift = on;
    return O3D_OK;
}

extern "C" int o3d_get_hostshift(void) { return o3d::hostshift_setting(); }

extern "C" int o3d_pipeline_plan(int nz, int periodic_z, int* z_bounds, int* issue_after,
                                 int* zfill) {
    if (nz < 1) return 0;
    o3d::Plan p;
    o3d::make_plan(nz, o3d::pipe_setting(), periodic_z != 0, p);
    for (int c = 0; c < p.C; ++c) {
        if (z_bounds) z_bounds[c] = p.z[c];
        if (issue_after) issue_after[c] = p.need[c];
        if (zfill) zfill[c] = p.zfill[c];
    }
    if (p.C && z_bounds) z_bounds[p.C] = p.z[p.C];
    return p.C;
}
