// io.cu -- the reference's binary field formats, written from device-resident state without
// stalling the time loop:
//   save_fields / read_fields   src/IOfunctions.f90:360-402 / :404-470   restart file
//       stream-access unformatted: time (f64) | nx ny nz (3 x i32) | x(nx) y(ny) z(nz) (f64) |
//       ux uy uz pp phi (nx*ny*nz f64 each, i fastest)
//   write_binary / write_all_data   src/visualization.f90:224-241 / :243-276   one raw f64
//       array per file: ux uy uz pp vort(= |curl u|) qcrit [phi if nscr] [nu_t if iles]
//
// z is the slowest index and the domain is decomposed into z slabs, so the slab a rank owns is
// ONE contiguous byte range of every field in the file: each rank pwrite()s its own range of the
// shared file -- no gather onto one rank.
//
// Asynchrony.  A field is (1) snapshotted on the session stream by the unpack kernel (padded ->
// contiguous, into a device staging buffer: after this the time loop may overwrite the field),
// (2) copied D2H on a separate I/O stream into pinned memory, (3) written by a host writer
// thread once the copy's event has fired.  Stages 2 and 3 overlap the following time steps.
// A small pool of staging buffers bounds the memory; acquiring one blocks only when all are
// still draining (back-pressure).  o3d_s_io_wait / o3d_session_destroy drain the queue.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>

#include "session.h"

namespace o3d {

namespace {
const int IO_NBUF = 3;
}

struct IoEngine {
    struct Buf {
        double* dev = nullptr;
        double* host = nullptr;  // pinned
        cudaEvent_t copied = nullptr;
        bool busy = false;
    };
    struct Job {
        int buf;
        int fd;
        long long offset;
        size_t bytes;
        bool close_fd;
    };
    Buf buf[IO_NBUF];
    cudaStream_t st_io = nullptr;
    cudaEvent_t snap = nullptr;
    int device = 0;
    std::thread worker;
    std::mutex m;
    std::condition_variable cv_job, cv_free;
    std::deque<Job> q;
    int in_flight = 0;
    bool stop = false;
    int err = 0;  // first errno seen by the writer
    std::string err_what;

    void run() {
        cudaSetDevice(device);
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_job.wait(lk, [&] { return stop || !q.empty(); });
                if (q.empty()) return;
                j = q.front();
                q.pop_front();
            }
            int e = 0;
            if (cudaEventSynchronize(buf[j.buf].copied) != cudaSuccess) e = EIO;
            const char* p = reinterpret_cast<const char*>(buf[j.buf].host);
            size_t left = j.bytes;
            long long off = j.offset;
            while (!e && left) {
                const ssize_t w = pwrite(j.fd, p, left, (off_t)off);
                if (w < 0) {
                    if (errno == EINTR) continue;
                    e = errno;
                    break;
                }
                p += w, off += w, left -= (size_t)w;
            }
            if (j.close_fd && close(j.fd) != 0 && !e) e = errno;
            {
                std::lock_guard<std::mutex> lk(m);
                if (e && !err) err = e, err_what = strerror(e);
                buf[j.buf].busy = false;
                --in_flight;
            }
            cv_free.notify_all();
        }
    }
};

namespace {

int io_engine(o3d_session* s, IoEngine** out) {
    if (s->io) {
        *out = s->io;
        return O3D_OK;
    }
    IoEngine* e = new IoEngine();
    cudaGetDevice(&e->device);
    const size_t bytes = (size_t)s->nloc * sizeof(double);
    cudaError_t ce = cudaStreamCreateWithFlags(&e->st_io, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->snap, cudaEventDisableTiming);
    for (int b = 0; b < IO_NBUF && ce == cudaSuccess; ++b) {
        ce = cudaMalloc(&e->buf[b].dev, bytes);
        if (ce == cudaSuccess) ce = cudaHostAlloc(&e->buf[b].host, bytes, cudaHostAllocDefault);
        if (ce == cudaSuccess)
            ce = cudaEventCreateWithFlags(&e->buf[b].copied, cudaEventDisableTiming);
    }
    s->io = e;
    if (ce != cudaSuccess) {
        set_error("I/O staging allocation failed: %s", cudaGetErrorString(ce));
        io_destroy(s);
        return O3D_ERR_CUDA;
    }
    e->worker = std::thread([e] { e->run(); });
    *out = e;
    return O3D_OK;
}

// queue this rank's slab of one padded field for writing at byte `offset` of fd
int queue_field(o3d_session* s, IoEngine* e, const double* d, int fd, long long offset,
                bool close_fd) {
    int b = -1;
    {
        std::unique_lock<std::mutex> lk(e->m);
        e->cv_free.wait(lk, [&] {
            for (int q = 0; q < IO_NBUF; ++q)
                if (!e->buf[q].busy) return true;
            return false;
        });
        for (int q = 0; q < IO_NBUF; ++q)
            if (!e->buf[q].busy) {
                b = q;
                break;
            }
        e->buf[b].busy = true;
        ++e->in_flight;
    }
    const size_t bytes = (size_t)s->nloc * sizeof(double);
    int rc = O3D_OK;
    if (launch_unpack(s->st, s->g, d, e->buf[b].dev)) rc = O3D_ERR_CUDA;
    if (!rc && (cudaEventRecord(e->snap, s->st) != cudaSuccess ||
                cudaStreamWaitEvent(e->st_io, e->snap, 0) != cudaSuccess ||
                cudaMemcpyAsync(e->buf[b].host, e->buf[b].dev, bytes, cudaMemcpyDeviceToHost,
                                e->st_io) != cudaSuccess ||
                cudaEventRecord(e->buf[b].copied, e->st_io) != cudaSuccess))
        rc = O3D_ERR_CUDA;
    if (rc) {
        {
            std::lock_guard<std::mutex> lk(e->m);
            e->buf[b].busy = false;
            --e->in_flight;
        }
        e->cv_free.notify_all();  // a waiter in o3d_s_io_wait / queue_field must see the release
        if (close_fd) close(fd);
        set_error("I/O snapshot of a field failed: %s", cudaGetErrorString(cudaGetLastError()));
        return rc;
    }
    {
        std::lock_guard<std::mutex> lk(e->m);
        e->q.push_back({b, fd, offset, bytes, close_fd});
    }
    e->cv_job.notify_one();
    return O3D_OK;
}

int open_out(const char* path, long long total_bytes, int* fd) {
    // no O_TRUNC: with several ranks writing their slabs into one file, a late truncation by one
    // rank would wipe what another already wrote.  Setting the exact final size is idempotent.
    *fd = open(path, O_WRONLY | O_CREAT, 0644);
    if (*fd < 0 || ftruncate(*fd, (off_t)total_bytes) != 0) {
        set_error("Error opening file: %s (%s)", path, strerror(errno));
        if (*fd >= 0) close(*fd);
        return O3D_ERR_IO;
    }
    return O3D_OK;
}

int pwrite_all(int fd, const void* p, size_t n, long long off) {
    const char* c = static_cast<const char*>(p);
    while (n) {
        const ssize_t w = pwrite(fd, c, n, (off_t)off);
        if (w < 0) {
            if (errno == EINTR) continue;
            return 1;
        }
        c += w, off += w, n -= (size_t)w;
    }
    return 0;
}

int pread_all(int fd, void* p, size_t n, long long off) {
    char* c = static_cast<char*>(p);
    while (n) {
        const ssize_t r = pread(fd, c, n, (off_t)off);
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) return 1;
        c += r, off += r, n -= (size_t)r;
    }
    return 0;
}

// one raw array per file (write_binary, src/visualization.f90:224-241)
int write_raw(o3d_session* s, IoEngine* e, const std::string& path, const double* d) {
    const o3d_config& c = s->cfg;
    const long long plane = (long long)c.nx * c.ny * 8;
    int fd;
    int rc = open_out(path.c_str(), plane * c.nz, &fd);
    if (rc) return rc;
    return queue_field(s, e, d, fd, plane * s->z0, true);
}

}  // namespace

void io_destroy(o3d_session* s) {
    IoEngine* e = s->io;
    if (!e) return;
    if (e->worker.joinable()) {
        {
            std::unique_lock<std::mutex> lk(e->m);
            e->cv_free.wait(lk, [&] { return e->in_flight == 0; });
            e->stop = true;
        }
        e->cv_job.notify_all();
        e->worker.join();
    }
    for (int b = 0; b < IO_NBUF; ++b) {
        if (e->buf[b].dev) cudaFree(e->buf[b].dev);
        if (e->buf[b].host) cudaFreeHost(e->buf[b].host);
        if (e->buf[b].copied) cudaEventDestroy(e->buf[b].copied);
    }
    if (e->snap) cudaEventDestroy(e->snap);
    if (e->st_io) cudaStreamDestroy(e->st_io);
    delete e;
    s->io = nullptr;
}

}  // namespace o3d

using namespace o3d;

extern "C" {

int o3d_s_io_wait(o3d_session* s) {
    if (!s) return O3D_ERR_INVALID;
    IoEngine* e = s->io;
    if (!e) return O3D_OK;
    std::unique_lock<std::mutex> lk(e->m);
    e->cv_free.wait(lk, [&] { return e->in_flight == 0; });
    if (e->err) {
        set_error("field output failed: %s", e->err_what.c_str());
        e->err = 0;
        return O3D_ERR_IO;
    }
    return O3D_OK;
}

int o3d_s_save_fields(o3d_session* s, const char* filename, double time, const double* x,
                      const double* y, const double* z) {
    if (!s || !filename || !x || !y || !z) return O3D_ERR_INVALID;
    const o3d_config& c = s->cfg;
    IoEngine* e;
    // the reference stops inside correct_velocity before any save_fields
    // (src/integration.f90:309-325): report a pending NaN / >1000 flag instead of writing
    int rc = o3d_sync(s);
    if (rc) return rc;
    if ((rc = io_engine(s, &e))) return rc;
    // src/IOfunctions.f90:396-399
    const long long hdr = 8 + 12 + 8ll * (c.nx + c.ny + c.nz);
    const long long fbytes = (long long)c.nx * c.ny * c.nz * 8;
    const int ids[5] = {O3D_F_UX, O3D_F_UY, O3D_F_UZ, O3D_F_PP, O3D_F_PHI};
    double* d[5];
    for (int f = 0; f < 5; ++f)
        if (!(d[f] = field(s, ids[f]))) return O3D_ERR_CUDA;
    int fd;
    if ((rc = open_out(filename, hdr + 5 * fbytes, &fd))) return rc;
    if (c.nranks <= 1 || c.rank == 0) {
        const int n3[3] = {c.nx, c.ny, c.nz};
        long long off = 0;
        int bad = pwrite_all(fd, &time, 8, off);
        off += 8;
        bad |= pwrite_all(fd, n3, 12, off);
        off += 12;
        bad |= pwrite_all(fd, x, 8ull * c.nx, off);
        off += 8ll * c.nx;
        bad |= pwrite_all(fd, y, 8ull * c.ny, off);
        off += 8ll * c.ny;
        bad |= pwrite_all(fd, z, 8ull * c.nz, off);
        if (bad) {
            set_error("Error writing file: %s (%s)", filename, strerror(errno));
            close(fd);
            return O3D_ERR_IO;
        }
    }
    const long long slab = (long long)c.nx * c.ny * 8 * s->z0;
    for (int f = 0; f < 5; ++f) {
        if ((rc = queue_field(s, e, d[f], fd, hdr + f * fbytes + slab, f == 4))) {
            // the five jobs share one descriptor that only the last one closes: drain what was
            // queued, close it here and do not leave a full-size, partly written restart file
            if (f < 4) {
                {
                    std::unique_lock<std::mutex> lk(e->m);
                    e->cv_free.wait(lk, [&] { return e->in_flight == 0; });
                }
                close(fd);
            }
            if (c.nranks <= 1 || c.rank == 0) unlink(filename);
            return rc;
        }
    }
    return O3D_OK;
}

int o3d_s_read_fields(o3d_session* s, const char* filename, double* time, double* x, double* y,
                      double* z) {
    if (!s || !filename || !time || !x || !y || !z) return O3D_ERR_INVALID;
    const o3d_config& c = s->cfg;
    int rc = o3d_s_io_wait(s);  // a restart file may be the one still being written
    if (rc) return rc;
    IoEngine* e;
    if ((rc = io_engine(s, &e))) return rc;
    const int fd = open(filename, O_RDONLY);
    if (fd < 0) {
        set_error("Error opening file: %s (%s)", filename, strerror(errno));
        return O3D_ERR_IO;
    }
    double t;
    int n3[3];
    if (pread_all(fd, &t, 8, 0) || pread_all(fd, n3, 12, 8)) {
        set_error("Error reading file: %s (short header)", filename);
        close(fd);
        return O3D_ERR_IO;
    }
    if (n3[0] != c.nx || n3[1] != c.ny || n3[2] != c.nz) {
        // src/IOfunctions.f90:452-459: reported, nothing is read
        set_error("number of cells are different in parameters and fields.bin: nx %d/%d ny %d/%d "
                  "nz %d/%d", c.nx, n3[0], c.ny, n3[1], c.nz, n3[2]);
        close(fd);
        return O3D_ERR_INVALID;
    }
    long long off = 20;
    int bad = pread_all(fd, x, 8ull * c.nx, off);
    off += 8ll * c.nx;
    bad |= pread_all(fd, y, 8ull * c.ny, off);
    off += 8ll * c.ny;
    bad |= pread_all(fd, z, 8ull * c.nz, off);
    off += 8ll * c.nz;
    const long long fbytes = (long long)c.nx * c.ny * c.nz * 8;
    const long long slab = (long long)c.nx * c.ny * 8 * s->z0;
    const int ids[5] = {O3D_F_UX, O3D_F_UY, O3D_F_UZ, O3D_F_PP, O3D_F_PHI};
    for (int f = 0; f < 5 && !bad; ++f) {
        bad = pread_all(fd, e->buf[0].host, (size_t)s->nloc * 8, off + f * fbytes + slab);
        if (!bad && (rc = o3d_upload(s, ids[f], e->buf[0].host))) {
            close(fd);
            return rc;
        }
    }
    close(fd);
    if (bad) {
        set_error("Error reading file: %s (truncated)", filename);
        return O3D_ERR_IO;
    }
    *time = t;
    return O3D_OK;
}

int o3d_s_write_binary(o3d_session* s, const char* filename, int fid) {
    if (!s || !filename || fid < 0 || fid >= O3D_F_COUNT) return O3D_ERR_INVALID;
    IoEngine* e;
    int rc = o3d_sync(s);
    if (rc) return rc;
    if ((rc = io_engine(s, &e))) return rc;
    double* d = field(s, phys_id(s, fid));
    if (!d) return O3D_ERR_CUDA;
    return write_raw(s, e, filename, d);
}

int o3d_s_write_all_data(o3d_session* s, const char* dir, int num) {
    if (!s || !dir) return O3D_ERR_INVALID;
    const o3d_config& c = s->cfg;
    IoEngine* e;
    int rc = o3d_sync(s);  // a diverged state is reported, not written (see o3d_s_save_fields)
    if (rc) return rc;
    if ((rc = io_engine(s, &e))) return rc;
    mkdir(dir, 0755);  // the reference's check_directories(); EEXIST is fine
    auto path = [&](const char* name) {
        return std::string(dir) + "/" + name + "_" + std::to_string(num) + ".bin";
    };
    // src/visualization.f90:251-274, same order
    const char* names[4] = {"ux", "uy", "uz", "pp"};
    const int ids[4] = {O3D_F_UX, O3D_F_UY, O3D_F_UZ, O3D_F_PP};
    for (int f = 0; f < 4; ++f) {
        double* d = field(s, ids[f]);
        if (!d) return O3D_ERR_CUDA;
        if ((rc = write_raw(s, e, path(names[f]), d))) return rc;
    }
    // vort = sqrt(rotx**2 + roty**2 + rotz**2) of rotational() and the Q criterion, computed
    // here (src/osinco3d_main.f90:130-133 calls both right before write_all_data)
    if ((rc = o3d_s_vorticity_magnitude(s, O3D_F_SCRATCH2))) return rc;
    if ((rc = write_raw(s, e, path("vort"), field(s, O3D_F_SCRATCH2)))) return rc;
    if ((rc = o3d_s_q_criterion(s, O3D_F_SCRATCH2))) return rc;
    if ((rc = write_raw(s, e, path("qcrit"), field(s, O3D_F_SCRATCH2)))) return rc;
    if (c.nscr == 1) {
        double* d = field(s, O3D_F_PHI);
        if (!d) return O3D_ERR_CUDA;
        if ((rc = write_raw(s, e, path("phi"), d))) return rc;
    }
    if (c.iles == 1) {
        double* d = field(s, O3D_F_NU_T);
        if (!d) return O3D_ERR_CUDA;
        if ((rc = write_raw(s, e, path("nu_t"), d))) return rc;
    }
    return O3D_OK;
}

}  // extern "C"
