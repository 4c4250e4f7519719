// session.h -- internal definition of o3d_session (device-resident state of one rank).
#pragma once
#include <functional>
#include <vector>

#include "../../include/o3d_b200.h"
#include "kernels.h"

namespace o3d {
struct Comm;  // z-slab halo exchange + reductions over NCCL (comm.cu)
struct MgHierarchy;  // level arrays + transfer tables of the V-cycle (multigrid.cu)
struct IoEngine;     // asynchronous field output: staging buffers, I/O stream, writer thread (io.cu)
struct Pipe;         // copy streams + staging slots of the pipelined host-pointer procedures (pipeline.cu)
struct PeerState;    // CUDA-IPC mappings of the other ranks' pp buffers and sync blocks (comm.cu)
}

struct o3d_session {
    o3d_config cfg;
    o3d::Geom g;           // local slab, padded layout
    int z0, nzl;           // owned global planes [z0, z0+nzl)
    long long nloc;        // interior points of this rank: nx*ny*nzl
    long long felems;      // doubles per padded field allocation
    cudaStream_t st;
    // halo exchanges run on their own stream so that they overlap interior compute: ev_ready is
    // recorded on st when the planes to send are final, ev_halo on st_comm when ghosts arrived
    cudaStream_t st_comm;
    cudaEvent_t ev_ready, ev_halo;
    int halo_pending;      // an exchange is in flight: comm_wait() before reading z ghosts
    o3d::Coef cx, cy, cz;

    // padded fields; history levels are logical views onto three physical buffers
    double* base[O3D_F_COUNT];       // allocation start (ghosts included)
    CUtensorMap tmap[O3D_F_COUNT];   // 40 x 14 x 1 boxes for the march engine
    CUtensorMap tmap_sor[O3D_F_COUNT];  // 36 x 20 x 1 boxes for the fused SOR pass (lazy)
    CUtensorMap tmap_st[O3D_F_COUNT];   // 32 x 8 x 1 boxes: stream operands of the march engine
    unsigned char tmap_sor_ok[O3D_F_COUNT];
    // ghost-cell state per field: which axes currently hold a valid closure and with which
    // parity bits (bit a: odd along axis a).  Producers / uploads reset gaxes to 0.
    // gaxes: 0x1 / 0x2 / 0x4 = faces of x / y / z valid (z including rank-boundary halos);
    // 0x8 = z faces valid on the wall (non-halo) sides; 0x10 = edges and corners valid too.
    unsigned gaxes[O3D_F_COUNT], gpar[O3D_F_COUNT];
    int lv[4][3];          // history: physical buffer per component (0..2 = fu?, 3 = fphi)

    double* stage_d;       // contiguous nloc doubles: H2D / D2H staging of one field

    // SOR
    o3d::SorCtrl* ctrl_d;
    o3d::SorCtrl* ctrl_h;  // pinned
    unsigned long long* seam_sync_d;  // grid-barrier / finish counters of sor_seam_fused_kernel
    // persistent SOR (sor_persist_kernel.cu): grid-barrier words; event recorded once the control
    // block of a finished solve is on its way to the host (the gated correction runs behind it)
    unsigned long long* persist_sync_d;
    cudaEvent_t ev_ctrl;
    int pp_phys;           // which of the two physical pp allocations is O3D_F_PP now (0: the original)
    o3d::PeerState* peers; // z slabs: peer-mapped neighbours, set up lazily by the first solve
    int peers_tried;
    unsigned long long peer_iter_base;  // iterations of all earlier peer-memory solves
    unsigned long long peer_solves;     // peer-memory solves so far
    int last_sor_path;     // bit 0: persistent kernel, bit 1: peer-memory halos (last solve)
    int sor_variant;       // 0: _0000, 1: _0011, 2: _111111
    int last_iters;
    double omega;

    int* flag_d;
    int* flag_h;           // pinned
    // o3d_step does not stall on the NaN / >1000 guard of correct_velocity: the flag copy is left
    // in flight and examined at the next host synchronisation (poll_flag)
    int flag_pending, diverged;
    // speculative projection (o3d_step; single rank, fused TMA SOR): the correction kernel is
    // queued right behind the first batch of SOR passes, gated on the device by the solver's
    // control block, so the GPU does not idle while the host polls for convergence.
    // spec_arm: o3d_step allows it for this solve; spec_state: 1 = launched, 2 = took effect
    int spec_arm, spec_state;
    double* partial;       // reduction scratch
    long long partial_n;
    double* scal_d;        // 64 device doubles
    double* scal_h;        // pinned mirror

    o3d::Comm* comm;
    o3d::MgHierarchy* mg;  // built lazily by mg_solve, cached across steps
    o3d::IoEngine* io;     // built lazily by the first output call
    o3d::Pipe* pipe;       // built lazily by the first pipelined host-pointer call
    int use_src;           // transeq source term uploaded to O3D_F_SCRATCH1

    // timers
    int timers_on;
    struct Span { cudaEvent_t a, b; int stage; };
    std::vector<Span> pending;
    std::vector<cudaEvent_t> free_events;
    double t_ms[6];
    long long t_cnt[6];
    cudaEvent_t sw_a, sw_b;  // stopwatch
    // O3D_TRACE=<step>: event timeline of one o3d_step across both streams, printed to stderr
    int trace_step, trace_on, step_count;
    struct Mark { cudaEvent_t e; const char* name; int comm; };
    std::vector<Mark> marks;
};

namespace o3d {

// module `initialization` state bound by o3d_schemes() (src/initialization.f90:226-304)
struct Schemes {
    int bound = 0;
    int flags[6] = {1, 1, 1, 1, 1, 1};  // nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn
    int bc[3] = {1, 1, 1};
    int sim2d = 0;
    int poisson_variant = 2;  // -1: null pointer
};
extern Schemes g_schemes;
extern int g_sor_order;
int ensure_device();
int poisson_variant_of(int bx1, int bxn, int by1, int byn);
int axis_bc(int b1, int bn, int* out);

enum { ST_RHS = 0, ST_DIV = 1, ST_SOR = 2, ST_CORR = 3, ST_TRANSEQ = 4, ST_HALO = 5 };

// lazily allocated, zero-initialised padded field; returns the interior origin (nullptr on OOM)
double* field(o3d_session* s, int id);
FieldRef fref(o3d_session* s, int id);
// logical history level (1..3) of component c (0..2 velocity, 3 scalar) -> field id
int hist_id(const o3d_session* s, int c, int level);
int phys_id(const o3d_session* s, int id);
// parity bits the reference's tables give a field (src/integration.f90:118-165): the velocity
// component normal to an axis is odd along it, everything else even
unsigned natural_parity(int id);
// mark the interior of a field as modified: its ghost cells are stale
void touch(o3d_session* s, int id);
// make the ghost cells of `n` fields valid on the axes in `axes` with parity `par[q]`
// (one fused launch + z-slab halo exchange where the z neighbour is another rank)
int ensure_ghosts(o3d_session* s, const int* ids, int n, const unsigned* par, unsigned axes,
                  bool defer = false);
int ensure_ghosts1(o3d_session* s, int id, unsigned par, unsigned axes, bool defer = false);
int ensure_ghosts_own_axis(o3d_session* s, const int* ids, const unsigned* par,
                           bool defer = false);
// x, y and wall-side z ghosts (plus edges / corners if `edges`) valid, WITHOUT any exchange
int ensure_local_ghosts(o3d_session* s, int id, unsigned par, bool edges);
const CUtensorMap* sor_tmap(o3d_session* s, int id);
int ensure_partial(o3d_session* s, long long n);
void fill_geom(o3d_session* s);
// call right after a synchronisation of the session stream
void poll_flag(o3d_session* s);

// timeline mark on the session stream (comm = 0) or the communication stream (comm = 1)
void trace_mark(o3d_session* s, int comm, const char* name);
void span_begin(o3d_session* s, int stage);
void span_end(o3d_session* s, int stage, long long count);

// gated launch of the projection correction behind the SOR passes already queued (api.cu)
int spec_correct_launch(o3d_session* s);

// the two halves of o3d_s_predict_velocity around its kernel launch (api.cu)
int rhs_prepare(o3d_session* s, int itime, RhsArgs& a, int* tgt);
void rhs_finish(o3d_session* s, const int* tgt, bool iles);

// Pipelined host-pointer procedures (pipeline.cu): the host arrays are cut into z chunks and
// upload / kernel / download of successive chunks overlap on three streams (full-duplex PCIe).
// pipe_chunks(nz): number of chunks for a grid of nz planes (0 = pipelining off / grid too thin).
int pipe_chunks(int nz);
int pipe_predict_velocity(o3d_session* s, int itime, double* const* up_h, const double* const* u_h,
                          double* const* f_h, double* nu_t_h);
int pipe_correct_velocity(o3d_session* s, double* const* u_h, const double* const* up_h,
                          const double* pp_h);
void pipe_destroy(o3d_session* s);

// Poisson solvers (poisson.cu)
int sor_solve(o3d_session* s, double* pp, const double* rhs, int* iters, double* dmax);
// fixed number of fused red-black sweeps on O3D_F_PP with the relaxation factor of `ctrl` and no
// exit tests (multigrid smoother, single rank); O3D_ERR_UNSUPPORTED: use the in-place sweeps
int sor_fixed_sweeps(o3d_session* s, const double* rhs, int sweeps, SorCtrl* ctrl);
int mg_solve(o3d_session* s, double* pp, const double* rhs, int nlevels, int npre, int npost,
             double tol, int* cycles, double* dmax);
void mg_destroy(o3d_session* s);
// drains the output queue, joins the writer thread, frees the staging buffers (io.cu)
void io_destroy(o3d_session* s);
// shared with poisson.cu
SorArgs make_sor_args(o3d_session* s, double* pp, const double* rhs);

// comm.cu
int comm_create(o3d_session* s);
void comm_destroy(o3d_session* s);
// exchange `width` ghost planes per side of `nf` fields given by their ALLOCATION BASE
// (no-op when nranks == 1); wrap: the slab ring is periodic in z
int comm_exchange(o3d_session* s, double* const* bases, int nf, int width, int wrap);
// same, on the communication stream, per-field widths; returns immediately.  comm_wait() makes the
// session stream wait for the ghosts (call it before the first kernel that reads z ghost planes)
int comm_exchange_async(o3d_session* s, double* const* bases, const int* widths, int nf, int wrap);
int comm_wait(o3d_session* s);
// planes at each end of a slab that the boundary launches of a split kernel cover (0: the slab
// is too thin or single-rank -> no split, blocking exchange)
int split_edge(const o3d_session* s);
int launch_overlapped(o3d_session* s, const std::function<int(cudaStream_t, int, int)>& launch);
int comm_allreduce(o3d_session* s, double* dev, int n, int op /* RED_* */);
// Collective, lazy: map the neighbours' two pp allocations and every rank's PeerBlock through CUDA
// IPC.  Returns 0 and fills `out` for this rank when the peer path is usable (all ranks agree),
// non-zero otherwise (the solver then keeps the NCCL path).  p_phys0 / p_phys1: allocation bases
// of this rank's two physical pp buffers.
int comm_peer_setup(o3d_session* s);
// PeerSync for a solve whose first iteration reads the buffer that is O3D_F_PP now
int comm_peer_args(o3d_session* s, PeerSync* out);
// swap the roles of O3D_F_PP and O3D_F_PP2 (O(1); after an odd number of ping-pong passes)
void swap_pp(o3d_session* s);
// rank r produced buf[first[r] .. first[r] + count[r]); replicate all chunks on every rank
int comm_allgather_chunks(o3d_session* s, double* buf, const long long* first,
                          const long long* count);
int nccl_unique_id(unsigned char* out128);

}  // namespace o3d
