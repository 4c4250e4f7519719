// kernels.h -- host-side launch interface of the CUDA kernels (internal to libo3d_b200).
// All field pointers address the INTERIOR origin (0,0,0) of a padded field (o3d_common.cuh).
#pragma once
#include <cuda.h>

#include <cstdlib>

#include "o3d_common.cuh"

namespace o3d {

// a padded field and its TMA tensor map (40 x 14 x 1 boxes, march.cuh)
struct FieldRef {
    double* p;
    const CUtensorMap* tm;   // 40 x 14 x 1 boxes (tile + 3-cell x/y halo)
    const CUtensorMap* tms;  // 32 x 8 x 1 boxes (tile only: stream operands)
};

// z-chunking shared by the z-marching kernels.  A CTA marches over one chunk of `zchunk` planes
// and pays per chunk the equivalent of `extra` additional planes of its normal per-plane cost
// (the stencil window / pipeline prologue is loaded but not computed on: for a kernel that moves
// `streams` values per point of which `nfz` carry a z window, 6 * nfz / streams planes).  CTAs are
// scheduled dynamically on `slots` resident positions (148 SMs x CTAs per SM), so in units of one
// plane per slot the run time is about
//     tiles * nch * (zchunk + extra) / slots   +   (zchunk + extra) / 2
// (throughput term + half a chunk of tail): more chunks shorten the tail and balance the SMs,
// fewer chunks save overhead planes.  The number of chunks minimises that estimate; measured
// sweeps over the chunk count at 256^3 and 512^3 (profiles/r1o_nch_sweep.txt) follow it.
inline int pick_zchunk_slots(int tiles_xy, int nz, int slots, double extra) {
    {   // tuning override: O3D_NCH_S (SOR pass), O3D_NCH_2 / O3D_NCH_3 (march kernels with 2 / 3
        // resident CTAs per SM) = number of z chunks
        const char* e = getenv(extra > 3.0 ? "O3D_NCH_S" : (slots == 148 * 2 ? "O3D_NCH_2" : "O3D_NCH_3"));
        if (e && atoi(e) > 0) {
            int nch = atoi(e);
            if (nch > nz / 8) nch = nz / 8 > 0 ? nz / 8 : 1;
            return (nz + nch - 1) / nch;
        }
    }
    int best = nz;
    double best_cost = 1e300;
    int max_chunks = nz / 8;
    if (max_chunks < 1) max_chunks = 1;
    if (max_chunks > 128) max_chunks = 128;
    for (int nch = 1; nch <= max_chunks; ++nch) {
        const int zc = (nz + nch - 1) / nch;
        const int real = (nz + zc - 1) / zc;
        const double work = (double)tiles_xy * real * (zc + extra);
        const double cost = work / slots + 0.5 * (zc + extra);
        if (cost < best_cost - 1e-9) best_cost = cost, best = zc;
    }
    return best;
}
// march kernels: nfz fields with a 7-plane z window out of `streams` values moved per point
inline int pick_zchunk(int tiles_xy, int nz, int ctas_per_sm, int nfz, int streams) {
    return pick_zchunk_slots(tiles_xy, nz, 148 * ctas_per_sm, 6.0 * nfz / streams);
}

// ---- padded-layout maintenance (ghost_kernels.cu) ----
// extra bits of GhostJob::axes: fill the z ghosts of one side only (pipeline.cu: the source planes
// of the other side have not been uploaded yet)
enum : unsigned { GHOST_Z_LO_ONLY = 0x10u, GHOST_Z_HI_ONLY = 0x20u };
struct GhostJob {
    double* p;
    unsigned par;   // bit a set: odd along axis a (der?i_11), else even (der?p_11);
                    // bit a + 4 set: Dirichlet wall along axis a (odd about the stored wall value)
    unsigned axes;  // bit a set: fill the ghosts of axis a (a = 0..2) | GHOST_Z_*_ONLY
};
struct GhostArgs {
    GhostJob job[6];
    int njobs;
};
int launch_fill_ghosts(cudaStream_t st, const Geom& g, const GhostArgs& a);
// faces + edges + corners, axis by axis (x, then y over the x ghosts, then z over both)
int launch_fill_ghosts_full(cudaStream_t st, const Geom& g, double* p, unsigned par);
int launch_pack(cudaStream_t st, const Geom& g, const double* contiguous, double* padded);
int launch_unpack(cudaStream_t st, const Geom& g, const double* padded, double* contiguous);
// planes [k0, k0 + nk) only; `chunk` = contiguous (nx, ny, nk)
int launch_pack_planes(cudaStream_t st, const Geom& g, const double* chunk, double* padded, int k0,
                       int nk);
int launch_unpack_planes(cudaStream_t st, const Geom& g, const double* padded, double* chunk,
                         int k0, int nk);

// ---- single-axis derivative, operator form (src/derivation.f90, der_type) ----
// the closure is whatever the ghost cells of f hold; zero != 0 -> der?_2dsim
int launch_der(cudaStream_t st, const Geom& g, int axis, int order, int zero, double d,
               const double* f, double* df);

// ---- fused RHS + nu_t + predictor (src/integration.f90:14-197) ----
struct RhsArgs {
    FieldRef u[3];        // ux, uy, uz with natural-parity ghosts
    const double* f2[3];  // history level 2 (previous step's f)
    const double* f3[3];  // history level 3 (may alias f1: read before written, per point)
    double* f1[3];        // new f
    double* up[3];        // u*
    double* nu_t;         // written only when iles
    Coef cx, cy, cz;
    double onere, adu, bdu, cdu, csd2;
    int iles;
};
int launch_rhs(cudaStream_t st, const Geom& g, const RhsArgs& a, int zmode = 0, int zedge = 0);

// ---- Smagorinsky nu_t alone (src/les_turbulence.f90:10-97) ----
int launch_nu_t(cudaStream_t st, const Geom& g, const FieldRef* u, const Coef& cx,
                const Coef& cy, const Coef& cz, double csd2, double* nu_t);

// ---- divergence (src/differential_operators.f90:7-38) [/dt -> Poisson rhs,
//      src/integration.f90:239]; f[0] needs x ghosts, f[1] y ghosts, f[2] z ghosts ----
int launch_div(cudaStream_t st, const Geom& g, const FieldRef* f, const Coef& cx, const Coef& cy,
               const Coef& cz, int divide_by_dt, double dt, double* out, int zmode = 0,
               int zedge = 0);

// ---- projection correction (src/integration.f90:257-330) ----
// flag: device int, OR-ed with 1 when a NaN or a value > 1000 is produced.
// gate != null: the launch does nothing unless gate->done is set, and reads pp_alt instead of pp
// when gate->iter is odd (queued behind a batch of ping-pong SOR passes, see sor_solve)
struct SorCtrl;
int launch_corr(cudaStream_t st, const Geom& g, const FieldRef& pp, const FieldRef* up,
                double* const* u, const Coef& cx, const Coef& cy, const Coef& cz, double dt,
                int* flag, int zmode = 0, int zedge = 0, const FieldRef* pp_alt = nullptr,
                const SorCtrl* gate = nullptr);

// ---- curl and Q criterion (src/differential_operators.f90:40-108) ----
int launch_rot(cudaStream_t st, const Geom& g, const FieldRef* u, const Coef& cx, const Coef& cy,
               const Coef& cz, double* rotx, double* roty, double* rotz);
int launch_qcrit(cudaStream_t st, const Geom& g, const FieldRef* u, const Coef& cx,
                 const Coef& cy, const Coef& cz, double* q);
// |curl u| as output by write_all_data (src/visualization.f90:258-259); even ghosts as launch_rot
int launch_vort(cudaStream_t st, const Geom& g, const FieldRef* u, const Coef& cx, const Coef& cy,
                const Coef& cz, double* vm);

// ---- scalar transport (src/integration.f90:332-468) ----
struct TranseqArgs {
    FieldRef phi;          // phi^n with even ghosts
    const double* u[3];
    const double* nu_t;    // used when iles
    const double* src;     // may be null
    const double* f2;
    const double* f3;      // may alias f1
    double* f1;
    double* phi_new;       // unclipped phi^{n+1} (different buffer than phi)
    double* partial;       // [3 * nblocks] per-CTA partial sums: old, clipped, weight
    Coef cx, cy, cz;
    double resc, sc, adu, bdu, cdu;
    int iles;
};
int transeq_blocks(const Geom& g);
int launch_transeq_rhs(cudaStream_t st, const Geom& g, const TranseqArgs& a);
// sums[0..2] = device doubles {sum_old, sum_clipped, sum_weight}; count = global N
int launch_transeq_clip(cudaStream_t st, const Geom& g, const double* phi_new, double* phi,
                        const double* sums, double count);

// ---- SOR (src/poisson.f90) ----
struct SorCtrl {
    unsigned long long dmax_bits;  // running max|p_new - p| of the current iteration
    double omega;
    double dmax_old;
    double dmax_last;
    int iter;         // sweeps completed
    int done;         // 1: dmax < eps, 2: stall exit, 3: kmax reached
    int pad0, pad1;
};
struct SorArgs {
    double* pp;
    const double* rhs;
    double oneondx2, oneondy2, oneondz2, A, invA;
    int mx, my;              // neighbour rule in x,y: BM_WRAP | BM_MIRROR
    int mz_lo, mz_hi;        // z: BM_WRAP | BM_MIRROR | BM_HALO
    int nx, ny, nz;
    long long sy, sz;        // padded strides
    int gz0;                 // global index of local plane 0 (colouring)
    int seam_x, seam_y, seam_z;  // odd periodic extents: last plane is a seam
    int gnz;
};
// one colour class: colour in {0,1}, seam_class in {0,1} (popcount parity of the seam mask);
// images != 0 (seam class only): also store the ghost images of every point written, so the
// sweep can follow the fused TMA pass on its output buffer
int launch_sor_rb(cudaStream_t st, const SorArgs& a, int colour, int seam_class, SorCtrl* ctrl,
                  int images = 0);
// single rank: both odd seam classes (with ghost images) and the end-of-iteration control in one
// cooperative launch; sync = 2 zero-initialised device counters that persist across launches
int launch_sor_seam_fused(cudaStream_t st, const SorArgs& a, SorCtrl* ctrl,
                          unsigned long long* sync, double eps, int kmax, int idyn, double factor);
// fused red+black iteration (one pass, ping-pong p_old -> p_new); needs a 2-colourable grid
// zmode / zedge: split launch as for the march kernels (0 = whole slab)
int launch_sor_fused(cudaStream_t st, const SorArgs& a, const double* p_old, double* p_new,
                     SorCtrl* ctrl, int zmode = 0, int zedge = 0);
// TMA-staged fused red+black iteration (sor_tma_kernel.cu): p_old and rhs are read through
// tensor maps with sor_tma_box_x() x sor_tma_box_y() x 1 boxes and must carry valid ghost cells
// (p_old: faces + edges, 2 deep; rhs: faces, 1 deep); p_new is written with its ghost images
int sor_tma_box_x();
int sor_tma_box_y();
int launch_sor_tma(cudaStream_t st, const SorArgs& a, const CUtensorMap* p_old_map,
                   const CUtensorMap* rhs_map, double* p_new, int bx, int by, int bz_lo,
                   int bz_hi, SorCtrl* ctrl, int zmode = 0, int zedge = 0);
// ---- persistent SOR: all iterations of a solve in one cooperative launch (sor_persist_kernel.cu)
// Synchronisation block of one rank of a z-slab run.  It lives in that rank's device memory and is
// mapped by every other rank of the node (CUDA IPC): neighbours count the ghost planes they have
// stored into this rank's pp buffers, all ranks deposit their residual maxima.
struct PeerBlock {
    unsigned long long halo_cnt[2];       // planes received into my low / high ghost planes (monotone)
    unsigned long long init_cnt[2];       // solves whose initial ghost planes have arrived (low / high)
    unsigned long long seam_cnt[2];       // odd seam sweeps (2 per iteration) the neighbour has delivered
    unsigned long long pad0[10];
    // residual maxima, flag-in-data: [global iteration & 1][2 * source rank + {0, 1}] =
    // {high / low 32 data bits << 32 | 32-bit tag (iteration + 1)}
    unsigned long long dmax_slot[2][32];
};
struct PeerSync {
    int nranks, rank;                     // nranks <= 1: single rank, nothing below is read
    int has_lo, has_hi;                   // a neighbour rank below / above (periodic wrap included)
    int lo_nz;                            // the lower neighbour's number of owned planes
    unsigned long long iter_base;         // iterations all ranks completed in earlier persistent solves
    unsigned long long solve_base;        // earlier peer-memory solves of this session
    int push_init;                        // 1: phase 0 of the kernel delivers the initial ghost planes
    double* lo_rhs;                       // the neighbours' right-hand-side fields (interior origins)
    double* hi_rhs;
    PeerBlock* mine;
    PeerBlock* lo;
    PeerBlock* hi;
    PeerBlock* all[16];                   // every rank's block (all[rank] == mine)
    double* lo_p[2];                      // the neighbours' ping-pong buffers (interior origins), in
    double* hi_p[2];                      //   the same role order as the local p0 / p1
};
// 0: launched; 1: CUDA error; 2: not applicable (no cooperative launch / slabs with seams) -> the
// caller uses the launch-per-pass path.  Runs until ctrl->done or max_iters iterations; iteration t
// reads p[first_src ^ (t & 1)].  sync: >= 32 device words (zeroed here).  fixed != 0: smoother
// mode, no exit tests.
int sor_persist_available();
int launch_sor_persist(cudaStream_t st, const SorArgs& a, const CUtensorMap* pmap0,
                       const CUtensorMap* pmap1, const CUtensorMap* rhs_map, double* p0, double* p1,
                       int first_src, int bx, int by, int bz_lo, int bz_hi, SorCtrl* ctrl,
                       unsigned long long* sync, int max_iters, double eps, int kmax, int idyn,
                       double factor, int fixed, const PeerSync* peer);
// end-of-iteration control: exits and dynamic omega, src/poisson.f90:110-122
int launch_sor_control(cudaStream_t st, SorCtrl* ctrl, double eps, int kmax, int idyn,
                       double factor);
// verification ordering: one hyperplane i+j+k = h of the lexicographic sweep
int launch_sor_wavefront(cudaStream_t st, const SorArgs& a, int h, SorCtrl* ctrl);

// ---- geometric multigrid (mg_kernels.cu; replaces src/poisson_multigrid.f90) ----
struct MgGrid {
    int nx, ny, nz;
    long long sy, sz;     // element strides (level 0: padded field; coarser levels: compact)
    int mx, my;           // neighbour rule in x, y: BM_WRAP | BM_MIRROR
    int mz_lo, mz_hi;     // z: BM_WRAP | BM_MIRROR | BM_HALO (level 0 of a z-slab run: the ghost
                          // planes of the padded field, filled by the halo exchange)
    double ox, oy, oz;    // 1/d^2 per axis, src/poisson.f90:42-47
    double A, invA;       // -(2ox + 2oy + 2oz), :48-51
};
// 1-D transfer tables between a level and the next coarser one (device pointers, per axis)
struct MgTables {
    const int* c0[3];     // [n_fine]   prolongation: lower coarse index
    const double* w[3];   // [n_fine]   weight of the upper coarse index
    const int* ridx[3];   // [n_coarse][4] restriction: fine indices
    const double* rw[3];  // [n_coarse][4] restriction: weights (row sum 1, may be 0)
};
int launch_mg_residual(cudaStream_t st, const MgGrid& g, const double* p, const double* rhs,
                       double* res, unsigned long long* maxbits);
// coarse planes [ck0, ck0 + nck) only (a z-slab rank restricts the coarse planes it owns; the
// z entries of t.ridx are then plane offsets relative to the rank's first fine plane)
int launch_mg_restrict(cudaStream_t st, const MgGrid& f, const MgGrid& c, const MgTables& t,
                       const double* res, double* rhs_c, double* p_c, int ck0, int nck);
// kz0: global index of the fine grid's plane 0 (z-slab runs; the z tables are global)
int launch_mg_prolong(cudaStream_t st, const MgGrid& f, const MgGrid& c, const MgTables& t,
                      const double* e, double* p, int kz0);
int launch_mg_coarse(cudaStream_t st, const MgGrid& g, double* p, double* rhs, int sweeps);

// ---- reductions over the interior ----
enum { RED_MIN = 0, RED_MAX = 1, RED_SUM = 2, RED_ABSMAX = 3, RED_MAXBITS = 10 };
int reduce_blocks(const Geom& g);
// partial: >= reduce_blocks(g) doubles of scratch; out: one device double
int launch_reduce(cudaStream_t st, const Geom& g, const double* f, int op, double* partial,
                  double* out);
// function_stats (src/functions.f90:27): out6 device doubles; partial >= 4 * 296 doubles
int launch_function_stats(cudaStream_t st, const Geom& g, const double* f, double* partial,
                          double* out6);
// calculate_residuals (src/utils.f90:93-160) over the interior points of the GLOBAL grid that
// this rank owns: out9 device doubles = 3 sums of squares | 3 maxima | 3 global linear indices
// (as doubles, -1 if none) of the last point attaining the maximum; partial >= 9 * 296 doubles
int launch_residuals(cudaStream_t st, const Geom& g, const double* const* unew,
                     const double* const* uold, double two_dt, double* partial, double* out9);
// sum `nparts` partial triples deterministically: out[c] = sum_b partial[c*nparts + b]
int launch_sum_partials(cudaStream_t st, const double* partial, int nparts, int ncomp,
                        double* out);
// per-step driver diagnostics of one velocity triple in one pass (vel_kernels.cu, DiagEpi):
// out13 = min, max, sum, first arg-max position (global array order, as a double) of
// divergence(odd = 1) | minval u[3] | maxval u[3] | maxval |u|[3]; u needs own-axis odd ghosts;
// partial >= 13 * diag_blocks(g) doubles
int diag_blocks(const Geom& g);
int launch_diag(cudaStream_t st, const Geom& g, const FieldRef* u, const Coef& cx, const Coef& cy,
                const Coef& cz, double* partial, double* out13);
// statistics_calc (src/utils.f90:243): 16 sums
int stats_blocks(const Geom& g);
int launch_stats(cudaStream_t st, const Geom& g, const FieldRef* u, const Coef& cx,
                 const Coef& cy, const Coef& cz, double xnu, double* partial);

}  // namespace o3d
