// kernels.h -- host-side launch interface of the CUDA kernels (internal to libo3d_b200).
#pragma once
#include "o3d_common.cuh"

namespace o3d {

// z-chunking heuristic shared by the z-marching kernels: enough CTAs for >= ~8 waves on
// 148 SMs x 2 resident CTAs, but chunks of at least 16 planes so the 6 extra window planes
// per chunk stay a small overhead.
inline int pick_zchunk(int tiles_xy, int nz) {
    const int target_ctas = 148 * 2 * 8;
    int nchunks = (target_ctas + tiles_xy - 1) / tiles_xy;
    int max_chunks = nz / 16;
    if (max_chunks < 1) max_chunks = 1;
    if (nchunks > max_chunks) nchunks = max_chunks;
    if (nchunks < 1) nchunks = 1;
    return (nz + nchunks - 1) / nchunks;
}

// ---- single-axis derivative, operator form (src/derivation.f90, der_type) ----
// mode/parity describe the closure: BM_WRAP (der?_00), BM_MIRROR + parity 0 (der?p_11) or
// parity 1 (der?i_11).  zero != 0 -> der?_2dsim.
int launch_der(cudaStream_t st, const Dims& g, int axis, int order, int parity, int zero,
               double d, const double* f, double* df);

// ---- fused RHS + nu_t + predictor (src/integration.f90:14-197) ----
struct RhsArgs {
    const double* u[3];   // ux, uy, uz
    const double* f2[3];  // history level 2 (previous step's f)
    const double* f3[3];  // history level 3 (may alias f1: read before written, per point)
    double* f1[3];        // new f
    double* up[3];        // u*
    double* nu_t;         // written only when iles
    Coef cx, cy, cz;
    double onere, adu, bdu, cdu, csd2;
    int iles;
    int zchunk;  // filled by the launcher
};
int launch_rhs(cudaStream_t st, const Dims& g, const RhsArgs& a);

// ---- Smagorinsky nu_t alone (src/les_turbulence.f90:10-97) ----
int launch_nu_t(cudaStream_t st, const Dims& g, const double* ux, const double* uy,
                const double* uz, const Coef& cx, const Coef& cy, const Coef& cz, double csd2,
                double* nu_t);

// ---- divergence (src/differential_operators.f90:7-38) [/dt -> Poisson rhs,
//      src/integration.f90:239] ----
int launch_div(cudaStream_t st, const Dims& g, const double* fx, const double* fy,
               const double* fz, const Coef& cx, const Coef& cy, const Coef& cz, int odd,
               int divide_by_dt, double dt, double* out);

// ---- projection correction (src/integration.f90:257-330) ----
// flag: device int, OR-ed with 1 when a NaN or a value > 1000 is produced.
int launch_corr(cudaStream_t st, const Dims& g, const double* pp, const double* const* up,
                double* const* u, const Coef& cx, const Coef& cy, const Coef& cz, double dt,
                int* flag);

// ---- curl and Q criterion (src/differential_operators.f90:40-108) ----
int launch_rot(cudaStream_t st, const Dims& g, const double* ux, const double* uy,
               const double* uz, const Coef& cx, const Coef& cy, const Coef& cz, double* rotx,
               double* roty, double* rotz);
int launch_qcrit(cudaStream_t st, const Dims& g, const double* ux, const double* uy,
                 const double* uz, const Coef& cx, const Coef& cy, const Coef& cz, double* q);

// ---- scalar transport (src/integration.f90:332-468) ----
struct TranseqArgs {
    const double* phi;     // phi^n (with z ghosts if BM_HALO)
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
    int zchunk;
};
int transeq_blocks(const Dims& g);
int launch_transeq_rhs(cudaStream_t st, const Dims& g, const TranseqArgs& a);
// sums[0..2] = device doubles {sum_old, sum_clipped, sum_weight}; count = global N
int launch_transeq_clip(cudaStream_t st, long long n, const double* phi_new, double* phi,
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
    int gz0;                 // global index of local plane 0 (colouring)
    int seam_x, seam_y, seam_z;  // odd periodic extents: last plane is a seam
    int gnz;
};
// one colour class: colour in {0,1}, seam_class in {0,1} (popcount parity of the seam mask)
int launch_sor_rb(cudaStream_t st, const SorArgs& a, int colour, int seam_class, SorCtrl* ctrl);
// end-of-iteration control: exits and dynamic omega, src/poisson.f90:110-122
int launch_sor_control(cudaStream_t st, SorCtrl* ctrl, double eps, int kmax, int idyn,
                       double factor);
// verification ordering: one hyperplane i+j+k = h of the lexicographic sweep
int launch_sor_wavefront(cudaStream_t st, const SorArgs& a, int h, SorCtrl* ctrl);

// ---- reductions ----
enum { RED_MIN = 0, RED_MAX = 1, RED_SUM = 2, RED_ABSMAX = 3, RED_MAXBITS = 10 };
int reduce_blocks(long long n);
// partial: >= reduce_blocks(n) doubles of scratch; out: one device double
int launch_reduce(cudaStream_t st, const double* f, long long n, int op, double* partial,
                  double* out);
// function_stats (src/functions.f90:27): out6 device doubles
int launch_function_stats(cudaStream_t st, const double* f, int nx, int ny, int nz,
                          double* partial, double* out6);
// sum `nparts` partial triples deterministically: out[c] = sum_b partial[c*nparts + b]
int launch_sum_partials(cudaStream_t st, const double* partial, int nparts, int ncomp,
                        double* out);
// statistics_calc (src/utils.f90:243): 17 sums
int stats_blocks(const Dims& g);
int launch_stats(cudaStream_t st, const Dims& g, const double* ux, const double* uy,
                 const double* uz, const Coef& cx, const Coef& cy, const Coef& cz, double xnu,
                 double* partial);

// ---- misc ----
int launch_fill(cudaStream_t st, double* p, long long n, double v);

}  // namespace o3d
