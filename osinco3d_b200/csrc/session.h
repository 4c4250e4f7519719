// session.h -- internal definition of o3d_session (device-resident state of one rank).
#pragma once
#include <vector>

#include "../../include/o3d_b200.h"
#include "kernels.h"

namespace o3d {
struct Comm;  // z-slab halo exchange + reductions over NCCL (comm.cu)
}

struct o3d_session {
    o3d_config cfg;
    o3d::Dims g;           // local slab
    int z0, nzl;           // owned global planes [z0, z0+nzl)
    long long plane, nloc; // nx*ny, nx*ny*nzl
    cudaStream_t st;
    o3d::Coef cx, cy, cz;

    // physical buffers: [0, O3D_F_COUNT) for plain fields; history levels are logical views
    double* base[O3D_F_COUNT];  // allocation start (3 ghost planes below plane 0)
    // history: physical buffer ids per component (0..2 = fux,fuy,fuz; 3 = fphi), logical level
    int lv[4][3];

    // SOR
    o3d::SorCtrl* ctrl_d;
    o3d::SorCtrl* ctrl_h;  // pinned
    int sor_variant;       // 0: _0000, 1: _0011, 2: _111111
    int last_iters;
    double omega;

    int* flag_d;
    int* flag_h;           // pinned
    double* partial;       // reduction scratch
    long long partial_n;
    double* scal_d;        // 64 device doubles
    double* scal_h;        // pinned mirror

    o3d::Comm* comm;
    int use_src;           // transeq source term uploaded to O3D_F_SCRATCH1

    // timers
    int timers_on;
    struct Span { cudaEvent_t a, b; int stage; };
    std::vector<Span> pending;
    std::vector<cudaEvent_t> free_events;
    double t_ms[6];
    cudaEvent_t sw_a, sw_b;  // stopwatch
    long long t_cnt[6];
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

// lazily allocated, zero-initialised field; returns pointer to owned plane 0 (nullptr on OOM)
double* field(o3d_session* s, int id);
// logical history level (1..3) of component c (0..2 velocity, 3 scalar) -> field id
int hist_id(const o3d_session* s, int c, int level);
int ensure_partial(o3d_session* s, long long n);
void fill_dims(o3d_session* s);

void span_begin(o3d_session* s, int stage);
void span_end(o3d_session* s, int stage, long long count);

// Poisson solvers (poisson.cu)
int sor_solve(o3d_session* s, double* pp, const double* rhs, int* iters, double* dmax);
int mg_solve(o3d_session* s, double* pp, const double* rhs, int nlevels, int npre, int npost,
             double tol, int* cycles, double* dmax);

// comm.cu
int comm_create(o3d_session* s);
void comm_destroy(o3d_session* s);
// exchange the 3 ghost planes per side of `nf` fields (no-op when nranks == 1)
int comm_exchange(o3d_session* s, double* const* fields, int nf, int width);
int comm_allreduce(o3d_session* s, double* dev, int n, int op /* RED_* */);
int comm_exchange_w(o3d_session* s, double* const* fields, int nf, int width, int wrap);
int nccl_unique_id(unsigned char* out128);

}  // namespace o3d
