/*
 * o3d_b200.h -- C ABI of libo3d_b200.so: the B200-native (sm_100a CUDA) implementation of
 * osinco3d's Chorin-projection time step.
 *
 * The reference (jojoledemago/osinco3d) is single-process Fortran 90 with no FFI layer; the
 * seam this library plugs into is the set of Fortran module-procedure interfaces of
 * `derivation`, `diffoper`, `les_turbulence`, `poisson`, `poisson_multigrid` and `integration`.
 * Every entry point below names the reference interface it replaces (file:line, paths
 * relative to the reference checkout).  INTEGRATION.md shows the ISO_C_BINDING shims.
 *
 * Conventions
 *  - plain C types only; arrays are Fortran-ordered real(8) (nx,ny,nz), i fastest, no padding,
 *    exactly the reference's allocatables (src/initialization.f90:157-168);
 *  - every function returns an int status (O3D_OK == 0).  The reference reports errors with
 *    print + stop; the Fortran shim turns a non-zero status into the same print + stop;
 *  - section A ("module procedures") takes HOST pointers and is a stateless drop-in: inputs
 *    are copied to the device, the CUDA kernels run, outputs are copied back.  BC dispatch
 *    follows module `initialization`: call o3d_schemes() once, as the reference calls
 *    schemes(), then the 12 pointer-named derivative entry points and the composite operators
 *    use the bound closures;
 *  - section B ("session") keeps all fields device-resident across calls (what a driver must
 *    use to run at GPU speed); it is also what bench.py times;
 *  - there is NO CPU fallback: without a CUDA device every compute entry point returns
 *    O3D_ERR_NO_DEVICE.
 *  - one host thread per process drives the library; one CUDA device per process.
 */
#ifndef O3D_B200_H
#define O3D_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define O3D_ABI_VERSION 1

/* ---- status codes ---- */
enum {
    O3D_OK = 0,
    O3D_ERR_INVALID = 1,       /* bad argument (null pointer, n < 7, unknown id ...)          */
    O3D_ERR_NO_DEVICE = 2,     /* no CUDA device / driver: there is no CPU fallback            */
    O3D_ERR_CUDA = 3,          /* a CUDA runtime call or kernel failed; see o3d_last_error()   */
    O3D_ERR_BC = 4,            /* boundary flags not PERIODIC/PERIODIC or FREE_SLIP/FREE_SLIP
                                  (schemes() "Unrecognized boundary" stop,
                                  src/initialization.f90:238-242) or null poisson_solver
                                  pointer (:283-301)                                           */
    O3D_ERR_ITSCHEME = 5,      /* "itscheme unrecognized" stop, src/integration.f90:99-104     */
    O3D_ERR_DIVERGED = 6,      /* NaN or max(u) > 1000 abort, src/integration.f90:309-325      */
    O3D_ERR_COMM = 7,          /* NCCL / multi-GPU set-up failure                              */
    O3D_ERR_UNSUPPORTED = 8,
    O3D_ERR_IO = 9             /* "Error opening file" (src/IOfunctions.f90:389-392,
                                  src/visualization.f90:231-234) or a failed read / write       */
};

/* boundary flags, src/initialization.f90 (PERIODIC = 0, FREE_SLIP = 1) */
enum { O3D_PERIODIC = 0, O3D_FREE_SLIP = 1 };

/* closure of one derivative routine in src/derivation.f90 */
enum {
    O3D_CLOSURE_00 = 0,   /* der?_00   periodic                    */
    O3D_CLOSURE_P11 = 1,  /* der?p_11  free-slip, even ghost       */
    O3D_CLOSURE_I11 = 2,  /* der?i_11  free-slip, odd ghost        */
    O3D_CLOSURE_2DSIM = 3,/* derz_2dsim / derzz_2dsim: zeros       */
    /* NEW, no counterpart in the reference: README.md:28 / the north star name a Dirichlet-x
     * closure, but the source accepts PERIODIC / FREE_SLIP only and stops otherwise
     * (src/initialization.f90:228-242).  Dirichlet wall on both faces of the axis: planes 1 and n
     * hold the prescribed values and the field is continued by odd reflection ABOUT them,
     * f(1-g) = 2 f(1) - f(1+g); der?i_11 is the f_wall = 0 special case.  Operator form only
     * (o3d_der); verified against manufactured solutions (tests/test_gpu_dirichlet.py). */
    O3D_CLOSURE_D11 = 4
};

/* SOR sweep ordering (DESIGN.md "Poisson").  The reference sweeps lexicographically
 * (src/poisson.f90:53-105), which has a loop-carried dependence. */
enum {
    O3D_SOR_RED_BLACK = 0,      /* fast path; converges to the same solution within eps        */
    O3D_SOR_LEXI_WAVEFRONT = 1  /* verification mode: hyperplane i+j+k sweep, reproduces the
                                   reference's lexicographic iterates bit for bit               */
};

const char* o3d_last_error(void);
int o3d_abi_version(void);
/* sizeof(o3d_config) as compiled (binding self-check for FFI mirrors of the struct) */
int o3d_config_size(void);
/* number of CUDA devices visible (0 if none / no driver) */
int o3d_device_count(void);
/* select the device used by this process (default 0; multi-GPU: LOCAL_RANK) */
int o3d_set_device(int device);
/* number of kernels this library has launched since load (bench.py "gpu_launches") */
long long o3d_kernel_launches(void);

/* pinned host memory helpers for drivers that want full-speed H2D/D2H */
int o3d_host_alloc(void** ptr, unsigned long long bytes);
int o3d_host_free(void* ptr);
int o3d_host_register(void* ptr, unsigned long long bytes);
int o3d_host_unregister(void* ptr);

/* =====================================================================================
 * A. Module procedures (HOST pointers, stateless).
 * ===================================================================================== */

/* Copy pipelining of o3d_predict_velocity / o3d_correct_velocity (process-wide; no counterpart in
 * the reference, whose arrays never leave the host).  chunks >= 2: the host arrays are cut into
 * that many z chunks and the upload of chunk j+1, the kernels on chunk j and the download of
 * chunk j-1 overlap, so both PCIe directions are busy at once; results are bitwise those of the
 * unpipelined call.  0 = off.  Default: environment variable O3D_PIPELINE, else 16.  Grids with
 * fewer than 8 planes per chunk use fewer chunks (or the plain path).  The overlap needs
 * page-locked host arrays (o3d_host_alloc / o3d_host_register); with pageable arrays the calls
 * are correct but the copies serialise. */
int o3d_set_pipeline(int chunks);
int o3d_get_pipeline(void);
/* Host-side history shift of the pipelined o3d_predict_velocity (process-wide).  The reference
 * ends predict_velocity with whole-array copies between the time levels of fux / fuy / fuz
 * (src/integration.f90:176-188: level 3 = old level 2, level 2 = level 1 = new f).  on = 1: only
 * level 1 is copied back from the device; levels 2 and 3 -- and nu_t = 0.d0 of a DNS call,
 * src/integration.f90:112 -- are produced in the host arrays by memcpy / memset on worker
 * threads (O3D_HOSTSHIFT_THREADS, default min(8, cores - 2)), chunk by chunk behind the
 * transfers: 6 (DNS: 7) of the 13 arrays of the call no longer cross PCIe.  The arrays are bit
 * for bit those of on = 0.  Default: environment variable O3D_HOSTSHIFT, else on when the host
 * has at least 10 hardware threads (with fewer workers than 8 the memcpy trails the transfers it
 * replaces: measured slower).  No effect on unpipelined calls. */
int o3d_set_hostshift(int on);
int o3d_get_hostshift(void);
/* The schedule a pipelined call would use for nz planes under the current setting (host logic
 * only, no device needed): returns the number of chunks C (0 = plain path); z_bounds[C+1] = plane
 * ranges [z[c], z[c+1]); issue_after[C] = the upload chunk after which the kernels on chunk c are
 * queued (chunks released by the same upload run in ascending order); zfill[C] = z ghost sides
 * filled right before chunk c (1 low, 2 high, 3 both).  Arrays may be NULL; size them for 64
 * chunks (65 bounds). */
int o3d_pipeline_plan(int nz, int periodic_z, int* z_bounds, int* issue_after, int* zfill);

/* schemes(), src/initialization.f90:226-304: binds the 12 derivative closures and the
 * Poisson solver variant from the boundary flags.  Returns O3D_ERR_BC for the combinations
 * the reference stops on. */
int o3d_schemes(int nbcx1, int nbcxn, int nbcy1, int nbcyn, int nbcz1, int nbczn, int sim2d);

/* der_type: subroutine(df, f, d), src/initialization.f90:86-91.  Fortran gets nx,ny,nz from
 * size(f,.); the C ABI passes them.  Generic form + the 20 concrete routines of
 * src/derivation.f90 + the 12 pointer names of src/initialization.f90:104-109. */
int o3d_der(int axis, int order, int closure, double* df, const double* f, double d, int nx,
            int ny, int nz);

#define O3D_DECL_DER(name) \
    int o3d_##name(double* df, const double* f, double d, int nx, int ny, int nz);
/* src/derivation.f90:6,62,111 | :165,219,269 | :323,377,427 */
O3D_DECL_DER(derx_00) O3D_DECL_DER(derxp_11) O3D_DECL_DER(derxi_11)
O3D_DECL_DER(dery_00) O3D_DECL_DER(deryp_11) O3D_DECL_DER(deryi_11)
O3D_DECL_DER(derz_00) O3D_DECL_DER(derzp_11) O3D_DECL_DER(derzi_11)
/* src/derivation.f90:497,593,544 | :642,691,740 | :789,885,836 */
O3D_DECL_DER(derxx_00) O3D_DECL_DER(derxxp_11) O3D_DECL_DER(derxxi_11)
O3D_DECL_DER(deryy_00) O3D_DECL_DER(deryyp_11) O3D_DECL_DER(deryyi_11)
O3D_DECL_DER(derzz_00) O3D_DECL_DER(derzzp_11) O3D_DECL_DER(derzzi_11)
/* src/derivation.f90:481,934 */
O3D_DECL_DER(derz_2dsim) O3D_DECL_DER(derzz_2dsim)
/* procedure pointers bound by o3d_schemes(), src/initialization.f90:104-109 */
O3D_DECL_DER(derxp) O3D_DECL_DER(derxxp) O3D_DECL_DER(derxi) O3D_DECL_DER(derxxi)
O3D_DECL_DER(deryp) O3D_DECL_DER(deryyp) O3D_DECL_DER(deryi) O3D_DECL_DER(deryyi)
O3D_DECL_DER(derzp) O3D_DECL_DER(derzzp) O3D_DECL_DER(derzi) O3D_DECL_DER(derzzi)
#undef O3D_DECL_DER

/* divergence(divf, fx,fy,fz, dx,dy,dz, nx,ny,nz, odd), src/differential_operators.f90:7 */
int o3d_divergence(double* divf, const double* fx, const double* fy, const double* fz,
                   double dx, double dy, double dz, int nx, int ny, int nz, int odd);
/* rotational(rotx,roty,rotz, ux,uy,uz, dx,dy,dz, nx,ny,nz), src/differential_operators.f90:40 */
int o3d_rotational(double* rotx, double* roty, double* rotz, const double* ux, const double* uy,
                   const double* uz, double dx, double dy, double dz, int nx, int ny, int nz);
/* calculate_Q_criterion(Q, ux,uy,uz, dx,dy,dz, nx,ny,nz), src/differential_operators.f90:79 */
int o3d_calculate_q_criterion(double* q, const double* ux, const double* uy, const double* uz,
                              double dx, double dy, double dz, int nx, int ny, int nz);

/* calculate_nu_t(nu_t, ux,uy,uz, dx,dy,dz, cs, delta), src/les_turbulence.f90:10.
 * stats6 (may be NULL) receives function_stats(nu_t) = {min,max,mean,imax,jmax,kmax} that the
 * reference prints at :89-90. */
int o3d_calculate_nu_t(double* nu_t, const double* ux, const double* uy, const double* uz,
                       double dx, double dy, double dz, double cs, double delta, int nx, int ny,
                       int nz, double* stats6);

/* predict_velocity(...), src/integration.f90:14-16.  fux/fuy/fuz are (nx,ny,nz,3) inout,
 * adt/bdt/cdt are the 3-vectors of src/initialization.f90:194-202. */
int o3d_predict_velocity(double* ux_pred, double* uy_pred, double* uz_pred, const double* ux,
                         const double* uy, const double* uz, double* fux, double* fuy,
                         double* fuz, double re, const double* adt, const double* bdt,
                         const double* cdt, int itime, int itscheme, double dx, double dy,
                         double dz, int nx, int ny, int nz, int iles, double cs, double delta,
                         double* nu_t);

/* poi_type, src/initialization.f90:93-102; the three SOR variants src/poisson.f90:6,132,257
 * and the `poisson_solver` pointer.  pp, omega are inout.  Extra outputs (may be NULL):
 * iters = value of the Fortran loop variable at exit, dmax = last max|p_new - p|.
 * The sweep ordering is the process-wide setting o3d_set_sor_order() (default RED_BLACK). */
int o3d_set_sor_order(int order);
int o3d_poisson_solver_0000(double* pp, const double* rhs, double dx, double dy, double dz,
                            int nx, int ny, int nz, double* omega, double eps, int kmax,
                            int idyn, int* iters, double* dmax);
int o3d_poisson_solver_0011(double* pp, const double* rhs, double dx, double dy, double dz,
                            int nx, int ny, int nz, double* omega, double eps, int kmax,
                            int idyn, int* iters, double* dmax);
int o3d_poisson_solver_111111(double* pp, const double* rhs, double dx, double dy, double dz,
                              int nx, int ny, int nz, double* omega, double eps, int kmax,
                              int idyn, int* iters, double* dmax);
int o3d_poisson_solver(double* pp, const double* rhs, double dx, double dy, double dz, int nx,
                       int ny, int nz, double* omega, double eps, int kmax, int idyn, int* iters,
                       double* dmax);

/* solve_poisson_multigrid(phi, rhs, dx,dy,dz, nx,ny,nz, nlevels,npre,npost, tol),
 * src/poisson_multigrid.f90:10.  The reference routine is undefined behaviour as called
 * (array-shape overrun); this entry point solves the SAME operator and
 * boundary rule as poisson_solver with geometric V-cycles until max|r|/|A| < tol. */
int o3d_solve_poisson_multigrid(double* phi, const double* rhs, double dx, double dy, double dz,
                                int nx, int ny, int nz, int nlevels, int npre, int npost,
                                double tol, int* cycles, double* dmax);

/* correct_pression(...), src/integration.f90:199-200 */
int o3d_correct_pression(double* pp, const double* ux_pred, const double* uy_pred,
                         const double* uz_pred, double dx, double dy, double dz, int nx, int ny,
                         int nz, double dt, double* omega, double eps, int kmax, int idyn,
                         int multigrid, int* iters, double* dmax);

/* correct_velocity(...), src/integration.f90:257-258.  Returns O3D_ERR_DIVERGED where the
 * reference would write_velocity_diverged() + stop (outputs are still written). */
int o3d_correct_velocity(double* ux, double* uy, double* uz, const double* ux_pred,
                         const double* uy_pred, const double* uz_pred, const double* pp,
                         double dt, double dx, double dy, double dz, int nx, int ny, int nz);

/* transeq(...), src/integration.f90:332-333.  src may be NULL (the reference never assigns
 * it, src/initialization.f90:159). fphi is (nx,ny,nz,3) inout. */
int o3d_transeq(double* phi, const double* ux, const double* uy, const double* uz,
                const double* src, double* fphi, double re, double sc, const double* adt,
                const double* bdt, const double* cdt, int itime, int itscheme, double dx,
                double dy, double dz, int nx, int ny, int nz, int iles, const double* nu_t);

/* statistics_calc(ux,uy,uz, nx,ny,nz, dx,dy,dz, re, t), src/utils.f90:243: out17 = the 17
 * columns the reference writes to stats.dat (src/IOfunctions.f90:504-552). */
int o3d_statistics_calc(const double* ux, const double* uy, const double* uz, int nx, int ny,
                        int nz, double dx, double dy, double dz, double re, double t,
                        double* out17);
/* function_stats(f,nx,ny,nz), src/functions.f90:27: {min,max,mean,imax,jmax,kmax} */
int o3d_function_stats(const double* f, int nx, int ny, int nz, double* stats6);

/* =====================================================================================
 * B. Session API: device-resident state for the time loop of src/osinco3d_main.f90:97-128.
 * ===================================================================================== */

typedef struct o3d_config {
    int nx, ny, nz;           /* GLOBAL grid (Domain namelist)                            */
    double dx, dy, dz;        /* xlx/(nx-1) etc., src/initialization.f90:182-184          */
    int nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d; /* BoundaryConditions namelist  */
    double re, sc, cs, delta; /* FlowParam / Scalar / LES; delta: initialization.f90:193  */
    double dt;
    double adt[3], bdt[3], cdt[3]; /* src/initialization.f90:194-202                      */
    int itscheme, iles, nscr;
    double omega, eps;        /* PoissonEq namelist; omega persists across steps          */
    int kmax, idyn, multigrid;
    int sor_order;            /* O3D_SOR_RED_BLACK | O3D_SOR_LEXI_WAVEFRONT               */
    int sor_check_every;      /* RED_BLACK only: host convergence poll interval (0 = auto)*/
    /* z-slab decomposition over the GPUs of one node (0/1 ranks = single GPU) */
    int rank, nranks;
    unsigned char nccl_id[128]; /* ncclUniqueId bytes, identical on all ranks, if nranks>1 */
    int reserved[8];
} o3d_config;

typedef struct o3d_session o3d_session;

/* multi-GPU bootstrap: rank 0 obtains 128 ncclUniqueId bytes, the launcher broadcasts them
 * (torch.distributed / MPI / a file) and every rank puts them in o3d_config.nccl_id */
int o3d_nccl_unique_id(unsigned char* out128);
/* the z-slab partition o3d_session_create applies: rank owns global planes [z0, z0+nz_local);
 * contiguous slabs, remainder planes to the low ranks.  Pure host arithmetic (no device). */
int o3d_slab_partition(int nz, int nranks, int rank, int* z0, int* nz_local);

/* fields a session owns (module `initialization` globals) */
enum {
    O3D_F_UX = 0, O3D_F_UY, O3D_F_UZ, O3D_F_PP, O3D_F_PHI,
    O3D_F_UX_PRED, O3D_F_UY_PRED, O3D_F_UZ_PRED, O3D_F_NU_T, O3D_F_RHS,
    O3D_F_FUX1, O3D_F_FUX2, O3D_F_FUX3, O3D_F_FUY1, O3D_F_FUY2, O3D_F_FUY3,
    O3D_F_FUZ1, O3D_F_FUZ2, O3D_F_FUZ3, O3D_F_FPHI1, O3D_F_FPHI2, O3D_F_FPHI3,
    O3D_F_DIVU, O3D_F_SCRATCH0, O3D_F_SCRATCH1, O3D_F_SCRATCH2,
    O3D_F_PP2, /* ping-pong partner of pp inside the fused red-black SOR (internal) */
    O3D_F_OLD_UX, O3D_F_OLD_UY, O3D_F_OLD_UZ, /* old_ux, old_uy, old_uz (src/utils.f90:165) */
    O3D_F_COUNT
};

/* reductions, src/functions.f90 / src/utils.f90:178 */
enum { O3D_RED_MIN = 0, O3D_RED_MAX = 1, O3D_RED_SUM = 2, O3D_RED_ABSMAX = 3 };

int o3d_session_create(const o3d_config* cfg, o3d_session** out);
int o3d_session_destroy(o3d_session* s);
/* local slab extent of this rank: planes [z0, z0+nz_local) of the global grid */
int o3d_session_slab(const o3d_session* s, int* z0, int* nz_local);
/* host <-> device copies of this rank's slab of one field (nx*ny*nz_local doubles).
 * History ids (O3D_F_FUX1..3 etc.) are LOGICAL levels 1..3 of fu?(:,:,:,1:3): the reference's
 * shifts fu(:,:,:,2) = fu(:,:,:,1), fu(:,:,:,3) = fu(:,:,:,2) (src/integration.f90:176-188) are
 * pointer rotations here, so after a step levels 1 and 2 hold the same data in ONE buffer.  An
 * upload to level 1 is redirected to the free buffer first (it never overwrites level 2 or 3);
 * a restart may therefore upload the levels in any order.
 * o3d_download also reports a pending NaN / >1000 guard (O3D_ERR_DIVERGED, data still
 * delivered), like o3d_sync. */
int o3d_upload(o3d_session* s, int field, const double* host);
int o3d_download(o3d_session* s, int field, double* host);
/* the same for local planes [k0, k0 + nk) only: host = a contiguous (nx, ny, nk) block.  Lets a
 * driver fill or read a slab in pieces without holding a whole-field host array (1024^3: 8.6 GB) */
int o3d_upload_planes(o3d_session* s, int field, const double* host, int k0, int nk);
int o3d_download_planes(o3d_session* s, int field, double* host, int k0, int nk);
/* Device fields are PADDED (ghost cells hold the boundary closure, DESIGN.md "Data layout"):
 * element (i,j,k) of a field lives at dptr[i + stride_j*j + stride_k*k].  o3d_device_ptr returns
 * the interior origin (0,0,0) for zero-copy callers; after writing through it call
 * o3d_mark_modified so that the ghost cells are refreshed before the next stencil. */
int o3d_device_ptr(o3d_session* s, int field, double** dptr);
int o3d_session_layout(const o3d_session* s, long long* stride_j, long long* stride_k);
int o3d_mark_modified(o3d_session* s, int field);

/* hot-path stages; same order and meaning as src/osinco3d_main.f90:105-115 */
int o3d_s_predict_velocity(o3d_session* s, int itime);
int o3d_s_correct_pression(o3d_session* s, int* iters, double* dmax);
int o3d_s_correct_velocity(o3d_session* s);
int o3d_s_transeq(o3d_session* s, int itime);
/* predict + pression + velocity [+ transeq].  The NaN / >1000 guard of correct_velocity
 * (src/integration.f90:309-325) does not stall the pipeline here: its flag is examined at the
 * next host synchronisation, so O3D_ERR_DIVERGED is returned by the FOLLOWING o3d_step (after
 * its Poisson solve) or by o3d_sync(), whichever comes first.  o3d_s_correct_velocity and the
 * stateless o3d_correct_velocity report it immediately. */
int o3d_step(o3d_session* s, int itime, int* iters, double* dmax);
/* asynchronous variant used by bench.py: enqueue only, no host synchronisation except the
 * SOR convergence polls; call o3d_sync() before reading results */
int o3d_sync(o3d_session* s);

/* diagnostics the reference driver runs every step (src/osinco3d_main.f90:116-128) */
int o3d_s_divergence(o3d_session* s, int fx, int fy, int fz, int dst, int odd);
int o3d_s_reduce(o3d_session* s, int field, int op, double* out);
int o3d_s_function_stats(o3d_session* s, int field, double* stats6);
int o3d_s_statistics(o3d_session* s, double t, double* out17);
/* Everything src/osinco3d_main.f90:116-128 prints per step, in two fused passes (one over u*,
 * one over u; 24 B/pt each) instead of 2 divergences + 2 function_stats + 9 reductions:
 *   out23[0..5]   function_stats(divergence(u*, odd=1)): min, max, mean, i, j, k of the max
 *   out23[6..11]  function_stats(divergence(u,  odd=1))          (src/functions.f90:27-63)
 *   out23[12..14] minval(ux, uy, uz)  out23[15..17] maxval   (print_velocity_values,
 *                                                             src/IOfunctions.f90:322)
 *   out23[18..20] cflx, cfly, cflz                           (compute_cfl, src/utils.f90:178)
 *   out23[21..22] minval(phi), maxval(phi) if nscr == 1      (print_scalar_values, :332) */
int o3d_s_step_diagnostics(o3d_session* s, double* out23);
/* old_values, src/utils.f90:165-176: old_u? = u? (device copies; only needed on the steps whose
 * residual is evaluated, src/osinco3d_main.f90:104,167) */
int o3d_s_old_values(o3d_session* s);
/* calculate_residuals, src/utils.f90:93-160, over old_u? and u?: out15 = res_u res_v res_w |
 * aa bb cc (= t_ref/u_ref * Linf) | (ia ja ka) (ib jb kb) (ic jc kc), 1-based GLOBAL indices of
 * the LAST interior point attaining each maximum, as the reference's loop leaves them.  The
 * "residue too high" stop (:155-158) is the caller's: res_u > 1e6. */
int o3d_s_calculate_residuals(o3d_session* s, double dt, double t_ref, double u_ref,
                              double* out15);
int o3d_s_rotational(o3d_session* s, int rotx, int roty, int rotz);
int o3d_s_q_criterion(o3d_session* s, int dst);
/* sqrt(rotx**2 + roty**2 + rotz**2), the "vort" array of write_all_data
 * (src/visualization.f90:258-259), in one fused kernel */
int o3d_s_vorticity_magnitude(o3d_session* s, int dst);

/* ---- field output in the reference's binary formats, straight from device-resident state.
 * Writes are ASYNCHRONOUS: the call snapshots the fields on the device (the time loop may go
 * on at once), the D2H copies run on a separate stream and a writer thread does the file I/O.
 * With z slabs every rank writes its own contiguous byte range of the shared file (call on all
 * ranks, same path on a shared file system).  o3d_s_io_wait blocks until everything queued has
 * reached the file and reports the first write error; o3d_session_destroy drains the queue. */
/* save_fields, src/IOfunctions.f90:360-402: time | nx ny nz | x y z | ux uy uz pp phi */
int o3d_s_save_fields(o3d_session* s, const char* filename, double time, const double* x,
                      const double* y, const double* z);
/* read_fields, src/IOfunctions.f90:404-470; a grid-size mismatch is O3D_ERR_INVALID and nothing
 * is read (:452-459).  x, y, z: GLOBAL coordinate arrays (nx, ny, nz) */
int o3d_s_read_fields(o3d_session* s, const char* filename, double* time, double* x, double* y,
                      double* z);
/* write_binary, src/visualization.f90:224-241: one raw f64 array */
int o3d_s_write_binary(o3d_session* s, const char* filename, int field);
/* write_all_data, src/visualization.f90:243-276: <dir>/ux_<num>.bin uy uz pp vort qcrit
 * [phi if nscr == 1] [nu_t if iles == 1]; vort and qcrit are computed here (the driver calls
 * rotational / calculate_Q_criterion just before, src/osinco3d_main.f90:130-133) */
int o3d_s_write_all_data(o3d_session* s, const char* dir, int num);
int o3d_s_io_wait(o3d_session* s);

/* how the LAST red-black solve of this session ran (bench.py / tests report it; no counterpart in
 * the reference): *persistent = 1 when all its iterations ran inside one cooperative launch
 * (exit tests and dynamic omega on the device, no host poll), *peer = 1 when, on z slabs, its ghost
 * planes and residual maxima travelled through peer-mapped memory instead of NCCL calls */
int o3d_s_sor_path(const o3d_session* s, int* persistent, int* peer);
/* current SOR relaxation factor (inout across steps, src/integration.f90:222,247) */
int o3d_get_omega(const o3d_session* s, double* omega);
int o3d_set_omega(o3d_session* s, double omega);
/* the Poisson controls are per-call arguments of correct_pression in the reference
 * (src/integration.f90:199-200); a resident driver pushes them into the session with this */
int o3d_session_set_poisson(o3d_session* s, double eps, int kmax, int idyn, int multigrid);
/* accumulated per-stage device time in ms since the last reset (CUDA events on the
 * session stream): out[0]=rhs/predictor out[1]=divergence out[2]=sor out[3]=correction
 * out[4]=transeq out[5]=halo exchange; counts[] = launches per stage */
int o3d_s_timers(o3d_session* s, double* ms6, long long* counts6, int reset);
int o3d_s_enable_timers(o3d_session* s, int on);
/* stopwatch on the session stream (CUDA events): bench.py brackets its timed region with these
 * because torch.cuda.Event only sees torch's current stream.  stop synchronises the stream. */
int o3d_s_stopwatch_start(o3d_session* s);
int o3d_s_stopwatch_stop(o3d_session* s, double* ms);

#ifdef __cplusplus
}
#endif
#endif
