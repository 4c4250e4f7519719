/*
 * o3d_oracle.h -- CPU ORACLE for the osinco3d Chorin-projection time step.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (osinco3d_b200/ + libo3d_b200.so) never links, imports or calls anything in oracle/.
 *
 * It is a plain-C restatement (no FMA contraction, same expression order) of the
 * reference's Fortran hot path; every function cites the reference file:line it follows
 * (paths relative to the reference checkout, src/...).  The reference itself cannot be
 * compiled in this image (no Fortran compiler, no FFTW), so there is no oracle/_ref.
 *
 * Pinning: tests/test_oracle_golden.py checks this oracle against the reference's shipped
 * statistics histories (examples/tgv_re1600_dns/tgv_stats_re1600_dns.dat rows 1-2,
 * examples/tgv_re2500_les/tgv_stats_re2500_les.dat row 1), committed as tests/golden/.
 *
 * Arrays are Fortran-ordered (nx,ny,nz), i fastest, 0-based here: idx = i + nx*(j + ny*k).
 */
#ifndef O3D_ORACLE_H
#define O3D_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* closure codes of one derivative routine (src/derivation.f90) */
enum { ORC_PERIODIC = 0, /* *_00   */
       ORC_EVEN = 1,     /* *p_11  */
       ORC_ODD = 2,      /* *i_11  */
       ORC_ZERO = 3      /* der*_2dsim */ };

typedef struct orc_grid {
    int nx, ny, nz;
    double dx, dy, dz;
    int bc[3];  /* per axis: 0 = PERIODIC, 1 = FREE_SLIP (src/initialization.f90:228-281) */
    int sim2d;
} orc_grid;

/* one derivative routine of src/derivation.f90: axis 0/1/2, order 1/2, closure code */
void orc_der(int axis, int order, int closure, double* df, const double* f, double d,
             int nx, int ny, int nz);

/* pointer binding of schemes() (src/initialization.f90:226-281): parity 0 = 'p', 1 = 'i' */
int orc_closure(const orc_grid* g, int axis, int parity);
void orc_derp(const orc_grid* g, int axis, int order, int parity, double* df, const double* f);

/* src/differential_operators.f90:7-38 */
void orc_divergence(const orc_grid* g, double* divf, const double* fx, const double* fy,
                    const double* fz, int odd);
/* src/differential_operators.f90:40-77 and :79-108 */
void orc_rotational(const orc_grid* g, double* rotx, double* roty, double* rotz,
                    const double* ux, const double* uy, const double* uz);
void orc_q_criterion(const orc_grid* g, double* q, const double* ux, const double* uy,
                     const double* uz);

/* src/les_turbulence.f90:10-97 */
void orc_calculate_nu_t(const orc_grid* g, double* nu_t, const double* ux, const double* uy,
                        const double* uz, double cs, double delta);

/* src/integration.f90:14-197. fux/fuy/fuz are (nx,ny,nz,3). returns 0, or 1 for the
 * "itscheme unrecognized" stop. */
int orc_predict_velocity(const orc_grid* g, double* ux_pred, double* uy_pred, double* uz_pred,
                         const double* ux, const double* uy, const double* uz, double* fux,
                         double* fuy, double* fuz, double re, const double* adt,
                         const double* bdt, const double* cdt, int itime, int itscheme, int iles,
                         double cs, double delta, double* nu_t);

/* src/poisson.f90:6-130 (_0000), :132-255 (_0011), :257-381 (_111111).
 * mirror[a] = 1 selects the mirrored neighbour rule on axis a; factor = 1.05 or 1.01.
 * Returns the iteration count at exit (Fortran `iter` after the loop, i.e. kmax+1 if the
 * loop ran out); *dmax_out = last dmax. */
int orc_poisson_sor(double* pp, const double* rhs, double dx, double dy, double dz, int nx,
                    int ny, int nz, const int* mirror, double factor, double* omega,
                    double eps, int kmax, int idyn, double* dmax_out);
/* poisson_solver pointer binding, src/initialization.f90:283-301. returns -1 if the
 * pointer would be null. */
int orc_poisson_solver(const orc_grid* g, double* pp, const double* rhs, double* omega,
                       double eps, int kmax, int idyn, double* dmax_out);

/* src/integration.f90:199-255 (multigrid == 0 branch only; see DESIGN.md for multigrid) */
int orc_correct_pression(const orc_grid* g, double* pp, const double* ux_pred,
                         const double* uy_pred, const double* uz_pred, double dt, double* omega,
                         double eps, int kmax, int idyn, double* dmax_out, double* rhs_out);

/* src/integration.f90:257-330. returns 1 if the NaN / >1000 abort would fire. */
int orc_correct_velocity(const orc_grid* g, double* ux, double* uy, double* uz,
                         const double* ux_pred, const double* uy_pred, const double* uz_pred,
                         const double* pp, double dt);

/* src/integration.f90:332-468. fphi is (nx,ny,nz,3); src may be NULL (== 0). */
int orc_transeq(const orc_grid* g, double* phi, const double* ux, const double* uy,
                const double* uz, const double* src, double* fphi, double re, double sc,
                const double* adt, const double* bdt, const double* cdt, int itime, int itscheme,
                int iles, const double* nu_t);

/* src/utils.f90:243-375; out[0..16] = the 17 columns of stats.dat (out[0] = t) */
void orc_statistics_calc(const orc_grid* g, const double* ux, const double* uy, const double* uz,
                         double re, double t, double* out);

/* src/functions.f90:27-65: out = {min, max, mean, imax, jmax, kmax} (1-based indices) */
void orc_function_stats(const double* f, int nx, int ny, int nz, double* out);

/* src/initialization.f90:194-202 */
/* utils.f90:93-160: res_u res_v res_w | aa bb cc | (ia ja ka) (ib jb kb) (ic jc kc) */
void orc_calculate_residuals(const double* u, const double* v, const double* w,
                             const double* old_u, const double* old_v, const double* old_w,
                             double dt, double t_ref, double u_ref, int nx, int ny, int nz,
                             double* out15);
void orc_ab_coefficients(double dt, double* adt, double* bdt, double* cdt);

/* initial conditions, src/initial_conditions.f90:103-175 (TGV). x0,y0,z0 origin. */
void orc_init_tgv(const orc_grid* g, double x0, double y0, double z0, double u0, double l0,
                  double ratio, int nscr, double delta, double* ux, double* uy, double* uz,
                  double* pp, double* phi);
/* src/initial_conditions.f90:329-395 (mixing layer) and :244-327 (coplanar jet) */
void orc_init_mixing_layer(const orc_grid* g, double x0, double y0, double z0, double u0,
                           double l0, double ratio, int nscr, double* ux, double* uy,
                           double* uz, double* pp, double* phi);
void orc_init_coplanar_jet(const orc_grid* g, double x0, double y0, double z0, double u0,
                           double l0, double ratio, int nscr, double* ux, double* uy,
                           double* uz, double* pp, double* phi);

/* ---- a whole run: the time loop of src/osinco3d_main.f90:97-128 (hot path only) ---- */
typedef struct orc_sim {
    orc_grid g;
    double re, sc, cs, delta, dt;
    double adt[3], bdt[3], cdt[3];
    int itscheme, iles, nscr;
    double omega, eps;
    int kmax, idyn;
    /* fields (owned) */
    double *ux, *uy, *uz, *pp, *phi, *ux_pred, *uy_pred, *uz_pred, *nu_t;
    double *fux, *fuy, *fuz, *fphi;
    /* per-step reports */
    int last_iters;
    double last_dmax;
    long total_iters;
} orc_sim;

orc_sim* orc_sim_create(const orc_grid* g, double re, double sc, double cs, double delta,
                        double dt, int itscheme, int iles, int nscr, double omega, double eps,
                        int kmax, int idyn);
void orc_sim_destroy(orc_sim* s);
/* one iteration of the main loop body, itime 1-based; returns non-zero on abort */
int orc_sim_step(orc_sim* s, int itime);
double* orc_sim_field(orc_sim* s, const char* name);

#ifdef __cplusplus
}
#endif
#endif
