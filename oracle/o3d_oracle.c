/*
 * o3d_oracle.c -- CPU ORACLE (test infrastructure; see o3d_oracle.h for the rules).
 *
 * Plain C99 restatement of the reference's Fortran hot path.  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math   (oracle/Makefile)
 * so that no FMA contraction or re-association happens: gfortran -O3 on baseline x86-64
 * (src/Makefile:15) emits neither, and expressions below keep the Fortran evaluation order
 * (left-to-right for equal precedence, parentheses honoured).
 *
 * Parity status: PINNED, two ways -- not against a compiled reference (none can be built here):
 *  (1) the reference's golden statistics files (tests/test_oracle_golden.py);
 *  (2) the reference SOURCE: every hot-path routine is translated statement by statement from
 *      /root/reference/src into NumPy and executed (tests/golden/f90np.py,
 *      make_hotpath_golden.py -> tests/golden/hotpath.npz); this file reproduces those vectors
 *      bit for bit -- stencils, operators, predictor, the three SOR solvers with their iteration
 *      counts and dynamic omega, correction, transeq (tests/test_oracle_reference_source.py).
 */
#include "o3d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define IDX(i, j, k) ((size_t)(i) + (size_t)nx * ((size_t)(j) + (size_t)ny * (size_t)(k)))

/* ------------------------------------------------------------------------------------
 * Stencils: src/derivation.f90.
 *
 * All 18 routines are "interior formula applied to a ghost-extended line":
 *   periodic  *_00  : f(p) = f(p +- n)                       (:34-56, :523-539)
 *   even      *p_11 : f(1-k) = +f(1+k), f(n+k) = +f(n-k)     (:87-105, :620-636)
 *   odd       *i_11 : f(1-k) = -f(1+k), f(n+k) = -f(n-k)     (:137-159, :571-587)
 * The source spells the boundary planes out as explicit sums/differences, e.g.
 * a*(f(5)+f(3)) for a*(f(5)-(-f(3))) at derivation.f90:140; x-(-y) == x+y and x+(-y) == x-y
 * bitwise in IEEE-754, so applying the interior expression to sign-carrying ghost values
 * reproduces them exactly.  The even first derivative assigns literal 0.d0 to planes 1 and
 * n (:87,:105).  x/y/z variants are textually identical up to the axis (checked
 * mechanically when this file was written).
 * ---------------------------------------------------------------------------------- */

/* value of line element q (0-based, may be out of [0,n)) under a closure */
static inline double ghost(const double* f, size_t base, size_t s, int q, int n, int closure) {
    if (q >= 0 && q < n) return f[base + (size_t)q * s];
    if (closure == ORC_PERIODIC) {
        q = (q < 0) ? q + n : q - n;
        return f[base + (size_t)q * s];
    }
    q = (q < 0) ? -q : 2 * (n - 1) - q; /* mirror about plane 1 / plane n */
    double v = f[base + (size_t)q * s];
    return (closure == ORC_ODD) ? -v : v;
}

void orc_der(int axis, int order, int closure, double* df, const double* f, double d, int nx,
             int ny, int nz) {
    const size_t N = (size_t)nx * ny * nz;
    if (closure == ORC_ZERO) { /* derz_2dsim / derzz_2dsim, derivation.f90:481-495, :934-948 */
        for (size_t m = 0; m < N; ++m) df[m] = 0.0;
        return;
    }
    const int n = (axis == 0) ? nx : (axis == 1) ? ny : nz;
    const size_t s = (axis == 0) ? 1 : (axis == 1) ? (size_t)nx : (size_t)nx * ny;
    double a, b, c;
    if (order == 1) { /* derivation.f90:26-30 */
        const double sixtyd = 60.0 * d;
        a = 1.0 / sixtyd;
        b = 9.0 / sixtyd;
        c = 45.0 / sixtyd;
    } else { /* derivation.f90:517-521 */
        const double twelvedsq = 12.0 * d * d;
        a = 1.0 / twelvedsq;
        b = 16.0 / twelvedsq;
        c = 30.0 / twelvedsq;
    }
    const int r = (order == 1) ? 3 : 2; /* stencil radius */
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int p = (axis == 0) ? i : (axis == 1) ? j : k;
                const size_t m = IDX(i, j, k);
                if (p >= r && p < n - r) {
                    if (order == 1) /* derivation.f90:43-47 */
                        df[m] = a * (f[m + 3 * s] - f[m - 3 * s]) - b * (f[m + 2 * s] - f[m - 2 * s]) +
                                c * (f[m + s] - f[m - s]);
                    else /* derivation.f90:529-533; "-a*x" is -(a*x) */
                        df[m] = -(a * (f[m - 2 * s] + f[m + 2 * s])) + b * (f[m - s] + f[m + s]) -
                                c * f[m];
                } else {
                    const size_t base = m - (size_t)p * s;
                    if (order == 1) {
                        if (closure == ORC_EVEN && (p == 0 || p == n - 1)) {
                            df[m] = 0.0; /* derivation.f90:87,:105 */
                        } else {
                            const double fp3 = ghost(f, base, s, p + 3, n, closure);
                            const double fm3 = ghost(f, base, s, p - 3, n, closure);
                            const double fp2 = ghost(f, base, s, p + 2, n, closure);
                            const double fm2 = ghost(f, base, s, p - 2, n, closure);
                            const double fp1 = ghost(f, base, s, p + 1, n, closure);
                            const double fm1 = ghost(f, base, s, p - 1, n, closure);
                            df[m] = a * (fp3 - fm3) - b * (fp2 - fm2) + c * (fp1 - fm1);
                        }
                    } else {
                        const double fm2 = ghost(f, base, s, p - 2, n, closure);
                        const double fp2 = ghost(f, base, s, p + 2, n, closure);
                        const double fm1 = ghost(f, base, s, p - 1, n, closure);
                        const double fp1 = ghost(f, base, s, p + 1, n, closure);
                        df[m] = -(a * (fm2 + fp2)) + b * (fm1 + fp1) - c * f[m];
                    }
                }
            }
}

/* schemes(), src/initialization.f90:226-281 */
int orc_closure(const orc_grid* g, int axis, int parity) {
    if (axis == 2 && g->sim2d == 1) return ORC_ZERO;
    if (g->bc[axis] == 0) return ORC_PERIODIC;
    return parity ? ORC_ODD : ORC_EVEN;
}

void orc_derp(const orc_grid* g, int axis, int order, int parity, double* df, const double* f) {
    const double d = (axis == 0) ? g->dx : (axis == 1) ? g->dy : g->dz;
    orc_der(axis, order, orc_closure(g, axis, parity), df, f, d, g->nx, g->ny, g->nz);
}

static double* newfield(size_t N) {
    double* p = (double*)malloc(N * sizeof(double));
    if (!p) abort();
    return p;
}

/* ------------------------------------------------------------------------------------
 * src/differential_operators.f90
 * ---------------------------------------------------------------------------------- */
void orc_divergence(const orc_grid* g, double* divf, const double* fx, const double* fy,
                    const double* fz, int odd) {
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    double *dfx = newfield(N), *dfy = newfield(N), *dfz = newfield(N);
    const int par = (odd == 0) ? 0 : 1; /* :25-33 */
    orc_derp(g, 0, 1, par, dfx, fx);
    orc_derp(g, 1, 1, par, dfy, fy);
    orc_derp(g, 2, 1, par, dfz, fz);
    for (size_t m = 0; m < N; ++m) divf[m] = dfx[m] + dfy[m] + dfz[m]; /* :35 */
    free(dfx);
    free(dfy);
    free(dfz);
}

void orc_rotational(const orc_grid* g, double* rotx, double* roty, double* rotz,
                    const double* ux, const double* uy, const double* uz) {
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    double *t1 = newfield(N), *t2 = newfield(N);
    orc_derp(g, 1, 1, 0, t1, uz); /* :64-66 */
    orc_derp(g, 2, 1, 0, t2, uy);
    for (size_t m = 0; m < N; ++m) rotx[m] = t1[m] - t2[m];
    orc_derp(g, 2, 1, 0, t1, ux); /* :68-70 */
    orc_derp(g, 0, 1, 0, t2, uz);
    for (size_t m = 0; m < N; ++m) roty[m] = t1[m] - t2[m];
    orc_derp(g, 0, 1, 0, t1, uy); /* :72-74 */
    orc_derp(g, 1, 1, 0, t2, ux);
    for (size_t m = 0; m < N; ++m) rotz[m] = t1[m] - t2[m];
    free(t1);
    free(t2);
}

void orc_q_criterion(const orc_grid* g, double* q, const double* ux, const double* uy,
                     const double* uz) {
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    double* d[9];
    for (int t = 0; t < 9; ++t) d[t] = newfield(N);
    /* :90-100 */
    orc_derp(g, 0, 1, 1, d[0], ux); /* duxdx */
    orc_derp(g, 1, 1, 1, d[1], uy); /* duydy */
    orc_derp(g, 2, 1, 1, d[2], uz); /* duzdz */
    orc_derp(g, 1, 1, 0, d[3], ux); /* duxdy */
    orc_derp(g, 2, 1, 0, d[4], ux); /* duxdz */
    orc_derp(g, 0, 1, 0, d[5], uy); /* duydx */
    orc_derp(g, 2, 1, 0, d[6], uy); /* duydz */
    orc_derp(g, 0, 1, 0, d[7], uz); /* duzdx */
    orc_derp(g, 1, 1, 0, d[8], uz); /* duzdy */
    for (size_t m = 0; m < N; ++m) /* :103-104 */
        q[m] = -(0.5 * (d[0][m] * d[0][m] + d[1][m] * d[1][m] + d[2][m] * d[2][m])) -
               d[3][m] * d[5][m] - d[4][m] * d[7][m] - d[6][m] * d[8][m];
    for (int t = 0; t < 9; ++t) free(d[t]);
}

/* ------------------------------------------------------------------------------------
 * src/les_turbulence.f90:10-97
 * ---------------------------------------------------------------------------------- */
void orc_calculate_nu_t(const orc_grid* g, double* nu_t, const double* ux, const double* uy,
                        const double* uz, double cs, double delta) {
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    double* d[9];
    for (int t = 0; t < 9; ++t) d[t] = newfield(N);
    /* :55-67 */
    orc_derp(g, 0, 1, 1, d[0], ux); /* duxdx  derxi */
    orc_derp(g, 1, 1, 0, d[1], ux); /* duxdy  deryp */
    orc_derp(g, 2, 1, 0, d[2], ux); /* duxdz  derzp */
    orc_derp(g, 0, 1, 0, d[3], uy); /* duydx  derxp */
    orc_derp(g, 1, 1, 1, d[4], uy); /* duydy  deryi */
    orc_derp(g, 2, 1, 0, d[5], uy); /* duydz  derzp */
    orc_derp(g, 0, 1, 0, d[6], uz); /* duzdx  derxp */
    orc_derp(g, 1, 1, 0, d[7], uz); /* duzdy  deryp */
    orc_derp(g, 2, 1, 1, d[8], uz); /* duzdz  derzi */
    const double csd = cs * delta;
    const double csd2 = csd * csd; /* (cs*delta)**2 */
    for (size_t m = 0; m < N; ++m) { /* :70-88 */
        const double s11 = d[0][m], s22 = d[4][m], s33 = d[8][m];
        const double s12 = 0.5 * (d[1][m] + d[3][m]);
        const double s13 = 0.5 * (d[2][m] + d[6][m]);
        const double s23 = 0.5 * (d[5][m] + d[7][m]);
        const double smag = sqrt(2.0 * (s11 * s11 + s22 * s22 + s33 * s33 +
                                        2.0 * (s12 * s12 + s13 * s13 + s23 * s23)));
        nu_t[m] = csd2 * smag;
    }
    for (int t = 0; t < 9; ++t) free(d[t]);
}

/* ------------------------------------------------------------------------------------
 * src/integration.f90:14-197
 * ---------------------------------------------------------------------------------- */
static int ab_select(const double* adt, const double* bdt, const double* cdt, int itime,
                     int itscheme, double* adu, double* bdu, double* cdu) {
    if (itscheme == 1 || itime == 1) { /* :84-88 */
        *adu = adt[0], *bdu = bdt[0], *cdu = cdt[0];
    } else if (itscheme == 2 || itime == 2) { /* :89-93 */
        *adu = adt[1], *bdu = bdt[1], *cdu = cdt[1];
    } else if (itscheme == 3) { /* :94-98 */
        *adu = adt[2], *bdu = bdt[2], *cdu = cdt[2];
    } else {
        return 1; /* :99-104 "unrecognized" + stop */
    }
    return 0;
}

/* one velocity component: integration.f90:118-134 (ux), :138-154 (uy), :158-174 (uz).
 * par[a] = parity of this component along axis a. */
static void predict_component(const orc_grid* g, double* u_pred, const double* uc,
                              const double* ux, const double* uy, const double* uz, double* fu,
                              const double* nu_eff, const int* par, double adu, double bdu,
                              double cdu, double** w) {
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    double *d1x = w[0], *d1y = w[1], *d1z = w[2], *d2x = w[3], *d2y = w[4], *d2z = w[5];
    orc_derp(g, 0, 1, par[0], d1x, uc);
    orc_derp(g, 1, 1, par[1], d1y, uc);
    orc_derp(g, 2, 1, par[2], d1z, uc);
    orc_derp(g, 0, 2, par[0], d2x, uc);
    orc_derp(g, 1, 2, par[1], d2y, uc);
    orc_derp(g, 2, 2, par[2], d2z, uc);
    double *f1 = fu, *f2 = fu + N, *f3 = fu + 2 * N;
    for (size_t m = 0; m < N; ++m) {
        f1[m] = nu_eff[m] * (d2x[m] + d2y[m] + d2z[m]) -
                (ux[m] * d1x[m] + uy[m] * d1y[m] + uz[m] * d1z[m]); /* :129-130 */
        u_pred[m] = uc[m] + adu * f1[m] + bdu * f2[m] + cdu * f3[m]; /* :132-134 */
    }
}

int orc_predict_velocity(const orc_grid* g, double* ux_pred, double* uy_pred, double* uz_pred,
                         const double* ux, const double* uy, const double* uz, double* fux,
                         double* fuy, double* fuz, double re, const double* adt,
                         const double* bdt, const double* cdt, int itime, int itscheme, int iles,
                         double cs, double delta, double* nu_t) {
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    double adu, bdu, cdu;
    if (ab_select(adt, bdt, cdt, itime, itscheme, &adu, &bdu, &cdu)) return 1;
    const double onere = 1.0 / re; /* :106 */
    if (iles == 1) /* :108-113 */
        orc_calculate_nu_t(g, nu_t, ux, uy, uz, cs, delta);
    else
        for (size_t m = 0; m < N; ++m) nu_t[m] = 0.0;
    double* nu_eff = newfield(N);
    for (size_t m = 0; m < N; ++m) nu_eff[m] = onere + nu_t[m]; /* :114 */
    double* w[6];
    for (int t = 0; t < 6; ++t) w[t] = newfield(N);
    static const int parx[3] = {1, 0, 0}, pary[3] = {0, 1, 0}, parz[3] = {0, 0, 1};
    predict_component(g, ux_pred, ux, ux, uy, uz, fux, nu_eff, parx, adu, bdu, cdu, w);
    predict_component(g, uy_pred, uy, ux, uy, uz, fuy, nu_eff, pary, adu, bdu, cdu, w);
    predict_component(g, uz_pred, uz, ux, uy, uz, fuz, nu_eff, parz, adu, bdu, cdu, w);
    double* fs[3] = {fux, fuy, fuz};
    if (itscheme == 2) { /* :177-180 */
        for (int c = 0; c < 3; ++c) memcpy(fs[c] + N, fs[c], N * sizeof(double));
    } else if (itscheme == 3) { /* :181-187 */
        for (int c = 0; c < 3; ++c) memcpy(fs[c] + 2 * N, fs[c] + N, N * sizeof(double));
        for (int c = 0; c < 3; ++c) memcpy(fs[c] + N, fs[c], N * sizeof(double));
    }
    for (int t = 0; t < 6; ++t) free(w[t]);
    free(nu_eff);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * src/poisson.f90: the three SOR copies differ only in the neighbour rule and `factor`.
 * ---------------------------------------------------------------------------------- */
static inline void nbr(int p, int n, int mirror, int* m1, int* p1) {
    if (p == 0) { /* poisson.f90:57-59 periodic / :197-199 mirrored */
        *m1 = mirror ? 1 : n - 1;
        *p1 = 1;
    } else if (p == n - 1) {
        *m1 = n - 2;
        *p1 = mirror ? n - 2 : 0;
    } else {
        *m1 = p - 1;
        *p1 = p + 1;
    }
}

int orc_poisson_sor(double* pp, const double* rhs, double dx, double dy, double dz, int nx,
                    int ny, int nz, const int* mirror, double factor, double* omega,
                    double eps, int kmax, int idyn, double* dmax_out) {
    /* poisson.f90:41-52 */
    const double dx2 = dx * dx, dy2 = dy * dy, dz2 = dz * dz;
    const double oneondx2 = 1.0 / dx2, oneondy2 = 1.0 / dy2, oneondz2 = 1.0 / dz2;
    const double twoondx2 = 2.0 * oneondx2, twoondy2 = 2.0 * oneondy2, twoondz2 = 2.0 * oneondz2;
    const double A = -(twoondx2 + twoondy2 + twoondz2);
    double dmax_old = 1609.0, dmax = 0.0;
    int iter;
    for (iter = 1; iter <= kmax; ++iter) { /* :53 */
        dmax = 0.0;
        for (int k = 0; k < nz; ++k) {
            int km1, kp1;
            nbr(k, nz, mirror[2], &km1, &kp1);
            for (int j = 0; j < ny; ++j) {
                int jm1, jp1;
                nbr(j, ny, mirror[1], &jm1, &jp1);
                for (int i = 0; i < nx; ++i) {
                    int im1, ip1;
                    nbr(i, nx, mirror[0], &im1, &ip1);
                    const size_t m = IDX(i, j, k);
                    /* :95-98 */
                    const double p_new = (-(oneondx2 * (pp[IDX(im1, j, k)] + pp[IDX(ip1, j, k)])) -
                                          oneondy2 * (pp[IDX(i, jm1, k)] + pp[IDX(i, jp1, k)]) -
                                          oneondz2 * (pp[IDX(i, j, km1)] + pp[IDX(i, j, kp1)]) +
                                          rhs[m]) /
                                         A;
                    const double d = fabs(p_new - pp[m]);      /* :100 */
                    dmax = (d > dmax) ? d : dmax;              /* :101 */
                    pp[m] = (1.0 - *omega) * pp[m] + *omega * p_new; /* :102 */
                }
            }
        }
        if (dmax < eps) break;                              /* :110 */
        if (fabs(dmax_old - dmax) < eps / 1000.0) break;    /* :111-114 */
        if (iter > 1 && idyn == 1) {                        /* :115-121 */
            if (dmax > dmax_old)
                *omega = *omega * (2.0 - factor);
            else if (dmax < 0.1 * dmax_old)
                *omega = fmin(*omega * factor, 2.0);
        }
        dmax_old = dmax;
    }
    if (dmax_out) *dmax_out = dmax;
    return iter;
}

int orc_poisson_solver(const orc_grid* g, double* pp, const double* rhs, double* omega,
                       double eps, int kmax, int idyn, double* dmax_out) {
    /* initialization.f90:283-301: only the x and y flags are inspected */
    int mirror[3];
    double factor;
    if (g->bc[0] == 0 && g->bc[1] == 0) {
        mirror[0] = 0, mirror[1] = 0, mirror[2] = 0, factor = 1.05; /* _0000, poisson.f90:35 */
    } else if (g->bc[0] == 0 && g->bc[1] == 1) {
        mirror[0] = 0, mirror[1] = 1, mirror[2] = 0, factor = 1.01; /* _0011, poisson.f90:162 */
    } else if (g->bc[0] == 1 && g->bc[1] == 1) {
        mirror[0] = 1, mirror[1] = 1, mirror[2] = 1, factor = 1.05; /* _111111, poisson.f90:287 */
    } else {
        return -1; /* pointer stays null */
    }
    return orc_poisson_sor(pp, rhs, g->dx, g->dy, g->dz, g->nx, g->ny, g->nz, mirror, factor,
                           omega, eps, kmax, idyn, dmax_out);
}

/* src/integration.f90:199-255 */
int orc_correct_pression(const orc_grid* g, double* pp, const double* ux_pred,
                         const double* uy_pred, const double* uz_pred, double dt, double* omega,
                         double eps, int kmax, int idyn, double* dmax_out, double* rhs_out) {
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    double* divu_pred = newfield(N);
    double* rhs = rhs_out ? rhs_out : newfield(N);
    orc_divergence(g, divu_pred, ux_pred, uy_pred, uz_pred, 1); /* :235-236 */
    for (size_t m = 0; m < N; ++m) rhs[m] = divu_pred[m] / dt;  /* :239 */
    const int it = orc_poisson_solver(g, pp, rhs, omega, eps, kmax, idyn, dmax_out); /* :247 */
    free(divu_pred);
    if (!rhs_out) free(rhs);
    return it;
}

/* src/integration.f90:257-330 */
int orc_correct_velocity(const orc_grid* g, double* ux, double* uy, double* uz,
                         const double* ux_pred, const double* uy_pred, const double* uz_pred,
                         const double* pp, double dt) {
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    double *dpdx = newfield(N), *dpdy = newfield(N), *dpdz = newfield(N);
    orc_derp(g, 0, 1, 0, dpdx, pp); /* :298-300 */
    orc_derp(g, 1, 1, 0, dpdy, pp);
    orc_derp(g, 2, 1, 0, dpdz, pp);
    int bad = 0;
    double mx = -HUGE_VAL, my = -HUGE_VAL, mz = -HUGE_VAL;
    for (size_t m = 0; m < N; ++m) { /* :304-306 */
        ux[m] = ux_pred[m] - dt * dpdx[m];
        uy[m] = uy_pred[m] - dt * dpdy[m];
        uz[m] = uz_pred[m] - dt * dpdz[m];
        if (ux[m] != ux[m] || uy[m] != uy[m] || uz[m] != uz[m]) bad = 1; /* contains_nan */
        if (ux[m] > mx) mx = ux[m];
        if (uy[m] > my) my = uy[m];
        if (uz[m] > mz) mz = uz[m];
    }
    if (mx > 1000. || my > 1000. || mz > 1000.) bad = 1; /* :310,:316,:322 */
    free(dpdx);
    free(dpdy);
    free(dpdz);
    return bad;
}

/* src/integration.f90:332-468 */
int orc_transeq(const orc_grid* g, double* phi, const double* ux, const double* uy,
                const double* uz, const double* src, double* fphi, double re, double sc,
                const double* adt, const double* bdt, const double* cdt, int itime, int itscheme,
                int iles, const double* nu_t) {
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    double adu, bdu, cdu;
    if (ab_select(adt, bdt, cdt, itime, itscheme, &adu, &bdu, &cdu)) return 1;
    double* d[6];
    for (int t = 0; t < 6; ++t) d[t] = newfield(N);
    orc_derp(g, 0, 1, 0, d[0], phi); /* :412-414 */
    orc_derp(g, 1, 1, 0, d[1], phi);
    orc_derp(g, 2, 1, 0, d[2], phi);
    orc_derp(g, 0, 2, 0, d[3], phi); /* :417-419 */
    orc_derp(g, 1, 2, 0, d[4], phi);
    orc_derp(g, 2, 2, 0, d[5], phi);
    double *f1 = fphi, *f2 = fphi + N, *f3 = fphi + 2 * N;
    const double resc = re * sc;
    for (size_t m = 0; m < N; ++m) {
        /* :403-409 */
        const double alpha_eff = (iles == 1) ? (1.0 / resc + nu_t[m] / sc) : (1.0 / resc);
        /* :422-423 */
        f1[m] = alpha_eff * (d[3][m] + d[4][m] + d[5][m]) -
                (ux[m] * d[0][m] + uy[m] * d[1][m] + uz[m] * d[2][m]) + (src ? src[m] : 0.0);
        /* :426 */
        phi[m] = phi[m] + adu * f1[m] + bdu * f2[m] + cdu * f3[m];
    }
    /* conservative clipping :432-450; sums in array-element order like the SUM intrinsic */
    const double count = (double)((long)g->nx * g->ny * g->nz);
    double s_old = 0.0;
    for (size_t m = 0; m < N; ++m) s_old += phi[m];
    const double phi_old_avg = s_old / count; /* :433 */
    double s_new = 0.0;
    for (size_t m = 0; m < N; ++m) {
        phi[m] = fmax(0.0, fmin(1.0, phi[m])); /* :436 */
        s_new += phi[m];
    }
    const double phi_new_avg = s_new / count; /* :439 */
    const double excess = phi_old_avg - phi_new_avg; /* :440 */
    double* weight = d[0];
    double s_w = 0.0;
    for (size_t m = 0; m < N; ++m) {
        weight[m] = fmin(phi[m], 1.0 - phi[m]); /* :443 */
        s_w += weight[m];
    }
    for (size_t m = 0; m < N; ++m) {
        const double wn = weight[m] / s_w;         /* :444 */
        phi[m] = phi[m] + excess * wn;             /* :447 */
        phi[m] = fmax(0.0, fmin(1.0, phi[m]));     /* :450 */
    }
    if (itscheme == 2) { /* :454-455 */
        memcpy(f2, f1, N * sizeof(double));
    } else if (itscheme == 3) { /* :456-458 */
        memcpy(f3, f2, N * sizeof(double));
        memcpy(f2, f1, N * sizeof(double));
    }
    for (int t = 0; t < 6; ++t) free(d[t]);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * src/utils.f90:243-375 and src/functions.f90:126-150 (average_3d_array: k innermost)
 * ---------------------------------------------------------------------------------- */
static double average_sq(const double* t, int nx, int ny, int nz) {
    double sum = 0.0;
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            for (int k = 0; k < nz; ++k) {
                const double v = t[IDX(i, j, k)];
                sum = sum + v * v;
            }
    return sum / (double)((long)nx * ny * nz);
}

void orc_statistics_calc(const orc_grid* g, const double* ux, const double* uy, const double* uz,
                         double re, double t, double* out) {
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    const size_t N = (size_t)nx * ny * nz;
    double* d[9];
    for (int q = 0; q < 9; ++q) d[q] = newfield(N);
    out[0] = t;
    out[5] = average_sq(ux, nx, ny, nz); /* utils.f90:277-279 */
    out[6] = average_sq(uy, nx, ny, nz);
    out[7] = average_sq(uz, nx, ny, nz);
    /* :282-291 (same parity table as calculate_nu_t) */
    orc_derp(g, 0, 1, 1, d[0], ux);
    orc_derp(g, 1, 1, 0, d[1], ux);
    orc_derp(g, 2, 1, 0, d[2], ux);
    orc_derp(g, 0, 1, 0, d[3], uy);
    orc_derp(g, 1, 1, 1, d[4], uy);
    orc_derp(g, 2, 1, 0, d[5], uy);
    orc_derp(g, 0, 1, 0, d[6], uz);
    orc_derp(g, 1, 1, 0, d[7], uz);
    orc_derp(g, 2, 1, 1, d[8], uz);
    for (int q = 0; q < 9; ++q) out[8 + q] = average_sq(d[q], nx, ny, nz); /* :294-302 */
    const double xnu = 1.0 / re;
    double dzeta = 0.0, eps = 0.0, e_k = 0.0;
    for (size_t m = 0; m < N; ++m) { /* :310-331, i fastest */
        e_k = e_k + 0.5 * (ux[m] * ux[m] + uy[m] * uy[m] + uz[m] * uz[m]);
        const double a = 2.0 * d[0][m], b = 2.0 * d[4][m], c = 2.0 * d[8][m];
        const double sxy = d[1][m] + d[3][m], sxz = d[2][m] + d[6][m], syz = d[5][m] + d[7][m];
        eps = eps + 0.5 * xnu *
                        (a * a + b * b + c * c + 2.0 * (sxy * sxy) + 2.0 * (sxz * sxz) +
                         2.0 * (syz * syz));
        const double wx = d[7][m] - d[5][m], wy = d[2][m] - d[6][m], wz = d[3][m] - d[1][m];
        dzeta = dzeta + 0.5 * (wx * wx + wy * wy + wz * wz);
    }
    const double cnt = (double)((long)nx * ny * nz);
    out[1] = e_k / cnt;   /* :334 */
    out[2] = eps / cnt;   /* :335 */
    out[4] = dzeta / cnt; /* :336 */
    /* :339-347 */
    orc_derp(g, 0, 2, 1, d[0], ux);
    orc_derp(g, 1, 2, 0, d[1], ux);
    orc_derp(g, 2, 2, 0, d[2], ux);
    orc_derp(g, 0, 2, 0, d[3], uy);
    orc_derp(g, 1, 2, 1, d[4], uy);
    orc_derp(g, 2, 2, 0, d[5], uy);
    orc_derp(g, 0, 2, 0, d[6], uz);
    orc_derp(g, 1, 2, 0, d[7], uz);
    orc_derp(g, 2, 2, 1, d[8], uz);
    double eps2 = 0.0;
    for (size_t m = 0; m < N; ++m) { /* :349-360 */
        const double t1 = (-xnu) * (ux[m] * (d[0][m] + d[1][m] + d[2][m]) +
                                    uy[m] * (d[3][m] + d[4][m] + d[5][m]) +
                                    uz[m] * (d[6][m] + d[7][m] + d[8][m]));
        eps2 = eps2 + t1;
    }
    out[3] = eps2 / cnt; /* :361 */
    for (int q = 0; q < 9; ++q) free(d[q]);
}

void orc_function_stats(const double* f, int nx, int ny, int nz, double* out) {
    double sum = 0.0, fmin_ = HUGE_VAL, fmax_ = -HUGE_VAL;
    int im = 1, jm = 1, km = 1;
    /* functions.f90:42-57; huge() vs HUGE_VAL only matters for all-inf input */
    fmin_ = 1.7976931348623157e308;
    fmax_ = -1.7976931348623157e308;
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const double v = f[IDX(i, j, k)];
                sum = sum + v;
                if (v < fmin_) fmin_ = v;
                if (v > fmax_) {
                    fmax_ = v;
                    im = i + 1, jm = j + 1, km = k + 1;
                }
            }
    out[0] = fmin_;
    out[1] = fmax_;
    out[2] = sum / (double)((long)nx * ny * nz);
    out[3] = im, out[4] = jm, out[5] = km;
}

void orc_calculate_residuals(const double* u, const double* v, const double* w,
                             const double* old_u, const double* old_v, const double* old_w,
                             double dt, double t_ref, double u_ref, int nx, int ny, int nz,
                             double* out15) {
    /* utils.f90:93-160: out = res_u res_v res_w | aa bb cc | ia ja ka ib jb kb ic jc kc */
    const double* nw[3] = {u, v, w};
    const double* od[3] = {old_u, old_v, old_w};
    for (int c = 0; c < 3; ++c) {
        double res = 0.0, linf = 0.0;
        /* :109-124 interior points 2 .. n-1 */
        for (int k = 1; k < nz - 1; ++k)
            for (int j = 1; j < ny - 1; ++j)
                for (int i = 1; i < nx - 1; ++i) {
                    const double a = fabs(od[c][IDX(i, j, k)] - nw[c][IDX(i, j, k)]) / (2.0 * dt);
                    res = res + a * a; /* (...)**2.d0 */
                    linf = a > linf ? a : linf;
                }
        /* :125-145: the LAST point whose value equals the maximum */
        int ia = 0, ja = 0, ka = 0;
        for (int k = 1; k < nz - 1; ++k)
            for (int j = 1; j < ny - 1; ++j)
                for (int i = 1; i < nx - 1; ++i) {
                    const double a = fabs(od[c][IDX(i, j, k)] - nw[c][IDX(i, j, k)]) / (2.0 * dt);
                    if (fabs(linf - a) <= 2.2250738585072014e-308) ia = i + 1, ja = j + 1, ka = k + 1;
                }
        /* :147-152; real(nx*ny*nz) is a default (single precision) real */
        const double cnt = (double)(float)(nx * ny * nz);
        out15[c] = (t_ref / u_ref) * sqrt((1.0 / cnt) * res);
        out15[3 + c] = (t_ref / u_ref) * linf;
        out15[6 + 3 * c] = ia, out15[7 + 3 * c] = ja, out15[8 + 3 * c] = ka;
    }
}

void orc_ab_coefficients(double dt, double* adt, double* bdt, double* cdt) {
    /* initialization.f90:194-202 */
    adt[0] = dt, bdt[0] = 0.0, cdt[0] = 0.0;
    adt[1] = 3.0 * dt / 2.0, bdt[1] = -1.0 * dt / 2.0, cdt[1] = 0.0;
    adt[2] = 23.0 * dt / 12.0, bdt[2] = -16.0 * dt / 12.0, cdt[2] = 5.0 * dt / 12.0;
}

/* ------------------------------------------------------------------------------------
 * initial conditions (inputs for the parity runs)
 * ---------------------------------------------------------------------------------- */
static void coords(const orc_grid* g, double x0, double y0, double z0, double** x, double** y,
                   double** z) {
    /* initialization.f90:211-219 */
    *x = (double*)malloc(sizeof(double) * g->nx);
    *y = (double*)malloc(sizeof(double) * g->ny);
    *z = (double*)malloc(sizeof(double) * g->nz);
    for (int i = 0; i < g->nx; ++i) (*x)[i] = x0 + (double)i * g->dx;
    for (int i = 0; i < g->ny; ++i) (*y)[i] = y0 + (double)i * g->dy;
    for (int i = 0; i < g->nz; ++i) (*z)[i] = z0 + (double)i * g->dz;
}

void orc_init_tgv(const orc_grid* g, double x0, double y0, double z0, double u0, double l0,
                  double ratio, int nscr, double delta, double* ux, double* uy, double* uz,
                  double* pp, double* phi) {
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    double *x, *y, *z;
    coords(g, x0, y0, z0, &x, &y, &z);
    for (int k = 0; k < nz; ++k) { /* initial_conditions.f90:141-153 */
        const double twoz = 2.0 * z[k];
        for (int j = 0; j < ny; ++j) {
            const double twoy = 2.0 * y[j];
            for (int i = 0; i < nx; ++i) {
                const double twox = 2.0 * x[i];
                const size_t m = IDX(i, j, k);
                ux[m] = ratio * u0 / l0 * sin(x[i]) * cos(y[j]) * cos(z[k]);
                uy[m] = -ratio * u0 / l0 * cos(x[i]) * sin(y[j]) * cos(z[k]);
                uz[m] = 0.0;
                pp[m] = 0.0625 * (cos(twox) + cos(twoy)) * (cos(twoz) + 2.0);
            }
        }
    }
    if (nscr == 1) { /* :155-170; note: uses the LES filter width `delta`, not `delt` */
        const double R = 0.25 * (x[nx - 1] - x[0]);
        const double cx = 0.5 * (x[0] + x[nx - 1]);
        const double cy = 0.5 * (y[0] + y[ny - 1]);
        const double cz = 0.5 * (z[0] + z[nz - 1]);
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) {
                    const double ddx = x[i] - cx, ddy = y[j] - cy, ddz = z[k] - cz;
                    const double dist = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
                    phi[IDX(i, j, k)] = 0.5 * (1.0 - tanh((dist - R) / delta));
                }
    }
    free(x);
    free(y);
    free(z);
}

void orc_init_mixing_layer(const orc_grid* g, double x0, double y0, double z0, double u0,
                           double l0, double ratio, int nscr, double* ux, double* uy,
                           double* uz, double* pp, double* phi) {
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    double *x, *y, *z;
    coords(g, x0, y0, z0, &x, &y, &z);
    const double tiny_value = 1.e-12; /* initial_conditions.f90:366-391 */
    double u1, u2;
    if (ratio < tiny_value && ratio > -tiny_value) {
        u2 = u0;
        u1 = 0.0;
    } else {
        u2 = u0 / (1.0 - ratio);
        u1 = u2 * ratio;
    }
    const double theta_o = 1.0 / (13.0 * l0);
    const double t1 = 0.5 * (u2 + u1), t2 = 0.5 * (u1 - u2);
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const size_t m = IDX(i, j, k);
                const double t3 = y[j] * log(2.0) / (2.0 * theta_o);
                ux[m] = x[i] * 0.0 + t1 - t2 * tanh(t3);
                uy[m] = y[j] * 0.0;
                uz[m] = z[k] * 0.0;
                if (nscr == 1) phi[m] = 0.5 - 0.5 * tanh(t3);
                pp[m] = 1.0;
            }
    free(x);
    free(y);
    free(z);
}

void orc_init_coplanar_jet(const orc_grid* g, double x0, double y0, double z0, double u0,
                           double l0, double ratio, int nscr, double* ux, double* uy,
                           double* uz, double* pp, double* phi) {
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    double *x, *y, *z;
    coords(g, x0, y0, z0, &x, &y, &z);
    /* initial_conditions.f90:284-323 */
    const double u2 = u0, u1 = u2 / ratio, u3 = 0.0 * u2;
    const double d1 = l0, d2 = 2.0 * d1, h1 = 0.5 * d1, h2 = 0.5 * d2;
    const double theta_1 = h1 / 10.0, theta_2 = h2 / 25.0, hm = 0.5 * (h1 + h2);
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const size_t m = IDX(i, j, k);
                if (fabs(y[j]) < hm) {
                    const double th = tanh((fabs(y[j]) - h1) / (2.0 * theta_1));
                    ux[m] = x[i] * 0.0 + 0.5 * (u1 + u2) + 0.5 * (u2 - u1) * th;
                    if (nscr == 1) phi[m] = 0.5 * (0.0 + 1.0) + 0.5 * (1.0 - 0.0) * th;
                } else {
                    const double th = tanh((fabs(y[j]) - h2) / (2.0 * theta_2));
                    ux[m] = 0.5 * (u2 + u3) + 0.5 * (u3 - u2) * th;
                    if (nscr == 1) phi[m] = 0.5 * (1.0 + 0.0) + 0.5 * (0.0 - 1.0) * th;
                }
                uy[m] = y[j] * 0.0;
                uz[m] = z[k] * 0.0;
                pp[m] = 0.0;
            }
    free(x);
    free(y);
    free(z);
}

/* ------------------------------------------------------------------------------------
 * the main loop body, src/osinco3d_main.f90:97-115 (hot path only)
 * ---------------------------------------------------------------------------------- */
orc_sim* orc_sim_create(const orc_grid* g, double re, double sc, double cs, double delta,
                        double dt, int itscheme, int iles, int nscr, double omega, double eps,
                        int kmax, int idyn) {
    orc_sim* s = (orc_sim*)calloc(1, sizeof(orc_sim));
    s->g = *g;
    s->re = re, s->sc = sc, s->cs = cs, s->delta = delta, s->dt = dt;
    orc_ab_coefficients(dt, s->adt, s->bdt, s->cdt);
    s->itscheme = itscheme, s->iles = iles, s->nscr = nscr;
    s->omega = omega, s->eps = eps, s->kmax = kmax, s->idyn = idyn;
    const size_t N = (size_t)g->nx * g->ny * g->nz;
    double** f1[] = {&s->ux, &s->uy, &s->uz, &s->pp, &s->phi, &s->ux_pred, &s->uy_pred,
                     &s->uz_pred, &s->nu_t};
    for (unsigned q = 0; q < sizeof(f1) / sizeof(f1[0]); ++q)
        *f1[q] = (double*)calloc(N, sizeof(double));
    double** f3[] = {&s->fux, &s->fuy, &s->fuz, &s->fphi};
    for (unsigned q = 0; q < 4; ++q) *f3[q] = (double*)calloc(3 * N, sizeof(double));
    return s;
}

void orc_sim_destroy(orc_sim* s) {
    if (!s) return;
    double* all[] = {s->ux, s->uy, s->uz, s->pp, s->phi, s->ux_pred, s->uy_pred, s->uz_pred,
                     s->nu_t, s->fux, s->fuy, s->fuz, s->fphi};
    for (unsigned q = 0; q < sizeof(all) / sizeof(all[0]); ++q) free(all[q]);
    free(s);
}

int orc_sim_step(orc_sim* s, int itime) {
    const orc_grid* g = &s->g;
    if (orc_predict_velocity(g, s->ux_pred, s->uy_pred, s->uz_pred, s->ux, s->uy, s->uz, s->fux,
                             s->fuy, s->fuz, s->re, s->adt, s->bdt, s->cdt, itime, s->itscheme,
                             s->iles, s->cs, s->delta, s->nu_t))
        return 1;
    s->last_iters = orc_correct_pression(g, s->pp, s->ux_pred, s->uy_pred, s->uz_pred, s->dt,
                                         &s->omega, s->eps, s->kmax, s->idyn, &s->last_dmax, 0);
    if (s->last_iters < 0) return 2;
    s->total_iters += s->last_iters;
    if (orc_correct_velocity(g, s->ux, s->uy, s->uz, s->ux_pred, s->uy_pred, s->uz_pred, s->pp,
                             s->dt))
        return 3;
    if (s->nscr == 1)
        if (orc_transeq(g, s->phi, s->ux, s->uy, s->uz, 0, s->fphi, s->re, s->sc, s->adt,
                        s->bdt, s->cdt, itime, s->itscheme, s->iles, s->nu_t))
            return 4;
    return 0;
}

double* orc_sim_field(orc_sim* s, const char* name) {
    struct {
        const char* n;
        double* p;
    } tab[] = {{"ux", s->ux},     {"uy", s->uy},           {"uz", s->uz},
               {"pp", s->pp},     {"phi", s->phi},         {"ux_pred", s->ux_pred},
               {"uy_pred", s->uy_pred}, {"uz_pred", s->uz_pred}, {"nu_t", s->nu_t},
               {"fux", s->fux},   {"fuy", s->fuy},         {"fuz", s->fuz},
               {"fphi", s->fphi}};
    for (unsigned q = 0; q < sizeof(tab) / sizeof(tab[0]); ++q)
        if (!strcmp(tab[q].n, name)) return tab[q].p;
    return 0;
}
