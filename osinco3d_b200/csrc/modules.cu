// modules.cu -- section A of include/o3d_b200.h: stateless, HOST-pointer drop-ins for the
// reference's Fortran module procedures.  Each call stages its arguments through a cached
// scratch session (upload -> CUDA kernels -> download); nothing is computed on the host.
#include <cstring>

#include "session.h"

using namespace o3d;

namespace {

o3d_session* g_scratch = nullptr;

void free_scratch() {
    if (g_scratch) o3d_session_destroy(g_scratch);
    g_scratch = nullptr;
}

// scratch session for an (nx,ny,nz) problem under the currently bound schemes()
int scratch(int nx, int ny, int nz, double dx, double dy, double dz, o3d_session** out) {
    int rc = ensure_device();
    if (rc) return rc;
    if (nx < 7 || ny < 7 || nz < 7) {
        set_error("grid extents must be >= 7: %d %d %d", nx, ny, nz);
        return O3D_ERR_INVALID;
    }
    const Schemes& sc = g_schemes;
    o3d_session* s = g_scratch;
    const bool reuse = s && s->cfg.nx == nx && s->cfg.ny == ny && s->cfg.nz == nz &&
                       s->cfg.nbcx1 == sc.flags[0] && s->cfg.nbcxn == sc.flags[1] &&
                       s->cfg.nbcy1 == sc.flags[2] && s->cfg.nbcyn == sc.flags[3] &&
                       s->cfg.nbcz1 == sc.flags[4] && s->cfg.nbczn == sc.flags[5] &&
                       s->cfg.sim2d == sc.sim2d;
    if (!reuse) {
        free_scratch();
        o3d_config c;
        memset(&c, 0, sizeof(c));
        c.nx = nx, c.ny = ny, c.nz = nz;
        c.dx = dx, c.dy = dy, c.dz = dz;
        c.nbcx1 = sc.flags[0], c.nbcxn = sc.flags[1], c.nbcy1 = sc.flags[2];
        c.nbcyn = sc.flags[3], c.nbcz1 = sc.flags[4], c.nbczn = sc.flags[5];
        c.sim2d = sc.sim2d;
        c.re = 1.0, c.sc = 1.0, c.dt = 1.0, c.itscheme = 1;
        c.omega = 1.0, c.eps = 1e-6, c.kmax = 1;
        c.nranks = 1;
        if ((rc = o3d_session_create(&c, &g_scratch))) return rc;
        static bool registered = false;
        if (!registered) {
            atexit(free_scratch);
            registered = true;
        }
        s = g_scratch;
    }
    s->cfg.dx = dx, s->cfg.dy = dy, s->cfg.dz = dz;
    s->cfg.nranks = 1, s->cfg.rank = 0;
    s->use_src = 0;
    fill_geom(s);
    *out = s;
    return O3D_OK;
}

int up(o3d_session* s, int fid, const double* h) { return o3d_upload(s, fid, h); }
int down(o3d_session* s, int fid, double* h) { return o3d_download(s, fid, h); }

void reset_history(o3d_session* s) {
    for (int c = 0; c < 4; ++c)
        for (int l = 0; l < 3; ++l) s->lv[c][l] = l;
}

void set_ab(o3d_session* s, const double* adt, const double* bdt, const double* cdt) {
    for (int q = 0; q < 3; ++q) s->cfg.adt[q] = adt[q], s->cfg.bdt[q] = bdt[q], s->cfg.cdt[q] = cdt[q];
}

int der_impl(int axis, int order, int closure, double* df, const double* f, double d, int nx,
             int ny, int nz) {
    if (!df || !f || axis < 0 || axis > 2 || (order != 1 && order != 2) || closure < 0 ||
        closure > O3D_CLOSURE_D11)
        return O3D_ERR_INVALID;
    o3d_session* s;
    int rc = scratch(nx, ny, nz, d, d, d, &s);
    if (rc) return rc;
    // the routine's own closure, whatever schemes() bound: put it into the ghost cells
    Geom g = s->g;
    const int mode = (closure == O3D_CLOSURE_00) ? BM_WRAP : BM_MIRROR;
    if (axis == 0) g.bx = mode;
    if (axis == 1) g.by = mode;
    if (axis == 2) g.bz_lo = g.bz_hi = mode;
    if ((rc = up(s, O3D_F_SCRATCH0, f))) return rc;
    double* src = field(s, O3D_F_SCRATCH0);
    double* dst = field(s, O3D_F_SCRATCH1);
    if (!src || !dst) return O3D_ERR_CUDA;
    const int zero = (closure == O3D_CLOSURE_2DSIM);
    if (!zero) {
        GhostArgs ga;
        ga.njobs = 1;
        ga.job[0].p = src;
        ga.job[0].par = (closure == O3D_CLOSURE_I11)   ? (1u << axis)
                        : (closure == O3D_CLOSURE_D11) ? (1u << (axis + 4))
                                                       : 0u;
        ga.job[0].axes = 1u << axis;
        if (launch_fill_ghosts(s->st, g, ga)) return O3D_ERR_CUDA;
    }
    touch(s, O3D_F_SCRATCH0);
    if (launch_der(s->st, g, axis, order, zero, d, src, dst)) {
        set_error("derivative kernel launch failed");
        return O3D_ERR_CUDA;
    }
    touch(s, O3D_F_SCRATCH1);
    return down(s, O3D_F_SCRATCH1, df);
}

// closure bound to a pointer name by schemes(): parity 0 = 'p', 1 = 'i'
int bound_closure(int axis, int parity) {
    if (axis == 2 && g_schemes.sim2d == 1) return O3D_CLOSURE_2DSIM;
    if (g_schemes.bc[axis] == O3D_PERIODIC) return O3D_CLOSURE_00;
    return parity ? O3D_CLOSURE_I11 : O3D_CLOSURE_P11;
}

int poisson_impl(int variant, double* pp, const double* rhs, double dx, double dy, double dz,
                 int nx, int ny, int nz, double* omega, double eps, int kmax, int idyn, int* iters,
                 double* dmax) {
    if (!pp || !rhs || !omega) return O3D_ERR_INVALID;
    if (variant < 0) {
        set_error("poisson_solver pointer is null for the bound boundary flags");
        return O3D_ERR_BC;
    }
    o3d_session* s;
    int rc = scratch(nx, ny, nz, dx, dy, dz, &s);
    if (rc) return rc;
    s->sor_variant = variant;
    s->cfg.eps = eps, s->cfg.kmax = kmax, s->cfg.idyn = idyn;
    s->cfg.sor_order = g_sor_order;
    s->cfg.sor_check_every = 0;
    s->omega = *omega;
    s->last_iters = 0;
    if ((rc = up(s, O3D_F_PP, pp))) return rc;
    if ((rc = up(s, O3D_F_RHS, rhs))) return rc;
    rc = sor_solve(s, field(s, O3D_F_PP), field(s, O3D_F_RHS), iters, dmax);
    if (rc) return rc;
    *omega = s->omega;
    return down(s, O3D_F_PP, pp);
}

}  // namespace

extern "C" {

int o3d_schemes(int nbcx1, int nbcxn, int nbcy1, int nbcyn, int nbcz1, int nbczn, int sim2d) {
    Schemes sc;
    int rc;
    if ((rc = axis_bc(nbcx1, nbcxn, &sc.bc[0]))) return rc;
    if ((rc = axis_bc(nbcy1, nbcyn, &sc.bc[1]))) return rc;
    sc.bc[2] = O3D_PERIODIC;
    if (sim2d == 0) {
        if ((rc = axis_bc(nbcz1, nbczn, &sc.bc[2]))) return rc;
    } else if (sim2d != 1) {
        // the reference leaves the z pointers unbound (src/initialization.f90:259-281)
        set_error("sim2d must be 0 or 1");
        return O3D_ERR_BC;
    }
    sc.flags[0] = nbcx1, sc.flags[1] = nbcxn, sc.flags[2] = nbcy1;
    sc.flags[3] = nbcyn, sc.flags[4] = nbcz1, sc.flags[5] = nbczn;
    if (sim2d == 1) {  // z flags are not inspected when sim2d == 1; keep the session valid
        if (!((nbcz1 == 0 && nbczn == 0) || (nbcz1 == 1 && nbczn == 1)))
            sc.flags[4] = sc.flags[5] = O3D_PERIODIC;
    }
    sc.sim2d = sim2d;
    sc.poisson_variant = poisson_variant_of(nbcx1, nbcxn, nbcy1, nbcyn);
    sc.bound = 1;
    g_schemes = sc;
    return O3D_OK;
}

int o3d_der(int axis, int order, int closure, double* df, const double* f, double d, int nx,
            int ny, int nz) {
    return der_impl(axis, order, closure, df, f, d, nx, ny, nz);
}

#define O3D_DEF_DER(name, axis, order, closure)                                        \
    int o3d_##name(double* df, const double* f, double d, int nx, int ny, int nz) {    \
        return der_impl(axis, order, closure, df, f, d, nx, ny, nz);                   \
    }
O3D_DEF_DER(derx_00, 0, 1, O3D_CLOSURE_00)
O3D_DEF_DER(derxp_11, 0, 1, O3D_CLOSURE_P11)
O3D_DEF_DER(derxi_11, 0, 1, O3D_CLOSURE_I11)
O3D_DEF_DER(dery_00, 1, 1, O3D_CLOSURE_00)
O3D_DEF_DER(deryp_11, 1, 1, O3D_CLOSURE_P11)
O3D_DEF_DER(deryi_11, 1, 1, O3D_CLOSURE_I11)
O3D_DEF_DER(derz_00, 2, 1, O3D_CLOSURE_00)
O3D_DEF_DER(derzp_11, 2, 1, O3D_CLOSURE_P11)
O3D_DEF_DER(derzi_11, 2, 1, O3D_CLOSURE_I11)
O3D_DEF_DER(derxx_00, 0, 2, O3D_CLOSURE_00)
O3D_DEF_DER(derxxp_11, 0, 2, O3D_CLOSURE_P11)
O3D_DEF_DER(derxxi_11, 0, 2, O3D_CLOSURE_I11)
O3D_DEF_DER(deryy_00, 1, 2, O3D_CLOSURE_00)
O3D_DEF_DER(deryyp_11, 1, 2, O3D_CLOSURE_P11)
O3D_DEF_DER(deryyi_11, 1, 2, O3D_CLOSURE_I11)
O3D_DEF_DER(derzz_00, 2, 2, O3D_CLOSURE_00)
O3D_DEF_DER(derzzp_11, 2, 2, O3D_CLOSURE_P11)
O3D_DEF_DER(derzzi_11, 2, 2, O3D_CLOSURE_I11)
O3D_DEF_DER(derz_2dsim, 2, 1, O3D_CLOSURE_2DSIM)
O3D_DEF_DER(derzz_2dsim, 2, 2, O3D_CLOSURE_2DSIM)
#undef O3D_DEF_DER

#define O3D_DEF_PTR(name, axis, order, parity)                                         \
    int o3d_##name(double* df, const double* f, double d, int nx, int ny, int nz) {    \
        return der_impl(axis, order, bound_closure(axis, parity), df, f, d, nx, ny, nz); \
    }
O3D_DEF_PTR(derxp, 0, 1, 0)
O3D_DEF_PTR(derxxp, 0, 2, 0)
O3D_DEF_PTR(derxi, 0, 1, 1)
O3D_DEF_PTR(derxxi, 0, 2, 1)
O3D_DEF_PTR(deryp, 1, 1, 0)
O3D_DEF_PTR(deryyp, 1, 2, 0)
O3D_DEF_PTR(deryi, 1, 1, 1)
O3D_DEF_PTR(deryyi, 1, 2, 1)
O3D_DEF_PTR(derzp, 2, 1, 0)
O3D_DEF_PTR(derzzp, 2, 2, 0)
O3D_DEF_PTR(derzi, 2, 1, 1)
O3D_DEF_PTR(derzzi, 2, 2, 1)
#undef O3D_DEF_PTR

int o3d_divergence(double* divf, const double* fx, const double* fy, const double* fz,
                   double dx, double dy, double dz, int nx, int ny, int nz, int odd) {
    if (!divf || !fx || !fy || !fz) return O3D_ERR_INVALID;
    o3d_session* s;
    int rc = scratch(nx, ny, nz, dx, dy, dz, &s);
    if (rc) return rc;
    if ((rc = up(s, O3D_F_UX_PRED, fx)) || (rc = up(s, O3D_F_UY_PRED, fy)) ||
        (rc = up(s, O3D_F_UZ_PRED, fz)))
        return rc;
    if ((rc = o3d_s_divergence(s, O3D_F_UX_PRED, O3D_F_UY_PRED, O3D_F_UZ_PRED, O3D_F_DIVU,
                               odd == 0 ? 0 : 1)))
        return rc;
    return down(s, O3D_F_DIVU, divf);
}

int o3d_rotational(double* rotx, double* roty, double* rotz, const double* ux, const double* uy,
                   const double* uz, double dx, double dy, double dz, int nx, int ny, int nz) {
    if (!rotx || !roty || !rotz || !ux || !uy || !uz) return O3D_ERR_INVALID;
    o3d_session* s;
    int rc = scratch(nx, ny, nz, dx, dy, dz, &s);
    if (rc) return rc;
    if ((rc = up(s, O3D_F_UX, ux)) || (rc = up(s, O3D_F_UY, uy)) || (rc = up(s, O3D_F_UZ, uz)))
        return rc;
    if ((rc = o3d_s_rotational(s, O3D_F_SCRATCH0, O3D_F_SCRATCH1, O3D_F_SCRATCH2))) return rc;
    if ((rc = down(s, O3D_F_SCRATCH0, rotx)) || (rc = down(s, O3D_F_SCRATCH1, roty)) ||
        (rc = down(s, O3D_F_SCRATCH2, rotz)))
        return rc;
    return O3D_OK;
}

int o3d_calculate_q_criterion(double* q, const double* ux, const double* uy, const double* uz,
                              double dx, double dy, double dz, int nx, int ny, int nz) {
    if (!q || !ux || !uy || !uz) return O3D_ERR_INVALID;
    o3d_session* s;
    int rc = scratch(nx, ny, nz, dx, dy, dz, &s);
    if (rc) return rc;
    if ((rc = up(s, O3D_F_UX, ux)) || (rc = up(s, O3D_F_UY, uy)) || (rc = up(s, O3D_F_UZ, uz)))
        return rc;
    if ((rc = o3d_s_q_criterion(s, O3D_F_SCRATCH0))) return rc;
    return down(s, O3D_F_SCRATCH0, q);
}

int o3d_calculate_nu_t(double* nu_t, const double* ux, const double* uy, const double* uz,
                       double dx, double dy, double dz, double cs, double delta, int nx, int ny,
                       int nz, double* stats6) {
    if (!nu_t || !ux || !uy || !uz) return O3D_ERR_INVALID;
    o3d_session* s;
    int rc = scratch(nx, ny, nz, dx, dy, dz, &s);
    if (rc) return rc;
    if ((rc = up(s, O3D_F_UX, ux)) || (rc = up(s, O3D_F_UY, uy)) || (rc = up(s, O3D_F_UZ, uz)))
        return rc;
    const double csd = cs * delta;
    double* out = field(s, O3D_F_NU_T);
    if (!out) return O3D_ERR_CUDA;
    const int ids[3] = {O3D_F_UX, O3D_F_UY, O3D_F_UZ};
    const unsigned par[3] = {0x1u, 0x2u, 0x4u};  // src/les_turbulence.f90:55-67
    if ((rc = ensure_ghosts(s, ids, 3, par, 0x7u))) return rc;
    FieldRef u[3] = {fref(s, O3D_F_UX), fref(s, O3D_F_UY), fref(s, O3D_F_UZ)};
    if (launch_nu_t(s->st, s->g, u, s->cx, s->cy, s->cz, csd * csd, out)) return O3D_ERR_CUDA;
    touch(s, O3D_F_NU_T);
    if (stats6 && (rc = o3d_s_function_stats(s, O3D_F_NU_T, stats6))) return rc;
    return down(s, O3D_F_NU_T, nu_t);
}

int o3d_predict_velocity(double* ux_pred, double* uy_pred, double* uz_pred, const double* ux,
                         const double* uy, const double* uz, double* fux, double* fuy,
                         double* fuz, double re, const double* adt, const double* bdt,
                         const double* cdt, int itime, int itscheme, double dx, double dy,
                         double dz, int nx, int ny, int nz, int iles, double cs, double delta,
                         double* nu_t) {
    if (!ux_pred || !uy_pred || !uz_pred || !ux || !uy || !uz || !fux || !fuy || !fuz || !adt ||
        !bdt || !cdt || !nu_t)
        return O3D_ERR_INVALID;
    o3d_session* s;
    int rc = scratch(nx, ny, nz, dx, dy, dz, &s);
    if (rc) return rc;
    s->cfg.re = re, s->cfg.itscheme = itscheme, s->cfg.iles = iles, s->cfg.cs = cs;
    s->cfg.delta = delta;
    set_ab(s, adt, bdt, cdt);
    reset_history(s);
    const size_t N = (size_t)nx * ny * nz;
    double* fh[3] = {fux, fuy, fuz};
    const int fb[3] = {O3D_F_FUX1, O3D_F_FUY1, O3D_F_FUZ1};
    if (iles != 1) {  // nu_t = 0.0d0, src/integration.f90:112
        double* nt = field(s, O3D_F_NU_T);
        if (!nt) return O3D_ERR_CUDA;
        O3D_CUDA_CHECK(cudaMemsetAsync(nt - interior_offset(s->g), 0,
                                       (size_t)s->felems * sizeof(double), s->st));
    }
    if (pipe_chunks(nz)) {  // z-chunk pipeline: both PCIe directions busy at once (pipeline.cu)
        double* const up_h[3] = {ux_pred, uy_pred, uz_pred};
        const double* const u_h[3] = {ux, uy, uz};
        return pipe_predict_velocity(s, itime, up_h, u_h, fh, nu_t);
    }
    if ((rc = up(s, O3D_F_UX, ux)) || (rc = up(s, O3D_F_UY, uy)) || (rc = up(s, O3D_F_UZ, uz)))
        return rc;
    for (int c = 0; c < 3; ++c) {
        // level 1 is overwritten before it is read (src/integration.f90:129); levels 2,3 are inputs
        if ((rc = up(s, fb[c] + 1, fh[c] + N)) || (rc = up(s, fb[c] + 2, fh[c] + 2 * N))) return rc;
    }
    if ((rc = o3d_s_predict_velocity(s, itime))) return rc;
    if ((rc = down(s, O3D_F_UX_PRED, ux_pred)) || (rc = down(s, O3D_F_UY_PRED, uy_pred)) ||
        (rc = down(s, O3D_F_UZ_PRED, uz_pred)) || (rc = down(s, O3D_F_NU_T, nu_t)))
        return rc;
    for (int c = 0; c < 3; ++c)
        for (int l = 0; l < 3; ++l)
            if ((rc = down(s, fb[c] + l, fh[c] + (size_t)l * N))) return rc;
    return O3D_OK;
}

int o3d_poisson_solver_0000(double* pp, const double* rhs, double dx, double dy, double dz,
                            int nx, int ny, int nz, double* omega, double eps, int kmax,
                            int idyn, int* iters, double* dmax) {
    return poisson_impl(0, pp, rhs, dx, dy, dz, nx, ny, nz, omega, eps, kmax, idyn, iters, dmax);
}
int o3d_poisson_solver_0011(double* pp, const double* rhs, double dx, double dy, double dz,
                            int nx, int ny, int nz, double* omega, double eps, int kmax,
                            int idyn, int* iters, double* dmax) {
    return poisson_impl(1, pp, rhs, dx, dy, dz, nx, ny, nz, omega, eps, kmax, idyn, iters, dmax);
}
int o3d_poisson_solver_111111(double* pp, const double* rhs, double dx, double dy, double dz,
                              int nx, int ny, int nz, double* omega, double eps, int kmax,
                              int idyn, int* iters, double* dmax) {
    return poisson_impl(2, pp, rhs, dx, dy, dz, nx, ny, nz, omega, eps, kmax, idyn, iters, dmax);
}
int o3d_poisson_solver(double* pp, const double* rhs, double dx, double dy, double dz, int nx,
                       int ny, int nz, double* omega, double eps, int kmax, int idyn, int* iters,
                       double* dmax) {
    return poisson_impl(g_schemes.poisson_variant, pp, rhs, dx, dy, dz, nx, ny, nz, omega, eps,
                        kmax, idyn, iters, dmax);
}

int o3d_solve_poisson_multigrid(double* phi, const double* rhs, double dx, double dy, double dz,
                                int nx, int ny, int nz, int nlevels, int npre, int npost,
                                double tol, int* cycles, double* dmax) {
    if (!phi || !rhs) return O3D_ERR_INVALID;
    o3d_session* s;
    int rc = scratch(nx, ny, nz, dx, dy, dz, &s);
    if (rc) return rc;
    if (g_schemes.poisson_variant < 0) return O3D_ERR_BC;
    s->sor_variant = g_schemes.poisson_variant;
    if ((rc = up(s, O3D_F_PP, phi)) || (rc = up(s, O3D_F_RHS, rhs))) return rc;
    if ((rc = mg_solve(s, field(s, O3D_F_PP), field(s, O3D_F_RHS), nlevels, npre, npost, tol,
                       cycles, dmax)))
        return rc;
    return down(s, O3D_F_PP, phi);
}

int o3d_correct_pression(double* pp, const double* ux_pred, const double* uy_pred,
                         const double* uz_pred, double dx, double dy, double dz, int nx, int ny,
                         int nz, double dt, double* omega, double eps, int kmax, int idyn,
                         int multigrid, int* iters, double* dmax) {
    if (!pp || !ux_pred || !uy_pred || !uz_pred || !omega) return O3D_ERR_INVALID;
    o3d_session* s;
    int rc = scratch(nx, ny, nz, dx, dy, dz, &s);
    if (rc) return rc;
    s->sor_variant = g_schemes.poisson_variant;
    s->cfg.dt = dt, s->cfg.eps = eps, s->cfg.kmax = kmax, s->cfg.idyn = idyn;
    s->cfg.multigrid = multigrid;
    s->cfg.sor_order = g_sor_order;
    s->cfg.sor_check_every = 0;
    s->omega = *omega;
    s->last_iters = 0;
    if ((rc = up(s, O3D_F_PP, pp)) || (rc = up(s, O3D_F_UX_PRED, ux_pred)) ||
        (rc = up(s, O3D_F_UY_PRED, uy_pred)) || (rc = up(s, O3D_F_UZ_PRED, uz_pred)))
        return rc;
    if ((rc = o3d_s_correct_pression(s, iters, dmax))) return rc;
    *omega = s->omega;
    return down(s, O3D_F_PP, pp);
}

int o3d_correct_velocity(double* ux, double* uy, double* uz, const double* ux_pred,
                         const double* uy_pred, const double* uz_pred, const double* pp,
                         double dt, double dx, double dy, double dz, int nx, int ny, int nz) {
    if (!ux || !uy || !uz || !ux_pred || !uy_pred || !uz_pred || !pp) return O3D_ERR_INVALID;
    o3d_session* s;
    int rc = scratch(nx, ny, nz, dx, dy, dz, &s);
    if (rc) return rc;
    s->cfg.dt = dt;
    if (pipe_chunks(nz)) {  // z-chunk pipeline (pipeline.cu)
        double* const u_h[3] = {ux, uy, uz};
        const double* const up_h[3] = {ux_pred, uy_pred, uz_pred};
        return pipe_correct_velocity(s, u_h, up_h, pp);
    }
    if ((rc = up(s, O3D_F_PP, pp)) || (rc = up(s, O3D_F_UX_PRED, ux_pred)) ||
        (rc = up(s, O3D_F_UY_PRED, uy_pred)) || (rc = up(s, O3D_F_UZ_PRED, uz_pred)))
        return rc;
    const int rcv = o3d_s_correct_velocity(s);
    if (rcv != O3D_OK && rcv != O3D_ERR_DIVERGED) return rcv;
    if ((rc = down(s, O3D_F_UX, ux)) || (rc = down(s, O3D_F_UY, uy)) || (rc = down(s, O3D_F_UZ, uz)))
        return rc;
    return rcv;
}

int o3d_transeq(double* phi, const double* ux, const double* uy, const double* uz,
                const double* src, double* fphi, double re, double sc, const double* adt,
                const double* bdt, const double* cdt, int itime, int itscheme, double dx,
                double dy, double dz, int nx, int ny, int nz, int iles, const double* nu_t) {
    if (!phi || !ux || !uy || !uz || !fphi || !adt || !bdt || !cdt || (iles == 1 && !nu_t))
        return O3D_ERR_INVALID;
    o3d_session* s;
    int rc = scratch(nx, ny, nz, dx, dy, dz, &s);
    if (rc) return rc;
    s->cfg.re = re, s->cfg.sc = sc, s->cfg.itscheme = itscheme, s->cfg.iles = iles;
    set_ab(s, adt, bdt, cdt);
    reset_history(s);
    const size_t N = (size_t)nx * ny * nz;
    if ((rc = up(s, O3D_F_PHI, phi)) || (rc = up(s, O3D_F_UX, ux)) || (rc = up(s, O3D_F_UY, uy)) ||
        (rc = up(s, O3D_F_UZ, uz)))
        return rc;
    if (iles == 1 && (rc = up(s, O3D_F_NU_T, nu_t))) return rc;
    if ((rc = up(s, O3D_F_FPHI2, fphi + N)) || (rc = up(s, O3D_F_FPHI3, fphi + 2 * N))) return rc;
    if (src) {
        if ((rc = up(s, O3D_F_SCRATCH1, src))) return rc;
        s->use_src = 1;
    }
    rc = o3d_s_transeq(s, itime);
    s->use_src = 0;
    if (rc) return rc;
    if ((rc = down(s, O3D_F_PHI, phi))) return rc;
    for (int l = 0; l < 3; ++l)
        if ((rc = down(s, O3D_F_FPHI1 + l, fphi + (size_t)l * N))) return rc;
    return O3D_OK;
}

int o3d_statistics_calc(const double* ux, const double* uy, const double* uz, int nx, int ny,
                        int nz, double dx, double dy, double dz, double re, double t,
                        double* out17) {
    if (!ux || !uy || !uz || !out17) return O3D_ERR_INVALID;
    o3d_session* s;
    int rc = scratch(nx, ny, nz, dx, dy, dz, &s);
    if (rc) return rc;
    s->cfg.re = re;
    if ((rc = up(s, O3D_F_UX, ux)) || (rc = up(s, O3D_F_UY, uy)) || (rc = up(s, O3D_F_UZ, uz)))
        return rc;
    return o3d_s_statistics(s, t, out17);
}

int o3d_function_stats(const double* f, int nx, int ny, int nz, double* stats6) {
    if (!f || !stats6) return O3D_ERR_INVALID;
    o3d_session* s;
    int rc = scratch(nx, ny, nz, 1.0, 1.0, 1.0, &s);
    if (rc) return rc;
    if ((rc = up(s, O3D_F_SCRATCH0, f))) return rc;
    return o3d_s_function_stats(s, O3D_F_SCRATCH0, stats6);
}

}  // extern "C"
