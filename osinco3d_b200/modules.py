"""Host-side mirror of the reference's Fortran module procedures (section A of
include/o3d_b200.h): same names, argument meaning and error behaviour as `derivation`,
`diffoper`, `les_turbulence`, `poisson`, `poisson_multigrid` and `integration` in
jojoledemago/osinco3d, over numpy arrays in Fortran order (nx,ny,nz), float64.

Everything is computed by the CUDA kernels in libo3d_b200.so; there is no CPU fallback.
Where the reference prints + stops, an O3DError carrying the C status code is raised.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import O3DError, check, lib  # noqa: F401


def _p(a):
    if a.dtype != np.float64 or not a.flags["F_CONTIGUOUS"]:
        raise ValueError("arrays must be float64 and Fortran-contiguous (nx,ny,nz)")
    return a.ctypes.data_as(L.dp)


def _like(a):
    return np.empty(a.shape, dtype=np.float64, order="F")


def _v3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def schemes(nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d=0):
    """initialization.schemes(), src/initialization.f90:226-304"""
    check(lib().o3d_schemes(nbcx1, nbcxn, nbcy1, nbcyn, nbcz1, nbczn, sim2d))


def set_sor_order(order):
    check(lib().o3d_set_sor_order(order))


def set_pipeline(chunks):
    """z-chunk copy pipelining of predict_velocity / correct_velocity (0 = off; default
    O3D_PIPELINE or 16); no counterpart in the reference.  Results are bitwise those of the
    unpipelined procedures."""
    check(lib().o3d_set_pipeline(int(chunks)))


def get_pipeline():
    return lib().o3d_get_pipeline()


def set_hostshift(on):
    """1: the pipelined predict_velocity downloads only history level 1 and makes levels 2 and 3
    (src/integration.f90:176-188) and the DNS nu_t = 0 in the host arrays with worker threads;
    bitwise the same arrays"""
    check(lib().o3d_set_hostshift(1 if on else 0))


def get_hostshift():
    return lib().o3d_get_hostshift()


def e2e_bytes_per_step(N, iles=0, itscheme=3):
    """(H2D, D2H) bytes one predict_velocity + correct_pression + correct_velocity chain moves for
    N grid points under the current settings (csrc/modules.cu, csrc/pipeline.cu)"""
    h2d = (3 + 6) + 4 + 4
    if get_hostshift() and get_pipeline() >= 2:
        down_pred = 3 + 3 + (1 if iles == 1 else 0)
    else:
        down_pred = 3 + 1 + 9
    d2h = down_pred + 1 + 3
    return h2d * N * 8, d2h * N * 8


def der(axis, order, closure, f, d):
    df = _like(f)
    nx, ny, nz = f.shape
    check(lib().o3d_der(axis, order, closure, _p(df), _p(f), C.c_double(d), nx, ny, nz))
    return df


def _make_der(name):
    def fn(f, d):
        df = _like(f)
        nx, ny, nz = f.shape
        check(getattr(lib(), "o3d_" + name)(_p(df), _p(f), C.c_double(d), nx, ny, nz))
        return df
    fn.__name__ = name
    fn.__doc__ = "derivation.%s(df, f, d) -> df   (src/derivation.f90 / initialization.f90:104-109)" % name
    return fn


DER_ROUTINES = ["derx_00", "derxp_11", "derxi_11", "dery_00", "deryp_11", "deryi_11",
                "derz_00", "derzp_11", "derzi_11", "derxx_00", "derxxp_11", "derxxi_11",
                "deryy_00", "deryyp_11", "deryyi_11", "derzz_00", "derzzp_11", "derzzi_11",
                "derz_2dsim", "derzz_2dsim"]
DER_POINTERS = ["derxp", "derxxp", "derxi", "derxxi", "deryp", "deryyp", "deryi", "deryyi",
                "derzp", "derzzp", "derzi", "derzzi"]
for _n in DER_ROUTINES + DER_POINTERS:
    globals()[_n] = _make_der(_n)


def divergence(fx, fy, fz, dx, dy, dz, odd=1):
    """diffoper.divergence, src/differential_operators.f90:7"""
    out = _like(fx)
    nx, ny, nz = fx.shape
    check(lib().o3d_divergence(_p(out), _p(fx), _p(fy), _p(fz), C.c_double(dx), C.c_double(dy),
                               C.c_double(dz), nx, ny, nz, odd))
    return out


def rotational(ux, uy, uz, dx, dy, dz):
    """diffoper.rotational, src/differential_operators.f90:40"""
    r = [_like(ux) for _ in range(3)]
    nx, ny, nz = ux.shape
    check(lib().o3d_rotational(_p(r[0]), _p(r[1]), _p(r[2]), _p(ux), _p(uy), _p(uz),
                               C.c_double(dx), C.c_double(dy), C.c_double(dz), nx, ny, nz))
    return r


def calculate_Q_criterion(ux, uy, uz, dx, dy, dz):
    """diffoper.calculate_Q_criterion, src/differential_operators.f90:79"""
    q = _like(ux)
    nx, ny, nz = ux.shape
    check(lib().o3d_calculate_q_criterion(_p(q), _p(ux), _p(uy), _p(uz), C.c_double(dx),
                                          C.c_double(dy), C.c_double(dz), nx, ny, nz))
    return q


def calculate_nu_t(ux, uy, uz, dx, dy, dz, cs, delta, want_stats=False):
    """les_turbulence.calculate_nu_t, src/les_turbulence.f90:10"""
    out = _like(ux)
    nx, ny, nz = ux.shape
    st = (C.c_double * 6)()
    check(lib().o3d_calculate_nu_t(_p(out), _p(ux), _p(uy), _p(uz), C.c_double(dx),
                                   C.c_double(dy), C.c_double(dz), C.c_double(cs),
                                   C.c_double(delta), nx, ny, nz, st if want_stats else None))
    return (out, list(st)) if want_stats else out


def ab_coefficients(dt):
    """adt/bdt/cdt of src/initialization.f90:194-202 (same operation order)"""
    adt = [dt, 3.0 * dt / 2.0, 23.0 * dt / 12.0]
    bdt = [0.0, -1.0 * dt / 2.0, -16.0 * dt / 12.0]
    cdt = [0.0, 0.0, 5.0 * dt / 12.0]
    return adt, bdt, cdt


def predict_velocity(ux, uy, uz, fux, fuy, fuz, re, adt, bdt, cdt, itime, itscheme, dx, dy, dz,
                     iles, cs, delta):
    """integration.predict_velocity, src/integration.f90:14.
    fux/fuy/fuz: (nx,ny,nz,3) inout.  returns ux_pred, uy_pred, uz_pred, nu_t"""
    nx, ny, nz = ux.shape
    up = [_like(ux) for _ in range(3)]
    nu_t = _like(ux)
    check(lib().o3d_predict_velocity(_p(up[0]), _p(up[1]), _p(up[2]), _p(ux), _p(uy), _p(uz),
                                     _p(fux), _p(fuy), _p(fuz), C.c_double(re), _v3(adt),
                                     _v3(bdt), _v3(cdt), itime, itscheme, C.c_double(dx),
                                     C.c_double(dy), C.c_double(dz), nx, ny, nz, iles,
                                     C.c_double(cs), C.c_double(delta), _p(nu_t)))
    return up[0], up[1], up[2], nu_t


def _poisson(name, pp, rhs, dx, dy, dz, omega, eps, kmax, idyn):
    nx, ny, nz = pp.shape
    om = C.c_double(omega)
    it = C.c_int(0)
    dmax = C.c_double(0.0)
    check(getattr(lib(), name)(_p(pp), _p(rhs), C.c_double(dx), C.c_double(dy), C.c_double(dz),
                               nx, ny, nz, C.byref(om), C.c_double(eps), kmax, idyn,
                               C.byref(it), C.byref(dmax)))
    return it.value, om.value, dmax.value


def poisson_solver_0000(pp, rhs, dx, dy, dz, omega, eps, kmax, idyn):
    """poisson.poisson_solver_0000, src/poisson.f90:6. pp in place; returns (iter, omega, dmax)"""
    return _poisson("o3d_poisson_solver_0000", pp, rhs, dx, dy, dz, omega, eps, kmax, idyn)


def poisson_solver_0011(pp, rhs, dx, dy, dz, omega, eps, kmax, idyn):
    """poisson.poisson_solver_0011, src/poisson.f90:132"""
    return _poisson("o3d_poisson_solver_0011", pp, rhs, dx, dy, dz, omega, eps, kmax, idyn)


def poisson_solver_111111(pp, rhs, dx, dy, dz, omega, eps, kmax, idyn):
    """poisson.poisson_solver_111111, src/poisson.f90:257"""
    return _poisson("o3d_poisson_solver_111111", pp, rhs, dx, dy, dz, omega, eps, kmax, idyn)


def poisson_solver(pp, rhs, dx, dy, dz, omega, eps, kmax, idyn):
    """the `poisson_solver` procedure pointer, src/initialization.f90:110,283-301"""
    return _poisson("o3d_poisson_solver", pp, rhs, dx, dy, dz, omega, eps, kmax, idyn)


def solve_poisson_multigrid(phi, rhs, dx, dy, dz, nlevels, npre, npost, tol):
    """poisson_multigrid.solve_poisson_multigrid, src/poisson_multigrid.f90:10"""
    nx, ny, nz = phi.shape
    cyc = C.c_int(0)
    dmax = C.c_double(0.0)
    check(lib().o3d_solve_poisson_multigrid(_p(phi), _p(rhs), C.c_double(dx), C.c_double(dy),
                                            C.c_double(dz), nx, ny, nz, nlevels, npre, npost,
                                            C.c_double(tol), C.byref(cyc), C.byref(dmax)))
    return cyc.value, dmax.value


def correct_pression(pp, ux_pred, uy_pred, uz_pred, dx, dy, dz, dt, omega, eps, kmax, idyn,
                     multigrid=0):
    """integration.correct_pression, src/integration.f90:199. pp in place;
    returns (iter, omega, dmax)"""
    nx, ny, nz = pp.shape
    om = C.c_double(omega)
    it = C.c_int(0)
    dmax = C.c_double(0.0)
    check(lib().o3d_correct_pression(_p(pp), _p(ux_pred), _p(uy_pred), _p(uz_pred),
                                     C.c_double(dx), C.c_double(dy), C.c_double(dz), nx, ny, nz,
                                     C.c_double(dt), C.byref(om), C.c_double(eps), kmax, idyn,
                                     multigrid, C.byref(it), C.byref(dmax)))
    return it.value, om.value, dmax.value


def correct_velocity(ux_pred, uy_pred, uz_pred, pp, dt, dx, dy, dz):
    """integration.correct_velocity, src/integration.f90:257. returns (ux,uy,uz, diverged)"""
    nx, ny, nz = pp.shape
    u = [_like(pp) for _ in range(3)]
    rc = check(lib().o3d_correct_velocity(_p(u[0]), _p(u[1]), _p(u[2]), _p(ux_pred), _p(uy_pred),
                                          _p(uz_pred), _p(pp), C.c_double(dt), C.c_double(dx),
                                          C.c_double(dy), C.c_double(dz), nx, ny, nz),
               allow=(L.ERR_DIVERGED,))
    return u[0], u[1], u[2], rc == L.ERR_DIVERGED


def transeq(phi, ux, uy, uz, fphi, re, sc, adt, bdt, cdt, itime, itscheme, dx, dy, dz, iles,
            nu_t=None, src=None):
    """integration.transeq, src/integration.f90:332. phi and fphi (nx,ny,nz,3) in place."""
    nx, ny, nz = phi.shape
    check(lib().o3d_transeq(_p(phi), _p(ux), _p(uy), _p(uz), None if src is None else _p(src),
                            _p(fphi), C.c_double(re), C.c_double(sc), _v3(adt), _v3(bdt),
                            _v3(cdt), itime, itscheme, C.c_double(dx), C.c_double(dy),
                            C.c_double(dz), nx, ny, nz, iles,
                            None if nu_t is None else _p(nu_t)))


def statistics_calc(ux, uy, uz, dx, dy, dz, re, t):
    """utils.statistics_calc, src/utils.f90:243 -> the 17 stats.dat columns"""
    nx, ny, nz = ux.shape
    out = (C.c_double * 17)()
    check(lib().o3d_statistics_calc(_p(ux), _p(uy), _p(uz), nx, ny, nz, C.c_double(dx),
                                    C.c_double(dy), C.c_double(dz), C.c_double(re),
                                    C.c_double(t), out))
    return np.array(list(out))


def function_stats(f):
    """functions.function_stats, src/functions.f90:27 -> [min,max,mean,imax,jmax,kmax]"""
    nx, ny, nz = f.shape
    out = (C.c_double * 6)()
    check(lib().o3d_function_stats(_p(f), nx, ny, nz, out))
    return list(out)
