"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under osinco3d_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libo3d_oracle.so")

PERIODIC, EVEN, ODD, ZERO = 0, 1, 2, 3
dp = C.POINTER(C.c_double)


class Grid(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
                ("bc", C.c_int * 3), ("sim2d", C.c_int)]


_SO_O3 = os.path.join(_HERE, "_build", "libo3d_oracle_O3.so")


def build(force=False):
    src = os.path.getmtime(os.path.join(_HERE, "o3d_oracle.c"))
    if force or any(not os.path.exists(f) or os.path.getmtime(f) < src for f in (_SO, _SO_O3)):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env=dict(os.environ, CC="gcc"))
    return _SO


_libs = {False: None, True: None}
_fast = False


def use_fast_build(on):
    """bench.py's CPU-baseline legs only: route lib() to the -O3 build (the reference's own
    optimisation level) instead of the strict -O2 parity build"""
    global _fast
    _fast = bool(on)


def lib():
    if _libs[_fast] is None:
        build()
        L = C.CDLL(_SO_O3 if _fast else _SO)
        L.orc_poisson_sor.restype = C.c_int
        L.orc_poisson_solver.restype = C.c_int
        L.orc_correct_pression.restype = C.c_int
        L.orc_sim_create.restype = C.c_void_p
        L.orc_sim_field.restype = dp
        L.orc_sim_step.restype = C.c_int
        _libs[_fast] = L
    return _libs[_fast]


def _p(a):
    assert a.dtype == np.float64 and a.flags["F_CONTIGUOUS"], "need float64 Fortran-ordered"
    return a.ctypes.data_as(dp)


def farray(shape):
    return np.zeros(shape, dtype=np.float64, order="F")


def grid(nx, ny, nz, dx, dy, dz, bc=(1, 1, 1), sim2d=0):
    g = Grid()
    g.nx, g.ny, g.nz = nx, ny, nz
    g.dx, g.dy, g.dz = dx, dy, dz
    g.bc[0], g.bc[1], g.bc[2] = bc
    g.sim2d = sim2d
    return g


def der(axis, order, closure, f, d):
    df = np.empty_like(f, order="F")
    nx, ny, nz = f.shape
    lib().orc_der(axis, order, closure, _p(df), _p(f), C.c_double(d), nx, ny, nz)
    return df


def divergence(g, fx, fy, fz, odd=1):
    out = np.empty_like(fx, order="F")
    lib().orc_divergence(C.byref(g), _p(out), _p(fx), _p(fy), _p(fz), odd)
    return out


def rotational(g, ux, uy, uz):
    r = [np.empty_like(ux, order="F") for _ in range(3)]
    lib().orc_rotational(C.byref(g), _p(r[0]), _p(r[1]), _p(r[2]), _p(ux), _p(uy), _p(uz))
    return r


def q_criterion(g, ux, uy, uz):
    q = np.empty_like(ux, order="F")
    lib().orc_q_criterion(C.byref(g), _p(q), _p(ux), _p(uy), _p(uz))
    return q


def calculate_residuals(u, v, w, old_u, old_v, old_w, dt, t_ref, u_ref):
    """utils.calculate_residuals, src/utils.f90:93-160 -> 15 values (see o3d_oracle.h)"""
    out = (C.c_double * 15)()
    nx, ny, nz = u.shape
    lib().orc_calculate_residuals(_p(u), _p(v), _p(w), _p(old_u), _p(old_v), _p(old_w),
                                  C.c_double(dt), C.c_double(t_ref), C.c_double(u_ref), nx, ny, nz,
                                  out)
    return np.array(list(out))


def calculate_nu_t(g, ux, uy, uz, cs, delta):
    out = np.empty_like(ux, order="F")
    lib().orc_calculate_nu_t(C.byref(g), _p(out), _p(ux), _p(uy), _p(uz), C.c_double(cs),
                             C.c_double(delta))
    return out


def ab_coefficients(dt):
    a, b, c = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_double * 3)()
    lib().orc_ab_coefficients(C.c_double(dt), a, b, c)
    return list(a), list(b), list(c)


def predict_velocity(g, ux, uy, uz, fux, fuy, fuz, re, dt, itime, itscheme, iles, cs, delta):
    """returns (ux_pred, uy_pred, uz_pred, nu_t); fu* (nx,ny,nz,3) updated in place."""
    a, b, c = ab_coefficients(dt)
    A, B, Cc = (C.c_double * 3)(*a), (C.c_double * 3)(*b), (C.c_double * 3)(*c)
    up = [np.empty_like(ux, order="F") for _ in range(3)]
    nu_t = np.empty_like(ux, order="F")
    rc = lib().orc_predict_velocity(C.byref(g), _p(up[0]), _p(up[1]), _p(up[2]), _p(ux), _p(uy),
                                    _p(uz), _p(fux), _p(fuy), _p(fuz), C.c_double(re), A, B, Cc,
                                    itime, itscheme, iles, C.c_double(cs), C.c_double(delta),
                                    _p(nu_t))
    if rc:
        raise RuntimeError("itscheme unrecognized")
    return up[0], up[1], up[2], nu_t


def poisson_solver(g, pp, rhs, omega, eps, kmax, idyn):
    """in-place on pp; returns (iters, omega_out, dmax)"""
    om = C.c_double(omega)
    dmax = C.c_double(0.0)
    it = lib().orc_poisson_solver(C.byref(g), _p(pp), _p(rhs), C.byref(om), C.c_double(eps),
                                  kmax, idyn, C.byref(dmax))
    return it, om.value, dmax.value


def correct_pression(g, pp, up, vp, wp, dt, omega, eps, kmax, idyn):
    om = C.c_double(omega)
    dmax = C.c_double(0.0)
    rhs = np.empty_like(pp, order="F")
    it = lib().orc_correct_pression(C.byref(g), _p(pp), _p(up), _p(vp), _p(wp), C.c_double(dt),
                                    C.byref(om), C.c_double(eps), kmax, idyn, C.byref(dmax),
                                    _p(rhs))
    return it, om.value, dmax.value, rhs


def correct_velocity(g, up, vp, wp, pp, dt):
    u = [np.empty_like(up, order="F") for _ in range(3)]
    bad = lib().orc_correct_velocity(C.byref(g), _p(u[0]), _p(u[1]), _p(u[2]), _p(up), _p(vp),
                                     _p(wp), _p(pp), C.c_double(dt))
    return u[0], u[1], u[2], bad


def transeq(g, phi, ux, uy, uz, fphi, re, sc, dt, itime, itscheme, iles, nu_t):
    a, b, c = ab_coefficients(dt)
    A, B, Cc = (C.c_double * 3)(*a), (C.c_double * 3)(*b), (C.c_double * 3)(*c)
    rc = lib().orc_transeq(C.byref(g), _p(phi), _p(ux), _p(uy), _p(uz), None, _p(fphi),
                           C.c_double(re), C.c_double(sc), A, B, Cc, itime, itscheme, iles,
                           _p(nu_t))
    if rc:
        raise RuntimeError("itscheme unrecognized")


def statistics_calc(g, ux, uy, uz, re, t):
    out = (C.c_double * 17)()
    lib().orc_statistics_calc(C.byref(g), _p(ux), _p(uy), _p(uz), C.c_double(re), C.c_double(t),
                              out)
    return np.array(list(out))


def function_stats(f):
    out = (C.c_double * 6)()
    nx, ny, nz = f.shape
    lib().orc_function_stats(_p(f), nx, ny, nz, out)
    return list(out)


def init_tgv(g, nscr=0, delta=None, u0=1.0, l0=1.0, ratio=1.0, origin=(0.0, 0.0, 0.0)):
    shape = (g.nx, g.ny, g.nz)
    ux, uy, uz, pp, phi = (farray(shape) for _ in range(5))
    if delta is None:
        delta = (g.dx * g.dy * g.dz) ** (1.0 / 3.0)
    lib().orc_init_tgv(C.byref(g), C.c_double(origin[0]), C.c_double(origin[1]),
                       C.c_double(origin[2]), C.c_double(u0), C.c_double(l0), C.c_double(ratio),
                       nscr, C.c_double(delta), _p(ux), _p(uy), _p(uz), _p(pp), _p(phi))
    return ux, uy, uz, pp, phi


def _init_generic(fn, g, nscr, u0, l0, ratio, origin):
    shape = (g.nx, g.ny, g.nz)
    ux, uy, uz, pp, phi = (farray(shape) for _ in range(5))
    fn(C.byref(g), C.c_double(origin[0]), C.c_double(origin[1]), C.c_double(origin[2]),
       C.c_double(u0), C.c_double(l0), C.c_double(ratio), nscr, _p(ux), _p(uy), _p(uz), _p(pp),
       _p(phi))
    return ux, uy, uz, pp, phi


def init_mixing_layer(g, nscr=0, u0=1.0, l0=1.0, ratio=0.0, origin=(0.0, 0.0, 0.0)):
    return _init_generic(lib().orc_init_mixing_layer, g, nscr, u0, l0, ratio, origin)


def init_coplanar_jet(g, nscr=0, u0=1.0, l0=1.0, ratio=3.0, origin=(0.0, 0.0, 0.0)):
    return _init_generic(lib().orc_init_coplanar_jet, g, nscr, u0, l0, ratio, origin)


class Sim:
    """The reference main loop (hot path only), state held by the oracle."""

    FIELDS3 = ("fux", "fuy", "fuz", "fphi")

    def __init__(self, g, re, dt, itscheme=3, iles=0, cs=0.0, delta=None, nscr=0, sc=1.0,
                 omega=1.8, eps=1e-6, kmax=10000, idyn=0):
        if delta is None:
            delta = (g.dx * g.dy * g.dz) ** (1.0 / 3.0)
        self.g = g
        self.delta = delta
        self.dt = dt
        self.re = re
        self._h = C.c_void_p(lib().orc_sim_create(
            C.byref(g), C.c_double(re), C.c_double(sc), C.c_double(cs), C.c_double(delta),
            C.c_double(dt), itscheme, iles, nscr, C.c_double(omega), C.c_double(eps), kmax, idyn))
        self.itime = 0

    def field(self, name):
        ptr = lib().orc_sim_field(self._h, name.encode())
        if not ptr:
            raise KeyError(name)
        shape = (self.g.nx, self.g.ny, self.g.nz) + ((3,) if name in self.FIELDS3 else ())
        n = int(np.prod(shape))
        flat = np.ctypeslib.as_array(ptr, shape=(n,))
        return flat.reshape(shape, order="F")

    def set(self, **fields):
        for k, v in fields.items():
            self.field(k)[...] = v

    def step(self):
        self.itime += 1
        rc = lib().orc_sim_step(self._h, self.itime)
        if rc:
            raise RuntimeError("oracle step aborted rc=%d" % rc)
        return self.last_iters

    class _S(C.Structure):
        pass

    @property
    def last_iters(self):
        return self._peek()[0]

    @property
    def last_dmax(self):
        return self._peek()[1]

    @property
    def omega(self):
        return self._peek()[2]

    def _peek(self):
        # layout of orc_sim up to the report fields (see o3d_oracle.h)
        class S(C.Structure):
            _fields_ = [("g", Grid), ("re", C.c_double), ("sc", C.c_double), ("cs", C.c_double),
                        ("delta", C.c_double), ("dt", C.c_double), ("adt", C.c_double * 3),
                        ("bdt", C.c_double * 3), ("cdt", C.c_double * 3), ("itscheme", C.c_int),
                        ("iles", C.c_int), ("nscr", C.c_int), ("omega", C.c_double),
                        ("eps", C.c_double), ("kmax", C.c_int), ("idyn", C.c_int),
                        ("ptrs", C.c_void_p * 13), ("last_iters", C.c_int),
                        ("last_dmax", C.c_double), ("total_iters", C.c_long)]
        s = C.cast(self._h, C.POINTER(S)).contents
        return s.last_iters, s.last_dmax, s.omega, s.total_iters

    def stats(self, t=None):
        if t is None:
            t = self.itime * self.dt
        return statistics_calc(self.g, self.field("ux"), self.field("uy"), self.field("uz"),
                               self.re, t)

    def close(self):
        if self._h:
            lib().orc_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# deterministic initial perturbation (ici = 1), host-side input generation only
# ---------------------------------------------------------------------------------------------
def dery1d(f, dy):
    """src/derivation.f90:950-992: 6th-order centred interior, one-sided / low-order closures"""
    n = f.size
    s = 60.0 * dy
    a, b, c = 1.0 / s, 9.0 / s, 45.0 / s
    df = np.empty(n)
    j = np.arange(3, n - 3)
    df[j] = a * (f[j + 3] - f[j - 3]) - b * (f[j + 2] - f[j - 2]) + c * (f[j + 1] - f[j - 1])
    df[0] = (-f[2] + 4.0 * f[1] - 3.0 * f[0]) / (2.0 * dy)
    df[1] = (f[2] - f[0]) / (2.0 * dy)
    df[2] = (-f[4] + 8.0 * f[3] - 8.0 * f[1] + f[0]) / (12.0 * dy)
    df[n - 3] = (-f[n - 1] + 8.0 * f[n - 2] - 8.0 * f[n - 4] + f[n - 5]) / (12.0 * dy)
    df[n - 2] = (f[n - 1] - f[n - 3]) / (2.0 * dy)
    df[n - 1] = (3.0 * f[n - 1] - 4.0 * f[n - 2] + f[n - 3]) / (2.0 * dy)
    return df


def normalize1d(f, base):
    """src/utils.f90:20-45"""
    rng = f.max() - f.min()
    if abs(rng) < 1e-12:
        return np.full_like(f, base)
    return base + (1.0 - base) * (f - f.min()) / rng


def add_oscillations_init(g, ux, uy, uz, u0, noise, typesim, x0, xlx):
    """src/initial_conditions.f90:554-629 (ici = 1): the deterministic perturbation used instead of
    the shipped clock-seeded FFTW noise (ici = 2), SURVEY 8d.  noise = (x, y, z) intensities."""
    u_base = dery1d(np.ascontiguousarray(ux[0, :, 0]), g.dy)           # calcul_u_base, utils.f90:9
    u_base = normalize1d(u_base, 0.0 if typesim == 5 else -1.0)
    x = x0 + g.dx * np.arange(g.nx)
    ub = u_base[None, :, None]
    pi = np.pi
    if typesim in (5, 3):
        sx = (np.sin(8 * pi * x / xlx) + np.sin(4 * pi * x / xlx) / 8.0
              + np.sin(2 * pi * x / xlx) / 16.0)[:, None, None]
        cxx = (np.cos(8 * pi * x / xlx) + np.cos(4 * pi * x / xlx) / 8.0
               + np.cos(2 * pi * x / xlx) / 16.0)[:, None, None]
        ux = ux + u0 * noise[0] * ub * sx
        uy = uy + u0 * noise[1] * ub * cxx + 0.0 * uz
    else:
        ph = np.sin(2 * pi * 9 * x / xlx)[:, None, None]
        ux = ux + u0 * noise[0] * ub * ph
        uy = uy + u0 * noise[1] * ub * ph
        uz = uz + u0 * noise[2] * ub * ph
    f = np.asfortranarray
    return f(ux), f(uy), f(uz)
