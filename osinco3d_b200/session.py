"""Device-resident session (section B of include/o3d_b200.h): the state of the reference's
main loop (src/osinco3d_main.f90:97-128) kept in HBM across calls.
"""
import ctypes as C
import os

import numpy as np

from . import _lib as L
from ._lib import Config, O3DError, check, lib  # noqa: F401
from .modules import ab_coefficients


def make_config(nx, ny, nz, dx, dy, dz, bc=(1, 1, 1), sim2d=0, re=1600.0, sc=1.0, cs=0.0,
                delta=None, dt=1e-3, itscheme=3, iles=0, nscr=0, omega=1.8, eps=1e-6,
                kmax=10000, idyn=0, multigrid=0, sor_order=L.SOR_RED_BLACK, sor_check_every=0,
                rank=0, nranks=1, nccl_id=None):
    """Config from the reference's namelist values (src/initialization.f90:117-125).
    bc = (x, y, z) with 0 = PERIODIC, 1 = FREE_SLIP on both faces of the axis."""
    c = Config()
    c.nx, c.ny, c.nz = nx, ny, nz
    c.dx, c.dy, c.dz = dx, dy, dz
    c.nbcx1 = c.nbcxn = bc[0]
    c.nbcy1 = c.nbcyn = bc[1]
    c.nbcz1 = c.nbczn = bc[2]
    c.sim2d = sim2d
    c.re, c.sc, c.cs = re, sc, cs
    # delta = (dx*dy*dz)**(1/3), src/initialization.f90:193
    c.delta = (dx * dy * dz) ** (1.0 / 3.0) if delta is None else delta
    c.dt = dt
    adt, bdt, cdt = ab_coefficients(dt)
    for q in range(3):
        c.adt[q], c.bdt[q], c.cdt[q] = adt[q], bdt[q], cdt[q]
    c.itscheme, c.iles, c.nscr = itscheme, iles, nscr
    c.omega, c.eps, c.kmax, c.idyn, c.multigrid = omega, eps, kmax, idyn, multigrid
    c.sor_order, c.sor_check_every = sor_order, sor_check_every
    c.rank, c.nranks = rank, nranks
    if nccl_id is not None:
        for q in range(128):
            c.nccl_id[q] = nccl_id[q]
    return c


def nccl_unique_id():
    buf = (C.c_ubyte * 128)()
    check(lib().o3d_nccl_unique_id(buf))
    return bytes(buf)


class Session:
    def __init__(self, cfg):
        self.cfg = cfg
        self._h = C.c_void_p()
        check(lib().o3d_session_create(C.byref(cfg), C.byref(self._h)))
        z0, nzl = C.c_int(), C.c_int()
        check(lib().o3d_session_slab(self._h, C.byref(z0), C.byref(nzl)))
        self.z0, self.nz_local = z0.value, nzl.value
        self.shape = (cfg.nx, cfg.ny, self.nz_local)
        self.itime = 0

    # -- data movement --
    def upload(self, name, host):
        a = np.asfortranarray(host, dtype=np.float64)
        if a.shape != self.shape:
            raise ValueError("shape %s != slab shape %s" % (a.shape, self.shape))
        check(lib().o3d_upload(self._h, L.FIELD_ID[name], a.ctypes.data_as(C.c_void_p)))

    def upload_planes(self, name, host, k0):
        """local planes [k0, k0 + host.shape[2]) of a field from a (nx, ny, nk) block"""
        a = np.asfortranarray(host, dtype=np.float64)
        if a.shape[:2] != self.shape[:2]:
            raise ValueError("plane shape %s != %s" % (a.shape[:2], self.shape[:2]))
        check(lib().o3d_upload_planes(self._h, L.FIELD_ID[name], a.ctypes.data_as(C.c_void_p), k0,
                                      a.shape[2]))

    def download_planes(self, name, k0, nk):
        out = np.empty(self.shape[:2] + (nk,), dtype=np.float64, order="F")
        check(lib().o3d_download_planes(self._h, L.FIELD_ID[name], out.ctypes.data_as(C.c_void_p),
                                        k0, nk))
        return out

    def upload_ptr(self, name, ptr):
        check(lib().o3d_upload(self._h, L.FIELD_ID[name], C.c_void_p(ptr)))

    def download(self, name, out=None):
        if out is None:
            out = np.empty(self.shape, dtype=np.float64, order="F")
        check(lib().o3d_download(self._h, L.FIELD_ID[name], out.ctypes.data_as(C.c_void_p)))
        return out

    def download_ptr(self, name, ptr):
        check(lib().o3d_download(self._h, L.FIELD_ID[name], C.c_void_p(ptr)))

    def device_ptr(self, name):
        p = C.c_void_p()
        check(lib().o3d_device_ptr(self._h, L.FIELD_ID[name], C.byref(p)))
        return p.value

    def set(self, **fields):
        for k, v in fields.items():
            self.upload(k, v)

    # -- stages (src/osinco3d_main.f90:105-115) --
    def predict_velocity(self, itime):
        check(lib().o3d_s_predict_velocity(self._h, itime))

    def correct_pression(self):
        it, dmax = C.c_int(0), C.c_double(0.0)
        check(lib().o3d_s_correct_pression(self._h, C.byref(it), C.byref(dmax)))
        return it.value, dmax.value

    def correct_velocity(self):
        check(lib().o3d_s_correct_velocity(self._h))

    def transeq(self, itime):
        check(lib().o3d_s_transeq(self._h, itime))

    def step(self):
        self.itime += 1
        it, dmax = C.c_int(0), C.c_double(0.0)
        check(lib().o3d_step(self._h, self.itime, C.byref(it), C.byref(dmax)))
        self.last_iters, self.last_dmax = it.value, dmax.value
        return it.value

    def sync(self):
        check(lib().o3d_sync(self._h))

    # -- diagnostics --
    def divergence(self, fx="ux", fy="uy", fz="uz", dst="divu", odd=1):
        check(lib().o3d_s_divergence(self._h, L.FIELD_ID[fx], L.FIELD_ID[fy], L.FIELD_ID[fz],
                                     L.FIELD_ID[dst], odd))

    def reduce(self, name, op):
        out = C.c_double(0.0)
        check(lib().o3d_s_reduce(self._h, L.FIELD_ID[name], op, C.byref(out)))
        return out.value

    def function_stats(self, name):
        out = (C.c_double * 6)()
        check(lib().o3d_s_function_stats(self._h, L.FIELD_ID[name], out))
        return list(out)

    def statistics(self, t=None):
        if t is None:
            t = self.itime * self.cfg.dt
        out = (C.c_double * 17)()
        check(lib().o3d_s_statistics(self._h, C.c_double(t), out))
        return np.array(list(out))

    def rotational(self, rx="scratch0", ry="scratch1", rz="scratch2"):
        check(lib().o3d_s_rotational(self._h, L.FIELD_ID[rx], L.FIELD_ID[ry], L.FIELD_ID[rz]))

    def q_criterion(self, dst="scratch0"):
        check(lib().o3d_s_q_criterion(self._h, L.FIELD_ID[dst]))

    def step_diagnostics(self):
        """the per-step prints of src/osinco3d_main.f90:116-128 in two fused passes -> dict"""
        o = (C.c_double * 23)()
        check(lib().o3d_s_step_diagnostics(self._h, o))
        o = list(o)
        return {"divu_pred": o[0:6], "divu": o[6:12], "umin": o[12:15], "umax": o[15:18],
                "cfl": o[18:21], "phi": o[21:23]}

    def old_values(self):
        """utils.old_values, src/utils.f90:165-176"""
        check(lib().o3d_s_old_values(self._h))

    def calculate_residuals(self, dt=None, t_ref=1.0, u_ref=1.0):
        """utils.calculate_residuals, src/utils.f90:93-160 -> 15 values (include/o3d_b200.h)"""
        out = (C.c_double * 15)()
        check(lib().o3d_s_calculate_residuals(self._h, C.c_double(self.cfg.dt if dt is None else dt),
                                              C.c_double(t_ref), C.c_double(u_ref), out))
        return np.array(list(out))

    def vorticity_magnitude(self, dst="scratch0"):
        check(lib().o3d_s_vorticity_magnitude(self._h, L.FIELD_ID[dst]))

    # -- field output in the reference's binary formats (asynchronous; io_wait() drains) --
    def save_fields(self, filename, time, x, y, z):
        """IOfunctions.save_fields, src/IOfunctions.f90:360-402"""
        x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
        dp = L.dp
        check(lib().o3d_s_save_fields(self._h, os.fsencode(filename), C.c_double(time),
                                      x.ctypes.data_as(dp), y.ctypes.data_as(dp),
                                      z.ctypes.data_as(dp)))

    def read_fields(self, filename):
        """IOfunctions.read_fields, src/IOfunctions.f90:404-470 -> (time, x, y, z)"""
        c = self.cfg
        x, y, z = np.zeros(c.nx), np.zeros(c.ny), np.zeros(c.nz)
        t = C.c_double(0.0)
        dp = L.dp
        check(lib().o3d_s_read_fields(self._h, os.fsencode(filename), C.byref(t),
                                      x.ctypes.data_as(dp), y.ctypes.data_as(dp),
                                      z.ctypes.data_as(dp)))
        return t.value, x, y, z

    def write_binary(self, filename, name):
        """visualization.write_binary, src/visualization.f90:224-241"""
        check(lib().o3d_s_write_binary(self._h, os.fsencode(filename), L.FIELD_ID[name]))

    def write_all_data(self, directory, num):
        """visualization.write_all_data, src/visualization.f90:243-276"""
        check(lib().o3d_s_write_all_data(self._h, os.fsencode(directory), num))

    def io_wait(self):
        check(lib().o3d_s_io_wait(self._h))

    @property
    def omega(self):
        v = C.c_double(0.0)
        check(lib().o3d_get_omega(self._h, C.byref(v)))
        return v.value

    @omega.setter
    def omega(self, v):
        check(lib().o3d_set_omega(self._h, C.c_double(v)))

    def sor_path(self):
        """(persistent, peer) of the last red-black solve: see o3d_s_sor_path"""
        a, b = C.c_int(0), C.c_int(0)
        check(lib().o3d_s_sor_path(self._h, C.byref(a), C.byref(b)))
        return bool(a.value), bool(b.value)

    def set_poisson(self, eps, kmax, idyn=0, multigrid=0):
        check(lib().o3d_session_set_poisson(self._h, C.c_double(eps), kmax, idyn, multigrid))

    def enable_timers(self, on=True):
        check(lib().o3d_s_enable_timers(self._h, 1 if on else 0))

    def timers(self, reset=False):
        ms = (C.c_double * 6)()
        cnt = (C.c_longlong * 6)()
        check(lib().o3d_s_timers(self._h, ms, cnt, 1 if reset else 0))
        names = ["rhs", "div", "sor", "corr", "transeq", "halo"]
        return {n: (ms[i], cnt[i]) for i, n in enumerate(names)}

    def stopwatch_start(self):
        check(lib().o3d_s_stopwatch_start(self._h))

    def stopwatch_stop(self):
        ms = C.c_double(0.0)
        check(lib().o3d_s_stopwatch_stop(self._h, C.byref(ms)))
        return ms.value

    def close(self):
        if self._h:
            lib().o3d_session_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedPool:
    """Page-locked host arrays (cudaHostAlloc through o3d_host_alloc) so that the host-pointer
    module procedures and upload/download run at full PCIe speed."""

    def __init__(self):
        self._ptrs = []

    def empty(self, shape):
        n = int(np.prod(shape))
        p = C.c_void_p()
        check(lib().o3d_host_alloc(C.byref(p), C.c_ulonglong(8 * n)))
        self._ptrs.append(p)
        buf = (C.c_double * n).from_address(p.value)
        a = np.frombuffer(buf, dtype=np.float64).reshape(shape, order="F")
        return a

    def array(self, src):
        a = self.empty(src.shape)
        a[...] = src
        return a

    def close(self):
        for p in self._ptrs:
            lib().o3d_host_free(p)
        self._ptrs = []
