"""ctypes binding of libo3d_b200.so (the C ABI in include/o3d_b200.h).

The product path has NO CPU fallback: if the shared library is missing the import fails
loudly, and every compute entry point returns O3D_ERR_NO_DEVICE without a CUDA device.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libo3d_b200.so")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_BC, ERR_ITSCHEME, ERR_DIVERGED, ERR_COMM, \
    ERR_UNSUPPORTED, ERR_IO = range(10)
ERR_NAMES = {1: "O3D_ERR_INVALID", 2: "O3D_ERR_NO_DEVICE", 3: "O3D_ERR_CUDA", 4: "O3D_ERR_BC",
             5: "O3D_ERR_ITSCHEME", 6: "O3D_ERR_DIVERGED", 7: "O3D_ERR_COMM",
             8: "O3D_ERR_UNSUPPORTED", 9: "O3D_ERR_IO"}

PERIODIC, FREE_SLIP = 0, 1
CLOSURE_00, CLOSURE_P11, CLOSURE_I11, CLOSURE_2DSIM = 0, 1, 2, 3
SOR_RED_BLACK, SOR_LEXI_WAVEFRONT = 0, 1
RED_MIN, RED_MAX, RED_SUM, RED_ABSMAX = 0, 1, 2, 3

FIELDS = ["ux", "uy", "uz", "pp", "phi", "ux_pred", "uy_pred", "uz_pred", "nu_t", "rhs",
          "fux1", "fux2", "fux3", "fuy1", "fuy2", "fuy3", "fuz1", "fuz2", "fuz3",
          "fphi1", "fphi2", "fphi3", "divu", "scratch0", "scratch1", "scratch2", "pp2",
          "old_ux", "old_uy", "old_uz"]
FIELD_ID = {n: i for i, n in enumerate(FIELDS)}


class Config(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
                ("nbcx1", C.c_int), ("nbcxn", C.c_int), ("nbcy1", C.c_int), ("nbcyn", C.c_int),
                ("nbcz1", C.c_int), ("nbczn", C.c_int), ("sim2d", C.c_int),
                ("re", C.c_double), ("sc", C.c_double), ("cs", C.c_double), ("delta", C.c_double),
                ("dt", C.c_double),
                ("adt", C.c_double * 3), ("bdt", C.c_double * 3), ("cdt", C.c_double * 3),
                ("itscheme", C.c_int), ("iles", C.c_int), ("nscr", C.c_int),
                ("omega", C.c_double), ("eps", C.c_double),
                ("kmax", C.c_int), ("idyn", C.c_int), ("multigrid", C.c_int),
                ("sor_order", C.c_int), ("sor_check_every", C.c_int),
                ("rank", C.c_int), ("nranks", C.c_int),
                ("nccl_id", C.c_ubyte * 128),
                ("reserved", C.c_int * 8)]


class O3DError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERR_NAMES.get(code, code), msg))
        self.code = code


_lib = None


def lib():
    """Load libo3d_b200.so; raise (never fall back) if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libo3d_b200.so is not built (%s). Run `python -m osinco3d_b200.build` or "
                "__graft_entry__.build(); there is no CPU fallback." % LIB_PATH)
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.o3d_last_error.restype = C.c_char_p
        _lib.o3d_kernel_launches.restype = C.c_longlong
        _lib.o3d_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_ulonglong]
        _lib.o3d_host_free.argtypes = [C.c_void_p]
        _lib.o3d_session_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        for name in ("o3d_session_destroy", "o3d_sync"):
            getattr(_lib, name).argtypes = [C.c_void_p]
        _lib.o3d_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.o3d_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.o3d_device_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        _lib.o3d_session_slab.argtypes = [C.c_void_p, ip, ip]
        _lib.o3d_session_layout.argtypes = [C.c_void_p, C.POINTER(C.c_longlong),
                                            C.POINTER(C.c_longlong)]
        _lib.o3d_mark_modified.argtypes = [C.c_void_p, C.c_int]
        _lib.o3d_s_predict_velocity.argtypes = [C.c_void_p, C.c_int]
        _lib.o3d_s_correct_pression.argtypes = [C.c_void_p, ip, dp]
        _lib.o3d_s_correct_velocity.argtypes = [C.c_void_p]
        _lib.o3d_s_transeq.argtypes = [C.c_void_p, C.c_int]
        _lib.o3d_step.argtypes = [C.c_void_p, C.c_int, ip, dp]
        _lib.o3d_s_divergence.argtypes = [C.c_void_p] + [C.c_int] * 5
        _lib.o3d_s_reduce.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
        _lib.o3d_s_function_stats.argtypes = [C.c_void_p, C.c_int, dp]
        _lib.o3d_s_statistics.argtypes = [C.c_void_p, C.c_double, dp]
        _lib.o3d_s_rotational.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib.o3d_s_q_criterion.argtypes = [C.c_void_p, C.c_int]
        _lib.o3d_s_vorticity_magnitude.argtypes = [C.c_void_p, C.c_int]
        _lib.o3d_s_old_values.argtypes = [C.c_void_p]
        _lib.o3d_s_step_diagnostics.argtypes = [C.c_void_p, dp]
        _lib.o3d_s_calculate_residuals.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double,
                                                   dp]
        _lib.o3d_s_save_fields.argtypes = [C.c_void_p, C.c_char_p, C.c_double, dp, dp, dp]
        _lib.o3d_s_read_fields.argtypes = [C.c_void_p, C.c_char_p, dp, dp, dp, dp]
        _lib.o3d_s_write_binary.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        _lib.o3d_s_write_all_data.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        _lib.o3d_s_io_wait.argtypes = [C.c_void_p]
        _lib.o3d_s_sor_path.argtypes = [C.c_void_p, ip, ip]
        _lib.o3d_upload_planes.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        _lib.o3d_download_planes.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        _lib.o3d_get_omega.argtypes = [C.c_void_p, dp]
        _lib.o3d_set_omega.argtypes = [C.c_void_p, C.c_double]
        _lib.o3d_session_set_poisson.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int,
                                                 C.c_int]
        _lib.o3d_s_timers.argtypes = [C.c_void_p, dp, C.POINTER(C.c_longlong), C.c_int]
        _lib.o3d_s_enable_timers.argtypes = [C.c_void_p, C.c_int]
        _lib.o3d_s_stopwatch_start.argtypes = [C.c_void_p]
        _lib.o3d_s_stopwatch_stop.argtypes = [C.c_void_p, dp]
        _lib.o3d_host_register.argtypes = [C.c_void_p, C.c_ulonglong]
        _lib.o3d_host_unregister.argtypes = [C.c_void_p]
        _lib.o3d_nccl_unique_id.argtypes = [C.POINTER(C.c_ubyte)]
    return _lib


def check(rc, allow=()):
    if rc != OK and rc not in allow:
        raise O3DError(rc, lib().o3d_last_error().decode(errors="replace"))
    return rc


def last_error():
    return lib().o3d_last_error().decode(errors="replace")
