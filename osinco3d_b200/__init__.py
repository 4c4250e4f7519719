"""osinco3d_b200 -- B200-native (sm_100a CUDA) Chorin-projection time step of osinco3d.

Product code.  The compute lives in libo3d_b200.so (csrc/, C ABI in include/o3d_b200.h);
this package is only the ctypes binding plus a host-side mirror of the reference's Fortran
module interfaces.  There is no CPU fallback and nothing here touches oracle/.
"""
from . import _lib  # noqa: F401
from ._lib import (CLOSURE_00, CLOSURE_2DSIM, CLOSURE_I11, CLOSURE_P11, FREE_SLIP,  # noqa: F401
                   PERIODIC, RED_ABSMAX, RED_MAX, RED_MIN, RED_SUM, SOR_LEXI_WAVEFRONT,
                   SOR_RED_BLACK, O3DError, lib)
from . import modules  # noqa: F401
from .session import PinnedPool, Session, make_config, nccl_unique_id  # noqa: F401


def device_count():
    return lib().o3d_device_count()


def kernel_launches():
    return lib().o3d_kernel_launches()
