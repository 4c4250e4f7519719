"""Host-side mirror of the reference interfaces (osinco3d_b200/modules.py, session.py) without a
device: argument validation, the configuration derived from the reference's namelist values,
field-id table == the C enum, error text plumbing."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_arrays_must_be_fortran_ordered_float64(built_lib):
    from osinco3d_b200 import modules as M
    f = np.zeros((8, 9, 10))                                   # C order
    with pytest.raises(ValueError):
        M.derx_00(f, 0.1)
    with pytest.raises(ValueError):
        M.derx_00(np.asfortranarray(f, dtype=np.float32), 0.1)
    with pytest.raises(ValueError):
        M.divergence(np.asfortranarray(f), f, np.asfortranarray(f), 0.1, 0.1, 0.1)


def test_make_config_follows_the_reference_initialisation(built_lib):
    """delta = (dx dy dz)^(1/3) (src/initialization.f90:193), AB coefficients (:194-202), both
    faces of an axis from one flag"""
    cfg = built_lib.make_config(16, 17, 18, 0.1, 0.2, 0.3, bc=(0, 1, 0), dt=2e-3, itscheme=3,
                                iles=1, cs=0.17, nscr=1, omega=1.9, eps=1e-5, kmax=500, idyn=1)
    assert (cfg.nx, cfg.ny, cfg.nz) == (16, 17, 18)
    assert (cfg.nbcx1, cfg.nbcxn, cfg.nbcy1, cfg.nbcyn, cfg.nbcz1, cfg.nbczn) == (0, 0, 1, 1, 0, 0)
    assert cfg.delta == (0.1 * 0.2 * 0.3) ** (1.0 / 3.0)
    dt = 2e-3
    assert list(cfg.adt) == [dt, 3.0 * dt / 2.0, 23.0 * dt / 12.0]
    assert list(cfg.bdt) == [0.0, -1.0 * dt / 2.0, -16.0 * dt / 12.0]
    assert list(cfg.cdt) == [0.0, 0.0, 5.0 * dt / 12.0]
    assert (cfg.rank, cfg.nranks) == (0, 1) and cfg.multigrid == 0
    assert built_lib.make_config(8, 8, 8, 1, 1, 1, delta=0.5).delta == 0.5


def test_field_ids_match_the_c_enum(built_lib):
    txt = open(os.path.join(ROOT, "include", "o3d_b200.h")).read()
    body = re.search(r"O3D_F_UX\s*=\s*0(.*?)O3D_F_COUNT", txt, re.S).group(0)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"O3D_F_([A-Z0-9_]+)", body)
    names = [n.lower() for n in names if n != "COUNT"]
    assert names == built_lib._lib.FIELDS, (names, built_lib._lib.FIELDS)


def test_error_codes_and_messages(built_lib):
    lib = built_lib.lib()
    txt = open(os.path.join(ROOT, "include", "o3d_b200.h")).read()
    for code, name in built_lib._lib.ERR_NAMES.items():
        m = re.search(r"%s\s*=\s*(\d+)" % name, txt)
        assert m and int(m.group(1)) == code, name
    # an invalid call leaves a readable message behind (no device needed for these)
    assert lib.o3d_set_sor_order(7) == built_lib._lib.ERR_INVALID
    assert lib.o3d_slab_partition(10, 0, 0, None, None) == built_lib._lib.ERR_INVALID
    with pytest.raises(built_lib.O3DError) as e:
        from osinco3d_b200 import modules as M
        M.schemes(1, 0, 1, 1, 1, 1)
    assert "Unrecognized" in str(e.value) and e.value.code == built_lib._lib.ERR_BC
    lib.o3d_schemes(1, 1, 1, 1, 1, 1, 0)


def test_slab_ranges_cover_the_grid(built_lib):
    z0, nzl = C.c_int(), C.c_int()
    lib = built_lib.lib()
    for nz, nr in ((2041, 8), (1024, 8), (129, 4), (81, 3)):
        planes = []
        for r in range(nr):
            assert lib.o3d_slab_partition(nz, nr, r, C.byref(z0), C.byref(nzl)) == 0
            planes += list(range(z0.value, z0.value + nzl.value))
        assert planes == list(range(nz))
