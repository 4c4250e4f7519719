"""CPU-side checks of the drop-in boundary: libo3d_b200.so builds (nvcc cross-compiles), loads
and exports every symbol include/o3d_b200.h declares; without a CUDA device every compute entry
point fails loudly with O3D_ERR_NO_DEVICE (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "o3d_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(o3d_[a-z0-9_]+)\s*\(", txt))
    # the O3D_DECL_DER(name) macro declares o3d_<name>
    names |= {"o3d_" + n for n in re.findall(r"O3D_DECL_DER\((\w+)\)", txt) if n != "name"}
    names.discard("o3d_")
    return sorted(names)


def test_header_declares_the_reference_interfaces():
    names = declared_symbols()
    # der_type routines + pointers (src/derivation.f90, src/initialization.f90:104-109)
    for n in ("derx_00", "derxp_11", "derxi_11", "deryy_00", "derzzi_11", "derz_2dsim", "derxp",
              "derzzi"):
        assert "o3d_" + n in names
    for n in ("o3d_divergence", "o3d_calculate_nu_t", "o3d_predict_velocity",
              "o3d_correct_pression", "o3d_correct_velocity", "o3d_transeq",
              "o3d_poisson_solver_0000", "o3d_poisson_solver_0011", "o3d_poisson_solver_111111",
              "o3d_solve_poisson_multigrid", "o3d_session_create", "o3d_step"):
        assert n in names
    assert len(names) > 70


def test_library_exports_every_declared_symbol(built_lib):
    lib = built_lib.lib()
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.o3d_abi_version() == 1


def test_config_struct_layout_matches_header(built_lib):
    """ctypes mirror of o3d_config must have the C layout (checked through a size probe)."""
    lib = built_lib.lib()
    assert hasattr(lib, "o3d_config_size")
    assert lib.o3d_config_size() == C.sizeof(built_lib._lib.Config)


def test_no_device_means_loud_failure(built_lib):
    if built_lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    from osinco3d_b200 import modules as M
    f = np.asfortranarray(np.zeros((8, 8, 8)))
    with pytest.raises(built_lib.O3DError) as e:
        M.derx_00(f, 0.1)
    assert e.value.code == built_lib._lib.ERR_NO_DEVICE
    with pytest.raises(built_lib.O3DError) as e:
        built_lib.Session(built_lib.make_config(16, 16, 16, 0.1, 0.1, 0.1))
    assert e.value.code == built_lib._lib.ERR_NO_DEVICE


def test_schemes_error_behaviour(built_lib):
    """schemes() stops on mixed boundary flags (src/initialization.f90:238-242); the C ABI
    returns O3D_ERR_BC instead.  No device needed."""
    from osinco3d_b200 import modules as M
    M.schemes(0, 0, 1, 1, 0, 0)
    with pytest.raises(built_lib.O3DError) as e:
        M.schemes(0, 1, 1, 1, 0, 0)
    assert e.value.code == built_lib._lib.ERR_BC
    with pytest.raises(built_lib.O3DError) as e:
        M.schemes(2, 2, 1, 1, 0, 0)
    assert e.value.code == built_lib._lib.ERR_BC
    M.schemes(1, 1, 1, 1, 1, 1)


def test_product_never_imports_oracle():
    """The product path must not route through the oracle."""
    pkg = os.path.join(ROOT, "osinco3d_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "oracle_py" not in txt and "o3d_oracle" not in txt, fn
                assert "import oracle" not in txt and "from oracle" not in txt, fn


def test_ctypes_argtypes_match_the_c_prototypes(built_lib):
    """every `argtypes` list osinco3d_b200/_lib.py declares agrees, parameter by parameter, with
    the prototype in include/o3d_b200.h (ctypes converts silently: a c_int declared where C takes
    a double would pass garbage)"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_fortran_shims_cpu import c_param_class, c_typed_prototypes
    lib = built_lib.lib()
    L = built_lib._lib
    protos = c_typed_prototypes()

    def cls(t):
        if t in (C.c_int,):
            return ("int", 0)
        if t is C.c_double:
            return ("double", 0)
        if t in (C.c_longlong, C.c_ulonglong, C.c_size_t):
            return ("i64", 0)
        if t is C.c_void_p:
            return ("ptr", 1)
        if t is C.c_char_p:
            return ("char", 1)
        if hasattr(t, "_type_"):                      # POINTER(x)
            inner = t._type_
            if inner is C.c_void_p:
                return ("ptr", 2)
            if inner is L.Config:
                return ("o3d_config", 1)
            if inner is C.c_ubyte:
                return ("unsigned char", 1)
            b, d = cls(inner)
            return (b, d + 1)
        raise AssertionError(t)

    checked = 0
    for name in protos:
        fn = getattr(lib, name, None)
        if fn is None or fn.argtypes is None:
            continue
        ret, args = protos[name]
        cparams = [] if args in ("", "void") else [c_param_class(a) for a in args.split(",")]
        assert len(cparams) == len(fn.argtypes), (name, args, fn.argtypes)
        for t, (cb, cd, _) in zip(fn.argtypes, cparams):
            pb, pd = cls(t)
            if pb == "ptr":
                assert cd == pd, (name, t, cb, cd)    # opaque handle / raw address
            else:
                assert (pb, pd) == (cb, cd), (name, t, (cb, cd))
            checked += 1
    assert checked > 120
