"""The Fortran side of the drop-in boundary, checked without a Fortran compiler (none in the
image): (1) the replacement modules in fortran/ define every hot-path procedure of the reference
modules they replace, under the reference's module and procedure names and with the reference's
dummy-argument lists (tests/golden/reference_signatures.json, extracted from the reference source
by tests/golden/make_signatures.py); (2) every bind(C) interface block of fortran/o3d_b200_c.f90
names a function that include/o3d_b200.h declares, with the same number of arguments."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import f90np  # noqa: E402

SIG = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_signatures.json")))
SHIM_OF = {"derivation": "derivation_b200.f90", "diffoper": "differential_operators_b200.f90",
           "les_turbulence": "les_turbulence_b200.f90", "poisson": "poisson_b200.f90",
           "poisson_multigrid": "poisson_multigrid_b200.f90", "integration": "integration_b200.f90"}
# reference procedures that are NOT on the hot path (dead code / internals of the replaced
# multigrid): the shims need not provide them
NOT_REPLACED = {"les_turbulence": {"calculate_tau_ij", "calculate_dtau_ij_dxj"},
                "poisson_multigrid": {"v_cycle", "gauss_seidel", "compute_residual",
                                      "restrict_full_weighting", "prolongation_add"}}


def shim_text(name):
    return open(os.path.join(ROOT, "fortran", name)).read()


def test_shims_keep_the_reference_module_and_procedure_interfaces():
    for mod, ref in SIG.items():
        text = shim_text(SHIM_OF[mod])
        assert re.search(r"^\s*module\s+%s\s*$" % mod, text, re.M | re.I), \
            "%s must define module %s" % (SHIM_OF[mod], mod)
        have = f90np.routines(text)
        for proc, dummies in ref["procedures"].items():
            if proc in NOT_REPLACED.get(mod, ()):
                continue
            assert proc in have, "%s lacks %s (%s)" % (SHIM_OF[mod], proc, ref["file"])
            got = have[proc].dummies
            assert len(got) == len(dummies), (proc, got, dummies)
            if mod == "derivation":
                # der_type(df, f, d): positional interface; the spacing dummy is dx / dy / dz
                assert got[:2] == dummies[:2] == ["df", "f"], (proc, got, dummies)
            else:
                assert got == dummies, (proc, got, dummies)


def c_prototypes():
    txt = open(os.path.join(ROOT, "include", "o3d_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(o3d_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    # O3D_DECL_DER(name) -> int o3d_name(double* df, const double* f, double d, int nx, int ny, int nz)
    for nm in re.findall(r"O3D_DECL_DER\((\w+)\)", txt):
        if nm != "name":
            protos["o3d_" + nm] = 6
    return protos


def test_bind_c_interfaces_match_the_c_header():
    protos = c_prototypes()
    text = "\n".join(f90np.logical_lines(shim_text("o3d_b200_c.f90")))
    found = re.findall(r"^function\s+(\w+)\s*\(([^)]*)\)\s*bind\s*\(\s*C\s*,\s*name\s*=\s*\"(\w+)\"\s*\)",
                       text, flags=re.M | re.I)
    assert len(found) > 40
    for fname, args, cname in found:
        assert fname == cname, (fname, cname)
        assert cname in protos, "%s is not declared in include/o3d_b200.h" % cname
        nargs = len([a for a in args.split(",") if a.strip()])
        assert nargs == protos[cname], (cname, nargs, protos[cname])


def test_config_type_mirrors_the_c_struct():
    """field order of type(o3d_config) == struct o3d_config == the ctypes mirror"""
    from osinco3d_b200 import _lib
    text = "\n".join(f90np.logical_lines(shim_text("o3d_b200_c.f90")))
    body = re.search(r"type,\s*bind\(C\)\s*::\s*o3d_config(.*?)end type", text, re.S | re.I).group(1)
    names = []
    for line in body.splitlines():
        if "::" in line:
            for ent in f90np.split_top(line.split("::", 1)[1]):
                names.append(re.match(r"\s*(\w+)", ent).group(1))
    assert names == [f[0] for f in _lib.Config._fields_]
