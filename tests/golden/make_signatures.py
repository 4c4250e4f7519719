#!/usr/bin/env python
"""Procedure names and dummy-argument lists of the reference modules that fortran/*_b200.f90
replace, read from /root/reference/src (build container only) -> reference_signatures.json.
tests/test_fortran_shims_cpu.py checks the shims against it (no Fortran compiler in the image)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90np  # noqa: E402

REF = "/root/reference/src"
MODULES = {"derivation": "derivation.f90", "diffoper": "differential_operators.f90",
           "les_turbulence": "les_turbulence.f90", "poisson": "poisson.f90",
           "poisson_multigrid": "poisson_multigrid.f90", "integration": "integration.f90"}


def typed_signatures(path):
    """{module: {procedure: [[dummy, typespec, kind, intent, rank], ...]}} of the module procedures
    of a Fortran file, from numpy.f2py's crackfortran (a second, independent parser: f90np reads
    names only).  kind is the text of the kind selector ('8', 'dp', 'c_double', None = default)."""
    import contextlib
    import io
    from numpy.f2py import crackfortran
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        crackfortran.verbose = 0
        tree = crackfortran.crackfortran([path])
    out = {}
    for m in tree:
        if m["block"] != "module":
            continue
        for p in m["body"]:
            if p["block"] not in ("subroutine", "function"):
                continue
            sig = []
            for a in p["args"]:
                v = p["vars"][a]
                ks = v.get("kindselector") or {}
                kind = ks.get("kind") or ks.get("*")
                sig.append([a, v.get("typespec"), str(kind) if kind else None,
                            (v.get("intent") or [None])[0], len(v.get("dimension") or [])])
            out.setdefault(m["name"], {})[p["name"]] = sig
    return out


def module_exports(path):
    """{module: {"vars": [...], "procedures": [...]}} (lower case) of a Fortran file"""
    import contextlib
    import io
    from numpy.f2py import crackfortran
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        crackfortran.verbose = 0
        tree = crackfortran.crackfortran([path])
    out = {}
    for m in tree:
        if m["block"] != "module":
            continue
        procs = sorted(p["name"].lower() for p in m["body"] if p["block"] in ("subroutine", "function"))
        # procedure-pointer / abstract-interface names are module entities too
        names = sorted(set(v.lower() for v in m.get("vars", {})) - set(procs))
        out[m["name"].lower()] = {"vars": names, "procedures": procs}
    return out


def build_order():
    import re
    mk = open(os.path.join(REF, "Makefile")).read()
    block = re.search(r"^SOURCES\s*=(.*?)(?:\n\s*\n|\n#)", mk, re.S | re.M).group(1)
    files = re.findall(r"\$\(SRCDIR\)/(\w+\.f90)", block)
    order = []
    for fn in files:
        text = "\n".join(f90np.logical_lines(open(os.path.join(REF, fn)).read()))
        defines = [m.lower() for m in re.findall(r"(?im)^\s*module\s+(\w+)\s*$", text)]
        uses = sorted(set(m.lower() for m in re.findall(r"(?im)^\s*use\s+(\w+)", text)))
        order.append({"file": fn, "defines": defines, "uses": uses})
    return order


def main():
    out = {}
    for mod, fn in MODULES.items():
        rs = f90np.routines(open(os.path.join(REF, fn)).read())
        out[mod] = {"file": "src/" + fn,
                    "procedures": {r.name: r.dummies for r in rs.values()},
                    "typed": typed_signatures(os.path.join(REF, fn))[mod]}
    # fortran/output_b200.f90 replaces two procedures that live in two other reference modules
    # (keys starting with "_" are not module -> shim entries)
    io = typed_signatures(os.path.join(REF, "IOfunctions.f90"))["iofunctions"]
    vis = typed_signatures(os.path.join(REF, "visualization.f90"))["visualization"]
    out["_output_b200"] = {"file": "src/IOfunctions.f90, src/visualization.f90",
                           "typed": {"save_fields": io["save_fields"],
                                     "write_all_data": vis["write_all_data"]}}
    # what the shims import from the reference modules they keep using (`use initialization`,
    # `use IOfunctions`): the names those modules export
    out["_exports"] = {}
    for fn in ("initialization.f90", "IOfunctions.f90"):
        for mod, ex in module_exports(os.path.join(REF, fn)).items():
            out["_exports"][mod] = dict(ex, file="src/" + fn)
    # compile order of the reference (SOURCES of src/Makefile: gfortran needs a module's .mod
    # before its first `use`) with, per file, the module it defines and the modules it uses
    out["_build"] = build_order()
    with open(os.path.join(HERE, "reference_signatures.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print({m: len(v.get("procedures", v.get("typed", v)) if isinstance(v, dict) else v)
           for m, v in out.items()})


if __name__ == "__main__":
    main()
