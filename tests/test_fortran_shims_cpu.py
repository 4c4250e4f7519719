"""The Fortran side of the drop-in boundary, checked without a Fortran compiler (none in the
image): (1) the replacement modules in fortran/ define every hot-path procedure of the reference
modules they replace, under the reference's module and procedure names and with the reference's
dummy-argument lists (tests/golden/reference_signatures.json, extracted from the reference source
by tests/golden/make_signatures.py), each dummy with the reference's type, kind, intent and
rank; (2) every bind(C) interface block of fortran/o3d_b200_c.f90 names a function that
include/o3d_b200.h declares, with the same number of arguments, and -- argument by argument -- what
a Fortran compiler would pass (by value / by reference, c_int / c_long_long / c_double / c_ptr,
intent(in) <-> const) is the C parameter type; (3) every call of a bind(C) function in the shim
bodies passes the declared number of arguments.  (2) and (3) use numpy.f2py's crackfortran as a
second, independent Fortran parser."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import f90np  # noqa: E402

_ALL = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_signatures.json")))
SIG = {k: v for k, v in _ALL.items() if not k.startswith("_")}
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


# ---------------------------------------------------------------------------------------------
# Typed check of the bind(C) interfaces with an independent Fortran parser (numpy.f2py's
# crackfortran): argument by argument, the C type a Fortran compiler would pass must be the
# parameter type of the prototype in include/o3d_b200.h.  A by-reference integer where C takes it
# by value, a c_int where C takes a long long, or an intent(out) array behind a `const double*`
# compiles and links silently -- this is the check a compiler + linker would NOT make.
# ---------------------------------------------------------------------------------------------
def c_typed_prototypes():
    txt = open(os.path.join(ROOT, "include", "o3d_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    ders = re.findall(r"O3D_DECL_DER\((\w+)\)", txt)
    der_ret, der_args = re.search(r"#define\s+O3D_DECL_DER\(name\)\s*\\\s*\n\s*(\w+)\s+"
                                  r"o3d_##name\(([^)]*)\)", txt).groups()
    txt = re.sub(r"^[ \t]*#(?:[^\n]*\\\n)*[^\n]*$", "", txt, flags=re.M)   # preprocessor lines
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(o3d_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt,
                         flags=re.S):
        protos[m.group(2)] = (m.group(1).strip(), m.group(3).strip())
    for nm in ders:
        if nm != "name":
            protos["o3d_" + nm] = (der_ret, der_args)
    return protos


def c_param_class(p):
    """'const double* adt' -> ('double', 1, True): base type, pointer depth, const-qualified"""
    const = bool(re.search(r"\bconst\b", p))
    q = re.sub(r"\bconst\b", "", p)
    depth = q.count("*") + (1 if "[" in q else 0)
    q = re.sub(r"\[[^\]]*\]", "", q.replace("*", " "))
    toks = [t for t in q.split() if t != "struct"]
    types = {"int", "double", "char", "void", "long", "unsigned", "size_t", "o3d_session",
             "o3d_config"}
    if len(toks) >= 2 and toks[-1] not in types:
        toks = toks[:-1]                      # the parameter name
    base = " ".join(toks)
    if base in ("long long", "unsigned long long", "size_t", "long long int"):
        base = "i64"
    return base, depth, const


def c_param_name(p):
    toks = re.sub(r"\[[^\]]*\]", "", p).replace("*", " ").split()
    return toks[-1].lower() if len(toks) >= 2 else None


def fortran_arg_class(v):
    """crackfortran variable -> (base, depth, read_only): what a bind(C) call passes for it"""
    by_value = "value" in v.get("attrspec", [])
    ts = v["typespec"]
    if ts == "integer":
        base = {"c_int": "int", "c_long_long": "i64", "c_size_t": "i64"}[v["kindselector"]["kind"]]
    elif ts == "real":
        assert v["kindselector"]["kind"] == "c_double"
        base = "double"
    elif ts == "character":
        assert v["charselector"]["kind"] == "c_char"
        base = "char"
    elif ts == "type":
        base = {"c_ptr": "ptr", "o3d_config": "o3d_config"}[v["typename"]]
    else:
        raise AssertionError("unexpected type %s" % ts)
    read_only = by_value or v.get("intent") == ["in"]
    if base == "ptr":                          # type(c_ptr): a C pointer, by value or by reference
        return "ptr", (1 if by_value else 2), read_only
    return base, (0 if by_value else 1), read_only


def test_bind_c_argument_types_match_the_c_prototypes():
    f2py = __import__("pytest").importorskip("numpy.f2py.crackfortran")
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        f2py.verbose = 0
        tree = f2py.crackfortran([os.path.join(ROOT, "fortran", "o3d_b200_c.f90")])
    mod = [b for b in tree if b["block"] == "module" and b["name"] == "o3d_b200_c"][0]
    funcs = [f for b in mod["body"] if b["block"] == "interface" for f in b["body"]]
    assert len(funcs) > 40
    protos = c_typed_prototypes()
    checked = 0
    for f in funcs:
        name = f["name"]
        assert name in protos, "%s is not declared in include/o3d_b200.h" % name
        ret, args = protos[name]
        cparams = [] if args in ("", "void") else [c_param_class(a) for a in args.split(",")]
        assert len(cparams) == len(f["args"]), (name, f["args"], args)
        # result: integer(c_int) <-> int; type(c_ptr) <-> any pointer
        rv = f["vars"][f.get("result") or name]
        rbase, rdepth, _ = c_param_class(ret + " r")
        if rv["typespec"] == "type":
            assert rv["typename"] == "c_ptr" and rdepth >= 1, (name, ret)
        else:
            assert rv["typespec"] == "integer" and rv["kindselector"]["kind"] == "c_int" and \
                (rbase, rdepth) == ("int", 0), (name, ret)
        # parameter NAMES agree too (the order of dx, dy, dz / nx, ny, nz / kmax, idyn is only
        # visible in the names); the session handle is `s` / `out` in C and `ses` here
        cnames = [] if args in ("", "void") else [c_param_name(a) for a in args.split(",")]
        for cn, a in zip(cnames, f["args"]):
            assert cn == a.lower() or (cn, a.lower()) in (("s", "ses"), ("out", "ses"),
                                                         ("out", "res")), (name, cn, a)
        for a, (cb, cd, cconst) in zip(f["args"], cparams):
            fb, fd, ro = fortran_arg_class(f["vars"][a])
            where = "%s(%s)" % (name, a)
            if fb == "ptr":
                # opaque handle / raw address: any C pointer of that depth (void*, o3d_session*,
                # double* ...; o3d_session** / void** by reference)
                assert cd == fd, (where, "pointer depth", cd, fd)
            else:
                assert (cb, cd) == (fb, fd), (where, "C has", (cb, cd), "Fortran passes", (fb, fd))
            if cd >= 1 and fb != "ptr":
                # const-correctness: what Fortran declares intent(in) is const in C and vice versa
                assert cconst == ro, (where, "const in C:", cconst, "intent(in) in Fortran:", ro)
            checked += 1
    assert checked > 250


def test_shim_call_sites_pass_the_declared_number_of_arguments():
    """every reference to a bind(C) function in the shim bodies (rc = o3d_xxx(...), call
    o3d_check(o3d_xxx(...))) has as many actual arguments as its interface declares -- the arity
    error a Fortran compiler would stop on"""
    f2py = __import__("pytest").importorskip("numpy.f2py.crackfortran")
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        f2py.verbose = 0
        tree = f2py.crackfortran([os.path.join(ROOT, "fortran", "o3d_b200_c.f90")])
    mod = [b for b in tree if b["block"] == "module" and b["name"] == "o3d_b200_c"][0]
    arity = {f["name"]: len(f["args"]) for b in mod["body"] if b["block"] == "interface"
             for f in b["body"]}
    calls = 0
    for fn in sorted(os.listdir(os.path.join(ROOT, "fortran"))):
        if not fn.endswith(".f90"):
            continue
        text = "\n".join(f90np.logical_lines(shim_text(fn)))
        if fn == "o3d_b200_c.f90":          # skip the interface block itself
            text = re.sub(r"(?is)^\s*interface\b.*?^\s*end interface", "", text, flags=re.M)
        for m in re.finditer(r"\b(o3d_[a-z0-9_]+)\s*\(", text, flags=re.I):
            name = m.group(1).lower()
            if name not in arity:
                continue
            line_start = text.rfind("\n", 0, m.start()) + 1
            if re.match(r"\s*(end\s+)?function\b", text[line_start:m.start() + 1], re.I):
                continue
            depth, i, n, any_arg = 1, m.end(), 0, False
            while depth and i < len(text):
                c = text[i]
                if c in "([":
                    depth += 1
                elif c in ")]":
                    depth -= 1
                elif c == "," and depth == 1:
                    n += 1
                elif not c.isspace():
                    any_arg = True
                i += 1
            nargs = n + 1 if any_arg else 0
            assert nargs == arity[name], (fn, name, nargs, arity[name])
            calls += 1
    assert calls > 40


def test_shim_dummy_arguments_have_the_reference_types_intents_and_ranks():
    """drop-in at the level a Fortran compiler checks an explicit interface: for every replaced
    module procedure, each dummy of the shim has the reference's type, kind, intent and rank
    (tests/golden/reference_signatures.json "typed", extracted from /root/reference/src by
    make_signatures.py with numpy.f2py's crackfortran; the shims are parsed the same way here)"""
    __import__("pytest").importorskip("numpy.f2py.crackfortran")
    from make_signatures import typed_signatures

    def norm(sig):
        name, ts, kind, intent, rank = sig
        kind = {"8": "double", "dp": "double", "c_double": "double", "c_int": "int",
                None: "default"}.get(kind, kind)
        return (name, ts, kind, intent, rank)

    checked = 0
    for mod, ref in SIG.items():
        shim = typed_signatures(os.path.join(ROOT, "fortran", SHIM_OF[mod]))[mod]
        for proc, sig in ref["typed"].items():
            if proc in NOT_REPLACED.get(mod, ()):
                continue
            assert proc in shim, (mod, proc)
            got = [norm(s) for s in shim[proc]]
            want = [norm(s) for s in sig]
            if mod == "derivation":
                # der_type(df, f, d): the spacing dummy is named dx / dy / dz in the reference
                got = [g[1:] if i == 2 else g for i, g in enumerate(got)]
                want = [w[1:] if i == 2 else w for i, w in enumerate(want)]
            assert got == want, (mod, proc, got, want)
            checked += len(want)
    # save_fields (src/IOfunctions.f90) and write_all_data (src/visualization.f90) -> output_b200
    shim = typed_signatures(os.path.join(ROOT, "fortran", "output_b200.f90"))["output_b200"]
    for proc, sig in _ALL["_output_b200"]["typed"].items():
        assert [norm(s) for s in shim[proc]] == [norm(s) for s in sig], proc
        checked += len(sig)
    assert checked > 170


def test_shims_only_use_names_the_reference_modules_export():
    """the shims keep using two unchanged reference modules (`use initialization`, `use
    IOfunctions`): every name imported with `only:` exists there, and every `call` in a shim body
    resolves to a procedure of the shim itself, of o3d_b200_c, or of a used reference module
    (tests/golden/reference_signatures.json "_exports", from /root/reference/src)"""
    f2py = __import__("pytest").importorskip("numpy.f2py.crackfortran")
    import contextlib
    import io
    exports = _ALL["_exports"]
    parsed = {}
    for fn in sorted(os.listdir(os.path.join(ROOT, "fortran"))):
        if fn.endswith(".f90"):
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                f2py.verbose = 0
                tree = f2py.crackfortran([os.path.join(ROOT, "fortran", fn)])
            parsed[fn] = [b for b in tree if b["block"] == "module"][0]
    shim_procs = {}
    for fn, mod in parsed.items():
        names = set()
        for b in mod["body"]:
            if b["block"] in ("subroutine", "function"):
                names.add(b["name"].lower())
            elif b["block"] == "interface":
                names |= {f["name"].lower() for f in b["body"]}
        shim_procs[mod["name"].lower()] = names
    n_only = n_calls = 0
    for fn, mod in parsed.items():
        uses = {k.lower(): v for k, v in (mod.get("use") or {}).items()}
        for used, spec in uses.items():
            if used in exports and spec.get("only"):
                known = set(exports[used]["vars"]) | set(exports[used]["procedures"])
                for local, remote in spec["map"].items():
                    assert remote.lower() in known, (fn, used, remote)
                    n_only += 1
        visible = set(shim_procs[mod["name"].lower()])
        for used in uses:
            if used in exports:
                visible |= set(exports[used]["procedures"])
            visible |= shim_procs.get(used, set())
        text = "\n".join(f90np.logical_lines(shim_text(fn)))
        for m in re.finditer(r"(?im)(?:^|[\s)])call\s+(\w+)", text):
            name = m.group(1).lower()
            if name in ("c_f_pointer",):          # iso_c_binding intrinsic
                continue
            assert name in visible, (fn, "call", name)
            n_calls += 1
    assert n_only >= 9 and n_calls > 60


_F_KEYWORDS = set("""if then else elseif endif end do enddo while call return stop print subroutine
function module contains use only implicit none integer real logical character type kind len intent
in out inout value save parameter dimension pointer target allocatable allocate deallocate result
bind c name interface import select case default exit cycle not and or eq ne lt gt le ge true false
procedure public private optional write read open close unit file status form access iostat
fmt""".split())
_F_INTRINSICS = set("""size int real dble abs max min maxval minval trim adjustl len_trim c_loc
c_associated c_f_pointer c_null_char c_null_ptr c_ptr c_int c_double c_long_long c_char c_sizeof
present shape reshape sum sqrt mod merge any all huge tiny epsilon associated null""".split())


def test_shim_bodies_reference_only_declared_names():
    """`implicit none` without a compiler: every identifier in the executable part of every shim
    procedure is a dummy, a local, an entity of the enclosing module, imported from a used module
    (the other shims, o3d_b200_c, or the reference's `initialization` / `IOfunctions` exports), a
    keyword or an intrinsic.  (Found: a bind(C) function called without an interface.)"""
    f2py = __import__("pytest").importorskip("numpy.f2py.crackfortran")
    import contextlib
    import io
    exports = _ALL["_exports"]
    mods = {}
    for fn in sorted(os.listdir(os.path.join(ROOT, "fortran"))):
        if fn.endswith(".f90"):
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                f2py.verbose = 0
                tree = f2py.crackfortran([os.path.join(ROOT, "fortran", fn)])
            mods[fn] = [b for b in tree if b["block"] == "module"][0]

    def module_names(m):
        names = set(v.lower() for v in m.get("vars", {}))
        for b in m["body"]:
            if b["block"] in ("subroutine", "function"):
                names.add(b["name"].lower())
            elif b["block"] == "interface":
                names |= {f["name"].lower() for f in b["body"]}
            elif b["block"] == "type":
                names.add(b["name"].lower())
        return names

    by_module = {m["name"].lower(): module_names(m) for m in mods.values()}

    def code(line):
        line = re.sub(r'"[^"]*"', '""', line)
        line = re.sub(r"'[^']*'", "''", line)
        return line.split("!")[0].lower()

    checked = 0
    for fn, m in mods.items():
        visible = set(by_module[m["name"].lower()])
        for used, spec in (m.get("use") or {}).items():
            u = used.lower()
            if spec.get("only"):
                visible |= set(k.lower() for k in spec["map"])
            elif u in exports:
                visible |= set(exports[u]["vars"]) | set(exports[u]["procedures"])
            elif u in by_module:
                visible |= by_module[u]
        routines = f90np.routines(shim_text(fn))
        for b in m["body"]:
            if b["block"] not in ("subroutine", "function"):
                continue
            body = routines[b["name"].lower()].body
            local = set(v.lower() for v in b.get("vars", {})) | {b["name"].lower()}
            for line in body:
                ln = code(line)
                if "::" in ln and re.match(r"\s*(integer|real|logical|character|type)\b", ln):
                    for ent in f90np.split_top(ln.split("::", 1)[1]):
                        local.add(re.match(r"\s*(\w+)", ent).group(1))
            unknown = set()
            for line in body:
                ln = re.sub(r"%\s*\w+", "", code(line))                    # type components
                ln = re.sub(r"\b\d+(\.\d*)?([de][+-]?\d+)?(_\w+)?", "", ln)  # numeric literals
                for tok in re.findall(r"[a-z_]\w*", ln):
                    if not (tok in _F_KEYWORDS or tok in _F_INTRINSICS or tok in local or
                            tok in visible):
                        unknown.add(tok)
            assert not unknown, (fn, b["name"], sorted(unknown))
            checked += 1
    assert checked >= 40


# ---------------------------------------------------------------------------------------------
# A small type checker for the calls the shims make into the bind(C) interfaces: the class
# (integer / real / character / type(c_ptr) / type(o3d_config)) and the scalar-vs-array shape of
# every actual argument that can be typed from its text (identifiers, array sections, int(...),
# literals, scalar arithmetic on named constants) against the dummy it is passed to -- what a
# compiler checks through the explicit interface, and what catches two swapped arguments.
# ---------------------------------------------------------------------------------------------
def _vclass(v):
    """crackfortran variable -> (type class, is_array)"""
    ts = v.get("typespec")
    if ts == "type":
        ts = "type(%s)" % v["typename"].lower()
    return ts, bool(v.get("dimension"))


def _decl_types(lines):
    """local declarations `type-spec [, attrs] :: a, b(3)` -> {name: (class, is_array)}"""
    out = {}
    for line in lines:
        m = re.match(r"\s*(integer|real|logical|character|type\s*\(\s*(\w+)\s*\))\b(.*?)::(.*)$",
                     line.split("!")[0], re.I)
        if not m:
            continue
        ts = m.group(1).lower()
        if ts.startswith("type"):
            ts = "type(%s)" % m.group(2).lower()
        dim = "dimension" in m.group(3).lower()
        for ent in f90np.split_top(m.group(4)):
            mm = re.match(r"\s*(\w+)\s*(\()?", ent)
            out[mm.group(1).lower()] = (ts, dim or bool(mm.group(2)))
    return out


def _actual_class(e, names):
    """(class, is_array) of an actual-argument expression, or None when not decidable here"""
    el = e.strip().lower()
    if re.fullmatch(r"[a-z_]\w*", el):
        return names.get(el)
    if re.match(r"^(int|size|len_trim|nint)\s*\(", el):
        return ("integer", False)
    if re.match(r"^(real|dble)\s*\(", el):
        return ("real", False)
    if re.match(r"^c_loc\s*\(", el) or el == "c_null_ptr":
        return ("type(c_ptr)", False)
    if re.fullmatch(r"[+-]?\d+(_\w+)?", el):
        return ("integer", False)
    if re.fullmatch(r"[+-]?\d+\.\d*([de][+-]?\d+)?(_\w+)?|[+-]?\d+[de][+-]?\d+", el):
        return ("real", False)
    if el.startswith('"') or el.startswith("'") or "//" in el:
        return ("character", False)
    m = re.match(r"^([a-z_]\w*)\s*\((.*)\)$", el)
    if m and m.group(1) in names:                 # array element / section
        cls, arr = names[m.group(1)]
        if arr:
            return (cls, ":" in m.group(2))
    if m and m.group(1) in names.get("__funcs__", {}):
        return names["__funcs__"][m.group(1)]
    m = re.match(r"^([a-z_]\w*)\s*([+\-*/])", el)
    if m and m.group(1) in names and not names[m.group(1)][1]:
        return (names[m.group(1)][0], False)      # scalar arithmetic: O3D_F_FPHI1 + 1
    return None


def test_shim_call_sites_pass_arguments_of_the_declared_type_and_shape():
    f2py = __import__("pytest").importorskip("numpy.f2py.crackfortran")
    import contextlib
    import io
    mods = {}
    for fn in sorted(os.listdir(os.path.join(ROOT, "fortran"))):
        if fn.endswith(".f90"):
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                f2py.verbose = 0
                tree = f2py.crackfortran([os.path.join(ROOT, "fortran", fn)])
            mods[fn] = [b for b in tree if b["block"] == "module"][0]
    cmod = mods["o3d_b200_c.f90"]
    ifaces = {f["name"].lower(): f for b in cmod["body"] if b["block"] == "interface"
              for f in b["body"]}
    cvars = {k.lower(): _vclass(v) for k, v in cmod.get("vars", {}).items()}

    def functions_of(m):
        return {b["name"].lower(): _vclass(b["vars"][b.get("result") or b["name"]])
                for b in m["body"] if b["block"] == "function"}

    checked = undecided = 0
    problems = []
    for fn, m in mods.items():
        modvars = {k.lower(): _vclass(v) for k, v in m.get("vars", {}).items()}
        funcs = functions_of(m)
        used = [u.lower() for u in (m.get("use") or {})]
        for other in mods.values():
            if other["name"].lower() in used:
                funcs.update(functions_of(other))
        routines = f90np.routines(shim_text(fn))
        for b in m["body"]:
            if b["block"] not in ("subroutine", "function"):
                continue
            body = routines[b["name"].lower()].body
            names = dict(cvars)
            names.update(modvars)
            names.update({k.lower(): _vclass(v) for k, v in b.get("vars", {}).items()})
            names.update(_decl_types(body))
            names["__funcs__"] = funcs
            text = "\n".join(ln.split("!")[0] for ln in body)
            for mm in re.finditer(r"\b(o3d_[a-z0-9_]+)\s*\(", text, re.I):
                nm = mm.group(1).lower()
                if nm not in ifaces:
                    continue
                end = f90np.matching_paren(text, mm.end() - 1)
                args = f90np.split_top(text[mm.end():end])
                f = ifaces[nm]
                assert len(args) == len(f["args"]), (fn, b["name"], nm)
                for a, d in zip(args, f["args"]):
                    want = _vclass(f["vars"][d])
                    got = _actual_class(a, names)
                    if got is None:
                        undecided += 1
                        continue
                    checked += 1
                    if got[0] != want[0]:
                        problems.append((fn, b["name"], nm, d, "type", a.strip(), got, want))
                    elif want[0] != "character" and got[1] != want[1]:
                        problems.append((fn, b["name"], nm, d, "shape", a.strip(), got, want))
    assert not problems, problems
    assert checked > 400 and undecided <= 5, (checked, undecided)


def test_shim_block_constructs_are_balanced():
    """if / do / select case constructs of every shim procedure open and close in order (the parse
    error a compiler reports first); one-line `if (...) statement` forms do not open a block"""
    n = 0
    for fn in sorted(os.listdir(os.path.join(ROOT, "fortran"))):
        if not fn.endswith(".f90"):
            continue
        for name, r in f90np.routines(shim_text(fn)).items():
            stack = []
            for line in r.body:
                ln = re.sub(r'"[^"]*"|\'[^\']*\'', '""', line).split("!")[0].strip().lower()
                if not ln:
                    continue
                if re.match(r"(\w+\s*:\s*)?if\s*\(.*\)\s*then$", ln):
                    stack.append("if")
                elif re.match(r"else\s*if\s*\(.*\)\s*then$", ln) or ln == "else":
                    assert stack and stack[-1] == "if", (fn, name, line)
                elif re.match(r"end\s*if$", ln):
                    assert stack and stack.pop() == "if", (fn, name, line)
                elif re.match(r"(\w+\s*:\s*)?do(\s+\w+\s*=.*|\s+while\s*\(.*\)|)$", ln):
                    stack.append("do")
                elif re.match(r"end\s*do$", ln):
                    assert stack and stack.pop() == "do", (fn, name, line)
                elif re.match(r"select\s+case\s*\(.*\)$", ln):
                    stack.append("select")
                elif re.match(r"end\s*select$", ln):
                    assert stack and stack.pop() == "select", (fn, name, line)
                elif re.match(r"(case\s*\(.*\)|case\s+default)$", ln):
                    assert stack and stack[-1] == "select", (fn, name, line)
                # every `(` of a statement closes on the same logical line
                assert ln.count("(") == ln.count(")"), (fn, name, line)
            assert not stack, (fn, name, stack)
            n += 1
    assert n >= 40


def c_enum_values():
    """every enumerator of include/o3d_b200.h with its value (implicit increments included)"""
    txt = open(os.path.join(ROOT, "include", "o3d_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    vals = {}
    for body in re.findall(r"\benum\s*\w*\s*\{(.*?)\}", txt, flags=re.S):
        nxt = 0
        for ent in body.split(","):
            ent = ent.strip()
            if not ent:
                continue
            m = re.match(r"(\w+)\s*(?:=\s*(-?\d+))?$", ent)
            assert m, ent
            if m.group(2) is not None:
                nxt = int(m.group(2))
            vals[m.group(1)] = nxt
            nxt += 1
    return vals


def test_fortran_constants_equal_the_c_enumerators():
    """integer(c_int), parameter :: O3D_* = value in o3d_b200_c.f90 == the enumerator of the same
    name in include/o3d_b200.h (field ids, reduction ops, status codes)"""
    enums = c_enum_values()
    assert enums["O3D_F_UX"] == 0 and enums["O3D_F_FUX2"] == enums["O3D_F_FUX1"] + 1
    text = "\n".join(f90np.logical_lines(shim_text("o3d_b200_c.f90")))
    n = 0
    for decl in re.findall(r"(?im)^\s*integer\s*\(\s*c_int\s*\)\s*,\s*parameter\s*::(.*)$", text):
        for ent in f90np.split_top(decl):
            name, value = [t.strip() for t in ent.split("=")]
            assert name.upper() in enums, name
            assert int(value) == enums[name.upper()], (name, value, enums[name.upper()])
            n += 1
    assert n >= 20


def test_open_session_fills_every_field_of_the_config():
    """o3d_open_session (integration_b200.f90) assigns each component of type(o3d_config): a field
    added to the struct and forgotten in the shim would reach the library uninitialised"""
    text = "\n".join(f90np.logical_lines(shim_text("o3d_b200_c.f90")))
    body = re.search(r"type,\s*bind\(C\)\s*::\s*o3d_config(.*?)end type", text, re.S | re.I).group(1)
    fields = set()
    for line in body.splitlines():
        if "::" in line:
            for ent in f90np.split_top(line.split("::", 1)[1]):
                fields.add(re.match(r"\s*(\w+)", ent).group(1).lower())
    r = f90np.routines(shim_text("integration_b200.f90"))["o3d_open_session"]
    assigned = set(m.lower() for line in r.body
                   for m in re.findall(r"\bc%(\w+)\s*=", line.split("!")[0]))
    assert assigned == fields, (sorted(fields - assigned), sorted(assigned - fields))


def test_derivation_shim_routines_forward_their_own_axis_order_and_closure():
    """der<axis>[<axis>][p|i]_<00|11|2dsim> -> o3d_der(axis, order, closure, ...) with the triple
    the NAME says (O3D_CLOSURE_00 = 0, _P11 = 1, _I11 = 2, _2DSIM = 3, include/o3d_b200.h)"""
    enums = c_enum_values()
    n = 0
    for name, r in f90np.routines(shim_text("derivation_b200.f90")).items():
        m = re.fullmatch(r"der([xyz])(\1?)([pi]?)_(00|11|2dsim)", name)
        if not m:
            assert name == "dery1d", name
            continue
        axis = "xyz".index(m.group(1))
        order = 2 if m.group(2) else 1
        closure = {"00": "O3D_CLOSURE_00", "2dsim": "O3D_CLOSURE_2DSIM"}.get(
            m.group(4), "O3D_CLOSURE_I11" if m.group(3) == "i" else "O3D_CLOSURE_P11")
        text = " ".join(r.body)
        call = re.search(r"o3d_der\(\s*(\d+)_c_int\s*,\s*(\d+)_c_int\s*,\s*(\d+)_c_int\s*,\s*df\s*,"
                         r"\s*f\s*,\s*d\s*,", text)
        assert call, name
        assert tuple(int(g) for g in call.groups()) == (axis, order, enums[closure]), name
        assert '"%s"' % name in text.lower()           # the error label names the routine
        n += 1
    assert n == 20


def test_shim_call_sites_pass_same_named_arguments_in_place():
    """what the type checker cannot see: two arguments of the same type swapped (dx <-> dy,
    nx <-> nz, kmax <-> idyn).  The bind(C) interfaces name their dummies like the reference's
    procedures do, so wherever a dummy of an interface has the name of a dummy of the calling
    shim procedure, the actual argument in that position must be that very variable (possibly
    wrapped: int(nx, c_int))."""
    f2py = __import__("pytest").importorskip("numpy.f2py.crackfortran")
    import contextlib
    import io
    mods = {}
    for fn in sorted(os.listdir(os.path.join(ROOT, "fortran"))):
        if fn.endswith(".f90"):
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                f2py.verbose = 0
                tree = f2py.crackfortran([os.path.join(ROOT, "fortran", fn)])
            mods[fn] = [b for b in tree if b["block"] == "module"][0]
    ifaces = {f["name"].lower(): f for b in mods["o3d_b200_c.f90"]["body"]
              if b["block"] == "interface" for f in b["body"]}
    n = 0
    for fn, m in mods.items():
        routines = f90np.routines(shim_text(fn))
        for b in m["body"]:
            if b["block"] not in ("subroutine", "function"):
                continue
            caller = set(a.lower() for a in b["args"])
            text = "\n".join(ln.split("!")[0] for ln in routines[b["name"].lower()].body)
            for mm in re.finditer(r"\b(o3d_[a-z0-9_]+)\s*\(", text, re.I):
                nm = mm.group(1).lower()
                if nm not in ifaces:
                    continue
                end = f90np.matching_paren(text, mm.end() - 1)
                actuals = [a.strip().lower() for a in f90np.split_top(text[mm.end():end])]
                for a, d in zip(actuals, [x.lower() for x in ifaces[nm]["args"]]):
                    if d not in caller:
                        continue
                    used = set(re.findall(r"[a-z_]\w*", a)) & caller
                    assert used == {d}, (fn, b["name"], nm, "dummy", d, "<- actual", a)
                    n += 1
    assert n > 200


def test_shims_compile_in_the_reference_makefile_order():
    """swap the six sources of src/Makefile's SOURCES for their _b200
    counterparts, put o3d_b200_c.f90 first (and output_b200.f90 behind integration).  gfortran
    needs a module's .mod before its first `use`: in that list every module a file uses --
    unchanged reference files included -- is defined by an EARLIER file
    (tests/golden/reference_signatures.json "_build": SOURCES order, defines / uses per file)."""
    build = _ALL["_build"]
    shim_for = {ref["file"].split("/")[-1]: SHIM_OF[mod] for mod, ref in SIG.items()}
    order = [("o3d_b200_c.f90", None)]
    for ent in build:
        if ent["file"] in shim_for:
            order.append((shim_for[ent["file"]], None))
        else:
            order.append((ent["file"], ent))
        if ent["file"] == "integration.f90":
            order.append(("output_b200.f90", None))
    assert len([f for f, e in order if e is None]) == 8       # all eight shim files are placed
    defined = {"iso_c_binding", "iso_fortran_env"}
    for fn, ent in order:
        if ent is None:
            text = "\n".join(f90np.logical_lines(shim_text(fn)))
            defines = [m.lower() for m in re.findall(r"(?im)^\s*module\s+(\w+)\s*$", text)]
            uses = set(m.lower() for m in re.findall(r"(?im)^\s*use\s+(\w+)", text))
        else:
            defines, uses = ent["defines"], set(ent["uses"])
        missing = uses - defined - set(defines)
        assert not missing, "%s uses %s before any earlier file defines it" % (fn, sorted(missing))
        defined |= set(defines)
    # the shims define exactly the modules of the files they replace
    for ent in build:
        if ent["file"] in shim_for:
            text = "\n".join(f90np.logical_lines(shim_text(shim_for[ent["file"]])))
            assert [m.lower() for m in re.findall(r"(?im)^\s*module\s+(\w+)\s*$", text)] == \
                ent["defines"], ent["file"]


def test_config_type_components_have_the_c_struct_types():
    """type(o3d_config) component by component against struct o3d_config of the header and the
    ctypes mirror: base type, byte size and array extent (a c_int where the struct has a double
    shifts every later field)"""
    f2py = __import__("pytest").importorskip("numpy.f2py.crackfortran")
    import contextlib
    import ctypes as C
    import io
    from osinco3d_b200 import _lib
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        f2py.verbose = 0
        tree = f2py.crackfortran([os.path.join(ROOT, "fortran", "o3d_b200_c.f90")])
    mod = [b for b in tree if b["block"] == "module"][0]
    ty = [b for b in mod["body"] if b["block"] == "type" and b["name"].lower() == "o3d_config"][0]
    fort = []
    for k in ty["varnames"]:
        v = ty["vars"][k]
        kind = v["kindselector"]["kind"]
        n = int(v["dimension"][0]) if v.get("dimension") else 1
        fort.append(({"c_int": ("int", 4), "c_double": ("double", 8),
                      "c_signed_char": ("char", 1)}[kind], n))
    # the C struct
    txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "o3d_b200.h")).read(),
                 flags=re.S)
    body = re.search(r"typedef\s+struct\s+o3d_config\s*\{(.*?)\}\s*o3d_config\s*;", txt, re.S).group(1)
    cstruct = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.match(r"(unsigned\s+char|int|double)\s+(.*)$", decl, re.S)
        assert m, decl
        base = {"int": ("int", 4), "double": ("double", 8), "unsigned char": ("char", 1)}[
            re.sub(r"\s+", " ", m.group(1))]
        for ent in m.group(2).split(","):
            mm = re.match(r"\s*(\w+)\s*(?:\[(\d+)\])?\s*$", ent)
            assert mm, ent
            cstruct.append((base, int(mm.group(2) or 1)))
    assert fort == cstruct
    # the ctypes mirror
    py = []
    for _, t in _lib.Config._fields_:
        n = 1
        if hasattr(t, "_length_"):
            n, t = t._length_, t._type_
        py.append(({C.c_int: ("int", 4), C.c_double: ("double", 8), C.c_ubyte: ("char", 1)}[t], n))
    assert py == cstruct
