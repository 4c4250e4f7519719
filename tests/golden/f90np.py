"""f90np -- a small Fortran-90 -> NumPy source translator, just large enough for the hot-path
routines of jojoledemago/osinco3d (modules derivation, diffoper, les_turbulence, poisson,
integration, functions.contains_nan, initialization.schemes).

Why: the reference cannot be compiled in the build image (no Fortran compiler), so its own
outputs cannot be produced by running it.  Instead the golden-vector scripts in this directory
read the reference SOURCE, translate each routine mechanically with this module and execute the
translation on seeded inputs: the vectors come from the reference's statements, not from anyone's
reading of them.  The translation is kept operation-for-operation:

* every Fortran expression is re-emitted fully parenthesised along Fortran's own parse tree
  (`-a*b` is `-(a*b)`, `**` binds tighter than unary minus, equal-precedence operators associate
  left to right), so Python evaluates the same IEEE-754 binary64 operations in the same order as
  gfortran without FMA contraction or -ffast-math;
* `x**n` with a small integer literal n is a multiplication chain, as gfortran expands it;
* `sum(a)` accumulates sequentially in array-element order (np.cumsum), as gfortran's loop does,
  NOT numpy's pairwise sum; `maxval/minval/abs/sqrt/max/min` are exact operations;
* whole-array assignments write into the existing array (dummy arguments alias the caller's);
* scalar dummy arguments that the callee modifies are copied back after a `call`.

Unsupported constructs raise at translation time -- nothing is silently skipped except I/O
(`print`, `write`), `use`, `implicit`, `deallocate` and the calls listed in SKIP_CALLS.
"""
import re

import numpy as np

SKIP_CALLS = {"print_nu_t_statistics", "write_velocity_diverged"}


class FortranStop(Exception):
    """the translated routine executed `stop`"""


# ------------------------------------------------------------------------------------------
# runtime helpers visible to the translated code
# ------------------------------------------------------------------------------------------
def _ipow(x, n):
    assert isinstance(n, int) and 1 <= n <= 4, n
    if n == 1:
        return x
    if n == 2:
        return x * x
    if n == 3:
        return (x * x) * x
    x2 = x * x
    return x2 * x2


def _seqsum(a):
    a = np.asarray(a, dtype=np.float64)
    return np.cumsum(a.ravel(order="F"))[-1]


def _fmax(*args):
    r = args[0]
    for a in args[1:]:
        r = np.maximum(r, a)
    return r


def _fmin(*args):
    r = args[0]
    for a in args[1:]:
        r = np.minimum(r, a)
    return r


def _mod(a, b):
    assert isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)) and a >= 0 < b
    return a % b


def _size(a, d=None):
    return a.size if d is None else a.shape[d - 1]


def _elemental(mfn, nfn):
    # scalars go through libm (math.*), which is what gfortran calls for a scalar argument;
    # numpy's vectorised loops may differ from libm in the last bit
    def fn(x):
        return nfn(x) if isinstance(x, np.ndarray) and x.ndim > 0 else mfn(float(x))
    return fn


def _copyback(fn, ret, pos, old):
    """value of the callee's scalar dummy argument #pos after the call (Fortran passes by
    reference); `old` if the callee never had such a local (e.g. an untranslated stub)"""
    names = getattr(fn, "__dummies__", None)
    if ret is None or names is None or pos >= len(names) or names[pos] not in ret:
        return old
    v = ret[names[pos]]
    return old if isinstance(v, np.ndarray) and v.ndim > 0 else v


import math  # noqa: E402

RUNTIME = {"_sin": _elemental(math.sin, np.sin), "_cos": _elemental(math.cos, np.cos),
           "_tanh": _elemental(math.tanh, np.tanh), "_log": _elemental(math.log, np.log),
           "_exp": _elemental(math.exp, np.exp), "_acos": _elemental(math.acos, np.arccos),
           "np": np, "_ipow": _ipow, "_seqsum": _seqsum, "_fmax": _fmax, "_fmin": _fmin,
           "_mod": _mod, "_size": _size, "_copyback": _copyback, "FortranStop": FortranStop}

INTRINSICS = {"abs": "np.abs", "sqrt": "np.sqrt", "max": "_fmax", "min": "_fmin", "mod": "_mod",
              "sum": "_seqsum", "maxval": "np.max", "minval": "np.min", "size": "_size",
              "isnan": "np.isnan", "dble": "float", "sin": "_sin", "cos": "_cos", "tanh": "_tanh",
              "log": "_log", "exp": "_exp", "acos": "_acos"}

# ------------------------------------------------------------------------------------------
# lexer / expression parser
# ------------------------------------------------------------------------------------------
TOKEN = re.compile(r"""
    (?P<num>(\d+\.\d*|\.\d+|\d+)([dDeE][+-]?\d+)?) |
    (?P<dotop>\.(and|or|not|true|false|eq|ne|lt|le|gt|ge)\.) |
    (?P<name>[A-Za-z_]\w*) |
    (?P<str>'[^']*'|"[^"]*") |
    (?P<op>\*\*|==|/=|<=|>=|=>|[-+*/<>=(),:])
""", re.X | re.I)

DOTOPS = {".eq.": "==", ".ne.": "/=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}


def lex(s):
    out, pos = [], 0
    s = s.strip()
    while pos < len(s):
        if s[pos].isspace():
            pos += 1
            continue
        m = TOKEN.match(s, pos)
        if not m:
            raise SyntaxError("cannot tokenise %r at %r" % (s, s[pos:]))
        kind = m.lastgroup
        text = m.group(kind)
        if kind == "dotop":
            text = text.lower()
            if text in DOTOPS:
                kind, text = "op", DOTOPS[text]
        out.append((kind, text))
        pos = m.end()
    return out


class Parser:
    """Fortran expression -> fully parenthesised Python.  `arrays` = names that are arrays in the
    current scope; anything else followed by '(' is a call."""

    def __init__(self, tokens, arrays):
        self.t, self.i, self.arrays = tokens, 0, arrays

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def take(self, text=None):
        k, v = self.peek()
        if text is not None and v != text:
            raise SyntaxError("expected %r, got %r in %r" % (text, v, self.t))
        self.i += 1
        return k, v

    def expr(self):                      # .or.
        a = self.and_()
        while self.peek()[1] == ".or.":
            self.take()
            a = "(%s or %s)" % (a, self.and_())
        return a

    def and_(self):
        a = self.not_()
        while self.peek()[1] == ".and.":
            self.take()
            a = "(%s and %s)" % (a, self.not_())
        return a

    def not_(self):
        if self.peek()[1] == ".not.":
            self.take()
            return "(not %s)" % self.not_()
        return self.cmp()

    def cmp(self):
        a = self.add()
        if self.peek()[1] in ("==", "/=", "<", "<=", ">", ">="):
            op = self.take()[1]
            a = "(%s %s %s)" % (a, "!=" if op == "/=" else op, self.add())
        return a

    def add(self):
        # Fortran: a leading sign applies to the whole first TERM:  -a*b == -(a*b)
        k, v = self.peek()
        if v in ("+", "-"):
            self.take()
            a = self.mul()
            a = "(-%s)" % a if v == "-" else a
        else:
            a = self.mul()
        while self.peek()[1] in ("+", "-"):
            op = self.take()[1]
            a = "(%s %s %s)" % (a, op, self.mul())
        return a

    def mul(self):
        a = self.pow_()
        while self.peek()[1] in ("*", "/"):
            op = self.take()[1]
            a = "(%s %s %s)" % (a, op, self.pow_())
        return a

    def pow_(self):
        a = self.primary()
        if self.peek()[1] == "**":
            self.take()
            k, v = self.peek()
            # integer literal, or a real literal with an integer value (x**2.d0: gfortran folds
            # pow(x, 2.0) to x*x, which is also what a correctly rounded pow returns)
            ev = float(re.sub(r"[dD]", "e", v)) if k == "num" else None
            if ev is None or ev != int(ev) or not 1 <= int(ev) <= 4:
                raise SyntaxError("only small integer-valued literal exponents are supported: %r" % (self.t,))
            self.take()
            a = "_ipow(%s, %d)" % (a, int(ev))
        return a

    def args(self):
        """after '(' : list of (kind, text) with kind 'expr' | 'range' | 'kw'"""
        out = []
        if self.peek()[1] == ")":
            self.take()
            return out
        while True:
            lo = hi = None
            # keyword argument  name = expr
            if self.peek()[0] == "name" and self.i + 1 < len(self.t) and self.t[self.i + 1][1] == "=":
                kw = self.take()[1]
                self.take("=")
                out.append(("kw", (kw.lower(), self.expr())))
            else:
                if self.peek()[1] != ":":
                    lo = self.expr()
                if self.peek()[1] == ":":
                    self.take()
                    if self.peek()[1] not in (",", ")"):
                        hi = self.expr()
                    out.append(("range", (lo, hi)))
                else:
                    out.append(("expr", lo))
            k, v = self.take()
            if v == ")":
                return out
            if v != ",":
                raise SyntaxError("expected , or ) in %r" % (self.t,))

    def primary(self):
        k, v = self.take()
        if k == "num":
            v = re.sub(r"[dD]", "e", v)
            if re.fullmatch(r"\d+", v):
                return v
            return repr(float(v))
        if k == "str":
            return repr(v[1:-1])
        if v == ".true.":
            return "True"
        if v == ".false.":
            return "False"
        if v == "(":
            a = self.expr()
            self.take(")")
            return "(%s)" % a
        if k == "name":
            name = v.lower()
            if self.peek()[1] != "(":
                return name
            self.take("(")
            args = self.args()
            if name in self.arrays:
                return "%s[%s]" % (name, index(args))
            if name == "huge":      # huge(x): a constant of x's kind (all reals here are binary64)
                return repr(float(np.finfo(np.float64).max))
            if name == "tiny":
                return repr(float(np.finfo(np.float64).tiny))
            if name == "real":      # real(x, kind=8) -> binary64; real(x) -> default real = binary32
                pos = [a for kd, a in args if kd == "expr"]
                kinds = [a for kd, a in args if kd == "kw"] + [("kind", p2) for p2 in pos[1:]]
                if kinds:
                    assert kinds[0][1].strip("()") == "8", kinds
                    return "float(%s)" % pos[0]
                return "float(np.float32(%s))" % pos[0]
            if name == "int":
                return "int(%s)" % args[0][1]
            fn = INTRINSICS.get(name, name)
            parts = []
            for kd, a in args:
                if kd == "expr":
                    parts.append(a)
                elif kd == "kw":
                    parts.append("%s=%s" % a)
                else:
                    raise SyntaxError("array section passed to %s" % name)
            return "%s(%s)" % (fn, ", ".join(parts))
        raise SyntaxError("unexpected token %r in %r" % (v, self.t))


def index(args):
    parts = []
    for kd, a in args:
        if kd == "expr":
            parts.append("(%s) - 1" % a)
        elif kd == "range":
            lo, hi = a
            parts.append("%s:%s" % ("" if lo is None else "(%s) - 1" % lo, "" if hi is None else hi))
        else:
            raise SyntaxError("keyword in an array subscript")
    return ", ".join(parts)


def expr_py(s, arrays):
    p = Parser(lex(s), arrays)
    out = p.expr()
    if p.i != len(p.t):
        raise SyntaxError("trailing tokens in %r" % s)
    return out


# ------------------------------------------------------------------------------------------
# statements
# ------------------------------------------------------------------------------------------
def logical_lines(text):
    """comment-free statements, continuation lines joined, ';' split"""
    out, cur = [], ""
    for raw in text.splitlines():
        line, q = "", None
        for ch in raw:                       # strip a trailing comment (not inside a string)
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "!":
                break
            line += ch
        line = line.strip()
        if not line:
            continue
        if line.startswith("&"):
            line = line[1:].lstrip()
        if line.endswith("&"):
            cur += line[:-1] + " "
            continue
        full = cur + line
        cur = ""
        depth, piece, q = 0, "", None
        for ch in full:                      # split on top-level ';'
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            if ch == ";" and depth == 0 and not q:
                out.append(piece.strip())
                piece = ""
            else:
                piece += ch
        if piece.strip():
            out.append(piece.strip())
    return out


def split_top(s, sep=","):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def matching_paren(s, start):
    depth = 0
    for i in range(start, len(s)):
        if s[i] == "(":
            depth += 1
        elif s[i] == ")":
            depth -= 1
            if depth == 0:
                return i
    raise SyntaxError("unbalanced parentheses in %r" % s)


DECL = re.compile(r"^(real|integer|logical|character|double\s+precision|procedure|type)\b", re.I)


class Routine:
    def __init__(self, kind, name, dummies, result, body):
        self.kind, self.name, self.dummies, self.result, self.body = kind, name, dummies, result, body


def routines(text):
    """every subroutine / function of a source file -> {name: Routine}"""
    out = {}
    lines = logical_lines(text)
    i = 0
    while i < len(lines):
        m = re.match(r"^(?:\w+\s+)*?(subroutine|function)\s+(\w+)\s*\(([^)]*)\)\s*(?:result\s*\(\s*(\w+)\s*\))?\s*$",
                     lines[i], re.I)
        if m and not lines[i].lower().startswith("end"):
            kind, name = m.group(1).lower(), m.group(2).lower()
            dummies = [a.strip().lower() for a in m.group(3).split(",") if a.strip()]
            j = i + 1
            body = []
            while not re.match(r"^end\s+%s\b" % kind, lines[j], re.I):
                body.append(lines[j])
                j += 1
            out[name] = Routine(kind, name, dummies, (m.group(4) or name).lower(), body)
            i = j
        i += 1
    return out


def translate(r, module_arrays=()):
    """Routine -> Python source of `def name(dummies)` that returns its locals().
    module_arrays: names of module-level arrays the routine reads through `use` (globals of the
    namespace the translation is executed in)"""
    arrays, alloc, params = set(a.lower() for a in module_arrays), [], []
    for st in r.body:                         # pass 1: declarations
        if not DECL.match(st) or "::" not in st:
            continue
        attrs, ents = st.split("::", 1)
        mdim = re.search(r"dimension\s*\(", attrs, re.I)
        dims_attr = None
        if mdim:
            e = matching_paren(attrs, mdim.end() - 1)
            dims_attr = attrs[mdim.end():e]
        is_param = re.search(r"\bparameter\b", attrs, re.I)
        allocatable = re.search(r"\ballocatable\b", attrs, re.I)
        if re.match(r"^character", attrs, re.I):
            continue
        for ent in split_top(ents):
            m = re.match(r"^(\w+)\s*(\((.*)\))?\s*(=\s*(.+))?$", ent)
            if not m:
                raise SyntaxError("declaration entity %r in %s" % (ent, r.name))
            nm = m.group(1).lower()
            dims = m.group(3) if m.group(2) else dims_attr
            if is_param and m.group(5):
                params.append((nm, m.group(5)))
            if dims is not None:
                arrays.add(nm)
                if nm not in r.dummies and not allocatable and ":" not in dims:
                    alloc.append((nm, dims))
    out = ["def %s(%s):" % (r.name, ", ".join(r.dummies))]
    ptr_targets = [re.match(r"^(\w+)\s*=>", st).group(1).lower() for st in r.body
                   if re.match(r"^\w+\s*=>", st)]
    if ptr_targets:
        out.append("    global " + ", ".join(sorted(set(ptr_targets))))
    for nm, dims in alloc:
        out.append("    %s = np.full((%s,), np.nan, order='F')"
                   % (nm, ", ".join(expr_py(d, arrays) for d in split_top(dims))))
    for nm, val in params:
        out.append("    %s = %s" % (nm, expr_py(val, arrays)))
    ind = [1]
    sel = []            # stack of select-case subjects: [expr, first_case_seen]
    loops = []          # stack of (loop variable, upper bound)

    def emit(s):
        out.append("    " * ind[0] + s)

    def simple(st):
        """one non-block statement"""
        low = st.lower()
        if re.match(r"^(print\b|write\s*\(|use\b|implicit\b|deallocate\b|intent\b|external\b)", low):
            emit("pass")
            return
        if low == "return":
            emit("return locals()")
            return
        if low == "exit":
            emit("break")
            return
        if low == "cycle":
            emit("continue")
            return
        if low == "stop" or low.startswith("stop "):
            emit("raise FortranStop(%r)" % r.name)
            return
        m = re.match(r"^allocate\s*\((.*)\)$", st, re.I)
        if m:
            for ent in split_top(m.group(1)):
                mm = re.match(r"^(\w+)\s*\((.*)\)$", ent)
                emit("%s = np.full((%s,), np.nan, order='F')"
                     % (mm.group(1).lower(), ", ".join(expr_py(d, arrays) for d in split_top(mm.group(2)))))
            return
        m = re.match(r"^call\s+(\w+)\s*(\((.*)\))?$", st, re.I)
        if m:
            fn = m.group(1).lower()
            if fn in SKIP_CALLS:
                emit("pass")
                return
            actual = split_top(m.group(3) or "")
            emit("_r = %s(%s)" % (fn, ", ".join(expr_py(a, arrays) for a in actual)))
            for pos, a in enumerate(actual):       # scalar variables passed by reference
                al = a.strip().lower()
                if re.fullmatch(r"[a-z_]\w*", al) and al not in arrays:
                    emit("%s = _copyback(%s, _r, %d, %s)" % (al, fn, pos, al))
            return
        m = re.match(r"^(\w+)\s*=>\s*(\w+)$", st)
        if m:
            emit("%s = %s" % (m.group(1).lower(), m.group(2).lower()))
            return
        # assignment: find the top-level '=' that is not part of ==, /=, <=, >=
        depth, pos, q = 0, -1, None
        for k, ch in enumerate(st):
            if q:
                if ch == q:
                    q = None
                continue
            if ch in "'\"":
                q = ch
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0:
                if st[k + 1:k + 2] == "=" or st[k - 1] in "=/<>":
                    continue
                pos = k
                break
        if pos < 0:
            raise SyntaxError("unsupported statement in %s: %r" % (r.name, st))
        lhs, rhs = st[:pos].strip(), st[pos + 1:].strip()
        rhs_py = expr_py(rhs, arrays)
        m = re.match(r"^(\w+)\s*\((.*)\)$", lhs)
        if m:
            nm = m.group(1).lower()
            if nm not in arrays:
                raise SyntaxError("assignment to a non-array reference %r in %s" % (lhs, r.name))
            p = Parser(lex("(" + m.group(2) + ")"), arrays)
            p.take("(")
            emit("%s[%s] = %s" % (nm, index(p.args()), rhs_py))
        else:
            nm = lhs.lower()
            emit("%s[...] = %s" % (nm, rhs_py) if nm in arrays else "%s = %s" % (nm, rhs_py))

    for st in r.body:                         # pass 2: executable statements
        low = st.lower()
        if DECL.match(st) and "::" in st:
            continue
        m = re.match(r"^if\s*\(", low)
        if m:
            e = matching_paren(st, st.index("("))
            cond, rest = st[st.index("(") + 1:e], st[e + 1:].strip()
            if rest.lower() == "then":
                emit("if %s:" % expr_py(cond, arrays))
                ind[0] += 1
                emit("pass")
            else:
                emit("if %s:" % expr_py(cond, arrays))
                ind[0] += 1
                simple(rest)
                ind[0] -= 1
            continue
        m = re.match(r"^else\s*if\s*\(", low)
        if m:
            e = matching_paren(st, st.index("("))
            ind[0] -= 1
            emit("elif %s:" % expr_py(st[st.index("(") + 1:e], arrays))
            ind[0] += 1
            emit("pass")
            continue
        if low == "else":
            ind[0] -= 1
            emit("else:")
            ind[0] += 1
            emit("pass")
            continue
        if re.match(r"^end\s*if$", low):
            ind[0] -= 1
            continue
        m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", st, re.I)
        if m:
            parts = split_top(m.group(2))
            lo, hi = expr_py(parts[0], arrays), expr_py(parts[1], arrays)
            step = expr_py(parts[2], arrays) if len(parts) > 2 else None
            if step is not None:
                raise SyntaxError("do loops with a stride are not supported (%s)" % r.name)
            emit("for %s in range(%s, (%s) + 1):" % (m.group(1).lower(), lo, hi))
            loops.append((m.group(1).lower(), hi))
            ind[0] += 1
            emit("pass")
            continue
        if re.match(r"^end\s*do$", low):
            # Fortran leaves the loop variable at hi + 1 when the loop runs to completion (the
            # SOR solvers report `iter` after their loop): Python's for-else does the same
            ind[0] -= 1
            var, hi = loops.pop()
            emit("else:")
            emit("    %s = (%s) + 1" % (var, hi))
            continue
        m = re.match(r"^select\s+case\s*\((.*)\)$", st, re.I)
        if m:
            sel.append([expr_py(m.group(1), arrays), False])
            continue
        m = re.match(r"^case\s*\((.*)\)$", st, re.I)
        if m:
            if sel[-1][1]:
                ind[0] -= 1
            emit("%s %s == %s:" % ("elif" if sel[-1][1] else "if", sel[-1][0],
                                   expr_py(m.group(1), arrays)))
            sel[-1][1] = True
            ind[0] += 1
            emit("pass")
            continue
        if re.match(r"^case\s+default$", low):
            if sel[-1][1]:
                ind[0] -= 1
            emit("else:" if sel[-1][1] else "if True:")
            sel[-1][1] = True
            ind[0] += 1
            emit("pass")
            continue
        if re.match(r"^end\s*select$", low):
            if sel.pop()[1]:
                ind[0] -= 1
            continue
        simple(st)
    if ind[0] != 1:
        raise SyntaxError("unbalanced blocks in %s" % r.name)
    out.append("    return locals()")
    return "\n".join(out)


def load(paths, names, namespace, module_arrays=()):
    """translate the routines `names` found in the Fortran files `paths` and define them in
    `namespace` (which must already hold RUNTIME); returns {name: python source}"""
    found = {}
    for p in paths:
        found.update(routines(open(p).read()))
    # the translated text comes from a source tree we do not control: it only ever needs these
    # builtins, so nothing else (open, exec, __import__ ...) is reachable from it
    namespace.setdefault("__builtins__", {"range": range, "float": float, "int": int,
                                          "locals": locals})
    src = {}
    for nm in names:
        r = found[nm.lower()]
        code = translate(r, module_arrays)
        exec(compile(code, "<f90np:%s>" % r.name, "exec"), namespace)
        fn = namespace[r.name]
        fn.__dummies__ = list(r.dummies)
        if r.kind == "function":
            res = r.result

            def wrap(_f=fn, _res=res):
                def call(*a):
                    return _f(*a)[_res]
                call.__dummies__ = _f.__dummies__
                return call
            namespace[r.name] = wrap()
        src[r.name] = code
    return src
