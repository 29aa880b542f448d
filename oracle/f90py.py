"""f90py -- run the REFERENCE'S OWN Fortran source for the hot-path routines, mechanically translated to Python.

TEST INFRASTRUCTURE (like everything under oracle/).  The Fortran reference cannot be compiled in this image (no Fortran front
end, no MPI, no NetCDF), so the C oracle (ufm_oracle.c) is a hand-written restatement.  To pin it to the reference itself and not
only to its author's reading of it, this module translates the text of selected SUBROUTINEs / FUNCTIONs of
/root/reference/src/*.f90 -- statement by statement, no knowledge of what they compute -- into Python functions that run on the
same arrays the oracle uses (reference field names, 1-based indexing through `FArray`), with single-rank MPI semantics
(`par%master` true, `CALL sync` and in-place all-reduces are no-ops, `partition_list` returns the whole range).  Everything is
IEEE double arithmetic evaluated in source order:

* `+ - * /` on numpy float64 scalars (division by zero gives Inf/NaN as in Fortran, not an exception);
* `x**y` with a REAL exponent is libm `pow` (what gfortran emits); with an INTEGER exponent it is repeated multiplication in
  the order of GCC's `powi` expansion;
* REAL literals without a kind suffix are single precision, as in Fortran (`1E-09` is not `1E-09_dp`);
* `SUM`, `MAXVAL`, `MINVAL` walk the section in array-element order; `NORM2` is the scaled 2-norm of libgfortran.

tests/test_reference_source.py runs the translated routines against the oracle on small meshes (bit for bit) and
tests/golden/make_reference_source_golden.py stores their outputs as golden vectors, so the pin survives where /root/reference is
not mounted (the GPU box).

Supported subset (what the hot-path routines use): SUBROUTINE / FUNCTION ... RESULT, declarations with INTENT / DIMENSION /
ALLOCATABLE, DO / DO WHILE / IF / ELSEIF / ELSE / CYCLE / EXIT / RETURN, one-line IF, assignments to scalars, elements, sections and
whole arrays, CALL with scalar OUT arguments, ALLOCATE / DEALLOCATE, derived-type components, the intrinsics listed in `_INTRINSICS`.
Anything else raises `Unsupported` with the offending line -- the translator never guesses.
"""
from __future__ import annotations

import math
import re

import numpy as np


class Unsupported(Exception):
    pass


# ------------------------------------------------------------------------------------------------------------------------
# run-time support
# ------------------------------------------------------------------------------------------------------------------------
class S:
    """A Fortran section bound pair `lo:hi` (1-based, inclusive); None = the array bound."""
    __slots__ = ("lo", "hi")

    def __init__(self, lo=None, hi=None):
        self.lo, self.hi = lo, hi


class FArray:
    """1-based view of a numpy array with Fortran subscripts.  Element access returns numpy scalars (IEEE semantics)."""
    __slots__ = ("a", "lb")

    def __init__(self, a, lb=None):
        self.a = a
        self.lb = tuple(lb) if lb is not None else (1,) * a.ndim   # lower bounds (DIMENSION(2:n) etc.)

    def _key(self, k):
        if not isinstance(k, tuple):
            k = (k,)
        if len(k) != self.a.ndim:
            raise IndexError(f"rank mismatch: {len(k)} subscripts for a rank-{self.a.ndim} array")
        out = []
        for q, n, b in zip(k, self.a.shape, self.lb):
            if isinstance(q, S):
                lo = b if q.lo is None else int(q.lo)
                hi = b + n - 1 if q.hi is None else int(q.hi)
                if lo < b or hi > b + n - 1:
                    raise IndexError(f"section {lo}:{hi} outside {b}:{b + n - 1}")
                out.append(slice(lo - b, hi - b + 1))
            else:
                q = int(q)
                if q < b or q > b + n - 1:
                    raise IndexError(f"subscript {q} outside {b}:{b + n - 1} (the reference is built with -fbounds-check)")
                out.append(q - b)
        return tuple(out)

    def __getitem__(self, k):
        return self.a[self._key(k)]

    def __setitem__(self, k, v):
        self.a[self._key(k)] = v.a if isinstance(v, FArray) else v

    def fill(self, v):
        self.a[...] = v.a if isinstance(v, FArray) else v

    # whole-array expressions (`D_deformation = 2._dp * H * D_deformation`, `SUM( vals)`): behave like the numpy array;
    # numpy's elementwise + - * / are the IEEE operations
    __array_priority__ = 100

    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)

    def __add__(self, o): return self.a + _unwrap(o)
    def __radd__(self, o): return _unwrap(o) + self.a
    def __sub__(self, o): return self.a - _unwrap(o)
    def __rsub__(self, o): return _unwrap(o) - self.a
    def __mul__(self, o): return self.a * _unwrap(o)
    def __rmul__(self, o): return _unwrap(o) * self.a
    def __truediv__(self, o): return self.a / _unwrap(o)
    def __rtruediv__(self, o): return _unwrap(o) / self.a
    def __neg__(self): return -self.a


def _unwrap(x):
    return x.a if isinstance(x, FArray) else x


class NS:
    """Derived-type stand-in: attribute names are case-insensitive; numpy arrays are wrapped in FArray on first access."""

    def __init__(self, **kw):
        object.__setattr__(self, "_d", {})
        for k, v in kw.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        self._d[k.lower()] = FArray(v) if isinstance(v, np.ndarray) else v

    def __getattr__(self, k):
        try:
            return self._d[k.lower()]
        except KeyError:
            raise AttributeError(f"component '{k}' is not available in this stand-in") from None


def _sp(x):
    """a single-precision literal, promoted"""
    return np.float64(np.float32(x))


def _div(a, b):
    a, b = _unwrap(a), _unwrap(b)
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        q = abs(int(a)) // abs(int(b))
        return q if (a >= 0) == (b >= 0) else -q
    return np.float64(a) / b if not isinstance(a, np.ndarray) and not isinstance(b, np.ndarray) else a / b


def _pow(a, b):
    a, b = _unwrap(a), _unwrap(b)
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):   # elementwise, each element through the scalar path (libm pow, not numpy's SIMD pow)
        aa, bb = np.broadcast_arrays(np.asarray(a), np.asarray(b))
        out = np.empty(aa.shape, np.float64)
        for idx in np.ndindex(aa.shape):
            out[idx] = _pow(aa[idx].item() if aa.dtype.kind in "iu" else aa[idx], bb[idx].item() if bb.dtype.kind in "iu" else bb[idx])
        return out
    if isinstance(b, (int, np.integer)):
        if isinstance(a, (int, np.integer)):
            return int(a) ** int(b)
        # GCC's powi expansion (tree-ssa-math-opts.c: powi_as_mults, optimal addition chains for small n)
        n = abs(int(b))
        if n == 0:
            r = np.float64(1.0)
        elif n == 1:
            r = a
        elif n == 2:
            r = a * a
        elif n == 3:
            r = (a * a) * a
        elif n == 4:
            t = a * a
            r = t * t
        else:
            raise Unsupported(f"integer power {b}")
        return r if b >= 0 else 1.0 / r
    return np.power(a, b) if isinstance(a, np.ndarray) else np.float64(math.pow(a, b)) if _finite_pow(a, b) else np.float64(np.power(np.float64(a), np.float64(b)))


def _finite_pow(a, b):
    try:
        math.pow(a, b)
        return True
    except (ValueError, OverflowError):
        return False


def _seq(x):
    a = np.asarray(x)
    return a.ravel(order="F")


def _sum(x):
    s = None
    for v in _seq(x):
        s = v if s is None else s + v
    return (np.float64(0.0) if np.asarray(x).dtype.kind == "f" else 0) if s is None else s


def _norm2(x):
    # libgfortran's norm2_r8 (m4/norm2.m4): the scaled sum of squares gfortran's NORM2 intrinsic calls at run time -- not hypot()
    scale, ssq = np.float64(1.0), np.float64(0.0)
    for v in _seq(x):
        if v != 0.0:
            av = abs(v)
            if scale < av:
                val = scale / av
                ssq = 1.0 + (ssq * val) * val      # C: `1 + result * val * val`
                scale = av
            else:
                val = av / scale
                ssq = ssq + val * val
    return scale * np.sqrt(ssq)


def _do_range(a, b, step=1):
    a, b, step = int(a), int(b), int(step)
    return range(a, b + (1 if step > 0 else -1), step)


def _real(x, kind=None):
    """REAL( x, dp) is a double; REAL( x) without a kind is DEFAULT real, i.e. single precision: the value is rounded to 24 bits
    (it is kept in a double here, which is what every later mixed-kind operation with a _dp operand would promote it to)"""
    if isinstance(x, (np.ndarray, FArray)):
        a = np.asarray(_unwrap(x), np.float64)
        return a if kind is not None else a.astype(np.float32).astype(np.float64)
    return np.float64(x) if kind is not None else np.float64(np.float32(x))


def _sign(a, b):
    return abs(a) if b >= 0 else -abs(a)


def _minmax(elementwise, scalar, args):
    args = [_unwrap(x) for x in args]
    if any(isinstance(x, np.ndarray) for x in args):   # MAX / MIN are elemental
        r = args[0]
        for x in args[1:]:
            r = elementwise(r, x)
        return r
    return scalar(args)


def _libm(fn):
    def f(x):
        x = _unwrap(x)
        if isinstance(x, np.ndarray):
            out = np.empty(x.shape, np.float64)
            for idx in np.ndindex(x.shape):
                out[idx] = f(x[idx])
            return out
        try:
            return np.float64(fn(x))
        except (ValueError, OverflowError):   # libm returns NaN / Inf where Python raises
            return np.float64(getattr(np, fn.__name__ if fn.__name__ not in ("atan", "asin", "acos") else "arc" + fn.__name__[1:])(np.float64(x)))
    return f


def _wrap(x):
    """an actual argument as the dummy sees it: Fortran subscripts from 1 (explicit- and assumed-shape dummies do not inherit bounds)"""
    if isinstance(x, np.ndarray):
        return FArray(x)
    if isinstance(x, FArray) and any(b != 1 for b in x.lb):
        return FArray(x.a)
    return x


def _alloc(dtype, *bounds):
    """zero array with the given extents; a bound is an extent n (1:n) or a pair (lo, hi)"""
    lb = [b[0] if isinstance(b, tuple) else 1 for b in bounds]
    shape = [int(b[1]) - int(b[0]) + 1 if isinstance(b, tuple) else int(b) for b in bounds]
    return FArray(np.zeros(tuple(max(n, 0) for n in shape), dtype=dtype, order="F"), [int(q) for q in lb])


def _dgtsv(dl, d, du, b):
    """CALL DGTSV( n, 1, dl, d, du, b, ldb, info) through a real LAPACK (scipy's); arrays are overwritten as LAPACK does"""
    from scipy.linalg import lapack
    du2, d2, du3, x, info = lapack.dgtsv(np.array(_unwrap(dl)), np.array(_unwrap(d)), np.array(_unwrap(du)), np.array(_unwrap(b)))
    _unwrap(b)[...] = x
    return int(info)


def _setc(obj, name, value):
    cur = obj._d.get(name) if isinstance(obj, NS) else getattr(obj, name, None)
    if isinstance(cur, FArray) and not isinstance(value, FArray):
        cur.fill(value)
    elif isinstance(cur, FArray) and isinstance(value, FArray) and cur.a.shape == value.a.shape:
        cur.fill(value)
    else:
        setattr(obj, name, value)


_RT = {
    "_wrap": _wrap, "_setc": _setc, "_alloc": _alloc, "_dgtsv": _dgtsv, "_S": S, "_FA": FArray, "_sp": _sp, "_div": _div, "_pow": _pow, "_sum": _sum, "_norm2": _norm2, "_do": _do_range, "_np": np,
    "abs": lambda x: np.abs(_unwrap(x)) if isinstance(_unwrap(x), np.ndarray) else abs(x), "sqrt": lambda x: np.sqrt(_unwrap(x)),
    # transcendental intrinsics go through libm (what gfortran calls), element by element -- numpy's own vectorised versions may
    # differ from libm in the last bit
    "exp": _libm(math.exp), "log": _libm(math.log), "tan": _libm(math.tan), "atan": _libm(math.atan), "sin": _libm(math.sin), "cos": _libm(math.cos),
    "asin": _libm(math.asin), "acos": _libm(math.acos), "atan2": lambda a, b: np.float64(math.atan2(a, b)),
    "max": lambda *a: _minmax(np.maximum, max, a), "min": lambda *a: _minmax(np.minimum, min, a), "real": _real, "dble": _real, "int": lambda x, k=None: int(x), "nint": lambda x: int(round(float(x))),
    "sum": _sum, "maxval": lambda x: np.max(np.asarray(x)), "minval": lambda x: np.min(np.asarray(x)), "norm2": _norm2,
    "size": lambda x, d=None: np.asarray(x).size if d is None else np.asarray(x).shape[int(d) - 1],
    "mod": lambda a, b: math.fmod(a, b) if isinstance(a, (float, np.floating)) else int(math.fmod(a, b)), "sign": _sign,
    "any": lambda x: bool(np.any(np.asarray(x))), "all": lambda x: bool(np.all(np.asarray(x))), "count": lambda x: int(np.count_nonzero(np.asarray(x))),
    "erf": lambda x: np.float64(math.erf(x)), "floor": lambda x: int(math.floor(x)), "ceiling": lambda x: int(math.ceil(x)),
    "trim": lambda x: x.rstrip(), "len_trim": lambda x: len(x.rstrip()), "present": lambda x: x is not None, "huge": lambda x: np.float64(np.finfo(np.float64).max),
}
_INTRINSICS = set(_RT) - {"_wrap", "_setc", "_alloc", "_dgtsv", "_S", "_FA", "_sp", "_div", "_pow", "_sum", "_norm2", "_do", "_np"}


# ------------------------------------------------------------------------------------------------------------------------
# source handling
# ------------------------------------------------------------------------------------------------------------------------
def _strip_comment(line):
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out).rstrip()


def logical_lines(text):
    """Comment-free statements with continuation lines joined; yields (first source line number, statement)."""
    cur, start = "", None
    for n, raw in enumerate(text.split("\n"), 1):
        line = _strip_comment(raw)
        if not line.strip():
            continue
        s = line.strip()
        if cur:
            if s.startswith("&"):
                s = s[1:].lstrip()
            cur += " " + s
        else:
            cur, start = s, n
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        yield start, cur.rstrip(";").rstrip()
        cur = ""


def unit_names(text):
    """Names of all SUBROUTINEs / FUNCTIONs of a source file, in file order."""
    pat = re.compile(r"^\s*(?:(?:PURE|ELEMENTAL|RECURSIVE)\s+)*(?:SUBROUTINE|FUNCTION)\s+(\w+)", re.I)
    return [m.group(1) for _, st in logical_lines(text) for m in [pat.match(st)] if m]


def extract_unit(text, name):
    """Statements of SUBROUTINE / FUNCTION `name` (case-insensitive), inclusive of its header and END line."""
    out, inside = [], False
    pat = re.compile(rf"^\s*(?:(?:PURE|ELEMENTAL|RECURSIVE)\s+)*(SUBROUTINE|FUNCTION)\s+{re.escape(name)}\b", re.I)
    end = re.compile(rf"^\s*END\s*(SUBROUTINE|FUNCTION)\b", re.I)
    for n, st in logical_lines(text):
        if not inside and pat.match(st):
            inside = True
        if inside:
            out.append((n, st))
            if end.match(st):
                return out
    raise KeyError(f"unit {name} not found")


# ------------------------------------------------------------------------------------------------------------------------
# expressions
# ------------------------------------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"""
    (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eEdD][+-]?\d+)?(?:_\w+)?)
  | (?P<dotop>\.(?:and|or|not|true|false|eq|ne|lt|le|gt|ge|eqv|neqv)\.)
  | (?P<name>[A-Za-z_]\w*)
  | (?P<op>\*\*|==|/=|<=|>=|//|[-+*/(),:%<>=\[\]])
  | (?P<ws>\s+)
""", re.X | re.I)


def tokenize(s):
    toks, i = [], 0
    while i < len(s):
        m = _TOKEN.match(s, i)
        if not m:
            raise Unsupported(f"cannot tokenize: {s[i:i + 30]!r}")
        i = m.end()
        k = m.lastgroup
        if k != "ws":
            toks.append((k, m.group(k)))
    return toks


_DOT = {".and.": " and ", ".or.": " or ", ".not.": " not ", ".true.": " True ", ".false.": " False ", ".eq.": "==", ".ne.": "!=", ".lt.": "<",
        ".le.": "<=", ".gt.": ">", ".ge.": ">=", ".eqv.": "==", ".neqv.": "!="}


class Expr:
    """Recursive-descent translation of a Fortran expression to a Python expression string."""

    def __init__(self, toks, ctx):
        self.t, self.i, self.ctx = toks, 0, ctx

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def expect(self, v):
        k, x = self.next()
        if x != v:
            raise Unsupported(f"expected {v!r}, got {x!r}")

    # precedence: .or. < .and. < .not. < comparison < +,- < *,/ < unary < **
    def parse(self):
        e = self.p_or()
        return e

    def p_or(self):
        e = self.p_and()
        while self.peek()[1] and self.peek()[1].lower() in (".or.", ".eqv.", ".neqv."):
            op = _DOT[self.next()[1].lower()]
            e = f"({e}{op}{self.p_and()})"
        return e

    def p_and(self):
        e = self.p_not()
        while self.peek()[1] and self.peek()[1].lower() == ".and.":
            self.next()
            e = f"({e} and {self.p_not()})"
        return e

    def p_not(self):
        if self.peek()[1] and self.peek()[1].lower() == ".not.":
            self.next()
            return f"(not {self.p_not()})"
        return self.p_cmp()

    def p_cmp(self):
        e = self.p_add()
        k, x = self.peek()
        xl = x.lower() if x else None
        if xl in ("==", "/=", "<", "<=", ">", ">=", ".eq.", ".ne.", ".lt.", ".le.", ".gt.", ".ge."):
            self.next()
            op = {"/=": "!="}.get(xl, _DOT.get(xl, xl))
            e = f"({e} {op} {self.p_add()})"
        return e

    def p_add(self):
        k, x = self.peek()
        if x in ("+", "-"):
            self.next()
            e = self.p_mul()
            e = f"(-{e})" if x == "-" else e
        else:
            e = self.p_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            e = f"({e} {op} {self.p_mul()})"
        return e

    def p_mul(self):
        e = self.p_pow()
        while self.peek()[1] in ("*", "/"):
            op = self.next()[1]
            r = self.p_pow()
            e = f"({e} * {r})" if op == "*" else f"_div({e}, {r})"
        return e

    def p_pow(self):
        b = self.p_unary()
        if self.peek()[1] == "**":
            self.next()
            e = self.p_pow_rhs()
            return f"_pow({b}, {e})"
        return b

    def p_pow_rhs(self):  # right-associative; a unary minus binds to the exponent
        k, x = self.peek()
        if x in ("+", "-"):
            self.next()
            e = self.p_pow_rhs()
            return f"(-{e})" if x == "-" else e
        b = self.p_unary()
        if self.peek()[1] == "**":
            self.next()
            return f"_pow({b}, {self.p_pow_rhs()})"
        return b

    def p_unary(self):
        k, x = self.peek()
        if x in ("+", "-"):
            self.next()
            e = self.p_unary()
            return f"(-{e})" if x == "-" else e
        return self.p_primary()

    def number(self, x):
        m = re.match(r"^((?:\d+\.\d*|\.\d+|\d+)(?:[eEdD][+-]?\d+)?)(?:_(\w+))?$", x)
        lit, kind = m.group(1), m.group(2)
        is_real = bool(re.search(r"[.eEdD]", lit))
        if not is_real:
            return lit.lstrip("0") or "0"
        if re.search(r"[dD]", lit):
            return f"_np.float64({lit.lower().replace('d', 'e')})"
        if kind:
            if kind.lower() != "dp":
                raise Unsupported(f"kind suffix _{kind}")
            return f"_np.float64({lit})"
        return f"_sp({lit})"

    def subscripts(self):
        """after '(' : list of subscript / argument strings up to the matching ')' """
        args = []
        if self.peek()[1] == ")":
            self.next()
            return args
        while True:
            # a section?  [expr] : [expr]
            lo = None
            if self.peek()[1] != ":":
                lo = self.parse()
            if self.peek()[1] == ":":
                self.next()
                hi = None
                if self.peek()[1] not in (",", ")"):
                    hi = self.parse()
                args.append(f"_S({lo}, {hi})")
            else:
                args.append(lo)
            k, x = self.next()
            if x == ")":
                return args
            if x != ",":
                raise Unsupported(f"unexpected {x!r} in subscript list")

    def p_primary(self):
        k, x = self.next()
        if k == "num":
            return self.number(x)
        if k == "str":
            return repr(x[1:-1])
        if k == "dotop":
            return _DOT[x.lower()].strip()
        if x == "(":
            e = self.parse()
            self.expect(")")
            return f"({e})"
        if x == "[":   # array constructor [a, b, c]
            items = []
            while True:
                items.append(self.parse())
                k2, x2 = self.next()
                if x2 == "]":
                    break
            return f"_np.array([{', '.join(items)}])"
        if k != "name":
            raise Unsupported(f"unexpected token {x!r}")
        name = x.lower()
        chain, is_component = name, False
        while True:
            k2, x2 = self.peek()
            if x2 == "%":
                self.next()
                chain += "." + self.next()[1].lower()
                is_component = True
            elif x2 == "(":
                self.next()
                args = self.subscripts()
                if is_component or self.ctx.is_array(chain):
                    chain = f"{chain}[{', '.join(args)}]"
                elif chain in self.ctx.functions or chain in _INTRINSICS:
                    chain = f"{chain}({', '.join(args)})"
                else:
                    raise Unsupported(f"'{chain}(...)' is neither a known array nor a known function")
                is_component = False
            else:
                break
        return self.ctx.rename(chain)


def translate_expr(s, ctx):
    p = Expr(tokenize(s), ctx)
    e = p.parse()
    if p.i != len(p.t):
        raise Unsupported(f"trailing tokens in expression: {s!r}")
    return e


# ------------------------------------------------------------------------------------------------------------------------
# statements
# ------------------------------------------------------------------------------------------------------------------------
_PY_KEYWORDS = {"lambda", "in", "is", "as", "def", "del", "for", "from", "global", "if", "import", "not", "or", "and", "pass", "print", "raise", "return",
                "try", "while", "with", "yield", "class", "None", "True", "False", "max", "min", "sum", "abs", "int", "all", "any"}

# single-rank MPI semantics: these CALLs do nothing
_NOOP_CALLS = {"sync", "mpi_allreduce", "mpi_bcast", "mpi_barrier", "mpi_reduce", "deallocate_shared"}


class Ctx:
    def __init__(self, functions, subroutine_outs):
        self.functions = functions            # names callable as functions
        self.subroutine_outs = subroutine_outs  # subroutine name -> positions of scalar OUT/INOUT dummies
        self.arrays = set()
        self.locals = {}

    def is_array(self, name):
        return name in self.arrays

    def rename(self, chain):
        head = chain.split(".", 1)[0].split("[", 1)[0].split("(", 1)[0]
        if head in _PY_KEYWORDS and head not in _INTRINSICS:
            return "v_" + chain
        return chain


def _split_top(s, sep=","):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _matching_paren(s, i):
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise Unsupported(f"unbalanced parentheses: {s!r}")


class Unit:
    """One translated SUBROUTINE / FUNCTION."""

    def __init__(self, statements, functions, subroutine_outs, module_constants=()):
        self.st = statements
        self.ctx = Ctx(functions, subroutine_outs)
        self.lines = []
        self.kind = self.name = None
        self.dummies, self.result = [], None
        self.decl = {}          # name -> dict(type=, dims=, intent=, allocatable=)
        self.construct_names, self.loop_depth_of, self.open_loops = set(), {}, 0
        self.module_constants = set(module_constants)

    # ---- declarations ----
    def parse_header(self, st):
        m = re.match(r"^(?:(?:PURE|ELEMENTAL|RECURSIVE)\s+)*(SUBROUTINE|FUNCTION)\s+(\w+)\s*(?:\((.*?)\))?\s*(?:RESULT\s*\(\s*(\w+)\s*\))?\s*$", st, re.I)
        if not m:
            raise Unsupported(f"header: {st}")
        self.kind, self.name = m.group(1).lower(), m.group(2).lower()
        self.dummies = [a.strip().lower() for a in (m.group(3) or "").split(",") if a.strip()]
        self.result = (m.group(4) or (self.name if self.kind == "function" else None))
        self.result = self.result.lower() if self.result else None

    def parse_decl(self, st):
        left, right = st.split("::", 1)
        attrs = _split_top(left)
        typ = attrs[0].strip().lower()
        info = {"type": typ, "dims": None, "intent": None, "allocatable": False, "parameter": False}
        for a in attrs[1:]:
            al = a.strip().lower()
            if al.startswith("dimension"):
                info["dims"] = a[a.index("(") + 1:_matching_paren(a, a.index("("))]
            elif al.startswith("intent"):
                info["intent"] = re.sub(r"\s", "", al)[7:-1]
            elif al == "allocatable":
                info["allocatable"] = True
            elif al == "parameter":
                info["parameter"] = True
            elif al in ("pointer", "target", "save", "optional"):
                pass
            else:
                raise Unsupported(f"attribute {a!r} in {st!r}")
        for ent in _split_top(right):
            init = None
            if "=" in ent and not ent.split("=", 1)[0].strip().endswith(("<", ">", "/")):
                ent, init = [q.strip() for q in ent.split("=", 1)]
            m = re.match(r"^(\w+)\s*(?:\((.*)\))?$", ent.strip())
            if not m:
                raise Unsupported(f"entity {ent!r}")
            d = dict(info)
            if m.group(2) is not None:
                d["dims"] = m.group(2)
            d["init"] = init
            self.decl[m.group(1).lower()] = d

    # ---- emit helpers ----
    def emit(self, depth, s):
        self.lines.append("    " * depth + s)
        if s.endswith(":") and not s.startswith("def "):
            self.lines.append("    " * (depth + 1) + "pass")   # a block may hold nothing but comments

    def ex(self, s):
        return translate_expr(s, self.ctx)

    def bounds(self, dims):
        """'2:C%NZ, 3' -> '(2, c.nz), 3' for _alloc"""
        out = []
        for q in _split_top(dims):
            parts = _split_top(q, ":")
            if len(parts) == 2 and ":" in q:
                out.append(f"({self.ex(parts[0])}, {self.ex(parts[1])})")
            else:
                out.append(self.ex(q))
        return ", ".join(out)

    def outs(self):
        """scalar dummies the caller must receive back"""
        r = []
        for k, a in enumerate(self.dummies):
            d = self.decl.get(a)
            if d and d["intent"] in ("out", "inout") and d["dims"] is None and not d["type"].startswith("type"):
                r.append((k, a))
        return r

    def ret(self):
        if self.kind == "function":
            return f"return {self.ctx.rename(self.result)}"
        o = self.outs()
        return "return (" + ", ".join(self.ctx.rename(a) for _, a in o) + ("," if len(o) == 1 else "") + ")" if o else "return None"

    def assignment(self, depth, st):
        # split at the top-level '=' that is not part of ==, /=, <=, >=
        depth_p, pos = 0, None
        for i, ch in enumerate(st):
            if ch in "([":
                depth_p += 1
            elif ch in ")]":
                depth_p -= 1
            elif ch == "=" and depth_p == 0:
                if st[i + 1:i + 2] == "=" or st[i - 1] in "=/<>":
                    continue
                pos = i
                break
        if pos is None:
            raise Unsupported(f"statement: {st}")
        lhs, rhs = st[:pos].strip(), st[pos + 1:].strip()
        r = self.ex(rhs)
        l = self.ex(lhs)
        base = lhs.strip().lower()
        if re.match(r"^\w+$", base) and self.ctx.is_array(base):
            self.emit(depth, f"{l}.fill({r})")            # whole-array assignment
        elif re.match(r"^\w+(\s*%\s*\w+)+$", lhs):
            # a bare component: whole-array assignment when the component is an array, plain assignment when it is a scalar --
            # only known at run time (the derived types are not parsed)
            parent, last = l.rsplit(".", 1)
            self.emit(depth, f"_setc({parent}, {last!r}, {r})")
        else:
            self.emit(depth, f"{l} = {r}")

    whole_components: set = set()

    def call(self, depth, st):
        m = re.match(r"^CALL\s+(\w+)\s*(?:\((.*)\))?\s*$", st, re.I)
        if not m:
            raise Unsupported(st)
        name = m.group(1).lower()
        args = _split_top(m.group(2) or "")
        if name in _NOOP_CALLS:
            self.emit(depth, "pass")
            return
        if name == "partition_list":      # ( ntot, i, n, i1, i2) with one rank: 1 .. ntot
            self.emit(depth, f"{self.ex(args[3])} = 1; {self.ex(args[4])} = {self.ex(args[0])}")
            return
        m = re.match(r"^allocate_shared_(int|dp|bool)_(\d)d$", name)
        if m:   # ( n1, .., nk, pointer, window): the shared-memory window of the reference is an ordinary array here
            nd = int(m.group(2))
            dt = {"int": "_np.int32", "dp": "_np.float64", "bool": "_np.bool_"}[m.group(1)]
            tgt = self.ex(args[nd])
            if nd == 0:
                self.emit(depth, f"{tgt} = {'0' if m.group(1) == 'int' else ('False' if m.group(1) == 'bool' else '_np.float64(0.0)')}")
            else:
                dims = ", ".join(f"int({self.ex(a)})" for a in args[:nd])
                self.emit(depth, f"{tgt} = _FA(_np.zeros(({dims},), dtype={dt}, order='F'))")
            return
        if name == "dgtsv":   # ( n, nrhs, dl, d, du, b, ldb, info)
            self.emit(depth, f"{self.ex(args[7])} = _dgtsv({self.ex(args[2])}, {self.ex(args[3])}, {self.ex(args[4])}, {self.ex(args[5])})")
            return
        if name == "mpi_abort":
            self.emit(depth, "raise RuntimeError('MPI_ABORT')")
            return
        if name not in self.ctx.subroutine_outs:
            raise Unsupported(f"CALL to untranslated routine {name}")
        a = [self.ex(x) for x in args]
        outs = self.ctx.subroutine_outs[name]
        call = f"{name}({', '.join(a)})"
        if outs:
            self.emit(depth, f"({', '.join(a[k] for k in outs)}{',' if len(outs) == 1 else ''}) = {call}")
        else:
            self.emit(depth, call)

    def statement(self, depth, st):
        m0 = re.match(r"^(\w+)\s*:\s*(DO\b.*|IF\s*\(.*)$", st, re.I)   # construct name, e.g. "viscosity_iteration: DO WHILE (...)"
        if m0 and not re.match(r"^\w+\s*::", st):
            self.construct_names.add(m0.group(1).lower())
            st = m0.group(2)
            if st.upper().startswith("DO"):
                self.loop_depth_of[m0.group(1).lower()] = self.open_loops + 1
        m0 = re.match(r"^(END\s*DO|END\s*IF|EXIT|CYCLE)\s+(\w+)$", st, re.I)
        if m0 and m0.group(2).lower() in self.construct_names:
            if m0.group(1).upper() in ("EXIT", "CYCLE") and self.loop_depth_of.get(m0.group(2).lower()) != self.open_loops:
                raise Unsupported(f"{m0.group(1)} of an outer construct: {st}")
            st = m0.group(1)
        u = st.upper()
        if re.match(r"^IF\s*\(", u):
            j = _matching_paren(st, st.index("("))
            cond, rest = st[st.index("(") + 1:j], st[j + 1:].strip()
            if rest.upper() == "THEN":
                self.emit(depth, f"if {self.ex(cond)}:")
                return depth + 1
            self.emit(depth, f"if {self.ex(cond)}:")
            self.statement(depth + 1, rest)
            return depth
        if re.match(r"^ELSE\s*IF\s*\(", u):
            j = _matching_paren(st, st.index("("))
            self.emit(depth - 1, f"elif {self.ex(st[st.index('(') + 1:j])}:")
            return depth
        if u == "ELSE":
            self.emit(depth - 1, "else:")
            return depth
        if re.match(r"^END\s*(IF|DO)\b", u):
            if re.match(r"^END\s*DO\b", u):
                self.open_loops -= 1
            return depth - 1
        m = re.match(r"^DO\s+WHILE\s*\(", u)
        if m:
            j = _matching_paren(st, st.index("("))
            self.emit(depth, f"while {self.ex(st[st.index('(') + 1:j])}:")
            self.open_loops += 1
            return depth + 1
        m = re.match(r"^DO\s+(\w+)\s*=\s*(.*)$", st, re.I)
        if m:
            self.open_loops += 1
            parts = _split_top(m.group(2))
            self.emit(depth, f"for {self.ctx.rename(m.group(1).lower())} in _do({', '.join(self.ex(p) for p in parts)}):")
            return depth + 1
        if u == "DO":
            self.emit(depth, "while True:")
            self.open_loops += 1
            return depth + 1
        if u == "CYCLE":
            self.emit(depth, "continue")
            return depth
        if u == "EXIT":
            self.emit(depth, "break")
            return depth
        if u == "RETURN":
            self.emit(depth, self.ret())
            return depth
        if u == "STOP" or u.startswith("STOP "):
            self.emit(depth, "raise RuntimeError('STOP')")
            return depth
        if u.startswith("CALL "):
            self.call(depth, st)
            return depth
        if u.startswith("WRITE") or u.startswith("PRINT"):
            self.emit(depth, "pass")
            return depth
        if u.startswith("ALLOCATE"):
            inner = st[st.index("(") + 1:_matching_paren(st, st.index("("))]
            for ent in _split_top(inner):
                ent = ent.strip()
                j = ent.rindex("(")
                target, dims = ent[:j].strip(), ent[j + 1:-1]
                name = target.lower()
                if "%" in target:       # a component: its type is not known here; the hot path only allocates REAL(dp) components
                    parent, last = self.ex(target).rsplit(".", 1)
                    self.emit(depth, f"setattr({parent}, {last!r}, _alloc(_np.float64, {self.bounds(dims)}))")
                else:
                    dt = "_np.int32" if self.decl[name]["type"].startswith("integer") or self.decl[name]["type"].startswith("logical") else "_np.float64"
                    self.emit(depth, f"{self.ctx.rename(name)} = _alloc({dt}, {self.bounds(dims)})")
            return depth
        if u.startswith("DEALLOCATE") or u.startswith("NULLIFY"):
            self.emit(depth, "pass")
            return depth
        self.assignment(depth, st)
        return depth

    def translate(self):
        self.parse_header(self.st[0][1])
        body = []
        for n, st in self.st[1:-1]:
            u = st.upper()
            if u.startswith("USE ") or u.startswith("IMPLICIT ") or u.startswith("EXTERNAL "):
                continue
            if "::" in st and re.match(r"^(INTEGER|REAL|LOGICAL|CHARACTER|TYPE\s*\(|DOUBLE PRECISION|COMPLEX)", u):
                self.parse_decl(st)
                continue
            body.append((n, st))
        for name, d in self.decl.items():
            if d["dims"] is not None:
                self.ctx.arrays.add(name)
        self.emit(0, f"def {self.name}({', '.join(self.ctx.rename(a) for a in self.dummies)}):")
        # local arrays with explicit shape; PARAMETERs and initialised locals
        for name, d in self.decl.items():
            if name in self.dummies:
                if d["dims"] is not None:   # an actual argument may be an array expression or a section (numpy): give it Fortran subscripts
                    self.emit(1, f"{self.ctx.rename(name)} = _wrap({self.ctx.rename(name)})")
                continue
            deferred = d["dims"] is not None and any(q.strip() == ":" for q in _split_top(d["dims"]))
            if d["dims"] is not None and not d["allocatable"] and not deferred:
                dt = "_np.float64" if d["type"].startswith("real") else "_np.int32"
                self.emit(1, f"{self.ctx.rename(name)} = _alloc({dt}, {self.bounds(d['dims'])})")
                if d.get("init"):
                    self.emit(1, f"{self.ctx.rename(name)}.fill({self.ex(d['init'])})")
            elif d["dims"] is None and d.get("init") is not None:
                self.emit(1, f"{self.ctx.rename(name)} = {self.ex(d['init'])}")
            elif d["dims"] is None and not d["type"].startswith("type") and not d["type"].startswith("character"):
                # Fortran locals start undefined; give OUT scalars and locals a defined placeholder so that `return` works
                self.emit(1, f"{self.ctx.rename(name)} = {'False' if d['type'].startswith('logical') else ('0' if d['type'].startswith('integer') else '_np.float64(0.0)')}")
        depth = 1
        for n, st in body:
            try:
                depth = self.statement(depth, st)
            except Unsupported as e:
                raise Unsupported(f"{self.name}, source line {n}: {e}") from None
        if depth != 1:
            raise Unsupported(f"{self.name}: unbalanced blocks")
        self.emit(1, self.ret())
        return "\n".join(self.lines)


class Program:
    """A set of translated units sharing one Python namespace."""

    def __init__(self, constants=None):
        self.ns = dict(_RT)
        self.ns.update({k.lower(): v for k, v in (constants or {}).items()})
        self.functions = set()
        self.subroutine_outs = {}
        self.sources = {}

    def declare(self, text, names):
        """Register the interfaces (function-ness, positions of scalar OUT dummies) of units that are translated later: call sites
        may precede their callee, in the same or in another file."""
        for name in names:
            st = extract_unit(text, name)
            pre = Unit(st, self.functions, self.subroutine_outs)
            pre.parse_header(st[0][1])
            for n, s in st[1:-1]:
                if "::" in s and re.match(r"^(INTEGER|REAL|LOGICAL|CHARACTER|TYPE\s*\(|DOUBLE PRECISION)", s.upper()):
                    pre.parse_decl(s)
            if pre.kind == "function":
                self.functions.add(pre.name)
            else:
                self.subroutine_outs[pre.name] = [k for k, _ in pre.outs()]
        return self

    def add(self, text, names, whole_components=()):
        """Translate the named units of `text` (any order)."""
        self.declare(text, names)
        for name in names:
            st = extract_unit(text, name)
            u = Unit(st, self.functions, self.subroutine_outs)
            u.whole_components = {c.lower() for c in whole_components}
            src = u.translate()
            self.sources[name.lower()] = src
            exec(compile(src, f"<f90py:{name}>", "exec"), self.ns)
        return self

    def __getattr__(self, k):
        try:
            return self.__dict__["ns"][k.lower()]
        except KeyError:
            raise AttributeError(k) from None
