"""Tracing of the user's `reaction-rates` section into the expression IR.

Mirrors `reactionRateExe` (PyREMOT/docs/rmtReaction.py:11-61):

* the base dict {"R_CONST","T","P","MoFri","SpCoi"} is merged *before* the
  user's VARS (`{**loopDict, **varDict}`, :39) — a user key that shadows `T`
  or `P` keeps the base position but takes the user's value;
* entries are evaluated in insertion order; only `types.FunctionType` values
  are called, with the dict of everything evaluated so far (:44-51); anything
  else is stored as a constant — numeric scalars become *kinetic-parameter
  slots* (uniform or per-instance, the natural sweep / estimation variables);
* RATES lambdas are then evaluated in insertion order; rate j belongs to
  reaction j by position (:56-58).

User lambdas reference `math` / `np` through their module globals.  They are
re-created with the same code object over a copy of their globals in which
those modules (and functions imported from them) are replaced by symbolic
shims, then called once with symbolic operands.
"""
import builtins
import math
import types

import numpy as np

from .expr import Graph, Sym, SymVec, TraceError, sym_pow

R_CONST = 8.314472  # PyREMOT/core/constants.py:8


# ----------------------------------------------------------------------------
# symbolic shims for math / numpy
# ----------------------------------------------------------------------------
def _any_sym(args):
    return any(isinstance(a, (Sym, SymVec)) for a in args)


def _unary(op, real):
    def f(x):
        if isinstance(x, Sym):
            return Sym(x.g, x.g.mk(op, x.n))
        if isinstance(x, SymVec):
            return SymVec([f(v) for v in x.items])
        return real(x)
    f.__name__ = op
    return f


def _log(x, base=None):
    if not _any_sym((x, base)):
        return math.log(x) if base is None else math.log(x, base)
    if base is None:
        return Sym(x.g, x.g.mk("log", x.n))
    g = x.g if isinstance(x, Sym) else base.g
    x, base = Sym.lift(g, x), Sym.lift(g, base)
    return Sym(g, g.mk("log", x.n))/Sym(g, g.mk("log", base.n))


def _pow(a, b):
    if not _any_sym((a, b)):
        return math.pow(a, b)
    return sym_pow(a, b)


def _minmax(op, real):
    def f(*args):
        if len(args) == 1:
            args = tuple(args[0])
        if not _any_sym(args):
            return real(*args)
        g = next(a.g for a in args if isinstance(a, Sym))
        acc = Sym.lift(g, args[0])
        for v in args[1:]:
            acc = Sym(g, g.mk(op, acc.n, Sym.lift(g, v).n))
        return acc
    return f


def _sum(v, *a, **k):
    if isinstance(v, SymVec):
        return v.sum()
    if isinstance(v, (list, tuple)) and _any_sym(v):
        acc = v[0]
        for t in v[1:]:
            acc = acc + t
        return acc
    return builtins.sum(v, *a) if not k and not isinstance(v, np.ndarray) else np.sum(v, *a, **k)


def _dot(a, b):
    a = a.items if isinstance(a, SymVec) else list(a)
    b = b.items if isinstance(b, SymVec) else list(b)
    acc = a[0]*b[0]
    for u, v in zip(a[1:], b[1:]):
        acc = acc + u*v
    return acc


def _array(v, *a, **k):
    if isinstance(v, SymVec):
        return v
    if isinstance(v, (list, tuple)) and _any_sym(v):
        return SymVec(v)
    return np.array(v, *a, **k)


_FUNCS = {
    "exp": _unary("exp", math.exp), "sqrt": _unary("sqrt", math.sqrt), "log10": _unary("log10", math.log10),
    "log2": _unary("log2", math.log2), "log1p": _unary("log1p", math.log1p), "expm1": _unary("expm1", math.expm1),
    "sin": _unary("sin", math.sin), "cos": _unary("cos", math.cos), "tan": _unary("tan", math.tan),
    "tanh": _unary("tanh", math.tanh), "sinh": _unary("sinh", math.sinh), "cosh": _unary("cosh", math.cosh),
    "atan": _unary("atan", math.atan), "asin": _unary("asin", math.asin), "acos": _unary("acos", math.acos),
    "fabs": _unary("abs", math.fabs), "log": _log, "pow": _pow,
}


class _Shim:
    """Attribute-compatible stand-in for the `math` or `numpy` module."""

    def __init__(self, real, extra):
        self.__dict__["_real"] = real
        self.__dict__["_tab"] = dict(_FUNCS)
        self._tab.update(extra)

    def __getattr__(self, k):
        t = self.__dict__["_tab"]
        if k in t:
            return t[k]
        return getattr(self.__dict__["_real"], k)


MATH_SHIM = _Shim(math, {})
NUMPY_SHIM = _Shim(np, {
    "abs": _unary("abs", np.abs), "absolute": _unary("abs", np.abs), "power": _pow, "float_power": _pow,
    "sum": _sum, "dot": _dot, "array": _array, "asarray": _array, "arctan": _unary("atan", np.arctan),
    "arcsin": _unary("asin", np.arcsin), "arccos": _unary("acos", np.arccos), "cbrt": _unary("cbrt", np.cbrt),
    "maximum": _minmax("max", np.maximum), "minimum": _minmax("min", np.minimum),
})

_BUILTIN_SHIMS = {"abs": _unary("abs", abs), "max": _minmax("max", max), "min": _minmax("min", min),
                  "sum": _sum, "pow": _pow}


def _swap(v):
    """Replacement for one global / closure value, or None to keep it."""
    if v is math:
        return MATH_SHIM
    if v is np:
        return NUMPY_SHIM
    if isinstance(v, types.ModuleType) and v.__name__ == "numpy.lib":
        return types.SimpleNamespace(math=MATH_SHIM)          # `from numpy.lib import math` idiom
    if isinstance(v, (types.BuiltinFunctionType, np.ufunc)) or callable(v):
        mod = getattr(v, "__module__", None)
        name = getattr(v, "__name__", None)
        if mod == "math" and name in MATH_SHIM._tab:
            return MATH_SHIM._tab[name]
        if isinstance(v, np.ufunc) or (mod or "").startswith("numpy"):
            if name in NUMPY_SHIM._tab:
                return NUMPY_SHIM._tab[name]
    return None


def rebind(fn):
    """Same code object, globals/closure with math & numpy replaced by shims."""
    g = dict(fn.__globals__)
    for k, v in list(g.items()):
        r = _swap(v)
        if r is not None:
            g[k] = r
    for name, shim in _BUILTIN_SHIMS.items():
        if name not in fn.__globals__:
            g[name] = shim                      # shadows the builtin for this function only
    closure = None
    if fn.__closure__:
        cells = []
        for c in fn.__closure__:
            try:
                v = c.cell_contents
            except ValueError:
                cells.append(c)
                continue
            r = _swap(v)
            cells.append(types.CellType(r) if r is not None else c)
        closure = tuple(cells)
    nf = types.FunctionType(fn.__code__, g, fn.__name__, fn.__defaults__, closure)
    nf.__kwdefaults__ = fn.__kwdefaults__
    return nf


# ----------------------------------------------------------------------------
# the traced kinetics
# ----------------------------------------------------------------------------
class KineticsIR:
    """Result of tracing: graph, rate outputs, parameter slots, partials."""

    def __init__(self, nc):
        self.nc = nc
        self.g = Graph()
        self.rates = []          # output nodes, one per reaction (by position)
        self.rate_names = []
        self.param_names = []    # scalar VARS entries, in VARS order
        self.param_defaults = []
        self.partials = None

    @property
    def nr(self):
        return len(self.rates)

    def input_nodes(self):
        g = self.g
        return ([g.input("T"), g.input("P")] + [g.input("y%d" % i) for i in range(self.nc)]
                + [g.input("C%d" % i) for i in range(self.nc)])

    def differentiate(self):
        """dR_j/d{T, P, y_i, C_i} as graph nodes (None where identically 0)."""
        g = self.g
        ins = self.input_nodes()
        zero = g.const(0.0)
        used = set()
        for n in g.topo(self.rates):
            if n.op == "in":
                used.add(n.name)
        out = {}
        for w in ins:
            if w.name not in used:
                out[w.name] = [None]*self.nr
                continue
            memo = {}
            row = []
            for r in self.rates:
                d = g.diff(r, w, memo)
                row.append(None if d is zero else d)
            out[w.name] = row
        self.partials = out
        return out

    def positive_species(self):
        """Species that must stay STRICTLY positive: those whose vanishing (on its own, every other species, T and P
        positive) makes a denominator zero or puts a non-positive argument into a logarithm / a negative or zero one
        into a non-integer or negative power — anywhere in the rates or in their partial derivatives (so the 1/sqrt
        of a square root's derivative counts).  There the kinetics have a pole: the reference raises (`math domain
        error`, `ZeroDivisionError`) or silently runs on the wrong side of it.

        Decided by a sign analysis of the DAG (abstract values = subsets of {-, 0, +}), not by mere dependence: the
        LHHW denominator `1 + K*p_i` depends on species i but is >= 1 when p_i -> 0, so a zero-feed species that only
        occurs there is NOT flagged and may sit at exactly 0 for the whole integration.  Kinetic-parameter
        slots carry the sign of their default value."""
        g = self.g
        if self.partials is None:
            self.differentiate()
        outs = list(self.rates) + [d for row in self.partials.values() for d in row if d is not None]
        order = g.topo(outs)
        NEG, ZERO, POS = 1, 2, 4
        FULL = NEG | ZERO | POS

        def bits(m):
            return [b for b in (NEG, ZERO, POS) if m & b]

        def flip(m):
            return (POS if m & NEG else 0) | (m & ZERO) | (NEG if m & POS else 0)

        def add(a, b):
            r = 0
            for x in bits(a):
                for y in bits(b):
                    r |= y if x == ZERO else x if (y == ZERO or x == y) else FULL
            return r

        def mul(a, b):
            r = 0
            for x in bits(a):
                for y in bits(b):
                    r |= ZERO if ZERO in (x, y) else (POS if x == y else NEG)
            return r

        def sgn(v):
            return POS if v > 0 else NEG if v < 0 else ZERO if v == 0 else FULL

        dep = {}
        for n in order:
            if n.op == "in":
                dep[n.id] = (1 << int(n.name[1:])) if (n.name[0] in "yC" and n.name[1:].isdigit()) else 0
            else:
                m = 0
                for a in n.args:
                    m |= dep[a.id]
                dep[n.id] = m
        flagged = []
        for i in range(self.nc):
            sg = {}
            need = False
            for n in order:
                a = [sg[x.id] for x in n.args]
                pole = False
                if n.op == "in":
                    if n.name.startswith("kp"):
                        v = sgn(self.param_defaults[int(n.name[2:])])
                    elif n.name in ("y%d" % i, "C%d" % i):
                        v = ZERO | POS
                    else:
                        v = POS
                elif n.op == "const":
                    v = sgn(n.value)
                elif n.op == "neg":
                    v = flip(a[0])
                elif n.op == "abs":
                    v = (a[0] & ZERO) | (POS if a[0] & (NEG | POS) else 0)
                elif n.op in ("exp", "exp10", "cosh"):
                    v = POS
                elif n.op == "sqrt":
                    pole = bool(a[0] & NEG)
                    v = a[0] & (ZERO | POS) or FULL
                elif n.op in ("cbrt", "sinh", "tanh", "atan", "asin", "expm1"):
                    v = a[0]
                elif n.op in ("log", "log10", "log2"):
                    pole = bool(a[0] & (NEG | ZERO))
                    v = FULL
                elif n.op == "log1p":
                    pole = bool(a[0] & NEG)
                    v = a[0] if not a[0] & NEG else FULL
                elif n.op == "add":
                    v = add(a[0], a[1])
                elif n.op == "sub":
                    v = add(a[0], flip(a[1]))
                elif n.op == "mul":
                    v = mul(a[0], a[1])
                elif n.op == "div":
                    pole = bool(a[1] & ZERO)
                    v = FULL if pole else mul(a[0], a[1])
                elif n.op == "pow":
                    expo = n.args[1]
                    pos_const = expo.op == "const" and expo.value > 0
                    pole = bool(a[0] & NEG) or (bool(a[0] & ZERO) and not pos_const)
                    v = POS if a[0] == POS else ((ZERO | POS) if (not a[0] & NEG and pos_const) else FULL)
                elif n.op == "powi":
                    k = int(n.value)
                    pole = k < 0 and bool(a[0] & ZERO)
                    if k == 0:
                        v = POS
                    elif k % 2 == 0:
                        v = (a[0] & ZERO) | (POS if a[0] & (NEG | POS) else 0)
                    else:
                        v = a[0]
                    if pole:
                        v = FULL
                elif n.op in ("min", "max"):
                    rank = {NEG: 0, ZERO: 1, POS: 2}
                    pick = max if n.op == "max" else min
                    v = 0
                    for x in bits(a[0]):
                        for y in bits(a[1]):
                            v |= pick(x, y, key=rank.get)
                else:                               # sin, cos, tan, acos, atan2, ...: no sign information
                    v = FULL
                sg[n.id] = v
                if pole and (dep[n.id] >> i) & 1:
                    need = True
            flagged.append(need)
        return flagged

    def evaluate(self, T, P, y, C, params=None):
        """Host interpreter of the traced rates (tests of the tracer only)."""
        env = {"T": T, "P": P}
        env.update({"y%d" % i: v for i, v in enumerate(y)})
        env.update({"C%d" % i: v for i, v in enumerate(C)})
        pv = self.param_defaults if params is None else params
        env.update({"kp%d" % k: v for k, v in enumerate(pv)})
        return self.g.evaluate(self.rates, env)

    def flops(self):
        alg, wt = self.g.count_flops(self.rates)
        res = {"rates_alg": alg, "rates_weighted": wt}
        if self.partials is not None:
            outs = list(self.rates) + [d for row in self.partials.values() for d in row if d is not None]
            alg, wt = self.g.count_flops(outs)
            res.update({"rates_jac_alg": alg, "rates_jac_weighted": wt})
        return res


def _is_scalar_param(v):
    return isinstance(v, (int, float, np.integer, np.floating)) and not isinstance(v, bool)


def trace_kinetics(varis, rates, nc):
    """Trace VARS/RATES (rmtReaction.py:11-61 semantics) into a KineticsIR."""
    ir = KineticsIR(nc)
    g = ir.g
    T, P = Sym(g, g.input("T")), Sym(g, g.input("P"))
    MoFri = SymVec([Sym(g, g.input("y%d" % i)) for i in range(nc)])
    SpCoi = SymVec([Sym(g, g.input("C%d" % i)) for i in range(nc)])
    merged = {"R_CONST": R_CONST, "T": T, "P": P, "MoFri": MoFri, "SpCoi": SpCoi}
    merged.update(varis)
    exe = {}
    for key, v in merged.items():
        if isinstance(v, types.FunctionType):
            try:
                val = rebind(v)(exe)
            except TraceError:
                raise
            except Exception as e:
                raise TraceError("VARS[%r] could not be traced: %s: %s" % (key, type(e).__name__, e)) from e
        elif key in varis and _is_scalar_param(v):
            k = len(ir.param_names)
            ir.param_names.append(key)
            ir.param_defaults.append(float(v))
            val = Sym(g, g.input("kp%d" % k))
        else:
            val = v
        exe[key] = val
    for key, f in rates.items():
        if not isinstance(f, types.FunctionType):
            raise TraceError("RATES[%r] must be a function of the variables dict" % key)
        try:
            val = rebind(f)(exe)
        except TraceError:
            raise
        except Exception as e:
            raise TraceError("RATES[%r] could not be traced: %s: %s" % (key, type(e).__name__, e)) from e
        ir.rates.append(Sym.lift(g, val).n)
        ir.rate_names.append(key)
    return ir
