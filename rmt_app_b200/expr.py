"""Order-preserving expression IR for user kinetics.

The reference evaluates the user's `VARS` / `RATES` sections by calling Python
lambdas with a dict (PyREMOT/docs/rmtReaction.py:11-61).  To run the same
expressions on the device they are *traced*: every lambda is called once with
symbolic operands (`Sym`) and the arithmetic it performs is recorded, in the
order Python performs it, into a hash-consed DAG (`Graph`).  The DAG is then
differentiated symbolically and printed as CUDA (see codegen.py).

Design notes
* No re-association: `a*b*c` is recorded as `(a*b)*c`, constants are folded
  only when *all* operands are constants (Python would have computed the same
  double).  SymPy is deliberately not used: it re-associates.
* Algebraic identities (x*0, x*1, x+0 ...) are applied only while building
  derivative nodes (`simplify=True`), never on the primal.
* Data-dependent control flow (`if x['T'] > 500`) cannot be traced and raises
  `TraceError` (use `max`/`min`/`abs` — those are recorded as ops).
"""
import math

__all__ = ["Graph", "Sym", "SymVec", "TraceError", "FLOP_WEIGHT"]


class TraceError(Exception):
    pass


UNARY = ("neg", "exp", "log", "log10", "log2", "sqrt", "exp10", "abs", "sin", "cos", "tan",
         "tanh", "sinh", "cosh", "atan", "asin", "acos", "expm1", "log1p", "cbrt")
BINARY = ("add", "sub", "mul", "div", "pow", "min", "max", "atan2")

# weighted FP64-instruction cost per op (SURVEY.md §8(d) convention: exp ~25,
# log ~35, div/sqrt ~10); "algorithmic" flops count every op as 1.
FLOP_WEIGHT = {
    "add": 1, "sub": 1, "mul": 1, "neg": 1, "abs": 1, "min": 1, "max": 1,
    "div": 10, "sqrt": 10, "cbrt": 30, "exp": 25, "exp10": 27, "expm1": 30, "log": 35, "log10": 37, "log2": 35,
    "log1p": 38, "pow": 70, "powi": 0, "sin": 40, "cos": 40, "tan": 60, "tanh": 40, "sinh": 40, "cosh": 40,
    "atan": 40, "asin": 40, "acos": 40, "atan2": 50,
}

_PYFUN = {
    "neg": lambda a: -a, "exp": math.exp, "log": math.log, "log10": math.log10, "log2": math.log2,
    "sqrt": math.sqrt, "exp10": lambda a: math.pow(10.0, a), "abs": abs, "sin": math.sin, "cos": math.cos,
    "tan": math.tan, "tanh": math.tanh, "sinh": math.sinh, "cosh": math.cosh, "atan": math.atan,
    "asin": math.asin, "acos": math.acos, "expm1": math.expm1, "log1p": math.log1p,
    "cbrt": lambda a: math.copysign(abs(a)**(1.0/3.0), a),
    "add": lambda a, b: a + b, "sub": lambda a, b: a - b, "mul": lambda a, b: a*b, "div": lambda a, b: a/b,
    "pow": math.pow, "min": min, "max": max, "atan2": math.atan2,
}


class Node:
    __slots__ = ("id", "op", "args", "value", "name")

    def __init__(self, id, op, args=(), value=None, name=None):
        self.id, self.op, self.args, self.value, self.name = id, op, args, value, name

    def __repr__(self):
        if self.op == "const":
            return repr(self.value)
        if self.op == "in":
            return self.name
        return "%s(%s)" % (self.op, ", ".join("n%d" % a.id for a in self.args))


class Graph:
    """Hash-consed expression DAG."""

    def __init__(self):
        self.nodes = []
        self._key = {}
        self.inputs = {}

    # -- construction ---------------------------------------------------------
    def _new(self, key, op, args=(), value=None, name=None):
        n = self._key.get(key)
        if n is None:
            n = Node(len(self.nodes), op, args, value, name)
            self.nodes.append(n)
            self._key[key] = n
        return n

    def const(self, v):
        v = float(v)
        # distinguish -0.0 / nan by repr
        return self._new(("const", repr(v)), "const", value=v)

    def input(self, name):
        n = self._new(("in", name), "in", name=name)
        self.inputs[name] = n
        return n

    def mk(self, op, *args, simplify=False):
        if all(a.op == "const" for a in args):
            try:
                return self.const(_PYFUN[op](*[a.value for a in args]))
            except (ValueError, ZeroDivisionError, OverflowError):
                if not simplify:
                    raise
        if simplify:
            s = self._simplify(op, args)
            if s is not None:
                return s
        return self._new((op,) + tuple(a.id for a in args), op, tuple(args))

    def powi(self, a, n, simplify=False):
        n = int(n)
        if a.op == "const":
            return self.const(math.pow(a.value, n))
        if simplify:
            if n == 0:
                return self.const(1.0)
            if n == 1:
                return a
        return self._new(("powi", a.id, n), "powi", (a,), value=n)

    def _simplify(self, op, args):
        def is_c(n, v):
            return n.op == "const" and n.value == v
        if op == "add":
            a, b = args
            if is_c(a, 0.0):
                return b
            if is_c(b, 0.0):
                return a
        elif op == "sub":
            a, b = args
            if is_c(b, 0.0):
                return a
            if is_c(a, 0.0):
                return self.mk("neg", b, simplify=True)
            if a is b:
                return self.const(0.0)
        elif op == "mul":
            a, b = args
            if is_c(a, 0.0) or is_c(b, 0.0):
                return self.const(0.0)
            if is_c(a, 1.0):
                return b
            if is_c(b, 1.0):
                return a
            if is_c(a, -1.0):
                return self.mk("neg", b, simplify=True)
            if is_c(b, -1.0):
                return self.mk("neg", a, simplify=True)
        elif op == "div":
            a, b = args
            if is_c(a, 0.0):
                return self.const(0.0)
            if is_c(b, 1.0):
                return a
        elif op == "neg":
            (a,) = args
            if a.op == "neg":
                return a.args[0]
            if is_c(a, 0.0):
                return a
        return None

    # -- symbolic differentiation --------------------------------------------
    def diff(self, node, wrt, memo=None):
        """d node / d wrt as a node of the same graph (forward accumulation,
        memoised per `wrt`; identities applied)."""
        if memo is None:
            memo = {}
        S = dict(simplify=True)
        zero, one = self.const(0.0), self.const(1.0)

        def d(n):
            r = memo.get(n.id)
            if r is not None:
                return r
            op = n.op
            if op == "const":
                r = zero
            elif op == "in":
                r = one if n is wrt else zero
            else:
                a = n.args[0]
                da = d(a)
                if op in ("add", "sub", "mul", "div", "pow", "min", "max", "atan2"):
                    b = n.args[1]
                    db = d(b)
                if op == "add":
                    r = self.mk("add", da, db, **S)
                elif op == "sub":
                    r = self.mk("sub", da, db, **S)
                elif op == "neg":
                    r = self.mk("neg", da, **S)
                elif op == "mul":
                    r = self.mk("add", self.mk("mul", da, b, **S), self.mk("mul", a, db, **S), **S)
                elif op == "div":
                    # (da - n*db)/b
                    r = self.mk("div", self.mk("sub", da, self.mk("mul", n, db, **S), **S), b, **S)
                elif op == "powi":
                    k = n.value
                    if da is zero:
                        r = zero
                    else:
                        r = self.mk("mul", self.mk("mul", self.const(float(k)), self.powi(a, k - 1, simplify=True), **S), da, **S)
                elif op == "pow":
                    # n = a**b : n*(db*log(a) + b*da/a)
                    t1 = self.mk("mul", db, self.mk("log", a), **S) if db is not zero else zero
                    t2 = self.mk("div", self.mk("mul", b, da, **S), a, **S) if da is not zero else zero
                    r = self.mk("mul", n, self.mk("add", t1, t2, **S), **S)
                elif op == "exp":
                    r = self.mk("mul", n, da, **S)
                elif op == "expm1":
                    r = self.mk("mul", self.mk("add", n, one), da, **S)
                elif op == "exp10":
                    r = self.mk("mul", self.mk("mul", n, self.const(math.log(10.0)), **S), da, **S)
                elif op == "log":
                    r = self.mk("div", da, a, **S)
                elif op == "log1p":
                    r = self.mk("div", da, self.mk("add", one, a), **S)
                elif op == "log10":
                    r = self.mk("div", da, self.mk("mul", a, self.const(math.log(10.0))), **S)
                elif op == "log2":
                    r = self.mk("div", da, self.mk("mul", a, self.const(math.log(2.0))), **S)
                elif op == "sqrt":
                    r = self.mk("div", da, self.mk("mul", self.const(2.0), n), **S)
                elif op == "cbrt":
                    r = self.mk("div", da, self.mk("mul", self.const(3.0), self.mk("mul", n, n)), **S)
                elif op == "sin":
                    r = self.mk("mul", self.mk("cos", a), da, **S)
                elif op == "cos":
                    r = self.mk("neg", self.mk("mul", self.mk("sin", a), da, **S), **S)
                elif op == "tan":
                    r = self.mk("mul", self.mk("add", one, self.mk("mul", n, n)), da, **S)
                elif op == "tanh":
                    r = self.mk("mul", self.mk("sub", one, self.mk("mul", n, n)), da, **S)
                elif op == "sinh":
                    r = self.mk("mul", self.mk("cosh", a), da, **S)
                elif op == "cosh":
                    r = self.mk("mul", self.mk("sinh", a), da, **S)
                elif op == "atan":
                    r = self.mk("div", da, self.mk("add", one, self.mk("mul", a, a)), **S)
                elif op == "asin":
                    r = self.mk("div", da, self.mk("sqrt", self.mk("sub", one, self.mk("mul", a, a))), **S)
                elif op == "acos":
                    r = self.mk("neg", self.mk("div", da, self.mk("sqrt", self.mk("sub", one, self.mk("mul", a, a))), **S), **S)
                elif op in ("abs", "min", "max", "atan2"):
                    if da is zero and (len(n.args) == 1 or db is zero):
                        r = zero
                    else:
                        raise TraceError("derivative of %s is not implemented for a state-dependent argument" % op)
                else:
                    raise TraceError("no derivative rule for " + op)
            memo[n.id] = r
            return r

        # iterative post-order to avoid recursion limits on long chains
        order = self.topo([node])
        for n in order:
            d(n)
        return memo[node.id]

    # -- traversal ------------------------------------------------------------
    def topo(self, outs):
        seen, order = set(), []
        stack = [(n, False) for n in reversed(outs)]
        while stack:
            n, done = stack.pop()
            if done:
                order.append(n)
                continue
            if n.id in seen:
                continue
            seen.add(n.id)
            stack.append((n, True))
            for a in reversed(n.args):
                if a.id not in seen:
                    stack.append((a, False))
        return order

    def depends_on(self, node, inputs):
        ids = {i.id for i in inputs}
        return any(n.id in ids for n in self.topo([node]))

    def count_flops(self, outs):
        """(algorithmic, weighted) flop counts of evaluating `outs` once."""
        alg = wt = 0
        for n in self.topo(outs):
            if n.op in ("const", "in"):
                continue
            if n.op == "powi":
                k = abs(n.value)
                muls = max(k.bit_length() - 1 + bin(k).count("1") - 1, 0) if k else 0
                alg += 1
                wt += muls + (10 if n.value < 0 else 0)
            else:
                alg += 1
                wt += FLOP_WEIGHT[n.op]
        return alg, wt

    def evaluate(self, outs, env):
        """Reference interpreter (host-side tests of the code generator only;
        the product never computes with it).  `env`: input name -> float."""
        val = {}
        for n in self.topo(outs):
            if n.op == "const":
                val[n.id] = n.value
            elif n.op == "in":
                val[n.id] = float(env[n.name])
            elif n.op == "powi":
                val[n.id] = math.pow(val[n.args[0].id], n.value)
            else:
                val[n.id] = _PYFUN[n.op](*[val[a.id] for a in n.args])
        return [val[o.id] for o in outs]


def _is_intlike(v):
    return float(v) == int(v) and abs(v) <= 64


class Sym:
    """Symbolic operand handed to user lambdas."""
    __slots__ = ("g", "n")
    __array_priority__ = 1000

    def __init__(self, g, n):
        self.g, self.n = g, n

    @staticmethod
    def lift(g, v):
        if isinstance(v, Sym):
            return v
        if isinstance(v, (bool, int, float)) or hasattr(v, "__float__"):
            return Sym(g, g.const(float(v)))
        raise TraceError("cannot use %r (%s) in a kinetics expression" % (v, type(v).__name__))

    def _bin(self, op, o, swap=False):
        o = Sym.lift(self.g, o)
        a, b = (o, self) if swap else (self, o)
        return Sym(self.g, self.g.mk(op, a.n, b.n))

    def __add__(self, o): return self._bin("add", o)
    def __radd__(self, o): return self._bin("add", o, True)
    def __sub__(self, o): return self._bin("sub", o)
    def __rsub__(self, o): return self._bin("sub", o, True)
    def __mul__(self, o): return self._bin("mul", o)
    def __rmul__(self, o): return self._bin("mul", o, True)
    def __truediv__(self, o): return self._bin("div", o)
    def __rtruediv__(self, o): return self._bin("div", o, True)
    def __neg__(self): return Sym(self.g, self.g.mk("neg", self.n))
    def __pos__(self): return self
    def __abs__(self): return Sym(self.g, self.g.mk("abs", self.n))

    def __pow__(self, o):
        return sym_pow(self, o)

    def __rpow__(self, o):
        return sym_pow(Sym.lift(self.g, o), self)

    def _nobool(self, *a):
        raise TraceError("data-dependent comparison/branch in a kinetics expression cannot be lowered to "
                         "straight-line device code; use min()/max()/abs()")
    __bool__ = __lt__ = __le__ = __gt__ = __ge__ = _nobool

    def __float__(self):
        if self.n.op == "const":
            return self.n.value
        raise TraceError("float() of a state-dependent value (use math.* through the traced shim)")

    def __repr__(self):
        return "Sym(%r)" % (self.n,)


def sym_pow(a, b):
    """`a ** b` / math.pow(a, b) with Python's double semantics."""
    g = a.g if isinstance(a, Sym) else b.g
    a, b = Sym.lift(g, a), Sym.lift(g, b)
    if b.n.op == "const" and _is_intlike(b.n.value):
        return Sym(g, g.powi(a.n, int(b.n.value)))
    if b.n.op == "const" and b.n.value == 0.5:
        return Sym(g, g.mk("sqrt", a.n))
    if a.n.op == "const" and a.n.value == 10.0:
        return Sym(g, g.mk("exp10", b.n))
    return Sym(g, g.mk("pow", a.n, b.n))


class SymVec:
    """Stand-in for the `MoFri` / `SpCoi` NumPy vectors of the kinetics dict."""

    def __init__(self, items):
        self.items = list(items)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return SymVec(self.items[i])
        if hasattr(i, "__index__"):
            return self.items[i.__index__()]
        raise TraceError("MoFri/SpCoi must be indexed with integers")

    def __len__(self):
        return len(self.items)

    def __iter__(self):
        return iter(self.items)

    def sum(self):
        acc = self.items[0]
        for v in self.items[1:]:
            acc = acc + v
        return acc

    def _ew(self, o, f):
        if isinstance(o, SymVec):
            return SymVec([f(a, b) for a, b in zip(self.items, o.items)])
        if hasattr(o, "__len__") and not isinstance(o, str):
            return SymVec([f(a, b) for a, b in zip(self.items, list(o))])
        return SymVec([f(a, o) for a in self.items])

    def __mul__(self, o): return self._ew(o, lambda a, b: a*b)
    def __rmul__(self, o): return self._ew(o, lambda a, b: b*a)
    def __add__(self, o): return self._ew(o, lambda a, b: a+b)
    def __radd__(self, o): return self._ew(o, lambda a, b: b+a)
    def __sub__(self, o): return self._ew(o, lambda a, b: a-b)
    def __rsub__(self, o): return self._ew(o, lambda a, b: b-a)
    def __truediv__(self, o): return self._ew(o, lambda a, b: a/b)
    def __rtruediv__(self, o): return self._ew(o, lambda a, b: b/a)
    def __pow__(self, o): return self._ew(o, lambda a, b: a**b)
