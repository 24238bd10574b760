"""Code generator: ModelSpec -> `rmt_model.cuh`.

The generated header is the model-specific half of the NVRTC translation
unit; the hand-written half (`csrc/rmt_kernels.cu`) `#include`s it.  It holds

* compile-time sizes (species, reactions, unknowns, kinetic-parameter slots),
* component / reaction literals from the component table,
* `rmt_rates()`      — the traced RATES as straight-line FP64 device code,
* `rmt_rates_jac()`  — the same plus dR_j/d{T, P, y_i, C_i} from symbolic
  differentiation of the traced DAG (shared sub-expressions emitted once),
* the Rosenbrock tableau as constexpr arrays,
* per-evaluation flop counts recomputed from the DAG.

Expression printing keeps Python's evaluation order; `x**k` / `math.pow(x,k)`
with small integer k becomes repeated multiplication (<= 1 ulp from the
correctly-rounded libm pow the reference calls), everything else maps to the
CUDA double-precision math library.
"""
from .expr import Graph
from .tableau import TABLEAUX

_CFUN = {
    "exp": "RMT_EXP", "log": "RMT_LOG", "log10": "RMT_LOG10", "log2": "log2", "sqrt": "RMT_SQRT", "exp10": "RMT_EXP10",
    "abs": "fabs", "sin": "sin", "cos": "cos", "tan": "tan", "tanh": "tanh", "sinh": "sinh", "cosh": "cosh",
    "atan": "atan", "asin": "asin", "acos": "acos", "expm1": "expm1", "log1p": "log1p", "cbrt": "cbrt",
    "pow": "pow", "min": "fmin", "max": "fmax", "atan2": "atan2",
}
_INFIX = {"add": "+", "sub": "-", "mul": "*", "div": "/"}


def lit(v):
    v = float(v)
    if v != v:
        return "(0.0/0.0)"
    if v in (float("inf"), float("-inf")):
        return "(%s1.0/0.0)" % ("-" if v < 0 else "")
    s = repr(v)
    if "e" not in s and "." not in s:
        s += ".0"
    return s


def _powi_expr(a, k):
    if k == 0:
        return "1.0"
    neg = k < 0
    k = abs(k)
    if k > 16:
        return "pow(%s, %s)" % (a, lit(-k if neg else k))
    if neg:
        return "RMT_DIV(1.0, %s)" % _powi_expr(a, k)
    # square-and-multiply over a fully parenthesised chain
    e = "*".join([a]*k) if k <= 3 else None
    if e is None:
        half = _powi_expr(a, k//2)
        e = "((%s)*(%s)%s)" % (half, half, "*" + a if k % 2 else "")
    return "(%s)" % e


class ConstPool:
    """Full-mantissa double literals are fetched from constant memory (a c[bank][off]
    operand of DFMA/DMUL) instead of being materialised with two UMOVs each; values
    whose low 32 bits are zero (1.0, 0.5, 1000.0 ...) fit the instruction's immediate."""

    def __init__(self, name="RMT_KC"):
        self.name, self.index, self.values = name, {}, []

    def ref(self, v):
        import struct
        v = float(v)
        bits = struct.unpack("<Q", struct.pack("<d", v))[0]
        if bits & 0xFFFFFFFF == 0 or v != v:
            s = lit(v)
            return "(" + s + ")" if v < 0 else s
        k = self.index.get(bits)
        if k is None:
            k = self.index[bits] = len(self.values)
            self.values.append(v)
        return "RMT_KL(%d, %s)" % (k, lit(v))

    def declaration(self):
        vals = self.values or [0.0]
        return ("__constant__ double %s[%d] = %s;\n"
                "#if RMT_USE_CBANK\n#define RMT_KL(k, v) %s[k]\n#else\n#define RMT_KL(k, v) (v)\n#endif"
                % (self.name, len(vals), _arr(vals), self.name))


def emit_dag(g: Graph, outs, in_name, indent="    ", pool=None):
    """Return (lines, refs): C statements computing every node reachable from
    `outs`, and the C expression naming each requested output."""
    lines = []
    name = {}
    for n in g.topo([o for o in outs if o is not None]):
        if n.op == "const":
            if pool is not None:
                name[n.id] = pool.ref(n.value)
                continue
            name[n.id] = lit(n.value)
            if n.value < 0:
                name[n.id] = "(" + name[n.id] + ")"
        elif n.op == "in":
            name[n.id] = in_name(n.name)
        else:
            a = [name[x.id] for x in n.args]
            if n.op == "div":
                e = "RMT_DIV(%s, %s)" % (a[0], a[1])
            elif n.op in _INFIX:
                e = "%s %s %s" % (a[0], _INFIX[n.op], a[1])
            elif n.op == "neg":
                e = "-%s" % a[0]
            elif n.op == "powi":
                e = _powi_expr(a[0], n.value)
            else:
                e = "%s(%s)" % (_CFUN[n.op], ", ".join(a))
            v = "t%d" % n.id
            lines.append("%sconst double %s = %s;" % (indent, v, e))
            name[n.id] = v
    return lines, [name[o.id] if o is not None else "0.0" for o in outs]


def _arr(vals, fmt=lit):
    return "{" + ", ".join(fmt(v) for v in vals) + "}"


def _arr2(rows, fmt=lit):
    return "{" + ", ".join(_arr(r, fmt) for r in rows) + "}"


def balance_flops(spec):
    """Operation count of the hand-written balance code in csrc/rmt_kernels.cu
    (`rmt_point` + `n1_eval` / the N2 node body), excluding the traced rates.
    Returns dict(rhs=(alg, weighted), jac=(alg, weighted)); the Jacobian count is
    for one evaluation of f AND its full n x n Jacobian.  Divisions weigh 10."""
    nc, nr = spec.nc, spec.nr
    nnz = int((spec.nu != 0).sum())
    energy = 0 if spec.iso else 1
    # un-scale, fractions, MW, EOS density
    mul_add = nc + 3 + nc + 2*nc + 1 + 1
    div = nc + 2
    # formation rates, Cp polynomials + means + mixture, heats of reaction
    mul_add += 2*nnz + 2 + 8*nc + 2*nc + energy*(8*nr + 2*nr + 1 + 3)
    if spec.model in ("N1", "M7"):
        mul_add += 3 + 7 + nc + energy*6      # velocities, Ergun, balances
        div += 5 + 1 + 1 + nc + energy*3
    else:
        mul_add += 7 + 4*nc + 3 + energy*12   # Ergun march, upwind convection, balances
        div += 1 + nc + 2 + energy*4
    rhs = (mul_add + div, mul_add + 10*div)
    n = spec.n if spec.model == "N1" else spec.n
    percol_ma = 10 + 2*nr + 2*nnz + 2*nc + energy*(4*nr + 10)
    percol_div = 2 + energy*3
    jma = mul_add + 2*nr*nc + 6*nc + 6*nr + n*percol_ma
    jdiv = div + 1 + n*percol_div
    out = {"rhs": rhs, "jac": (jma + jdiv, jma + 10*jdiv)}
    if spec.model in ("N1", "M7"):
        # system form in reaction extents (`n1_eval_sys`): nr directions nu_k (contractions over the non-zeros of
        # nu_k) + P and T directions; rows = nr reaction rows, the Ergun row and the energy row
        sma = mul_add + 2*nr*nc + 6*nc + 6*nr
        for k in range(nr):
            nzk = int((spec.nu[k] != 0).sum())
            sma += 2*nzk + 8 + nr*(4*nzk + 4) + 10 + 3*nr + energy*(2*nr + 10)
        sma += (spec.n - nc)*(10 + nr + 3*nr + energy*(4*nr + 10))
        sdiv = div + 1 + (nr + spec.n - nc)*percol_div
        out["sys"] = (sma + sdiv, sma + 10*sdiv)
    return out


def model_flops(spec, reduced=False):
    """Per-axial-point flop counts: traced rates + balance.  sys_*: one evaluation of the
    integrator's g and A = dg/dx (the full f and J unless it works in reaction extents)."""
    k = spec.kin.flops()
    b = balance_flops(spec)
    sys_ = b["sys"] if (reduced and "sys" in b) else b["jac"]
    return {
        "sys_alg": k["rates_jac_alg"] + sys_[0], "sys_weighted": k["rates_jac_weighted"] + sys_[1],
        "rates_alg": k["rates_alg"], "rates_weighted": k["rates_weighted"],
        "rates_jac_alg": k["rates_jac_alg"], "rates_jac_weighted": k["rates_jac_weighted"],
        "rhs_alg": k["rates_alg"] + b["rhs"][0], "rhs_weighted": k["rates_weighted"] + b["rhs"][1],
        "jac_alg": k["rates_jac_alg"] + b["jac"][0], "jac_weighted": k["rates_jac_weighted"] + b["jac"][1],
    }


def use_extents(spec):
    """Integrate the steady-state models in reaction extents (nr + 2 unknowns instead of nc + 2)?  The
    species balances of N1 (pbHomoReactor.py:3283-3289) and M7 (pbReactor.py:1545-1551) are c(y)*nu^T R(y),
    so the Rosenbrock stage increments lie in the range of nu^T and the iterates are the same either way;
    it pays when there are fewer reactions than species."""
    return spec.model in ("N1", "M7") and spec.nr < spec.nc


def system_size(spec, reduced=None):
    """Dimension of the integrator's linear systems."""
    if reduced is None:
        reduced = use_extents(spec)
    return spec.nr + (spec.n - spec.nc) if reduced else spec.n


def generate_model_header(spec, tableau="rodas4", reduced=None, lanes=1):
    kin, g = spec.kin, spec.kin.g
    nc, nr, nkp = spec.nc, spec.nr, spec.nkp
    if reduced is None:
        reduced = use_extents(spec)
    if reduced and spec.model not in ("N1", "M7"):
        raise ValueError("reaction-extent form exists for the steady-state models only")

    def in_name(nm):
        if nm in ("T", "P"):
            return nm
        if nm[0] == "y":
            return "y[%s]" % nm[1:]
        if nm[0] == "C":
            return "C[%s]" % nm[1:]
        if nm.startswith("kp"):
            return "kp[%s]" % nm[2:]
        raise KeyError(nm)

    L = []
    A = L.append
    A("// generated by rmt_app_b200.codegen — do not edit")
    A("// model %s, components %s, %s" % (spec.model, ",".join(spec.compList), "iso-thermal" if spec.iso else "non-iso-thermal"))
    for j, r in enumerate(spec.reactions):
        A("// reaction %d: %s   rate: RATES[%r]" % (j, r, kin.rate_names[j]))
    A("#pragma once")
    A("#define RMT_MODEL_%s 1" % spec.model)
    A("#define RMT_NC %d" % nc)
    A("#define RMT_NR %d" % nr)
    A("#define RMT_NKP %d" % nkp)
    A("#define RMT_ISO %d" % (1 if spec.iso else 0))
    A("#define RMT_NIN %d" % spec.nin)
    A("// integrator unknowns: 1 = reaction extents (nr + non-species unknowns), 0 = the full state")
    A("#ifndef RMT_REDUCED")
    A("#define RMT_REDUCED %d" % (1 if reduced else 0))
    A("#endif")
    if spec.model == "M9" and lanes not in (0, 1):
        raise ValueError("M9 marches the velocity through the kinetics node by node: one lane per reactor, or the stage pipeline (0)")
    if spec.model in ("N2", "M9"):
        if lanes not in (0, 1, 2, 4, 8, 16, 32):
            raise ValueError("lanes per reactor must be a power of two <= 32, or 0 = stage-pipelined mapping")
        A("// threads per reactor of the dynamic integrator (nodes evaluated in parallel); 0 = one thread per reactor and")
        A("// Rosenbrock stage role (stage pipeline, block = 32 x (stages + 1))")
        A("#ifndef RMT_N2_G")
        A("#define RMT_N2_G %d" % lanes)
        A("#endif")
    for k, nm in enumerate(kin.param_names):
        A("// kinetic parameter slot %d = VARS[%r] (default %r)" % (k, nm, kin.param_defaults[k]))

    # which inputs do the rates depend on (lets the balance code skip zero blocks)
    P = kin.partials
    dep = {
        "T": any(d is not None for d in P["T"]),
        "P": any(d is not None for d in P["P"]),
        "Y": any(d is not None for i in range(nc) for d in P["y%d" % i]),
        "C": any(d is not None for i in range(nc) for d in P["C%d" % i]),
    }
    for k, v in dep.items():
        A("#define RMT_RATES_DEP_%s %d" % (k, 1 if v else 0))

    pos = kin.positive_species()
    A("// species that must stay > 0 for the traced kinetics to be defined (denominators, log, sqrt, pow)")
    A("__device__ constexpr bool RMT_POSITIVE[RMT_NC] = %s;" % _arr(pos, fmt=lambda v: "true" if v else "false"))
    fl = model_flops(spec)
    A("// flops per evaluation at one axial point, recomputed from the traced DAG + the balance code")
    A("// (ALG: 1 per +,-,*,/,sqrt,exp,log,pow;  WT: FP64-instruction weighted, see expr.FLOP_WEIGHT)")
    A("#define RMT_FLOPS_RHS_ALG %d" % fl["rhs_alg"])
    A("#define RMT_FLOPS_RHS_WT %d" % fl["rhs_weighted"])
    A("#define RMT_FLOPS_JAC_ALG %d" % fl["jac_alg"])
    A("#define RMT_FLOPS_JAC_WT %d" % fl["jac_weighted"])
    A("#define RMT_FLOPS_RATES_ALG %d" % fl["rates_alg"])
    A("#define RMT_FLOPS_RATES_WT %d" % fl["rates_weighted"])
    A("#define RMT_FLOPS_RATESJAC_ALG %d" % fl["rates_jac_alg"])
    A("#define RMT_FLOPS_RATESJAC_WT %d" % fl["rates_jac_weighted"])
    A("")
    A("// component table rows in compList order (PyREMOT/data/componentData.py)")
    comps = spec.components
    A("__device__ constexpr double RMT_MW[RMT_NC] = %s;" % _arr([c.MW for c in comps]))
    A("__device__ constexpr double RMT_CP[RMT_NC][4] = %s;" % _arr2([c.cp for c in comps]))
    from .componentdb import Tref
    A("__device__ constexpr double RMT_CPREF[RMT_NC] = %s;  // Cp_i(Tref)" % _arr([c.cp_at(Tref) for c in comps]))
    A("__device__ constexpr int RMT_VISC_EQ[RMT_NC] = %s;" % _arr([c.visc_eq for c in comps], fmt=lambda v: str(int(v))))
    A("__device__ constexpr double RMT_VISC[RMT_NC][4] = %s;" % _arr2([c.visc for c in comps]))
    A("// stoichiometry nu[j][i], standard heats of reaction [J/mol], dCp_j(T) cubic")
    A("__device__ constexpr double RMT_NU[RMT_NR][RMT_NC] = %s;" % _arr2(spec.nu))
    A("__device__ constexpr double RMT_DH25[RMT_NR] = %s;" % _arr(spec.dH25))
    A("__device__ constexpr double RMT_DCP[RMT_NR][4] = %s;" % _arr2(spec.dcp))
    A("// constant-bank mirrors used as instruction operands (the constexpr copies drive compile-time structure)")
    A("__constant__ double RMT_cMW[RMT_NC] = %s;" % _arr([c.MW for c in comps]))
    A("__constant__ double RMT_cCP[RMT_NC][4] = %s;" % _arr2([c.cp for c in comps]))
    A("__constant__ double RMT_cCPREF[RMT_NC] = %s;" % _arr([c.cp_at(Tref) for c in comps]))
    A("__constant__ double RMT_cDH25[RMT_NR] = %s;" % _arr(spec.dH25))
    A("__constant__ double RMT_cDCP[RMT_NR][4] = %s;" % _arr2(spec.dcp))
    A("")

    # ---- rates -----------------------------------------------------------------
    decl_at = len(L)
    A("// traced RATES (rmtReaction.py:11-61 semantics): R[j] in mol/(m^3 s)")
    A("__device__ __forceinline__ void rmt_rates(const double T, const double P, const double (&y)[RMT_NC],")
    A("        const double (&C)[RMT_NC], const double* __restrict__ kp, double (&R)[RMT_NR])")
    A("{")
    pool = ConstPool()
    lines, refs = emit_dag(g, kin.rates, in_name, pool=pool)
    L.extend(lines)
    for j, r in enumerate(refs):
        A("    R[%d] = %s;" % (j, r))
    A("}")
    A("")
    A("// rates + partial derivatives (symbolic differentiation of the same DAG)")
    A("__device__ __forceinline__ void rmt_rates_jac(const double T, const double P, const double (&y)[RMT_NC],")
    A("        const double (&C)[RMT_NC], const double* __restrict__ kp, double (&R)[RMT_NR],")
    A("        double (&dRdT)[RMT_NR], double (&dRdP)[RMT_NR], double (&dRdy)[RMT_NR][RMT_NC], double (&dRdC)[RMT_NR][RMT_NC])")
    A("{")
    outs = list(kin.rates) + list(P["T"]) + list(P["P"])
    for i in range(nc):
        outs += list(P["y%d" % i])
    for i in range(nc):
        outs += list(P["C%d" % i])
    lines, refs = emit_dag(g, outs, in_name, pool=pool)
    L.extend(lines)
    it = iter(refs)
    for j in range(nr):
        A("    R[%d] = %s;" % (j, next(it)))
    for j in range(nr):
        A("    dRdT[%d] = RMT_DERIV(%s);" % (j, next(it)))
    for j in range(nr):
        A("    dRdP[%d] = RMT_DERIV(%s);" % (j, next(it)))
    for i in range(nc):
        for j in range(nr):
            A("    dRdy[%d][%d] = RMT_DERIV(%s);" % (j, i, next(it)))
    for i in range(nc):
        for j in range(nr):
            A("    dRdC[%d][%d] = RMT_DERIV(%s);" % (j, i, next(it)))
    A("}")
    A("")

    L.insert(decl_at, "// literal pool of the traced kinetics (constant bank operands)\n" + pool.declaration())

    # ---- Rosenbrock tableau ------------------------------------------------------
    tab = TABLEAUX[tableau]
    s = tab["stages"]

    def full(rows):
        return [[(rows[i][j] if j < len(rows[i]) else 0.0) for j in range(s)] for i in range(s)]
    A("// Rosenbrock tableau: %s (rmt_app_b200/tableau.py)" % tab["name"])
    A("#define RMT_ROS_S %d" % s)
    A("#define RMT_ROS_ORDER %d" % tab["order"])
    A("__device__ constexpr double RMT_ROS_GAMMA = %s;" % lit(tab["gamma"]))
    A("__device__ constexpr double RMT_ROS_A[RMT_ROS_S][RMT_ROS_S] = %s;" % _arr2(full(tab["a"])))
    A("__device__ constexpr double RMT_ROS_C[RMT_ROS_S][RMT_ROS_S] = %s;" % _arr2(full(tab["c"])))
    A("__device__ constexpr double RMT_ROS_M[RMT_ROS_S] = %s;" % _arr(tab["m"]))
    A("__device__ constexpr double RMT_ROS_E[RMT_ROS_S] = %s;" % _arr(tab["e"]))
    dense = tab["dense"] if tab.get("dense") else [[0.0]*s, [0.0]*s]
    from .tableau import new_function_flags
    newf = new_function_flags(tab)
    A("#define RMT_ROS_DENSE %d" % (1 if tab.get("dense") else 0))
    A("__device__ constexpr double RMT_ROS_D[2][RMT_ROS_S] = %s;" % _arr2(dense))
    A("// stage i evaluates f at a new argument (0: same argument as the previous stage, whose f is re-used)")
    A("#define RMT_ROS_REUSE %d" % (0 if all(newf) else 1))
    A("__device__ constexpr int RMT_ROS_NEWF[RMT_ROS_S] = %s;" % _arr(newf, fmt=lambda v: str(int(v))))
    A("__constant__ int RMT_cROS_NEWF[RMT_ROS_S] = %s;" % _arr(newf, fmt=lambda v: str(int(v))))
    A("__constant__ double RMT_cROS_A[RMT_ROS_S][RMT_ROS_S] = %s;" % _arr2(full(tab["a"])))
    A("__constant__ double RMT_cROS_C[RMT_ROS_S][RMT_ROS_S] = %s;" % _arr2(full(tab["c"])))
    A("__constant__ double RMT_cROS_M[RMT_ROS_S] = %s;" % _arr(tab["m"]))
    A("__constant__ double RMT_cROS_E[RMT_ROS_S] = %s;" % _arr(tab["e"]))
    A("__constant__ double RMT_cROS_D[2][RMT_ROS_S] = %s;" % _arr2(dense))
    A("")
    return "\n".join(L) + "\n"
