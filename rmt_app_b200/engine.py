"""Host front-end: modelInput (+ optional per-instance sweep) -> device arrays
-> librmtb200 kernels -> result arrays.

Mirrors what `runN1` / `runN2` do around their `solve_ivp` call
(PyREMOT/docs/pbHomoReactor.py:2694-3015, :3319-3704) with the per-solve setup,
the integration and the un-scaling all executed on the GPU.  PyTorch is used
for device memory, pinned host staging and the current CUDA stream only.
"""
import threading

import numpy as np

from . import capi
from .codegen import generate_model_header, model_flops, system_size, use_extents
from .model import SCALAR_INPUTS, ModelSpec

# The reference's only "config system" for the path: module-level mutable dict
# PyREMOT/solvers/solSetting.py:30-39.  Same keys, same defaults, same usage
# (callers mutate it to change grid sizes).
solverSetting = {
    "N1": {"zNo": 100},
    "N2": {"zNo": 20, "rNo": 5, "tNo": 5, "timesNo": 5},
    "M9": {"zNo": 30},          # runM3 (model M7) takes its number of output points from here (pbReactor.py:1283)
    "S2": {"tNo": 10, "zNo": 100, "rNo": 7, "timesNo": 5},      # grid of the dynamic model M9 (runM5, pbReactor.py:2072)
}

# SciPy defaults the reference inherits by never passing tolerances
# (pbHomoReactor.py:2931-2932; scipy/integrate/_ivp/ivp.py)
DEFAULT_RTOL, DEFAULT_ATOL = 1e-3, 1e-6

_lock = threading.Lock()
_compiled = {}


def _torch():
    import torch
    return torch


class nvtx_range:
    """NVTX range around a host-side phase (trace / compile / H2D / setup / solve / D2H) — visible in Nsight Systems
    timelines; a no-op pair of calls (~0.2 us) when no profiler is attached or CUDA is absent."""
    __slots__ = ("name", "on")

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        try:
            _torch().cuda.nvtx.range_push("rmt:" + self.name)
            self.on = True
        except Exception:
            self.on = False
        return self

    def __exit__(self, *a):
        if self.on:
            try:
                _torch().cuda.nvtx.range_pop()
            except Exception:
                pass
        return False


EXACT_MATH_OPTS = ("-DRMT_EXACT_MATH=1", "-DRMT_EXACT_DIV=1")


class CompiledModel:
    def __init__(self, spec, block, method="rodas4", reduced=None, lanes=1, exact_math=False):
        self.spec = spec
        self.block = block
        self.method = method
        # IEEE special-value semantics (exp(-Inf) = 0, x/Inf = 0, NaN propagation) instead of the branch-free
        # device math: libdevice exp/log/sqrt/pow and IEEE division, at ~1.5x the integrator time
        self.exact_math = bool(exact_math)
        self.opts = EXACT_MATH_OPTS if self.exact_math else ()
        self.reduced = use_extents(spec) if reduced is None else bool(reduced)
        self.m = system_size(spec, self.reduced)          # unknowns of the integrator's linear systems
        self.lanes = int(lanes) if spec.model in ("N2", "M9") else 1     # dynamic models: threads per reactor
        self.header = generate_model_header(spec, tableau=method, reduced=self.reduced, lanes=self.lanes)
        self.flops = model_flops(spec, self.reduced)
        self.module = None

    def load(self, device):
        if self.module is None:
            capi.init(device)
            with nvtx_range("nvrtc_compile_or_cache"):
                cubin = capi.cached_cubin(self.header, block=self.block, extra_opts=self.opts)
            with nvtx_range("module_load"):
                self.module = capi.Module(cubin)
        return self.module

    def key(self):
        """Identity of the loaded cubin (see capi.cubin_key)."""
        return capi.cubin_key(self.header, block=self.block, extra_opts=self.opts)


def default_block(spec, stages=6, reduced=None):
    """Integrator block size.  Per-thread shared memory is (m^2 + s*m) doubles (LU + stage vectors).
    One block per SM, as many warps as fit: the warps of a block run in lockstep (one barrier per step
    attempt) so that they share instruction-cache lines — measured 2x faster than two independent
    128-thread blocks per SM.  256 threads leave ~250 registers per thread (no spills); when the linear
    systems are small enough for 384 threads (reaction-extent form) the 168-register build spills a few
    words but 12 warps hide more latency — measured 16.0 vs 18.4 ms per 2^20 config-3 reactors."""
    if spec.model in ("N2", "M9"):
        return 64
    m = system_size(spec, reduced)
    per_thread = 8*(m*m + stages*m)
    fit = (227*1024 - 1024)//per_thread
    if fit >= 384:
        return 384
    return int(max(32, min(256, (fit//32)*32)))


class _NoFastKey(Exception):
    """A value the cheap identity cannot describe: compile_model falls back to tracing (always correct)."""


def _value_sig(v, depth=0):
    """Hashable description of a value the tracer would bake into the graph as a constant."""
    import types
    if isinstance(v, (float, np.floating)):
        return (type(v).__name__, float(v).hex())            # -0.0 != 0.0, nan == nan
    if isinstance(v, (int, str, bool, complex, bytes, type(None))):
        return (type(v).__name__, v)
    if isinstance(v, (np.integer, np.bool_)):
        return (type(v).__name__, v.item())
    if isinstance(v, np.ndarray):
        if v.size > 4096 or v.dtype == object:
            raise _NoFastKey()
        return ("ndarray", v.shape, str(v.dtype), v.tobytes())
    if isinstance(v, (list, tuple)):
        if len(v) > 4096 or depth > 4:
            raise _NoFastKey()
        return (type(v).__name__, tuple(_value_sig(x, depth + 1) for x in v))
    if isinstance(v, dict):
        if len(v) > 4096 or depth > 4:
            raise _NoFastKey()
        return ("dict", tuple((repr(k), _value_sig(x, depth + 1)) for k, x in v.items()))
    if isinstance(v, types.ModuleType):
        return ("module", v.__name__)
    if isinstance(v, types.FunctionType):
        if depth > 4:
            raise _NoFastKey()
        return ("function", _fn_sig(v, depth + 1))
    if isinstance(v, (types.BuiltinFunctionType, np.ufunc, type)):
        return ("callable", getattr(v, "__module__", None), getattr(v, "__qualname__", getattr(v, "__name__", None)))
    raise _NoFastKey()


def _code_names(code):
    """Global / attribute names a code object (and the code objects nested in it) can look up."""
    import types
    names = set(code.co_names)
    for c in code.co_consts:
        if isinstance(c, types.CodeType):
            names |= _code_names(c)
    return names


def _fn_sig(f, depth=0):
    """Identity of a kinetics lambda for the compile cache: its code object AND everything the tracer would read
    through it and bake into the graph — closure cells, the module-level globals its code names, defaults.  A
    parameter-estimation loop that rebinds a global pre-exponential, or mutates an array captured in a closure,
    therefore gets a new key (the key used to hold only id(f.__globals__) and primitive cells)."""
    cells = []
    for c in (getattr(f, "__closure__", None) or ()):
        try:
            cells.append(_value_sig(c.cell_contents, depth + 1))
        except ValueError:                      # empty cell
            cells.append(("empty",))
    globs = []
    g = f.__globals__
    for name in sorted(_code_names(f.__code__)):
        if name in g:
            globs.append((name, _value_sig(g[name], depth + 1)))
    defaults = _value_sig(f.__defaults__, depth + 1) if f.__defaults__ else None
    kwdefaults = _value_sig(f.__kwdefaults__, depth + 1) if f.__kwdefaults__ else None
    return (f.__code__, tuple(cells), tuple(globs), defaults, kwdefaults)


def n2_lanes(B, zNo, sm_count=148, n=7):
    """Threads per reactor of the N2 integrator (a power of two <= 32; the nodes of a reactor are spread over them).
    The block forward substitution gives lane r the rows r, r + lanes, ... of the n + 1 rows of every node's update,
    so up to n + 1 lanes are busy in it; more lanes only pay while the ensemble is too small to fill the GPU with
    node evaluations.  Measured (12 500 x 200 nodes, n = 7): 0.195 / 0.160 / 0.202 s with 4 / 8 / 16 lanes;
    50 000 x 50 nodes: 0.150 / 0.144 s with 4 / 8; one 50-node reactor: 9.2 / 5.3 ms with 8 / 32 lanes."""
    cap = 1
    while cap < min(n + 1, 32):
        cap *= 2                                  # smallest power of two >= n + 1
    if B*32 <= 2*sm_count*256:                    # small ensemble: as many lanes as the grid has nodes for
        lanes = 32
    else:
        lanes = cap
        while lanes > 1 and B*lanes > 16*sm_count*256:
            lanes //= 2
    while lanes > 1 and lanes >= 2*zNo:
        lanes //= 2
    return lanes


# Stage-pipelined N2 kernel (lanes = 0, rmt_kernels.cu "stage pipeline"): a block serves 64 reactors — two warps per
# role, roles = Jacobian | stages 1-2 (+ LU) | stages 3-4 | stages 5-6 — one block per SM, so an ensemble is worked off in
# rounds of sm_count*64 = 9 472 reactors; the lanes kernel (8 lanes per reactor, 592 resident blocks of 8 reactors) has
# rounds of 4 736.  Which one is faster depends on how the ensemble fills those rounds, so the choice is made from the
# measured times of both (one B200, methanol model, 200 nodes, period 0.5 s; tools/n2_lanes.py, gpurun_out s15 sweep):
N2_PIPELINE_REACTORS_PER_BLOCK = 64
N2_PIPELINE_MIN_B = 2048
_N2_T_LANES = ((1024, .0335), (2048, .0368), (3072, .0435), (4096, .0466), (6144, .0789), (8192, .0831), (9472, .0940),
               (11000, .1147), (12500, .1215), (14000, .1278), (16000, .1524), (18944, .1674), (22000, .1979), (28416, .2452))
_N2_T_PIPE = ((1024, .0684), (2048, .0724), (4096, .0743), (6144, .0748), (8192, .0785), (9472, .0806), (11000, .1331),
              (14000, .1425), (16000, .1463), (18944, .1543), (22000, .2031), (28416, .2230))


def _interp(table, x):
    xs, ys = zip(*table)
    return float(np.interp(x, xs, ys))


def n2_kernel_costs(B, zNo, sm_count=148):
    """(lanes kernel, stage pipeline): estimated seconds for B reactors x zNo nodes, interpolated from the measured
    table above and continued beyond it (lanes: 8.63 us per reactor; pipeline: 74.5 ms per full round of 9 472 plus
    50 + 28*fill ms for a partly filled one), scaled with the node count.  Short grids favour the pipeline a little
    more than the scaling says (50 000 x 50 nodes: 0.100 vs 0.131 s measured)."""
    per_round = sm_count*N2_PIPELINE_REACTORS_PER_BLOCK
    if B <= _N2_T_LANES[-1][0]:
        tl, tp = _interp(_N2_T_LANES, B), _interp(_N2_T_PIPE, B)
    else:
        full, frac = divmod(B/per_round, 1.0)
        tl = B*8.63e-6
        tp = 0.0745*full + ((0.050 + 0.028*frac) if frac > 0 else 0.0)
    scale = zNo/200.0
    return tl*scale, tp*scale*(0.85 if zNo <= 100 else 1.0)


def n2_use_pipeline(B, zNo, sm_count=148):
    """Stage pipeline or lanes kernel for an N2 ensemble?  (M9: see compile_model_n2.)"""
    if B < N2_PIPELINE_MIN_B or zNo < 8:
        return False
    tl, tp = n2_kernel_costs(B, zNo, sm_count)
    return tp < tl


def n2_block(B, sm_count=148, lanes=1):
    """Threads per block of the N2 integrator (`lanes` threads per reactor, lockstep blocks).  The kernel keeps a
    78-row record per thread plus a hand-over record per reactor in shared memory ((n + 1) n + 3 n + 1 rows of
    block + 1 doubles; 1 kB per reactor) and needs 255 registers, so 256 threads are resident per SM whatever the block
    size; 64-thread blocks measured best (12 500 x 200 nodes, 8 lanes: 0.148 s with 64, 0.160 s with 128, 0.163 s with
    32) — a block barrier per node group then waits for two warps, not four.  Small ensembles take 32-thread blocks
    so that they spread over all SMs."""
    threads = B*lanes
    return 64 if threads >= sm_count*64*3//4 else 32


def compile_model_n2(modelInput, B, zNo, method=None):
    """compile_model with the launch shape (lanes per reactor, block size) for an ensemble of B reactors."""
    n_node = len(modelInput["feed"]["components"]["shell"]) + (0 if modelInput["operating-conditions"].get("process-type") == "iso-thermal" else 1)
    lanes = n2_lanes(B, zNo, n=n_node) if modelInput["model"] == "N2" else 1       # M9: the velocity march is sequential
    # M9: the lanes kernel has one lane per reactor, the pipeline four threads — taken from a few hundred reactors on
    if (modelInput["model"] == "N2" and n2_use_pipeline(B, zNo)) or (modelInput["model"] == "M9" and B >= 256 and zNo >= 8):
        from .tableau import TABLEAUX
        m = method or choose_method(modelInput, rtol=0.0)          # like compile_model: Rodas4 unless solver-config says otherwise
        from .tableau import new_function_flags
        if not all(new_function_flags(TABLEAUX[m])[1:]):            # a stage re-using f (Ros4): lanes kernel
            return compile_model(modelInput, block=n2_block(B, lanes=lanes), method=method, lanes=lanes)
        S = TABLEAUX[m]["stages"]
        if S % 2 == 0:                                              # two stages per role, two warps per role
            return compile_model(modelInput, block=32*(1 + S//2)*(N2_PIPELINE_REACTORS_PER_BLOCK//32), method=m, lanes=0)
    return compile_model(modelInput, block=n2_block(B, lanes=lanes), method=method, lanes=lanes)


# step-size controller per tableau {safety, max shrink, max growth, kappa, PI beta, initial-step factor}
# (None = the library default, tuned for Rodas4); tuned on 2^20 config-3 reactors (tools/method_compare.py,
# tools/safety_probe.py).  Ros4 with Hairer's standard safety factor 0.9: 43.2 accepted + 4.3 rejected steps per solve,
# median / p99 / max outlet error 7.3e-4 / 1.1e-3 / 1.4e-3 at rtol 1e-3 (safety 0.8: 48.4 + 3.3 steps, 3.8e-4 / 6.0e-4 /
# 9.2e-4, 7 % slower; 0.95: 11 rejections per solve, slower again).  The reference's LSODA at the same tolerances has a
# median error of 7e-4 on the 36 corner cases.
METHOD_CTRL = {"rodas4": None, "rodas3": None, "ros4": [0.9, 5.0, 6.0, 1.0, 0.0, 0.02]}


def choose_method(modelInput, rtol=None, n_eval=1, dense=True, method=None):
    """Integrator tableau.  `solver-config.method` (extension key): "rodas4" | "ros4" | "rodas3" | "auto".
    auto (default): Ros4 (4 stages, 3 RHS evaluations, order 4, no dense output) when only the end state is
    wanted at a loose tolerance (rtol >= 5e-4) — measured 15 % faster than Rodas4 there with a tighter error
    tail — and Rodas4(3) (6 stages, stiffly accurate, dense output) otherwise: below rtol ~3e-4 Ros4's order
    reduction on this stiff problem costs more steps than its cheaper step saves."""
    from .tableau import TABLEAUX
    m = method or modelInput.get("solver-config", {}).get("method", "auto")
    if m == "auto":
        rt = DEFAULT_RTOL if rtol is None else rtol
        m = "ros4" if (rt >= 5e-4 and (n_eval == 1 or not dense)) else "rodas4"
    if m not in TABLEAUX:
        raise ValueError("solver-config.method must be one of %s or \"auto\" (got %r)" % (sorted(TABLEAUX), m))
    return m


def _fast_key(modelInput, block):
    """Cheap identity of the model *structure* (no tracing): same components,
    reactions, process type and the same code objects in VARS/RATES."""
    import types
    rr = modelInput["reaction-rates"]
    sig = []
    from .kinetics import _is_scalar_param
    for k, v in rr["VARS"].items():
        if isinstance(v, types.FunctionType):
            sig.append((k, _fn_sig(v)))
        elif _is_scalar_param(v):
            sig.append((k, "param"))            # a kinetic-parameter slot: its value is a run-time input, not baked
        else:
            sig.append((k, _value_sig(v)))      # arrays / lists / bools / strings are baked into the graph
    for k, v in rr["RATES"].items():
        sig.append((k, _fn_sig(v)) if isinstance(v, types.FunctionType) else (k, _value_sig(v)))
    return (modelInput["model"], tuple(modelInput["feed"]["components"]["shell"]),
            modelInput["operating-conditions"].get("process-type"), tuple(modelInput["reactions"].values()),
            tuple(sig), block)


_fast = {}


def compile_model(modelInput, block=None, method=None, reduced=None, lanes=1, exact_math=None):
    """Trace + generate + (lazily) NVRTC-compile; cached per model structure and integrator tableau.
    `method` None resolves solver-config.method for a dense-output solve (Rodas4 unless stated);
    `reduced` None integrates in reaction extents whenever nr < nc (codegen.use_extents);
    `exact_math` None reads solver-config["exact-math"] (extension key, default False): libdevice math and IEEE
    division with their special-value semantics instead of the branch-free device versions."""
    if method is None:
        method = choose_method(modelInput, rtol=0.0)
    if exact_math is None:
        exact_math = bool(modelInput.get("solver-config", {}).get("exact-math", False))
    try:
        fk = _fast_key(modelInput, block) + (method, reduced, lanes, exact_math)
        cm = _fast.get(fk)
        if cm is not None:
            return cm
    except Exception:
        fk = None
    with nvtx_range("trace_and_codegen"):
        cm = _compile_model(modelInput, block, method, reduced, lanes, exact_math)
    if fk is not None:
        if len(_fast) > 256:
            _fast.clear()
        _fast[fk] = cm
    return cm


def _compile_model(modelInput, block, method, reduced=None, lanes=1, exact_math=False):
    from .tableau import TABLEAUX
    spec = ModelSpec(modelInput)
    if reduced is None:
        reduced = use_extents(spec)
    blk = block or default_block(spec, TABLEAUX[method]["stages"], reduced)
    key = spec.key("b%d%s%sg%d%s" % (blk, method, "x" if reduced else "", lanes, "e" if exact_math else ""))
    with _lock:
        cm = _compiled.get(key)
        if cm is None:
            cm = CompiledModel(spec, blk, method, reduced, lanes, exact_math)
            _compiled[key] = cm
    return cm


# ----------------------------------------------------------------------------------
# inputs
# ----------------------------------------------------------------------------------
def uniform_inputs(spec, modelInput):
    """The nin scalars of one reactor in device row order (rmt_b200.h, rmt_setup)."""
    oc, feed, rs, eh = (modelInput["operating-conditions"], modelInput["feed"], modelInput["reactor"],
                        modelInput["external-heat"])
    conc = np.asarray(feed["concentration"], dtype=np.float64).ravel()
    if conc.size != spec.nc:
        raise ValueError("feed.concentration has %d entries for %d components" % (conc.size, spec.nc))
    vals = [float(oc["temperature"]), float(oc["pressure"])] + [float(c) for c in conc]
    vals += [float(feed["volumetric-flowrate"]), float(rs["ReInDi"]), float(rs["ReLe"]), float(rs["PaDi"]),
             float(rs["BeVoFr"]), float(eh["OvHeTrCo"]), float(eh["MeTe"]),
             float(feed.get("mixture-viscosity", 0.0) or 0.0),      # read by M7/M9 only (pbReactor.py:1235, :2069)
             float(eh.get("EfHeTrAr", 0.0) or 0.0),                 # used by M7/M9 only; N1/N2 overwrite it with 4/ReInDi
             float(rs.get("CaDe", 0.0) or 0.0),                     # catalyst density and heat capacity: M9's energy
             float(rs.get("CaSpHeCa", 0.0) or 0.0)]                 # balance only (pbReactor.py:2619)
    # scalar VARS entries of THIS modelInput (the compiled model is shared by every input with the same structure)
    varis = modelInput["reaction-rates"]["VARS"]
    vals += [float(varis[name]) for name in spec.kin.param_names]
    return np.array(vals, dtype=np.float64)


def sweep_rows(spec, sweep, B):
    """sweep dict -> (rows [n_rows, B] float64, row_map [nin] int32)."""
    nin = spec.nin
    row_map = -np.ones(nin, dtype=np.int32)
    rows = []
    scalar_index = {name: 2 + spec.nc + k for k, name in enumerate(SCALAR_INPUTS[2:])}
    scalar_index["temperature"], scalar_index["pressure"] = 0, 1
    kp_index = {name: 2 + spec.nc + len(SCALAR_INPUTS) - 2 + k for k, name in enumerate(spec.kin.param_names)}
    for key, val in (sweep or {}).items():
        a = np.asarray(val, dtype=np.float64)
        if key == "concentration":
            if a.shape != (B, spec.nc):
                raise ValueError("sweep['concentration'] must have shape (B, nc) = (%d, %d)" % (B, spec.nc))
            for i in range(spec.nc):
                row_map[2 + i] = len(rows)
                rows.append(a[:, i])
            continue
        if a.shape != (B,):
            raise ValueError("sweep[%r] must have shape (B,) = (%d,)" % (key, B))
        if key in scalar_index:
            q = scalar_index[key]
        elif key in kp_index:
            q = kp_index[key]
        else:
            raise KeyError("sweep key %r is neither an operating/feed/reactor input %r nor a scalar VARS entry %r"
                           % (key, ["temperature", "pressure", "concentration"] + list(SCALAR_INPUTS[2:]),
                              spec.kin.param_names))
        row_map[q] = len(rows)
        rows.append(a)
    R = np.ascontiguousarray(np.stack(rows, axis=0)) if rows else np.zeros((0, B))
    return R, row_map


# ----------------------------------------------------------------------------------
# buffers
# ----------------------------------------------------------------------------------
class Workspace:
    """Grow-only pinned-host and device buffers reused across ensemble calls.

    Results returned from a call that was given a workspace are views into its
    pinned buffers and stay valid until the next call with the same workspace."""

    def __init__(self):
        self._bufs = {}
        self._pipe = None

    def pipeline(self, device):
        """Two (stream, private Workspace) pairs for the copy/compute pipeline of large host-side ensembles:
        chunk c runs on pair c % 2, so consecutive chunks never share device or staging buffers."""
        torch = _torch()
        if self._pipe is None or self._pipe[0] != str(device):
            self._pipe = (str(device), [(torch.cuda.Stream(device=device), Workspace()) for _ in range(2)])
        return self._pipe[1]

    def get(self, name, shape, dtype, device=None, pinned=False):
        torch = _torch()
        need = int(np.prod(shape)) if len(shape) else 1
        key = (name, str(dtype), str(device), pinned)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < need:
            cap = max(need, 1)
            if device is not None:
                buf = torch.empty((cap,), dtype=dtype, device=device)
            else:
                buf = torch.empty((cap,), dtype=dtype, pin_memory=pinned)
            self._bufs[key] = buf
        return buf[:need].view(*shape)


def _sweep_plan(spec, sweep, B):
    """Map sweep keys to device input rows: [(value, column or None)], row_map."""
    nin = spec.nin
    row_map = -np.ones(nin, dtype=np.int32)
    scalar_index = {name: 2 + spec.nc + k for k, name in enumerate(SCALAR_INPUTS[2:])}
    scalar_index["temperature"], scalar_index["pressure"] = 0, 1
    kp_index = {name: 2 + spec.nc + len(SCALAR_INPUTS) - 2 + k for k, name in enumerate(spec.kin.param_names)}
    plan = []
    for key, val in (sweep or {}).items():
        if key == "concentration":
            if tuple(val.shape) != (B, spec.nc):
                raise ValueError("sweep['concentration'] must have shape (B, nc) = (%d, %d)" % (B, spec.nc))
            for i in range(spec.nc):
                row_map[2 + i] = len(plan)
                plan.append((val, i))
            continue
        if tuple(np.shape(val)) != (B,):
            raise ValueError("sweep[%r] must have shape (B,) = (%d,)" % (key, B))
        if key in scalar_index:
            q = scalar_index[key]
        elif key in kp_index:
            q = kp_index[key]
        else:
            raise KeyError("sweep key %r is neither an operating/feed/reactor input %r nor a scalar VARS entry %r"
                           % (key, ["temperature", "pressure", "concentration"] + list(SCALAR_INPUTS[2:]),
                              spec.kin.param_names))
        row_map[q] = len(plan)
        plan.append((val, None))
    return plan, row_map


def sweep_rows_into(spec, sweep, B, ws):
    """Like sweep_rows but writes straight into a pinned staging buffer."""
    torch = _torch()
    plan, row_map = _sweep_plan(spec, sweep, B)
    n_rows = len(plan)
    if n_rows == 0:
        return None, 0, row_map
    stage = ws.get("h_rows", (n_rows, B), torch.float64, pinned=True)
    view = stage.numpy()
    for r, (val, col) in enumerate(plan):
        if torch.is_tensor(val):
            stage[r].copy_(val if col is None else val[:, col])
        else:
            a = np.asarray(val)
            np.copyto(view[r], a if col is None else a[:, col], casting="same_kind")
    return stage, n_rows, row_map


def sweep_rows_to_device(spec, sweep, B, ws, dev):
    """Inputs -> device rows [n_rows][B].  NumPy arrays are staged through one pinned buffer; torch
    tensors (pinned host memory or already on the device) are copied directly, a [B, nc] concentration
    block in one transfer followed by a device-side transpose.  Returns (d_rows, n_rows, row_map, h2d_bytes)."""
    torch = _torch()
    plan, row_map = _sweep_plan(spec, sweep, B)
    n_rows = len(plan)
    if n_rows == 0:
        return None, 0, row_map, 0
    d_rows = ws.get("d_rows", (n_rows, B), torch.float64, device=dev)
    h2d = 0
    np_rows = [r for r, (val, _) in enumerate(plan) if not torch.is_tensor(val)]
    if np_rows:
        stage = ws.get("h_rows", (len(np_rows), B), torch.float64, pinned=True)
        view = stage.numpy()
        for k, r in enumerate(np_rows):
            val, col = plan[r]
            a = np.asarray(val)
            np.copyto(view[k], a if col is None else a[:, col], casting="same_kind")
        if np_rows == list(range(np_rows[0], np_rows[0] + len(np_rows))):
            d_rows[np_rows[0]:np_rows[0] + len(np_rows)].copy_(stage, non_blocking=True)
        else:
            for k, r in enumerate(np_rows):
                d_rows[r].copy_(stage[k], non_blocking=True)
        h2d += stage.numel()*8
    done = set()
    for r, (val, col) in enumerate(plan):
        if not torch.is_tensor(val) or id(val) in done:
            continue
        if col is None:
            d_rows[r].copy_(val.to(torch.float64) if val.dtype != torch.float64 else val, non_blocking=True)
            h2d += 0 if val.is_cuda else B*8
        else:
            done.add(id(val))
            blk = val if val.is_cuda else ws.get("d_conc", (B, spec.nc), torch.float64, device=dev)
            if not val.is_cuda:
                blk.copy_(val, non_blocking=True)
                h2d += val.numel()*8
            d_rows[r - col:r - col + spec.nc].copy_(blk.t())
    return d_rows, n_rows, row_map, h2d


# ----------------------------------------------------------------------------------
# N1 ensemble
# ----------------------------------------------------------------------------------
class N1Result:
    """Arrays of one ensemble solve (host numpy unless `keep_on_device`)."""
    __slots__ = ("out", "status", "stats", "z_eval", "objective", "n", "nc", "out_mode", "flops", "consts",
                 "h2d_bytes", "d2h_bytes")


# ensembles at least this large with host-side inputs are solved as a copy/compute pipeline
PIPELINE_MIN_B = 1 << 18
# fractions of the ensemble per pipeline chunk: a small first chunk starts the GPU early, a small last one
# leaves little to copy back after the last kernel; every extra launch costs a kernel tail (0.3-0.6 ms).
# Measured on 2^20 config-3 reactors, pinned inputs (tools/pipe_probe.py): one launch 19.04 ms, halves 18.78,
# (1/4, 1/2, 1/4) 18.48, (1/8, 3/4, 1/8) 17.89, eight equal chunks 20.46; the kernel alone takes 16.25 ms.
PIPELINE_SPLIT = (0.125, 0.75, 0.125)


def pipeline_cuts(B, split=None):
    """Chunk boundaries [0, ..., B] of the copy/compute pipeline: fractions `split` of the ensemble, rounded to
    multiples of 1024 reactors, every chunk non-empty."""
    split = PIPELINE_SPLIT if split is None else split
    cuts = [0]
    for f in split[:-1]:
        nxt = min(B, cuts[-1] + max(1024, int(round(f*B/1024))*1024))
        if nxt > cuts[-1]:
            cuts.append(nxt)
    if cuts[-1] < B:
        cuts.append(B)
    return cuts


def _slice_sweep(sweep, b0, b1):
    return {k: v[b0:b1] for k, v in (sweep or {}).items()}


def _host_side(sweep):
    torch = _torch()
    return all(not (torch.is_tensor(v) and v.is_cuda) for v in (sweep or {}).values())


def n1_solve_ensemble(cm, modelInput, sweep=None, B=1, z_eval=None, rtol=None, atol=None, out_mode=1,
                      dense=True, max_steps=100000, objective_ref=None, device=None, keep_on_device=False,
                      workspace=None, ctrl=None, want_stats=True, pipeline=None):
    """Solve B independent steady-state reactors on the current CUDA device.

    Per call: stage the varying inputs in pinned memory -> H2D -> `rmt_setup`
    (per-reactor constants) -> `rmt_n1_solve` -> D2H.  Large ensembles whose inputs and outputs
    live on the host (`pipeline` None: B >= PIPELINE_MIN_B) are cut into three chunks on two streams,
    so that the copies of one chunk run under the integrator kernel of another; the result does not
    depend on the chunking (every reactor is an independent solve).  Returns an N1Result with
    out[n_eval][rows][B].  Raises capi.RmtError when the CUDA library/driver is unavailable (no CPU
    path exists)."""
    torch = _torch()
    if not torch.cuda.is_available():
        raise capi.RmtError("rmt_app_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    spec = cm.spec
    mod = cm.load(dev.index)
    sc = modelInput.get("solver-config", {})
    rtol = float(sc.get("rtol", DEFAULT_RTOL) if rtol is None else rtol)
    atol = float(sc.get("atol", DEFAULT_ATOL) if atol is None else atol)
    if z_eval is None:
        z_eval = np.array([1.0])
    z_eval = np.ascontiguousarray(z_eval, dtype=np.float64)
    uniform = uniform_inputs(spec, modelInput)
    ws = workspace if workspace is not None else Workspace()
    n, nc = spec.n, spec.nc
    out_rows = 2*n + nc if out_mode == 2 else n
    if ctrl is None:
        ctrl = METHOD_CTRL.get(cm.method)
    res = N1Result()
    res.z_eval, res.n, res.nc, res.out_mode, res.flops = z_eval, n, nc, out_mode, cm.flops
    if pipeline is None:
        pipeline = B >= PIPELINE_MIN_B
    pipeline = bool(pipeline) and not keep_on_device and bool(sweep) and _host_side(sweep) and B >= 3*1024

    def device_part(sub, Bc, w):
        """H2D + setup + integrator for one (sub-)ensemble on the current stream; device tensors."""
        stream = torch.cuda.current_stream().cuda_stream
        with nvtx_range("h2d"):
            d_rows, n_rows, row_map, h2d = sweep_rows_to_device(spec, sub, Bc, w, dev)
        d_consts = w.get("d_consts", (mod.info.nconst, Bc), torch.float64, device=dev)
        d_out = w.get("d_out", (z_eval.size, out_rows, Bc), torch.float64, device=dev)
        d_status = w.get("d_status", (Bc,), torch.int32, device=dev)
        d_stats = w.get("d_stats", (4, Bc), torch.int32, device=dev)
        d_obj = w.get("d_obj", (Bc,), torch.float64, device=dev) if objective_ref is not None else None
        with nvtx_range("setup"):
            mod.setup(Bc, d_rows, n_rows, row_map, uniform, d_consts, stream=stream)
        with nvtx_range("n1_solve"):
            mod.n1_solve(Bc, d_consts, z_eval, rtol, atol, d_out, d_status, d_stats, max_steps=max_steps, dense=dense,
                         out_mode=out_mode, obj_ref=objective_ref, d_obj=d_obj, ctrl=ctrl, stream=stream)
        return d_consts, d_out, d_status, d_stats, d_obj, h2d

    with torch.cuda.device(dev):
        if keep_on_device:
            res.consts, res.out, res.status, res.stats, res.objective, res.h2d_bytes = device_part(sweep, B, ws)
            res.d2h_bytes = 0
            return res
        h_out = ws.get("h_out", (z_eval.size, out_rows, B), torch.float64, pinned=True)
        h_status = ws.get("h_status", (B,), torch.int32, pinned=True)
        h_stats = ws.get("h_stats", (4, B), torch.int32, pinned=True) if want_stats else None
        h_obj = ws.get("h_obj", (B,), torch.float64, pinned=True) if objective_ref is not None else None
        res.d2h_bytes = h_out.numel()*8 + h_status.numel()*4 + (h_stats.numel()*4 if want_stats else 0) \
            + (h_obj.numel()*8 if h_obj is not None else 0)
        if not pipeline:
            res.consts, d_out, d_status, d_stats, d_obj, res.h2d_bytes = device_part(sweep, B, ws)
            with nvtx_range("d2h"):
                h_out.copy_(d_out, non_blocking=True)
                h_status.copy_(d_status, non_blocking=True)
                if want_stats:
                    h_stats.copy_(d_stats, non_blocking=True)
                if d_obj is not None:
                    h_obj.copy_(d_obj, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        else:
            # chunk c runs on stream c % 2 with that stream's own device buffers: H2D(c+1) and D2H(c-1) overlap
            # the integrator kernel of chunk c, and the blocks of the next kernel fill the tail of this one
            pipe = ws.pipeline(dev)
            cuts = pipeline_cuts(B)
            ready = torch.cuda.Event()
            ready.record()
            res.h2d_bytes = 0
            res.consts = None
            staged = {}
            for c in range(len(cuts) - 1):
                b0, b1 = cuts[c], cuts[c + 1]
                st, w = pipe[c % 2]
                if c - 2 in staged:
                    staged[c - 2].synchronize()       # the pinned staging buffer of this stream is free again
                st.wait_event(ready)
                with torch.cuda.stream(st):
                    _, d_out, d_status, d_stats, d_obj, h2d = device_part(_slice_sweep(sweep, b0, b1), b1 - b0, w)
                    staged[c] = torch.cuda.Event()
                    staged[c].record()
                    res.h2d_bytes += h2d
                    for e in range(z_eval.size):
                        for r in range(out_rows):
                            h_out[e, r, b0:b1].copy_(d_out[e, r], non_blocking=True)
                    h_status[b0:b1].copy_(d_status, non_blocking=True)
                    if want_stats:
                        for r in range(4):
                            h_stats[r, b0:b1].copy_(d_stats[r], non_blocking=True)
                    if d_obj is not None:
                        h_obj[b0:b1].copy_(d_obj, non_blocking=True)
            for st, _ in pipe:
                st.synchronize()
        res.out, res.status = h_out.numpy(), h_status.numpy()
        res.stats = None if h_stats is None else h_stats.numpy()
        res.objective = None if h_obj is None else h_obj.numpy()
    return res


def n1_rhs_batch(cm, modelInput, Y, sweep=None, jac=False, device=None, system=False):
    """modelEquationN1 (and optionally its Jacobian) at states Y [B][n]; every
    instance shares `modelInput` unless `sweep` varies inputs.  Returns
    (F [B][n], J [B][n][n] or None, consts [nconst][B]).  `system=True` returns the integrator's
    own form instead (rmt_n1_sys): (g [B][m], A [B][m][m], consts)."""
    torch = _torch()
    if not torch.cuda.is_available():
        raise capi.RmtError("rmt_app_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    spec = cm.spec
    mod = cm.load(dev.index)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    B, n = Y.shape
    assert n == spec.n
    uniform = uniform_inputs(spec, modelInput)
    rows, row_map = sweep_rows(spec, sweep, B)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        d_rows = torch.from_numpy(rows).to(dev) if rows.shape[0] else None
        d_consts = torch.empty((mod.info.nconst, B), dtype=torch.float64, device=dev)
        d_y = torch.from_numpy(np.ascontiguousarray(Y.T)).to(dev)
        d_f = torch.empty((n, B), dtype=torch.float64, device=dev)
        mod.setup(B, d_rows, rows.shape[0], row_map, uniform, d_consts, stream=stream)
        if system:
            m = mod.info.m
            d_g = torch.empty((m, B), dtype=torch.float64, device=dev)
            d_A = torch.empty((m*m, B), dtype=torch.float64, device=dev)
            mod.n1_sys(B, d_consts, d_y, d_g, d_A, stream=stream)
            return d_g.cpu().numpy().T.copy(), d_A.cpu().numpy().T.reshape(B, m, m), d_consts.cpu().numpy()
        if jac:
            d_J = torch.empty((n*n, B), dtype=torch.float64, device=dev)
            mod.n1_jac(B, d_consts, d_y, d_f, d_J, stream=stream)
            J = d_J.cpu().numpy().T.reshape(B, n, n)
        else:
            mod.n1_rhs(B, d_consts, d_y, d_f, stream=stream)
            J = None
        return d_f.cpu().numpy().T.copy(), J, d_consts.cpu().numpy()


# ----------------------------------------------------------------------------------
# N2 ensemble (dynamic model, method of lines)
# ----------------------------------------------------------------------------------
class N2Result:
    __slots__ = ("out", "status", "stats", "zNo", "tNo", "n", "nc", "out_mode", "flops", "h2d_bytes", "d2h_bytes")


def n2_solve_ensemble(cm, modelInput, sweep=None, B=1, zNo=None, tNo=None, period=None, rtol=None, atol=None,
                      out_mode=1, max_steps=1000000, device=None, keep_on_device=False, workspace=None, ctrl=None):
    """Integrate B independent dynamic reactors over [0, period]; returns the state at the end of
    each of the tNo slabs, out[tNo][rows][zNo][B] (runN2's dataPack list, pbHomoReactor.py:3589-3696)."""
    torch = _torch()
    if not torch.cuda.is_available():
        raise capi.RmtError("rmt_app_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    spec = cm.spec
    assert spec.model in ("N2", "M9")
    mod = cm.load(dev.index)
    sc = modelInput.get("solver-config", {})
    rtol = float(sc.get("rtol", DEFAULT_RTOL) if rtol is None else rtol)
    atol = float(sc.get("atol", DEFAULT_ATOL) if atol is None else atol)
    grid = solverSetting["N2" if spec.model == "N2" else "S2"]       # runN2 :3436-3440 / runM5 :2072, :2145
    zNo = int(grid["zNo"] if zNo is None else zNo)
    tNo = int(grid["tNo"] if tNo is None else tNo)
    period = float(modelInput["operating-conditions"]["period"] if period is None else period)
    uniform = uniform_inputs(spec, modelInput)
    ws = workspace if workspace is not None else Workspace()
    n, nc = spec.n, spec.nc
    out_rows = 2*n + nc if out_mode == 2 else n
    res = N2Result()
    res.zNo, res.tNo, res.n, res.nc, res.out_mode, res.flops = zNo, tNo, n, nc, out_mode, cm.flops
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        d_rows, n_rows, row_map, res.h2d_bytes = sweep_rows_to_device(spec, sweep, B, ws, dev)
        d_consts = ws.get("d_consts", (mod.info.nconst, B), torch.float64, device=dev)
        d_out = ws.get("d_out", (tNo, out_rows, zNo, B), torch.float64, device=dev)
        d_status = ws.get("d_status", (B,), torch.int32, device=dev)
        d_stats = ws.get("d_stats", (4, B), torch.int32, device=dev)
        d_work = ws.get("d_work", (mod.n2_work_doubles(B, zNo),), torch.float64, device=dev)
        with nvtx_range("setup"):
            mod.setup(B, d_rows, n_rows, row_map, uniform, d_consts, stream=stream)
        with nvtx_range("n2_solve"):
            mod.n2_solve(B, zNo, tNo, period, d_consts, rtol, atol, d_out, d_status, d_stats, d_work,
                         max_steps=max_steps, out_mode=out_mode, ctrl=ctrl, stream=stream)
        if keep_on_device:
            res.out, res.status, res.stats = d_out, d_status, d_stats
            res.d2h_bytes = 0
        else:
            with nvtx_range("d2h"):
                res.out = d_out.cpu().numpy()
                res.status = d_status.cpu().numpy()
                res.stats = d_stats.cpu().numpy()
            res.d2h_bytes = res.out.nbytes + res.status.nbytes + res.stats.nbytes
    return res


def n2_rhs_batch(cm, modelInput, Y, zNo, sweep=None, device=None):
    """modelEquationN2 at states Y [B][n*zNo] (variable-major per instance, like the reference's
    flattened state).  Returns F [B][n*zNo]."""
    torch = _torch()
    if not torch.cuda.is_available():
        raise capi.RmtError("rmt_app_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    spec = cm.spec
    mod = cm.load(dev.index)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    B = Y.shape[0]
    assert Y.shape[1] == spec.n*zNo
    uniform = uniform_inputs(spec, modelInput)
    rows, row_map = sweep_rows(spec, sweep, B)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        d_rows = torch.from_numpy(rows).to(dev) if rows.shape[0] else None
        d_consts = torch.empty((mod.info.nconst, B), dtype=torch.float64, device=dev)
        d_y = torch.from_numpy(np.ascontiguousarray(Y.T)).to(dev)            # [n*zNo][B]
        d_f = torch.empty_like(d_y)
        mod.setup(B, d_rows, rows.shape[0], row_map, uniform, d_consts, stream=stream)
        mod.n2_rhs(B, zNo, d_consts, d_y, d_f, stream=stream)
        return d_f.cpu().numpy().T.copy()
