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
from .codegen import generate_model_header, model_flops
from .model import SCALAR_INPUTS, ModelSpec

# The reference's only "config system" for the path: module-level mutable dict
# PyREMOT/solvers/solSetting.py:30-39.  Same keys, same defaults, same usage
# (callers mutate it to change grid sizes).
solverSetting = {
    "N1": {"zNo": 100},
    "N2": {"zNo": 20, "rNo": 5, "tNo": 5, "timesNo": 5},
}

# SciPy defaults the reference inherits by never passing tolerances
# (pbHomoReactor.py:2931-2932; scipy/integrate/_ivp/ivp.py)
DEFAULT_RTOL, DEFAULT_ATOL = 1e-3, 1e-6

_lock = threading.Lock()
_compiled = {}


def _torch():
    import torch
    return torch


class CompiledModel:
    def __init__(self, spec, block):
        self.spec = spec
        self.block = block
        self.header = generate_model_header(spec)
        self.flops = model_flops(spec)
        self.module = None

    def load(self, device):
        if self.module is None:
            capi.init(device)
            cubin = capi.cached_cubin(self.header, block=self.block)
            self.module = capi.Module(cubin)
        return self.module


def default_block(spec):
    """Integrator block size: per-thread shared memory is (n^2 + s*n) doubles;
    pick the largest warp multiple <= 128 that lets two blocks share an SM."""
    if spec.model != "N1":
        return 64
    per_thread = 8*(spec.n*spec.n + 6*spec.n)
    for b in (128, 96, 64, 32):
        if 2*(b*per_thread + 1024) <= 227*1024:
            return b
    return 32


def compile_model(modelInput, block=None):
    """Trace + generate + (lazily) NVRTC-compile; cached per model structure."""
    spec = ModelSpec(modelInput)
    blk = block or default_block(spec)
    key = spec.key("b%d" % blk)
    with _lock:
        cm = _compiled.get(key)
        if cm is None:
            cm = CompiledModel(spec, blk)
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
             float(rs["BeVoFr"]), float(eh["OvHeTrCo"]), float(eh["MeTe"])]
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
# N1 ensemble
# ----------------------------------------------------------------------------------
class N1Result:
    """Arrays of one ensemble solve (host numpy unless `keep_on_device`)."""
    __slots__ = ("out", "status", "stats", "z_eval", "objective", "n", "nc", "out_mode", "flops", "seconds")


def n1_solve_ensemble(cm, modelInput, sweep=None, B=1, z_eval=None, rtol=None, atol=None, out_mode=1,
                      dense=True, max_steps=100000, objective_ref=None, device=None, keep_on_device=False,
                      pinned=None, ctrl=None):
    """Solve B independent steady-state reactors on the current CUDA device.

    Returns an N1Result with out[n_eval][rows][B].  Raises capi.RmtError when the
    CUDA library/driver is unavailable (no CPU path exists)."""
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
    rows, row_map = sweep_rows(spec, sweep, B)
    n, nc = spec.n, spec.nc
    out_rows = 2*n + nc if out_mode == 2 else n
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        if rows.shape[0]:
            h_rows = torch.from_numpy(rows)
            if pinned is not None:
                pinned[:rows.shape[0]].copy_(h_rows)
                h_rows = pinned[:rows.shape[0]]
            d_rows = h_rows.to(dev, non_blocking=True)
        else:
            d_rows = None
        d_consts = torch.empty((mod.info.nconst, B), dtype=torch.float64, device=dev)
        d_out = torch.empty((z_eval.size, out_rows, B), dtype=torch.float64, device=dev)
        d_status = torch.empty((B,), dtype=torch.int32, device=dev)
        d_stats = torch.empty((4, B), dtype=torch.int32, device=dev)
        d_obj = torch.empty((B,), dtype=torch.float64, device=dev) if objective_ref is not None else None
        mod.setup(B, d_rows, rows.shape[0], row_map, uniform, d_consts, stream=stream)
        mod.n1_solve(B, d_consts, z_eval, rtol, atol, d_out, d_status, d_stats, max_steps=max_steps, dense=dense,
                     out_mode=out_mode, obj_ref=objective_ref, d_obj=d_obj, ctrl=ctrl, stream=stream)
        res = N1Result()
        res.z_eval, res.n, res.nc, res.out_mode, res.flops = z_eval, n, nc, out_mode, cm.flops
        if keep_on_device:
            res.out, res.status, res.stats, res.objective = d_out, d_status, d_stats, d_obj
        else:
            res.out = d_out.cpu().numpy()
            res.status = d_status.cpu().numpy()
            res.stats = d_stats.cpu().numpy()
            res.objective = None if d_obj is None else d_obj.cpu().numpy()
    return res


def n1_rhs_batch(cm, modelInput, Y, sweep=None, jac=False, device=None):
    """modelEquationN1 (and optionally its Jacobian) at states Y [B][n]; every
    instance shares `modelInput` unless `sweep` varies inputs.  Returns
    (F [B][n], J [B][n][n] or None, consts [nconst][B])."""
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
        if jac:
            d_J = torch.empty((n*n, B), dtype=torch.float64, device=dev)
            mod.n1_jac(B, d_consts, d_y, d_f, d_J, stream=stream)
            J = d_J.cpu().numpy().T.reshape(B, n, n)
        else:
            mod.n1_rhs(B, d_consts, d_y, d_f, stream=stream)
            J = None
        return d_f.cpu().numpy().T.copy(), J, d_consts.cpu().numpy()
