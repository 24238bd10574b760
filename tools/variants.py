"""Time kernel variants (NVRTC -D options) on the config-3 workload."""
import os, sys, time, itertools
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, torch
from rmt_app_b200 import engine, capi

B = int(os.environ.get("B", 1 << 20))
base = cases.methanol_readme_input(); sw = cases.config3_sweep(B)
if os.environ.get("SORT"):      # experiment: reactors ordered by a stiffness proxy
    key = {"T": sw["temperature"], "P": sw["pressure"], "TP": sw["temperature"] + 1e-5*sw["pressure"]}[os.environ["SORT"]]
    o = np.argsort(key); sw = {k: v[o] for k, v in sw.items()}
base["solver-config"]["method"] = os.environ.get("METHOD", "rodas4")
cm = engine.compile_model(base)
# step-size controller: the tableau's tuned setting (what rmtExeBatch and bench.py use) unless CTRL overrides it
CTRL = [float(v) for v in os.environ["CTRL"].split(",")] if os.environ.get("CTRL") else engine.METHOD_CTRL.get(cm.method)
capi.init(0)
ws = engine.Workspace()
h_rows, n_rows, row_map = engine.sweep_rows_into(cm.spec, sw, B, ws)
uniform = engine.uniform_inputs(cm.spec, base)
d_rows = h_rows.cuda()
stream = torch.cuda.current_stream().cuda_stream
ref = None
variants = [v.split(",") for v in sys.argv[1:]] or [["128", "RMT_SYNC=0"], ["128", "RMT_SYNC=1"], ["128", "RMT_SYNC=2"],
                                                   ["256", "RMT_SYNC=1"], ["256", "RMT_SYNC=2"], ["64", "RMT_SYNC=1"]]
for v in variants:
    block, defs = int(v[0]), v[1:]
    t0 = time.time()
    cubin, log = capi.nvrtc_compile(cm.header, block=block, extra_opts=["-D" + d for d in defs])
    tc = time.time() - t0
    try:
        mod = capi.Module(cubin)
    except capi.RmtError as e:
        print(v, "load failed:", e); continue
    info = mod.info
    d_consts = torch.empty((info.nconst, B), dtype=torch.float64, device="cuda")
    d_out = torch.empty((1, info.n, B), dtype=torch.float64, device="cuda")
    d_status = torch.empty((B,), dtype=torch.int32, device="cuda"); d_stats = torch.empty((4, B), dtype=torch.int32, device="cuda")
    mod.setup(B, d_rows, n_rows, row_map, uniform, d_consts, stream=stream)
    z = np.array([1.0])
    for _ in range(2):
        mod.n1_solve(B, d_consts, z, 1e-3, 1e-6, d_out, d_status, d_stats, ctrl=CTRL, stream=stream)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        mod.n1_solve(B, d_consts, z, 1e-3, 1e-6, d_out, d_status, d_stats, ctrl=CTRL, stream=stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/3
    # stand-alone RHS / Jacobian kernels
    n = info.n
    d_y = torch.rand((n, B), dtype=torch.float64, device="cuda")*0.5 + 0.25
    d_y[n - 2] = 1.0; d_y[n - 1] = 0.1
    d_f = torch.empty((n, B), dtype=torch.float64, device="cuda"); d_J = torch.empty((n*n, B), dtype=torch.float64, device="cuda")
    tk = {}
    for nm, fn in (("rhs", lambda: mod.n1_rhs(B, d_consts, d_y, d_f, stream=stream)), ("jac", lambda: mod.n1_jac(B, d_consts, d_y, d_f, d_J, stream=stream))):
        for _ in range(3): fn()
        a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(20): fn()
        a1.record(); torch.cuda.synchronize()
        tk[nm] = a0.elapsed_time(a1)/20
    out = d_out.cpu().numpy(); ok = int((d_status == 0).sum())
    if ref is None:
        ref = out
    same = np.array_equal(out, ref)
    dev = np.nanmax(np.abs(out - ref)/np.abs(ref))
    print("%-40s %8.2f ms  %.2f Msolves/s  ok %d  identical %s maxdev %.1e  (compile %.0fs, att/solve %.1f) rhs %.4f ms (%.0f GB/s) jac %.4f ms (%.0f GB/s)" % (
        ",".join(v), ms, B/ms/1e3, ok, same, dev, tc, d_stats[3].double().mean().item(),
        tk["rhs"], 8e-6*(info.nconst + 2*n)*B/tk["rhs"], tk["jac"], 8e-6*(info.nconst + 2*n + n*n)*B/tk["jac"]))
    mod.close()
