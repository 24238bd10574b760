"""N2 ensemble: integrator tableau comparison (time per ensemble, steps, error against a tight solve)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, torch
from rmt_app_b200 import engine

B = int(os.environ.get("B", 12500)); Z = int(os.environ.get("Z", 200)); NE = int(os.environ.get("NE", 256))
mi = cases.methanol_readme_input("N2")
sw = cases.config3_sweep(B, 20240613)
swe = {k: v[:NE] for k, v in sw.items()}
ref = engine.n2_solve_ensemble(engine.compile_model(mi, block=engine.n2_block(NE)), mi, swe, NE, zNo=Z, tNo=5, period=0.5,
                               rtol=1e-8, atol=1e-11).out
for method in sys.argv[1:] or ["rodas4", "ros4"]:
    ctrl = engine.METHOD_CTRL.get(method)
    cm = engine.compile_model(mi, block=engine.n2_block(B), method=method)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        res = engine.n2_solve_ensemble(cm, mi, sw, B, zNo=Z, tNo=5, period=0.5, keep_on_device=True, ctrl=ctrl)
        torch.cuda.synchronize(); dt = time.time() - t0
    st = res.stats.cpu().numpy(); ok = int((res.status == 0).sum())
    cme = engine.compile_model(mi, block=engine.n2_block(NE), method=method)
    got = engine.n2_solve_ensemble(cme, mi, swe, NE, zNo=Z, tNo=5, period=0.5, ctrl=ctrl).out
    rel = np.abs(got - ref)/np.maximum(np.abs(ref), 1e-3*np.max(np.abs(ref), axis=(0, 2), keepdims=True))
    per = rel.reshape(-1, NE).max(axis=0)
    print("%-7s B=%d zNo=%d block %d: %.3fs ok %d  steps %.1f rej %.1f | err vs tight: median %.1e p99 %.1e max %.1e" % (
        method, B, Z, cm.block, dt, ok, st[0].mean(), st[1].mean(), np.median(per), np.percentile(per, 99), per.max()))
