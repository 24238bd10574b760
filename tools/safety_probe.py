"""Step-size controller safety factor: attempts, time and error against a converged solve (Ros4, default tolerance)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, torch
from rmt_app_b200 import engine
B = int(os.environ.get("B", 1 << 20))
sw = cases.config3_sweep(B)
base = cases.methanol_readme_input()
ws = engine.Workspace()
dsw = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in sw.items()}
R = engine.n1_solve_ensemble(engine.compile_model(base), base, sw, B, rtol=1e-10, atol=1e-13, workspace=engine.Workspace()).out[0].copy()
cm = engine.compile_model(base, method="ros4")
for ctrl in ([0.8, 5, 6, 1, 0, 0.03], [0.85, 5, 6, 1, 0, 0.03], [0.9, 5, 6, 1, 0, 0.03], [0.95, 5, 6, 1, 0, 0.03], [0.9, 5, 4, 1, 0, 0.03],
             [0.9, 5, 6, 1, 0, 0.02], [0.9, 3, 6, 1, 0, 0.03]):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        r = engine.n1_solve_ensemble(cm, base, dsw, B, rtol=1e-3, atol=1e-6, ctrl=ctrl, keep_on_device=True, workspace=ws)
        torch.cuda.synchronize(); dt = time.time() - t0
    out = r.out.cpu().numpy()[0]; st = r.stats.cpu().numpy(); ok = r.status.cpu().numpy() == 0
    e = (np.abs(out[:, ok] - R[:, ok])/np.abs(R[:, ok])).max(axis=0)
    print("%-28s %6.2f ms | acc %.1f rej %.2f | err med %.2e p99 %.2e max %.2e | fails %d" % (
        ctrl, dt*1e3, st[0].mean(), st[1].mean(), np.median(e), np.quantile(e, 0.99), e.max(), int((~ok).sum())), flush=True)
