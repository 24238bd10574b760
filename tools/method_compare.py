import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, torch
from rmt_app_b200 import engine
B = int(os.environ.get("B", 1 << 20))
sw = cases.config3_sweep(B)
base4 = cases.methanol_readme_input(); cm4 = engine.compile_model(base4)
ws = engine.Workspace()
ref = engine.n1_solve_ensemble(cm4, base4, sw, B, rtol=1e-10, atol=1e-13, workspace=engine.Workspace())
R = ref.out[0].copy()
def run(tag, method, rtol=1e-3, atol=1e-6, ctrl=None):
    mi = cases.methanol_readme_input(); mi["solver-config"]["method"] = method
    cm = engine.compile_model(mi)
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        r = engine.n1_solve_ensemble(cm, mi, sw, B, rtol=rtol, atol=atol, ctrl=ctrl, keep_on_device=True, workspace=ws)
        torch.cuda.synchronize(); dt = time.time() - t0
    out = r.out.cpu().numpy()[0]; st = r.stats.cpu().numpy(); status = r.status.cpu().numpy()
    ok = status == 0
    e = (np.abs(out[:, ok] - R[:, ok])/np.abs(R[:, ok])).max(axis=0)
    print("%-36s %7.1f ms | acc %.1f rej %.1f | err med %.2e p99 %.2e max %.2e | fails %d" % (
        tag, dt*1e3, st[0].mean(), st[1].mean(), np.median(e), np.quantile(e, 0.99), e.max(), int((~ok).sum())))
run("rodas4 default", "rodas4")
run("rodas3 default ctrl", "rodas3")
for beta in (0.0, 0.06, 0.1):
    for safe in (0.8, 0.9):
        for kappa in (1.0, 0.5, 0.25):
            run("rodas3 b%.2f s%.1f k%.2f" % (beta, safe, kappa), "rodas3", ctrl=[safe, 5.0, 6.0, kappa, beta, 0.1])
run("rodas4 rtol1e-6", "rodas4", 1e-6, 1e-9); run("rodas3 rtol1e-6", "rodas3", 1e-6, 1e-9)
