"""Sweep step-size-controller settings on a config-3 sample (GPU)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, torch
from rmt_app_b200 import engine

B = int(os.environ.get("B", 65536))
base = cases.methanol_readme_input()
sw = cases.config3_sweep(B)
cm = engine.compile_model(base)
ref = engine.n1_solve_ensemble(cm, base, sw, B, rtol=1e-10, atol=1e-13)
print("reference run: ok %d/%d, steps mean %.0f" % ((ref.status == 0).sum(), B, ref.stats[0].mean()))
R = ref.out[0]          # [n][B]

def run(tag, rtol=1e-3, atol=1e-6, ctrl=None, show_fail=False):
    torch.cuda.synchronize(); t0 = time.time()
    r = engine.n1_solve_ensemble(cm, base, sw, B, rtol=rtol, atol=atol, ctrl=ctrl)
    torch.cuda.synchronize(); dt = time.time() - t0
    ok = (r.status == 0) & (ref.status == 0)
    e = np.abs(r.out[0][:, ok] - R[:, ok])/np.abs(R[:, ok])
    emax = e.max(axis=0)
    st = r.stats
    fails = {int(k): int((r.status == k).sum()) for k in np.unique(r.status) if k != 0}
    print("%-34s acc %.1f rej %.1f max %d | err med %.2e p99 %.2e max %.2e | T err p99 %.2e | fails %s | %.0f ms" % (
        tag, st[0].mean(), st[1].mean(), (st[0] + st[1]).max(), np.median(emax), np.quantile(emax, 0.99), emax.max(),
        np.quantile(e[-1], 0.99), fails, dt*1e3))
    if show_fail and fails:
        idx = np.nonzero(r.status != 0)[0][:8]
        for i in idx:
            print("   fail inst %d status %d T0 %.3f P0 %.1f C0 %s steps %s" % (
                i, r.status[i], sw["temperature"][i], sw["pressure"][i], np.array2string(sw["concentration"][i], precision=4), st[:, i]))
    return r

run("library default", show_fail=True)
for h0f in (1.0, 0.3, 0.1, 0.03, 0.01):
    run("old ctrl h0f %.2f" % h0f, ctrl=[0.9, 5.0, 6.0, 1.0, 0.0, h0f], show_fail=True)
    run("PI b0.08 s0.8 h0f %.2f" % h0f, ctrl=[0.8, 5.0, 6.0, 1.0, 0.08, h0f], show_fail=True)
for beta in (0.06, 0.08, 0.10):
    for kappa in (1.0, 0.5):
        run("PI b%.2f s0.8 k%.2f h0f 0.1" % (beta, kappa), ctrl=[0.8, 5.0, 6.0, kappa, beta, 0.1])
for rt in (1e-2, 1e-4, 1e-5, 1e-6, 1e-8):
    run("rtol %g default ctrl" % rt, rtol=rt, atol=rt*1e-3, show_fail=True)
