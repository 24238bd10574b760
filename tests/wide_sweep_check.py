"""Robustness probe: a much wider operating box than config 3 (T0 450-650 K, P0 1-10 MPa, H2/COx 0.5-5, CO2/COx 0.05-0.95,
coolant offset +-30 K): failure counts, step statistics, and a subsample against the oracle at tight tolerance."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cases
import pyremot_oracle as O
from rmt_app_b200 import rmtExeBatch

B = int(os.environ.get("B", 1 << 18))
rng = np.random.default_rng(99)
T0 = rng.uniform(450, 650, B); P0 = rng.uniform(1e6, 1e7, B)
r = rng.uniform(0.5, 5.0, B); c = rng.uniform(0.05, 0.95, B)
ytr = 1e-5
yH2 = r/(1 + r); yCOx = 1/(1 + r)
y = np.stack([yH2*(1 - 3*ytr), c*yCOx*(1 - 3*ytr), np.full(B, ytr), (1 - c)*yCOx*(1 - 3*ytr), np.full(B, ytr), np.full(B, ytr)], axis=1)
y /= y.sum(axis=1, keepdims=True)
C0 = y*(P0/(8.314472*T0))[:, None]
sw = {"temperature": T0, "pressure": P0, "concentration": C0, "MeTe": T0 + rng.uniform(-30, 30, B)}
base = cases.methanol_readme_input("N1")
for rtol, atol in ((1e-3, 1e-6), (1e-6, 1e-9)):
    t0 = time.time(); res = rmtExeBatch(base, sw, rtol=rtol, atol=atol); dt = time.time() - t0
    st = res["stats"]
    print("rtol %g: %.3f s, status counts %s, accepted mean %.1f max %d, rejected mean %.2f, T_out range %.1f-%.1f" % (
        rtol, dt, dict(zip(*np.unique(res["status"], return_counts=True))), st[0].mean(), st[0].max(), st[1].mean(),
        np.nanmin(res["dataYs"][:, 7]), np.nanmax(res["dataYs"][:, 7])), flush=True)
tight = rmtExeBatch(base, sw, rtol=1e-9, atol=1e-12)
idx = rng.choice(B, 24, replace=False)
worst = 0.0
for i in idx:
    ref = O.rmtExe(cases.instance_input(base, sw, int(i)), method="LSODA", rtol=1e-10, atol=1e-12)["resModel"][0]["dataYs"][:, -1]
    worst = max(worst, float(np.max(np.abs(tight["dataYs"][i] - ref)/np.abs(ref))))
print("tight vs oracle on 24 random reactors: worst relative deviation %.2e; all tight converged: %s" % (worst, bool(tight["success"].all())))
d = np.abs(res["dataYs"] - tight["dataYs"])/np.abs(tight["dataYs"])
print("rtol 1e-6 vs 1e-9: median %.1e p99 %.1e max %.1e" % (np.median(d.max(axis=1)), np.percentile(d.max(axis=1), 99), d.max()))
