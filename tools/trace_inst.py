import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, torch
from rmt_app_b200 import engine, capi
B = 65536
base = cases.methanol_readme_input(); sw = cases.config3_sweep(B)
cm = engine.compile_model(base)
for inst in [int(a) for a in sys.argv[1:]]:
    cm.load(0)
    tr = torch.zeros((400, 4), dtype=torch.float64, device="cuda")
    capi.debug_trace(tr, 400, inst)
    r = engine.n1_solve_ensemble(cm, base, sw, B)
    capi.debug_trace(None)
    t = tr.cpu().numpy()
    n = int(r.stats[0, inst] + r.stats[1, inst])
    print("inst", inst, "status", r.status[inst], "stats", r.stats[:, inst], "out", r.out[0][:, inst])
    for k in range(min(n + 1, 400)):
        print("  %3d t=%.6f h=%.3e err=%.3e %s" % (k, t[k, 0], t[k, 1], t[k, 2], "ACC" if t[k, 3] else "rej"))
