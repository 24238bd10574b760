"""e2e time of rmtExeBatch-style solves for different pipeline splits (pinned inputs, 2^20 config-3 reactors)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, torch
from rmt_app_b200 import engine

B = 1 << 20
base = cases.methanol_readme_input("N1")
sw = cases.config3_sweep(B)
ps = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in sw.items()}
cm = engine.compile_model(base, method="ros4")
ws = engine.Workspace()

def run(pipeline, n=10):
    for _ in range(2):
        engine.n1_solve_ensemble(cm, base, ps, B, workspace=ws, want_stats=False, pipeline=pipeline)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        engine.n1_solve_ensemble(cm, base, ps, B, workspace=ws, want_stats=False, pipeline=pipeline)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0)/n*1e3

print("one launch: %.2f ms" % run(False))
for split in [(0.5, 0.5), (0.25, 0.5, 0.25), (0.125, 0.75, 0.125), (0.0625, 0.875, 0.0625), (0.1, 0.3, 0.3, 0.3), (0.125,)*8]:
    engine.PIPELINE_SPLIT = split
    print(split, "%.2f ms" % run(True))
# kernel-only time per ensemble size (device-resident inputs)
d = {k: v.cuda() for k, v in ps.items()}
for Bc in (B, B//2, B//4, B//8):
    sub = {k: v[:Bc] for k, v in d.items()}
    for _ in range(2):
        r = engine.n1_solve_ensemble(cm, base, sub, Bc, workspace=ws, keep_on_device=True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        r = engine.n1_solve_ensemble(cm, base, sub, Bc, workspace=ws, keep_on_device=True)
    e1.record(); torch.cuda.synchronize()
    print("device-resident B=%d: %.3f ms  (%.2f Msolves/s)" % (Bc, e0.elapsed_time(e1)/5, Bc/(e0.elapsed_time(e1)/5)/1e3))
