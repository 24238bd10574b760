"""N2 ensemble: lanes-per-reactor / block-size sweep; checks that every variant returns the same bits."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, torch
from rmt_app_b200 import engine

B = int(os.environ.get("B", 12500)); Z = int(os.environ.get("Z", 200))
mi = cases.methanol_m9_input() if os.environ.get("MODEL") == "M9" else cases.methanol_readme_input("N2")
PERIOD = float(os.environ.get("PERIOD", mi["operating-conditions"]["period"] if os.environ.get("MODEL") == "M9" else 0.5))
sw = cases.config3_sweep(B, 20240613) if B > 1 else None
if os.environ.get("MODEL") == "M9" and B > 1:       # M9's units differ (kmol/m^3): sweep the feed temperature only
    sw = {"temperature": 523.0 + np.random.default_rng(7).uniform(-10.0, 10.0, B)}
ref = None
for v in sys.argv[1:]:
    parts = v.split(",")
    lanes, blk, defs = int(parts[0]), int(parts[1]), parts[2:]
    cm = engine.compile_model(mi, block=blk, lanes=lanes)
    if defs:        # extra -D options: bypass the cubin cache
        from rmt_app_b200 import capi
        capi.init(0)
        cubin, _ = capi.nvrtc_compile(cm.header, block=blk, extra_opts=["-D" + d for d in defs])
        import copy
        cm = copy.copy(cm); cm.module = capi.Module(cubin)
    try:
        ws = engine.Workspace()
        dsw = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in sw.items()} if sw else None
        dts = []
        for rep in range(int(os.environ.get("REPS", 4))):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            res = engine.n2_solve_ensemble(cm, mi, dsw, B, zNo=Z, tNo=5, period=PERIOD, keep_on_device=True, workspace=ws)
            e1.record(); torch.cuda.synchronize(); dts.append(e0.elapsed_time(e1)*1e-3)
        dt = min(dts[1:]) if len(dts) > 1 else dts[0]
    except Exception as e:
        print(v, "failed:", str(e)[:300]); continue
    st = res.stats.cpu().numpy(); ok = int((res.status == 0).sum())
    out = res.out.cpu().numpy()
    if ref is None:
        ref = out
    same = np.array_equal(out, ref)
    dev = float(np.nanmax(np.abs(out - ref)/np.abs(ref)))
    print(",".join(defs), end=" ")
    print("lanes %2d block %3d B=%d zNo=%d: %.4fs ok %d steps %.3f rej %.3f identical %s maxdev %.1e" % (
        lanes, blk, B, Z, dt, ok, st[0].mean(), st[1].mean(), same, dev), flush=True)
