"""Where the time of one sharded parameter-estimation population (BASELINE config 4, 65 536 sets) goes: phases of
ensemble.rmtExeBatchSharded timed with a device synchronisation after each (which serialises them: the sum is an upper
bound of the call's wall time)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, torch
from rmt_app_b200 import engine, ensemble

B = int(os.environ.get("B", 65536))
base = cases.methanol_readme_input("N1"); base["reaction-rates"] = cases.methanol_kinetics_param(1171.2)
pop = cases.config4_population(B)
ppop = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in pop.items()}
cm = engine.compile_model(base, method="ros4")
nominal = engine.n1_solve_ensemble(cm, base, None, 1, rtol=1e-9, atol=1e-12).out[0, :, 0]
ws = engine.Workspace()
for _ in range(3):
    r = ensemble.rmtExeBatchSharded(base, ppop, B, objective_ref=nominal, workspace=ws)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    r = ensemble.rmtExeBatchSharded(base, ppop, B, objective_ref=nominal, workspace=ws)
torch.cuda.synchronize()
print("whole call: %.3f ms" % ((time.perf_counter() - t0)/20*1e3))
t0 = time.perf_counter()
for _ in range(20):
    r = ensemble.rmtExeBatchSharded(base, ppop, B, objective_ref=nominal, workspace=ws, gather=False)
torch.cuda.synchronize()
print("statistics only: %.3f ms" % ((time.perf_counter() - t0)/20*1e3))

# phases
dev = torch.device("cuda", 0); mod = cm.load(0); spec = cm.spec; n = spec.n
stream = torch.cuda.current_stream().cuda_stream
def timed(f, reps=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): out = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0)/reps*1e3, out
t, cmx = timed(lambda: engine.compile_model(base, method=engine.choose_method(base, 1e-3, 1))); print("compile_model lookup %.3f ms" % t)
t, rows = timed(lambda: engine.sweep_rows_to_device(spec, ppop, B, ws, dev)); print("inputs -> device (%d rows) %.3f ms" % (rows[1], t))
d_rows, n_rows, row_map, _ = rows
d_consts = ws.get("d_consts", (mod.info.nconst, B), torch.float64, device=dev)
uni = engine.uniform_inputs(spec, base)
t, _ = timed(lambda: mod.setup(B, d_rows, n_rows, row_map, uni, d_consts, stream=stream)); print("rmt_setup %.3f ms" % t)
lay = ensemble.PackLayout(B, 1, n + 1)
pack = ws.get("d_pack", (lay.length,), torch.float64, device=dev)
prow, status, tail = lay.views(pack, 0)
d_stats = ws.get("d_stats", (4, B), torch.int32, device=dev)
ctrl = engine.METHOD_CTRL.get(cm.method)
t, _ = timed(lambda: mod.n1_solve_population(B, d_consts, 1.0, 1e-3, 1e-6, prow[:n], status, d_stats, nominal, prow[n], tail, index_offset=0, ctrl=ctrl, stream=stream))
print("rmt_n1_solve_population %.3f ms" % t)
t, _ = timed(lambda: lay.unpack_device(pack.view(1, -1), ws, transpose=True)); print("unpack_device + D2H %.3f ms" % t)
