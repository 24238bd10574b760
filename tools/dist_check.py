"""Multi-GPU check (run under torchrun): config-4 style population, sharded, NCCL objective reduction
and gather, compared on rank 0 with an unsharded solve of the same population."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, torch, torch.distributed as dist
from rmt_app_b200 import engine, ensemble

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
B = int(os.environ.get("B", 65536))
base = cases.methanol_readme_input("N1")
base["reaction-rates"] = cases.methanol_kinetics_param(1171.2)
pop = cases.config4_population(B)
cm = engine.compile_model(base, method=engine.choose_method(base, engine.DEFAULT_RTOL, 1))   # what rmtExeBatchSharded picks
nominal = engine.n1_solve_ensemble(cm, base, None, 1).out[0, :, 0]
ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
r = ensemble.rmtExeBatchSharded(base, pop, B, objective_ref=nominal)       # warm-up
if world > 1: dist.barrier()
torch.cuda.synchronize(); ev0.record()
r = ensemble.rmtExeBatchSharded(base, pop, B, objective_ref=nominal)
ev1.record(); torch.cuda.synchronize()
ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ok = True
if rank == 0:
    full = engine.n1_solve_ensemble(cm, base, pop, B, objective_ref=nominal)
    ok = (np.array_equal(full.out[0].T, r["dataYs"]) and np.array_equal(full.objective, r["objective"])
          and abs(r["objective_sum"] - full.objective.sum()) <= 1e-9*abs(full.objective.sum())
          and r["objective_min"] == full.objective.min() and r["objective_argmin"] == int(full.objective.argmin())
          and r["failed"] == int((full.status != 0).sum()))
    print(json.dumps({"world": world, "B": B, "ok": bool(ok), "ms_sharded_incl_gather": float(ms.item()),
                      "objective_min": r["objective_min"], "objective_argmin": r["objective_argmin"],
                      "objective_sum": r["objective_sum"], "failed": r["failed"]}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
