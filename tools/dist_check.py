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
# the same population through the C ABI's own NCCL transport (rmt_comm_*: no torch.distributed on the data path; the
# unique id travels over the torch.distributed store here, any side channel would do)
ok_comm = True
if world > 1:
    from rmt_app_b200 import capi
    ids = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = capi.Comm(world, rank, ids[0])
    rc = ensemble.rmtExeBatchSharded(base, pop, B, objective_ref=nominal, comm=comm)
    ok_comm = (np.array_equal(rc["dataYs"], r["dataYs"]) and np.array_equal(rc["objective"], r["objective"])
               and rc["objective_min"] == r["objective_min"] and rc["objective_argmin"] == r["objective_argmin"]
               and rc["failed"] == r["failed"] and comm.nccl_version() > 0)
    comm.close()
# dynamic model, sharded: every rank integrates its block, ONE packed gather of the final profiles
mi2 = cases.methanol_testfile_input("N2")
B2 = 37
sw2 = {"temperature": np.linspace(505.0, 545.0, B2)}
r2 = ensemble.rmtExeBatchN2Sharded(mi2, sw2, B2, zNo=16, tNo=2, gather="final")
ok_n2 = r2["failed"] == 0 and r2["dataYs"].shape == (B2, 7, 16)
if rank == 0:
    cm2 = engine.compile_model_n2(mi2, B2, 16)
    full2 = engine.n2_solve_ensemble(cm2, mi2, sw2, B2, zNo=16, tNo=2)
    want = np.moveaxis(full2.out[-1], -1, 0)                          # [B][rows][zNo]
    ok_n2 = ok_n2 and np.allclose(r2["dataYs"], want, rtol=1e-9, atol=0) and np.array_equal(r2["status"], full2.status)
flags = torch.tensor([float(ok), float(ok_comm), float(ok_n2)], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
ok_all = bool(flags.min().item() > 0.5)
if rank == 0:
    print(json.dumps({"ok": ok_all, "ok_population": bool(flags[0].item() > 0.5), "ok_rmt_comm_transport": bool(flags[1].item() > 0.5),
                      "ok_n2_sharded": bool(flags[2].item() > 0.5)}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
