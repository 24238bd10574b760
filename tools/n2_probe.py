import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, torch
from rmt_app_b200 import engine, rmtExe, solverSetting
G = os.path.join(ROOT, "tests/golden")
g = np.load(os.path.join(G, "n2_rhs_reference.npz"))
for name, mk, z in [("ch4_z20", lambda: cases.ch4_input("N2"), 20), ("methanol_testfile_z20", lambda: cases.methanol_testfile_input("N2"), 20),
                    ("methanol_readme_z50", lambda: cases.methanol_readme_input("N2"), 50)]:
    mi = mk(); cm = engine.compile_model(mi)
    Y, F = g[name + "__rhs_Y"], g[name + "__rhs_F"]
    Fg = engine.n2_rhs_batch(cm, mi, Y, z)
    n = cm.spec.n
    Fr, Fq = F.reshape(len(F), n, z), Fg.reshape(len(F), n, z)
    sc = np.max(np.abs(Fr), axis=1, keepdims=True)
    print(name, "rhs err/node-scale", np.max(np.abs(Fq - Fr)/sc), "first state", np.max(np.abs(Fq[0] - Fr[0])/sc[0]))
# solves
for nm, mk, z, ref, tols in [("ch4", lambda: cases.ch4_input("N2"), 20, "n2_sol_ch4_tight_reference.npz", [(1e-3, 1e-6), (1e-6, 1e-9), (1e-9, 1e-12)]),
                             ("m20", lambda: cases.methanol_testfile_input("N2"), 20, "n2_sol_m20_tight_reference.npz", [(1e-3, 1e-6), (1e-6, 1e-9), (1e-8, 1e-11)]),
                             ("m50", lambda: cases.methanol_readme_input("N2"), 50, "n2_sol_m50_bdf_reference.npz", [(1e-3, 1e-6), (1e-6, 1e-9)])]:
    r = np.load(os.path.join(G, ref))
    solverSetting["N2"]["zNo"] = z
    for rtol, atol in tols:
        mi = mk(); mi["solver-config"].update(rtol=rtol, atol=atol)
        cm = engine.compile_model(mi, block=engine.n2_block(1))
        t0 = time.time()
        res = engine.n2_solve_ensemble(cm, mi, None, 1, out_mode=1)
        dt = time.time() - t0
        rel = np.abs(res.out[..., 0] - r["dataYs"])/np.abs(r["dataYs"])
        print("%s rtol %g: status %d stats %s  %.2fs | vs %s: max rel all slabs %.2e, last slab outlet %.2e, T prof %.2e" % (
            nm, rtol, res.status[0], res.stats[:, 0], dt, ref, rel.max(), rel[-1][:, -1].max(), rel[-1][-1].max()))
# small ensemble timing
mi = cases.methanol_readme_input("N2")
for B, z, blk in [(1024, 50, None), (4096, 50, None), (4096, 200, None), (4096, 200, 64), (12500, 200, None), (12500, 200, 128), (12500, 200, 256), (50000, 50, None)]:
    cm = engine.compile_model(mi, block=blk or engine.n2_block(B))
    sw = cases.config3_sweep(B, 20240613)
    torch.cuda.synchronize(); t0 = time.time()
    res = engine.n2_solve_ensemble(cm, mi, sw, B, zNo=z, tNo=5, period=0.5, keep_on_device=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    st = res.stats.cpu().numpy(); ok = int((res.status == 0).sum())
    print("N2 ensemble B=%d zNo=%d block %d: %.2fs -> %.0f inst/s, ok %d, steps mean %.0f rej %.1f" % (B, z, cm.block, dt, B/dt, ok, st[0].mean(), st[1].mean()))
