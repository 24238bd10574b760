"""First-contact probe for a GPU box: RHS/Jacobian/solve parity + timing."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
import torch
from rmt_app_b200 import engine, rmtExe, rmtExeBatch, capi

d = np.load(os.path.join(ROOT, "tests/golden/n1_reference.npz"))
for name, mi in [("methanol_readme", cases.methanol_readme_input()), ("methanol_testfile", cases.methanol_testfile_input()),
                 ("ch4_noniso", cases.ch4_input()), ("ch4_iso", cases.ch4_input("N1", "iso-thermal"))]:
    t0 = time.time()
    cm = engine.compile_model(mi)
    Y, F = d[name + "__rhs_Y"], d[name + "__rhs_F"]
    Fg, Jg, consts = engine.n1_rhs_batch(cm, mi, Y, jac=True)
    print(name, "compile+rhs %.1fs" % (time.time() - t0))
    scale = np.max(np.abs(F), axis=1, keepdims=True)
    print("  rhs max err / max|f| :", np.max(np.abs(Fg - F)/scale), " elementwise rel:", np.max(np.abs(Fg - F)/np.maximum(np.abs(F), 1e-300)))
    print("  GaMiVi", consts[10, 0], d[name + "__const_GaMiVi"], " Gh", consts[9, 0], d[name + "__da_GaHeCoTe0"])
    # Jacobian vs central differences of the GPU RHS itself
    B, n = Y.shape
    Jfd = np.zeros((B, n, n))
    for j in range(n):
        h = 1e-6*np.maximum(np.abs(Y[:, j]), 1e-3)
        Yp, Ym = Y.copy(), Y.copy(); Yp[:, j] += h; Ym[:, j] -= h
        Fp, _, _ = engine.n1_rhs_batch(cm, mi, Yp); Fm, _, _ = engine.n1_rhs_batch(cm, mi, Ym)
        Jfd[:, :, j] = (Fp - Fm)/(2*h)[:, None]
    js = np.max(np.abs(Jfd), axis=(1, 2), keepdims=True)
    print("  jac max |J-Jfd| / max|J| :", np.max(np.abs(Jg - Jfd)/js))
    # solves
    for rtol, atol in [(1e-3, 1e-6), (1e-6, 1e-9), (1e-9, 1e-12)]:
        mi2 = dict(mi); mi2["solver-config"] = dict(mi["solver-config"], rtol=rtol, atol=atol)
        t0 = time.time(); r = rmtExe(mi2); dt = time.time() - t0
        dp = r["resModel"][0]
        tight = d[name + "__tight_LSODA__dataYs"]
        dflt = d[name + "__default__dataYs"]
        e = np.abs(dp["dataYs"] - tight)/np.abs(tight)
        print("  rtol %g: outlet rel err vs tight ref %.3e  profile max %.3e | ref default err %.3e | stats %s  %.3fs" % (
            rtol, e[:, -1].max(), e.max(), (np.abs(dflt - tight)/np.abs(tight)).max(), dp["solverStats"], dt))

# sweep timing
base = cases.methanol_readme_input()
for B in (1024, 65536, 1 << 20):
    sw = cases.config3_sweep(B)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        r = rmtExeBatch(base, sw)
        torch.cuda.synchronize(); dt = time.time() - t0
    st = r["stats"]
    print("B=%d e2e %.3fs -> %.0f solves/s ; ok=%d/%d ; steps mean %.1f max %d rej mean %.1f" % (
        B, dt, B/dt, int(r["success"].sum()), B, st[0].mean(), st[0].max(), st[1].mean()))
cm = engine.compile_model(base)
print("fp64 peak TFLOP/s:", cm.module.fp64_peak())
