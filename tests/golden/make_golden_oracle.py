#!/usr/bin/env python
"""Converged N2 solutions from the ORACLE (tests/golden/n2_sol_*_oracle_tight.npz).

The reference's own N2 path needs minutes per default-tolerance solve (56-446 s,
see make_golden.py logs) and hours at rtol=1e-10, so the converged solutions used
for the north_star 1e-6 check of the dynamic model come from the oracle port, whose
RHS is pinned to the reference's modelEquationN2 to 1e-12 (test_oracle_golden_n2.py)
and which calls the same scipy.integrate.solve_ivp with the same slab structure.

usage: python tests/golden/make_golden_oracle.py m20 | m50 | ch4 | m9 | config5
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import cases  # noqa: E402
import pyremot_oracle as O  # noqa: E402

CFG = {
    "m20": (lambda: cases.methanol_testfile_input("N2"), 20, "BDF", 1e-10, 1e-13),
    "m50": (lambda: cases.methanol_readme_input("N2"), 50, "BDF", 1e-9, 1e-12),
    "ch4": (lambda: cases.ch4_input("N2"), 20, "LSODA", 1e-11, 1e-13),
}

def m9_tight():
    """Converged M9 solution (12 nodes, 3 slabs: the grid of m9_reference.npz) from the oracle port, whose RHS and
    default-tolerance run are pinned to the reference (test_oracle_golden_m9.py)."""
    O.solverSetting["S2"].update(zNo=12, tNo=3)
    t0 = time.time()
    o = O.M9Oracle(cases.methanol_m9_input())
    res = o.solve(method="BDF", rtol=1e-10, atol=1e-13)
    dps = res["dataPack"]
    np.savez_compressed(os.path.join(HERE, "m9_sol_oracle_tight.npz"),
                        dataYs=np.array([d["dataYs"] for d in dps]), solY=np.array([d["solY"] for d in dps]),
                        dataTime=np.array([d["dataTime"] for d in dps]), zNo=np.array(12), nfev=np.array(o.nfev),
                        wall=np.array(time.time() - t0))
    print("m9 done in %.0f s, nfev %d" % (time.time() - t0, o.nfev), flush=True)


def n2_block_sparsity(nvar, zNo):
    """Variable-major state [var][node]: node k's equations read node k and node k-1 (upwind).  The pressure march
    couples a node to every node upstream as well (:3979) — a weak, smooth dependence that is left out of the
    finite-difference Jacobian on purpose: it only affects the Newton convergence rate of BDF, the converged
    solution is checked against the full right-hand side."""
    from scipy.sparse import lil_matrix
    S = lil_matrix((nvar*zNo, nvar*zNo), dtype=np.int8)
    for k in range(zNo):
        for r in range(nvar):
            for c in range(nvar):
                S[r*zNo + k, c*zNo + k] = 1
                if k > 0:
                    S[r*zNo + k, c*zNo + k - 1] = 1
    return S.tocsr()


def config5_tight(indices=(0, 6789, 12499), zNo=200, seed=20240613, B=12500):
    """Converged 200-node solutions of three instances of BASELINE config 5's per-GPU share (config-3 distributions,
    seed 20240613, period 0.5 s, 5 slabs) -> n2_sol_config5_z200_oracle_tight.npz."""
    from concurrent.futures import ProcessPoolExecutor
    t0 = time.time()
    with ProcessPoolExecutor(len(indices)) as ex:
        res = list(ex.map(_config5_one, [(i, zNo, seed, B) for i in indices]))
    np.savez_compressed(os.path.join(HERE, "n2_sol_config5_z200_oracle_tight.npz"),
                        index=np.array(indices), dataYs=np.array([r[0] for r in res]), solY=np.array([r[1] for r in res]),
                        dataTime=res[0][2], nfev=np.array([r[3] for r in res]), zNo=np.array(zNo), seed=np.array(seed),
                        B=np.array(B), method=np.array("BDF"), rtol=np.array(1e-9), atol=np.array(1e-12),
                        wall=np.array(time.time() - t0))
    print("config5 z200 done in %.0f s, nfev %s" % (time.time() - t0, [r[3] for r in res]), flush=True)


def _config5_one(arg):
    i, zNo, seed, B = arg
    O.solverSetting["N2"]["zNo"] = zNo
    base = cases.methanol_readme_input("N2")
    sw = cases.config3_sweep(B, seed)
    o = O.N2Oracle(cases.instance_input(base, sw, i))
    res = o.solve(method="BDF", rtol=1e-9, atol=1e-12, jac_sparsity=n2_block_sparsity(o.varNo, zNo))
    dps = res["dataPack"]
    return (np.array([d["dataYs"] for d in dps]), np.array([d["solY"] for d in dps]),
            np.array([d["dataTime"] for d in dps]), o.nfev)


if __name__ == "__main__":
    if "m9" in sys.argv[1:]:
        m9_tight()
        sys.argv.remove("m9")
    if "config5" in sys.argv[1:]:
        config5_tight()
        sys.argv.remove("config5")
    for which in sys.argv[1:]:
        mk, zNo, method, rtol, atol = CFG[which]
        O.solverSetting["N2"]["zNo"] = zNo
        t0 = time.time()
        o = O.N2Oracle(mk())
        res = o.solve(method=method, rtol=rtol, atol=atol)
        dps = res["dataPack"]
        np.savez_compressed(os.path.join(HERE, "n2_sol_%s_oracle_tight.npz" % which),
                            dataYs=np.array([d["dataYs"] for d in dps]), solY=np.array([d["solY"] for d in dps]),
                            dataTime=np.array([d["dataTime"] for d in dps]), zNo=np.array(zNo),
                            method=np.array(method), rtol=np.array(rtol), atol=np.array(atol), nfev=np.array(o.nfev),
                            wall=np.array(time.time() - t0))
        print(which, "done in %.0f s, nfev %d" % (time.time() - t0, o.nfev), flush=True)
