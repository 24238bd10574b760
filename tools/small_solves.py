"""Tiny solves of every kernel family and launch shape (N1 extents / full state / dense output, M7, N2 with 1-32 lanes
and ragged node counts, M9) — a quick end-to-end check, small enough to run under a memory checker where one is available."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from rmt_app_b200 import engine, rmtExe, rmtExeBatch, rmtExeBatchN2, solverSetting

base = cases.methanol_readme_input("N1")
sw = cases.config3_sweep(200, seed=5)
r = rmtExeBatch(base, sw)                                   # Ros4, extents
assert r["success"].all()
r = rmtExeBatch(base, sw, profile=True)                     # Rodas4, dense output
assert r["success"].all()
cm = engine.compile_model(base, reduced=False)
a = engine.n1_solve_ensemble(cm, base, sw, 200)             # full-state path
assert (a.status == 0).all()
Y = np.tile(np.r_[np.array(base["feed"]["concentration"])/max(base["feed"]["concentration"]), 1.0, 0.0], (5, 1))
engine.n1_rhs_batch(engine.compile_model(base), base, Y, jac=True)
engine.n1_rhs_batch(engine.compile_model(base), base, Y, system=True)
print("N1 ok")
rmtExeBatch(cases.methanol_m7_input(), {"temperature": np.array([520.0, 530.0])})
print("M7 ok")
mi = cases.ch4_input("N2")
for B, z in ((1, 12), (5, 21), (40, 16)):
    rb = rmtExeBatchN2(mi, {"temperature": np.linspace(940, 990, B)}, zNo=z, tNo=2)
    assert rb["success"].all()
for lanes, block in ((1, 64), (4, 32), (8, 64), (32, 32)):
    c2 = engine.compile_model(cases.methanol_testfile_input("N2"), block=block, lanes=lanes)
    q = engine.n2_solve_ensemble(c2, cases.methanol_testfile_input("N2"), None, 1, zNo=21, tNo=2, period=0.05)
    assert q.status[0] == 0
c3 = engine.compile_model(cases.methanol_testfile_input("N2"), block=256, lanes=0)      # stage pipeline: 64 reactors x 4 roles per block
q = engine.n2_solve_ensemble(c3, cases.methanol_testfile_input("N2"), {"temperature": np.linspace(500, 540, 70)}, 70, zNo=21, tNo=2, period=0.05)
assert (q.status == 0).all()
print("N2 ok")
old = dict(solverSetting["S2"]); solverSetting["S2"].update(zNo=12, tNo=2)
m9 = cases.methanol_m9_input(period=0.2)
rmtExe(m9)
rmtExeBatchN2(m9, {"temperature": np.array([523.0, 527.0, 531.0])})
solverSetting["S2"].update(old)
c9 = engine.compile_model(m9, block=256, lanes=0)
q = engine.n2_solve_ensemble(c9, m9, {"temperature": np.linspace(520.0, 530.0, 40)}, 40, zNo=12, tNo=2, period=0.2)
assert (q.status == 0).all()
print("M9 ok")
