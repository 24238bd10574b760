"""Model M7 (dimensional steady-state twin of N1) — oracle against the reference-generated fixture."""
import os

import numpy as np

import cases
import pyremot_oracle as O
from conftest import GOLDEN


def test_m7_rhs_and_solution():
    g = np.load(os.path.join(GOLDEN, "m7_reference.npz"))
    mi = cases.methanol_m7_input()
    o = O.M7Oracle(mi)
    F = np.array([o.rhs(0.0, y) for y in g["rhs_Y"]])
    assert np.max(np.abs(F - g["rhs_F"])/np.maximum(np.abs(g["rhs_F"]), 1e-300)) < 1e-13
    res = O.rmtExe(mi)["resModel"]
    np.testing.assert_allclose(res["dataYs"], g["default__dataYs"], rtol=1e-10)
    assert res["nfev"] == int(g["default__nfev_wall"][0])
    assert len(res["XYList"]) == 7 and res["dataList"][6]["leg"] == "Temperature"
    tight = O.rmtExe(mi, method="LSODA", rtol=1e-10, atol=1e-12)["resModel"]
    np.testing.assert_allclose(tight["dataYs"], g["tight__dataYs"], rtol=1e-9)
