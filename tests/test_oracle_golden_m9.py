"""Model M9 (dimensional dynamic twin of N2, pbReactor.runM5) — oracle against the reference-generated fixture."""
import os

import numpy as np

import cases
import pyremot_oracle as O
from conftest import GOLDEN


def test_m9_rhs_and_slab_states():
    g = np.load(os.path.join(GOLDEN, "m9_reference.npz"))
    old = dict(O.solverSetting["S2"])
    O.solverSetting["S2"].update(zNo=12, tNo=3)           # the grid the fixture was generated on
    try:
        mi = cases.methanol_m9_input()
        o = O.M9Oracle(mi)
        F = np.array([o.rhs(0.0, y) for y in g["rhs_Y"]])
        assert np.max(np.abs(F - g["rhs_F"])/np.maximum(np.abs(g["rhs_F"]), 1e-300)) < 1e-13
        res = O.rmtExe(mi)["resModel"]
        assert o.__class__.__name__ == "M9Oracle"
        states = np.array([p["solY"] for p in res["dataPack"]])
        np.testing.assert_allclose(states, g["default__slab_end_states"], rtol=1e-10)
        np.testing.assert_allclose(np.array([xy[1] for xy in res["XYList"]]), g["default__T_profiles"], rtol=1e-10)
        np.testing.assert_array_equal(res["XYList"][0][0], g["default__x"])
        assert [d["leg"] for d in res["dataList"]] == list(g["default__legends"])
    finally:
        O.solverSetting["S2"].update(old)
