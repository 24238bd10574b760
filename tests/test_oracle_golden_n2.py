"""N2 (dynamic, method of lines) oracle against reference-generated fixtures."""
import os

import numpy as np
import pytest

import cases
import pyremot_oracle as O
from conftest import GOLDEN

RHS_CASES = {
    "methanol_testfile_z20": (lambda: cases.methanol_testfile_input("N2"), 20),
    "methanol_readme_z50": (lambda: cases.methanol_readme_input("N2"), 50),
    "ch4_z20": (lambda: cases.ch4_input("N2"), 20),
}


@pytest.fixture()
def zno():
    old = O.solverSetting["N2"]["zNo"]
    yield
    O.solverSetting["N2"]["zNo"] = old


@pytest.mark.parametrize("name", list(RHS_CASES))
def test_n2_rhs_known_answers(name, zno):
    g = np.load(os.path.join(GOLDEN, "n2_rhs_reference.npz"))
    mk, z = RHS_CASES[name]
    O.solverSetting["N2"]["zNo"] = z
    o = O.N2Oracle(mk())
    Y, F = g[name + "__rhs_Y"], g[name + "__rhs_F"]
    assert Y.shape[1] == o.varNo*z
    Fo = np.array([o.rhs(0.0, y) for y in Y])
    scale = np.maximum(np.abs(F), 1e-12*np.max(np.abs(F), axis=1, keepdims=True))
    assert np.max(np.abs(Fo - F)/scale) < 1e-12


def test_n2_ch4_solution(zno):
    g = np.load(os.path.join(GOLDEN, "n2_sol_ch4_reference.npz"))
    O.solverSetting["N2"]["zNo"] = int(g["zNo"])
    res = O.rmtExe(cases.ch4_input("N2"))["resModel"]
    assert len(res["dataPack"]) == 5
    for i, dp in enumerate(res["dataPack"]):
        np.testing.assert_allclose(dp["dataYs"], g["dataYs"][i], rtol=1e-9)
        np.testing.assert_allclose(dp["dataTime"], g["dataTime"][i])
        np.testing.assert_allclose(dp["dataXs"], g["dataXs"])
    # SURVEY App. B.5 value
    np.testing.assert_allclose(g["dataYs"][-1][:, -1],
                               [0.32318841102818, 0.238955222553162, 0.437856366418658, 218.4977721820071], rtol=1e-9)


def test_n2_methanol_fixture_matches_survey():
    g = np.load(os.path.join(GOLDEN, "n2_sol_m50_bdf_reference.npz"))
    np.testing.assert_allclose(g["dataYs"][-1][:, -1],
                               [0.4532165918878848, 0.2611244390409547, 0.02404786268633952, 0.2388594005756856,
                                0.006871246210421956, 0.01588045959871348, 621.0909995321149], rtol=1e-6)


@pytest.mark.slow
def test_n2_methanol_first_slab_bdf(zno):
    """First slab of the zNo=20 methanol case with BDF (the reference needs 56 s for all five)."""
    g = np.load(os.path.join(GOLDEN, "n2_sol_m20_bdf_reference.npz"))
    O.solverSetting["N2"]["zNo"] = 20
    o = O.N2Oracle(cases.methanol_testfile_input("N2"))
    from scipy.integrate import solve_ivp
    sol = solve_ivp(lambda t, y: o.rhs(t, y), [0, 0.1], o.IV, method="BDF", t_eval=np.linspace(0, 0.1, 5))
    np.testing.assert_allclose(sol.y[:, -1], g["soly_last"][0], rtol=1e-8, atol=1e-12)


def test_n2_isothermal_rhs_and_solution(zno):
    """process-type "iso-thermal": nc unknowns per node, temperature frozen (pbHomoReactor.py:3873-3914, :3638)."""
    g = np.load(os.path.join(GOLDEN, "n2_iso_reference.npz"))
    z = int(g["zNo"])
    O.solverSetting["N2"]["zNo"] = z
    mi = cases.ch4_input("N2", "iso-thermal")
    o = O.N2Oracle(mi)
    assert o.varNo == 3 and g["rhs_Y"].shape[1] == 3*z
    Fo = np.array([o.rhs(0.0, y) for y in g["rhs_Y"]])
    F = g["rhs_F"]
    scale = np.maximum(np.abs(F), 1e-12*np.max(np.abs(F), axis=1, keepdims=True))
    assert np.max(np.abs(Fo - F)/scale) < 1e-12
    res = O.rmtExe(mi)["resModel"]
    for i, dp in enumerate(res["dataPack"]):
        np.testing.assert_allclose(dp["dataYs"], g["default__dataYs"][i], rtol=1e-9)
