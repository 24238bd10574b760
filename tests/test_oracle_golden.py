"""The CPU oracle against numbers produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest

import cases
import pyremot_oracle as O
from conftest import GOLDEN

N1_CASES = {
    "methanol_readme": lambda: cases.methanol_readme_input("N1"),
    "methanol_testfile": lambda: cases.methanol_testfile_input("N1"),
    "ch4_noniso": lambda: cases.ch4_input("N1", "non-iso-thermal"),
    "ch4_iso": lambda: cases.ch4_input("N1", "iso-thermal"),
}


@pytest.mark.parametrize("name", list(N1_CASES))
def test_n1_rhs_known_answers(golden_n1, name):
    o = O.N1Oracle(N1_CASES[name]())
    Y, F = golden_n1[name + "__rhs_Y"], golden_n1[name + "__rhs_F"]
    Fo = np.array([o.rhs(0.0, y) for y in Y])
    assert np.max(np.abs(Fo - F)/np.maximum(np.abs(F), 1e-300)) < 1e-13


@pytest.mark.parametrize("name", list(N1_CASES))
def test_n1_setup_constants(golden_n1, name):
    o = O.N1Oracle(N1_CASES[name]())
    g = lambda k: golden_n1[name + "__" + k]
    np.testing.assert_allclose(o.GaMiVi, g("const_GaMiVi"), rtol=1e-14)
    np.testing.assert_allclose(o.StHeRe25, g("const_StHeRe25"), rtol=1e-14)
    np.testing.assert_allclose(o.GaHeCoTe0, g("da_GaHeCoTe0"), rtol=1e-14)
    np.testing.assert_allclose(o.GaMaCoTe0, g("da_GaMaCoTe0"), rtol=1e-14)
    np.testing.assert_allclose(o.GaDe0, g("bc_GaDe0"), rtol=1e-14)
    np.testing.assert_allclose(o.GaCpMeanMix0, g("bc_GaCpMeanMix0"), rtol=1e-14)
    np.testing.assert_allclose(o.CrSeAr, g("const_CrSeAr"), rtol=1e-15)
    np.testing.assert_allclose(o.a, g("exhe_EfHeTrAr"), rtol=1e-15)


@pytest.mark.parametrize("name", list(N1_CASES))
@pytest.mark.parametrize("ivp", ["default", "BDF"])
def test_n1_solution_same_integrator(golden_n1, name, ivp):
    """Same SciPy integrator on a bit-identical RHS reproduces the reference's dataPack."""
    mi = N1_CASES[name]()
    mi["solver-config"]["ivp"] = ivp
    dp = O.rmtExe(mi)["resModel"][0]
    ref = golden_n1["%s__%s__dataYs" % (name, ivp)]
    assert dp["dataYs"].shape == ref.shape
    np.testing.assert_allclose(dp["dataYs"], ref, rtol=1e-10, atol=0)
    if ivp == "default":
        for k in ("dataXs", "dataYCons1", "dataYCons2", "dataYTemp1", "dataYTemp2"):
            np.testing.assert_allclose(np.asarray(dp[k]), golden_n1[name + "__" + k], rtol=1e-10, atol=1e-300)
        assert list(golden_n1[name + "__labelList"]) == dp["labelList"]
        assert list(golden_n1[name + "__indexList"]) == dp["indexList"]
        assert dp["nfev"] == int(golden_n1["%s__default__nfev_njev_wall" % name][0])


def test_n1_tight_tolerance(golden_n1):
    mi = cases.methanol_readme_input("N1")
    dp = O.rmtExe(mi, method="LSODA", rtol=1e-10, atol=1e-12)["resModel"][0]
    np.testing.assert_allclose(dp["dataYs"], golden_n1["methanol_readme__tight_LSODA__dataYs"], rtol=1e-9)
    # the reference's own solvers agree with each other at this tolerance
    a, b = golden_n1["methanol_readme__tight_LSODA__dataYs"], golden_n1["methanol_readme__tight_BDF__dataYs"]
    assert np.max(np.abs(a - b)/np.abs(a)) < 5e-8


def test_survey_appendix_b_numbers(golden_n1):
    """Cross-check the fixture against the numbers printed in SURVEY.md App. B."""
    out = golden_n1["methanol_readme__default__dataYs"][:, -1]
    np.testing.assert_allclose(out, [0.4530051402722371, 0.2615038024945862, 0.02404580582309218, 0.2384800253074776,
                                     0.006904390387159711, 0.01606083571544745, 4992662.964385521, 620.8566399345168],
                               rtol=1e-12)
    np.testing.assert_allclose(golden_n1["methanol_testfile__rhs_F"][0],
                               [-3.7201689861515614e+00, -2.1931170600484968e+00, 4.3185423763574233e+00,
                                1.4295910969969647e+00, -3.4873246695663203e+00, 2.1254253163089261e+00,
                                -1.8092791793935929e-03, 2.7384600289671184e+00], rtol=1e-12)
    np.testing.assert_allclose(golden_n1["rates_ka__R"], [-18.74827321861196, 85.01229621255123, 13586.06348190731],
                               rtol=1e-12)


def test_rates_known_answer(golden_n1):
    inp = golden_n1["rates_ka__inputs"]
    kin = cases.methanol_kinetics(1982*(1 - 0.39))
    R = O.reaction_rate_exe((inp[0], inp[1], inp[2:8], inp[8:14]), kin["VARS"], kin["RATES"])
    np.testing.assert_allclose(R, golden_n1["rates_ka__R"], rtol=1e-15)


def test_corner_sweep_instances(golden_corners):
    """Three of the 36 config-3 corners through the oracle."""
    base = cases.methanol_readme_input("N1")
    sw = cases.config3_corners()
    for k in ("temperature", "pressure", "concentration"):
        np.testing.assert_array_equal(sw[k], golden_corners["sweep_" + k])
    for i in (0, 17, 35):
        dp = O.rmtExe(cases.instance_input(base, sw, i))["resModel"][0]
        np.testing.assert_allclose(dp["dataYs"], golden_corners["default_dataYs"][i], rtol=1e-9)
        assert dp["nfev"] == int(golden_corners["stats"][i, 0])


def test_component_properties():
    g = np.load(os.path.join(GOLDEN, "component_props_reference.npz"))
    syms = list(g["symbols"])
    assert O.rmtCom() == str(g["rmtCom"])
    for k, T in enumerate(g["Ts"]):
        cp = [O.cp_component(s, T) for s in syms]
        np.testing.assert_allclose(cp, g["cp"][k], rtol=1e-15)
        mu = O.gas_viscosity(syms, T)
        np.testing.assert_allclose(mu, g["mu"][k], rtol=1e-15)
        np.testing.assert_allclose(O.wilke_mixture(12, mu, g["wilke_y"], g["MW"]), g["wilke"][k], rtol=1e-14)
    np.testing.assert_allclose([O._DB[s][0] for s in syms], g["MW"])
    np.testing.assert_allclose([O._DB[s][1] for s in syms], g["dHf25"])
    reactions = {"R%d" % i: str(r) for i, r in enumerate(g["reactions"])}
    np.testing.assert_allclose([O.standard_enthalpy_of_reaction(r) for r in reactions.values()], g["dH25"], rtol=1e-13)
    _, vec = O.parse_reactions(reactions)
    flat = np.array([[j, syms.index(s), v] for j, r in enumerate(vec) for s, v in r], float)
    np.testing.assert_array_equal(flat, g["stoich"])


def test_unknown_component_raises():
    mi = cases.ch4_input("N1")
    mi["feed"]["components"]["shell"] = ["CH4", "C2H4", "Xe"]
    with pytest.raises(Exception, match="Component database is not up to date"):
        O.rmtExe(mi)
