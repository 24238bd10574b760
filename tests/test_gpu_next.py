"""GPU tests of the rows SURVEY 8(f) lists as "next": estimation driver, N2 batch API."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def test_differential_evolution_recovers_kinetic_parameters():
    from rmt_app_b200 import differential_evolution, engine
    base = cases.methanol_readme_input("N1")
    base["reaction-rates"] = cases.methanol_kinetics_param(1171.2)
    truth = {"k01": 35.45*1.3, "E3": -5.2940e4*0.985}
    data_in = cases.methanol_readme_input("N1")
    data_in["reaction-rates"] = cases.methanol_kinetics_param(1171.2, truth)
    cm = engine.compile_model(data_in)
    outlet = engine.n1_solve_ensemble(cm, data_in, None, 1, rtol=1e-8, atol=1e-11).out[0, :, 0]
    res = differential_evolution(base, {"k01": (20.0, 60.0), "E3": (-5.6e4, -5.0e4)}, outlet, popsize=1024,
                                 generations=40, seed=1, rtol=1e-6, atol=1e-9)
    assert res["fun"] < 1e-9
    assert res["x"]["k01"] == pytest.approx(truth["k01"], rel=2e-3)
    assert res["x"]["E3"] == pytest.approx(truth["E3"], rel=2e-4)
    assert res["history"][-1] <= res["history"][0] and res["nsolves"] == 1024*41


def test_rmtexebatch_n2_matches_single_runs():
    from rmt_app_b200 import rmtExe, rmtExeBatchN2, solverSetting
    old = dict(solverSetting["N2"])
    try:
        solverSetting["N2"].update(zNo=16, tNo=3)
        mi = cases.ch4_input("N2")
        B = 40
        rng = np.random.default_rng(4)
        sw = {"temperature": rng.uniform(920, 990, B), "pressure": rng.uniform(2.5e5, 3.5e5, B)}
        r = rmtExeBatchN2(mi, sw)
        assert r["dataYs"].shape == (B, 3, 4, 16) and r["success"].all()
        np.testing.assert_allclose(r["dataTime"], [10/3, 20/3, 10.0])
        j = 7
        one = dict(mi)
        one["operating-conditions"] = dict(mi["operating-conditions"], temperature=float(sw["temperature"][j]),
                                           pressure=float(sw["pressure"][j]))
        packs = rmtExe(one)["resModel"]["dataPack"]
        for i in range(3):
            np.testing.assert_allclose(r["dataYs"][j, i], packs[i]["dataYs"], rtol=1e-12)
    finally:
        solverSetting["N2"].update(old)
