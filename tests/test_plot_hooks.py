"""Plot hooks: rmt_app_b200/plotting.py draws what the reference's plotResultsSteadyState /
plotResultsDynamic draw (solvers/solResultAnalysis.py:307-459) — compared call by call with a recording made by
running the REFERENCE's hooks on the same result dictionaries (tests/golden/make_golden.py plots)."""
import json
import os
import sys
import types

import numpy as np
import pytest

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
from make_golden import PlotRecorder, synthetic_packs  # noqa: E402  (recorder + inputs shared with the fixture generator)


@pytest.fixture()
def recorder(monkeypatch):
    rec = PlotRecorder()
    plt = types.ModuleType("matplotlib.pyplot")
    rec.install(plt)
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    return rec


def test_hooks_draw_what_the_reference_draws(recorder):
    from rmt_app_b200 import plotting
    want = json.load(open(os.path.join(GOLDEN, "plot_calls_reference.json")))
    steady, steady_iso, dyn = synthetic_packs()
    plotting.plotResultsSteadyState(steady)
    assert recorder.calls == want["steady"]
    recorder.calls = []
    plotting.plotResultsSteadyState(steady_iso)
    assert recorder.calls == want["steady_iso"]
    for seed in (7, 8):                                    # the slab choice follows NumPy's global generator like the reference's
        recorder.calls = []
        np.random.seed(seed)
        plotting.plotResultsDynamic(dyn, 6)
        assert recorder.calls == want["dynamic_seed%d" % seed]
    # first and last slab always, two interior ones
    np.random.seed(0)
    sel = plotting.select_slabs(6)
    assert sel[0] == 0 and sel[-1] == 5 and len(sel) == 4 and sel == sorted(sel) and len(set(sel)) == 4


def test_missing_matplotlib_is_a_notice_not_an_error(monkeypatch, capsys):
    from rmt_app_b200 import plotting
    monkeypatch.setitem(sys.modules, "matplotlib", None)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", None)
    plotting.plotResultsSteadyState(synthetic_packs()[0])
    assert "matplotlib is not installed" in capsys.readouterr().out


@pytest.mark.gpu
def test_display_result_true_draws_through_rmtExe(recorder):
    """solver-config.display-result == "True" (README.md:225-228): N1 draws three figures, N2 two per selected slab."""
    import cases
    from rmt_app_b200 import rmtExe, solverSetting
    mi = cases.methanol_readme_input("N1")
    mi["solver-config"] = dict(mi["solver-config"], **{"display-result": "True"})
    res = rmtExe(mi)["resModel"][0]
    calls = recorder.calls
    assert [c[0] for c in calls].count("show") == 3
    plots = [c for c in calls if c[0] == "plot"]
    assert [c[1] for c in plots] == res["labelList"]
    assert plots[-1][6] == pytest.approx(res["dataYs"][-1, -1]) and plots[-2][5] == pytest.approx(5e6)
    assert calls[6 + 0][1].startswith("Steady-State Modeling N1, computation-time ")
    assert [c[1] for c in calls if c[0] == "ylabel"] == ["Concentration (mol/$m^3$)", "Pressure (bar)", "Temperature (K)"]
    recorder.calls = []
    mi2 = cases.ch4_input("N2")
    mi2["solver-config"] = dict(mi2["solver-config"], **{"display-result": "True"})
    np.random.seed(3)
    res2 = rmtExe(mi2)["resModel"]
    tNo = solverSetting["N2"]["tNo"]
    shows = [c for c in recorder.calls if c[0] == "show"]
    assert len(shows) == 2*4 and tNo == 5                   # first + two interior + last slab, two figures each
    titles = [c[1] for c in recorder.calls if c[0] == "title"]
    assert titles[0].endswith(" at t=%s" % (res2["dataPack"][0]["dataTime"],))
    assert titles[-1].endswith(" at t=%s" % (res2["dataPack"][-1]["dataTime"],))
