"""GPU parity tests for the dynamic method-of-lines model N2 (run with -m gpu)."""
import os

import numpy as np
import pytest

import cases
import pyremot_oracle as O
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

RHS_CASES = {
    "methanol_testfile_z20": (lambda: cases.methanol_testfile_input("N2"), 20),
    "methanol_readme_z50": (lambda: cases.methanol_readme_input("N2"), 50),
    "ch4_z20": (lambda: cases.ch4_input("N2"), 20),
}


@pytest.fixture()
def n2_settings():
    from rmt_app_b200 import solverSetting
    old = dict(solverSetting["N2"]), O.solverSetting["N2"]["zNo"]
    yield solverSetting
    solverSetting["N2"].update(old[0])
    O.solverSetting["N2"]["zNo"] = old[1]


@pytest.mark.parametrize("name", list(RHS_CASES))
def test_n2_rhs_parity_with_reference_golden(name):
    from rmt_app_b200 import engine
    g = np.load(os.path.join(GOLDEN, "n2_rhs_reference.npz"))
    mk, z = RHS_CASES[name]
    mi = mk()
    cm = engine.compile_model(mi)
    Y, F = g[name + "__rhs_Y"], g[name + "__rhs_F"]
    Fg = engine.n2_rhs_batch(cm, mi, Y, z)
    # node-wise scale: the near-equilibrium cancellation of N1 applies per node
    n = cm.spec.n
    Fr, Fq = F.reshape(len(F), n, z), Fg.reshape(len(F), n, z)
    scale = np.max(np.abs(Fr), axis=1, keepdims=True)
    assert np.max(np.abs(Fq - Fr)/scale) < 2e-9
    # the initial state (feed composition everywhere) is well conditioned
    assert np.max(np.abs(Fq[0] - Fr[0])/np.maximum(np.abs(Fr[0]), 1e-9*scale[0])) < 1e-10


def test_n2_ch4_tight_matches_reference_tight(n2_settings):
    """Level 2 for the dynamic model: all five slab states within 1e-6 of the reference's own
    tight-tolerance LSODA run (tests/golden/n2_sol_ch4_tight_reference.npz)."""
    from rmt_app_b200 import rmtExe
    g = np.load(os.path.join(GOLDEN, "n2_sol_ch4_tight_reference.npz"))
    n2_settings["N2"]["zNo"] = int(g["zNo"])
    mi = cases.ch4_input("N2")
    mi["solver-config"].update(rtol=1e-9, atol=1e-12)
    res = rmtExe(mi)["resModel"]
    assert len(res["dataPack"]) == 5
    for i, dp in enumerate(res["dataPack"]):
        ref = g["dataYs"][i]
        assert dp["dataYs"].shape == ref.shape
        assert np.max(np.abs(dp["dataYs"] - ref)/np.abs(ref)) < 1e-6, i
        np.testing.assert_allclose(dp["dataTime"], g["dataTime"][i])
        np.testing.assert_allclose(dp["dataXs"], g["dataXs"])
        assert dp["labelList"] == ["CH4", "C2H4", "H2", "Temperature"]
        assert dp["dataYCons1"].shape == (3, 20) and dp["dataYCons2"].shape == (3, 20)
        assert np.asarray(dp["dataYTemp1"]).shape == (20,) and dp["dataYTemp2"].shape == (1, 20)


def test_n2_methanol_z20_against_converged_oracle_and_reference(n2_settings):
    from rmt_app_b200 import rmtExe
    conv = np.load(os.path.join(GOLDEN, "n2_sol_m20_oracle_tight.npz"))
    n2_settings["N2"]["zNo"] = 20
    mi = cases.methanol_testfile_input("N2")
    mi["solver-config"].update(rtol=1e-8, atol=1e-11)
    res = rmtExe(mi)["resModel"]
    for i, dp in enumerate(res["dataPack"]):
        ref = conv["dataYs"][i]
        rel = np.abs(dp["dataYs"] - ref)/np.abs(ref)
        assert rel[-1].max() < 1e-6, (i, rel[-1].max())            # temperature profile
        assert rel[:, -1].max() < 1e-6, (i, rel[:, -1])            # outlet mole fractions
        assert rel.max() < 1e-5
    # the reference's own default runs sit within their tolerance of our converged answer
    for f in ("n2_sol_m20_lsoda_reference.npz", "n2_sol_m20_bdf_reference.npz", "n2_sol_m20_tight_reference.npz"):
        r = np.load(os.path.join(GOLDEN, f))
        ours = res["dataPack"][-1]["dataYs"]
        rel = np.abs(ours - r["dataYs"][-1])/np.abs(r["dataYs"][-1])
        assert rel[:, -1].max() < (2e-5 if "tight" in f else 5e-3), (f, rel[:, -1])


def test_n2_methanol_z50_default_vs_reference_bdf(n2_settings):
    """BASELINE config 2 (50 nodes): default tolerance against the reference's BDF run (446 s there)."""
    from rmt_app_b200 import rmtExe
    r = np.load(os.path.join(GOLDEN, "n2_sol_m50_bdf_reference.npz"))
    n2_settings["N2"]["zNo"] = 50
    res = rmtExe(cases.methanol_readme_input("N2"))["resModel"]
    ours = res["dataPack"][-1]["dataYs"]
    rel = np.abs(ours - r["dataYs"][-1])/np.abs(r["dataYs"][-1])
    assert rel[:, -1].max() < 5e-3 and rel[-1].max() < 2e-3


@pytest.mark.parametrize("case", ["methanol", "ch4"])
def test_n2_lanes_per_reactor_give_the_same_solution(n2_settings, case):
    """The integrator spreads the nodes of a reactor over 1..32 lanes; the sequential parts (pressure march,
    block substitution, error sum) are handed from lane to lane in node order, so the result does not depend
    on the lane count beyond the compiler's choice of fused operations (<= 1e-11 relative), including node
    counts that are not a multiple of the lane count."""
    from rmt_app_b200 import engine
    mi = cases.methanol_testfile_input("N2") if case == "methanol" else cases.ch4_input("N2")
    B, zNo = 5, 21
    rng = np.random.default_rng(3)
    T0 = mi["operating-conditions"]["temperature"]
    sw = {"temperature": T0*rng.uniform(0.98, 1.02, B)}
    period = float(mi["operating-conditions"]["period"])
    ref = None
    for lanes, block in ((1, 64), (4, 32), (8, 64), (32, 32)):
        cm = engine.compile_model(mi, block=block, lanes=lanes)
        r = engine.n2_solve_ensemble(cm, mi, sw, B, zNo=zNo, tNo=3, period=period)
        assert (r.status == 0).all(), (lanes, r.status)
        if ref is None:
            ref = r
            continue
        np.testing.assert_allclose(r.out, ref.out, rtol=1e-11, atol=0)
        np.testing.assert_array_equal(r.stats[0], ref.stats[0])
    assert engine.n2_lanes(1, 50) == 32 and engine.n2_lanes(12500, 200) == 8 and engine.n2_lanes(10**7, 50) == 1
    assert engine.n2_lanes(1, 4) == 4


@pytest.mark.parametrize("case,zNo", [("methanol", 21), ("ch4", 12), ("methanol", 50), ("ch4iso", 16)])
def test_n2_stage_pipeline_kernel_equals_the_lanes_kernel(n2_settings, case, zNo):
    """The stage-pipelined mapping (one thread per reactor and pair of Rosenbrock stages, rmt_kernels.cu "stage
    pipeline") integrates the same method with the same per-node arithmetic as the lanes kernel; only the order of the
    linear algebra differs (LU solves instead of products with the explicit inverse blocks): same step sequence, states
    equal to <= 1e-9 — with more reactors than one block holds, a partly filled block and a reactor that fails (NaN
    feed)."""
    from rmt_app_b200 import engine
    mi = {"methanol": lambda: cases.methanol_testfile_input("N2"), "ch4": lambda: cases.ch4_input("N2"),
          "ch4iso": lambda: cases.ch4_input("N2", "iso-thermal")}[case]()
    B = 150
    rng = np.random.default_rng(11)
    T0 = mi["operating-conditions"]["temperature"]
    sw = {"temperature": T0*rng.uniform(0.97, 1.03, B)}
    sw["temperature"][17] = np.nan
    period = float(mi["operating-conditions"]["period"])
    lanes = engine.compile_model(mi, block=64, lanes=8)
    pipe = engine.compile_model(mi, block=256, lanes=0)
    assert pipe is not lanes and pipe.load(0).info.lanes == 0
    a = engine.n2_solve_ensemble(lanes, mi, sw, B, zNo=zNo, tNo=3, period=period)
    b = engine.n2_solve_ensemble(pipe, mi, sw, B, zNo=zNo, tNo=3, period=period)
    ok = np.arange(B) != 17
    assert (a.status[ok] == 0).all() and (b.status[ok] == 0).all() and a.status[17] != 0 and b.status[17] != 0
    assert np.isnan(b.out[..., 17]).all()
    np.testing.assert_array_equal(a.stats[:, ok], b.stats[:, ok])
    np.testing.assert_allclose(b.out[..., ok], a.out[..., ok], rtol=1e-9, atol=0)
    # out_mode 2 (raw | C_i | dataYs rows) through the same kernel
    a2 = engine.n2_solve_ensemble(lanes, mi, sw, B, zNo=zNo, tNo=2, period=period, out_mode=2)
    b2 = engine.n2_solve_ensemble(pipe, mi, sw, B, zNo=zNo, tNo=2, period=period, out_mode=2)
    np.testing.assert_allclose(b2.out[..., ok], a2.out[..., ok], rtol=1e-9, atol=0)
    # the launch-shape policy: full rounds of 148 x 64 reactors go to the pipeline, small or badly filling ensembles do not
    assert engine.n2_use_pipeline(9472, 200) and engine.n2_use_pipeline(100000, 200)
    assert not engine.n2_use_pipeline(12500, 200) and not engine.n2_use_pipeline(500, 200) and not engine.n2_use_pipeline(9472, 4)


def test_n2_stage_pipeline_lanes_pick_up_further_reactors(n2_settings):
    """More reactors than the resident blocks have lanes (148 x 64 = 9 472): lanes that finish draw the next reactor from
    the queue while their neighbours are still integrating — new instance, pending slab output of the old one and the norm
    pass of the new one in the same block step.  Every reactor must come out as from the lanes kernel."""
    from rmt_app_b200 import engine
    mi = cases.ch4_input("N2")
    B, zNo = 9472 + 777, 12
    rng = np.random.default_rng(5)
    sw = {"temperature": rng.uniform(900.0, 1000.0, B), "k0": 7.2e-4*rng.uniform(0.5, 2.0, B)}
    lanes = engine.compile_model(mi, block=64, lanes=8)
    pipe = engine.compile_model(mi, block=256, lanes=0)
    a = engine.n2_solve_ensemble(lanes, mi, sw, B, zNo=zNo, tNo=3, period=5.0, keep_on_device=True)
    oa, sa, ta = a.out.clone(), a.status.clone(), a.stats.clone()
    b = engine.n2_solve_ensemble(pipe, mi, sw, B, zNo=zNo, tNo=3, period=5.0, keep_on_device=True)
    assert bool((sa == 0).all()) and bool((b.status == 0).all())
    assert bool((ta == b.stats).all())
    rel = ((b.out - oa).abs()/oa.abs()).max().item()
    assert rel < 1e-9, rel


def test_n2_stage_pipeline_kernel_against_the_converged_oracle(n2_settings):
    """Config 5's grid (200 nodes, period 0.5 s, 5 slabs) through the stage-pipelined kernel, the fixture instances
    planted in an ensemble of 9 472 reactors (one full round of 148 blocks x 64): temperature profiles and outlets within
    1e-6 of the converged oracle runs — the same bar as the lanes kernel's test below."""
    from rmt_app_b200 import engine
    conv = np.load(os.path.join(GOLDEN, "n2_sol_config5_z200_oracle_tight.npz"))
    zNo, idx = int(conv["zNo"]), [int(i) for i in conv["index"]]
    mi = cases.methanol_readme_input("N2")
    full = cases.config3_sweep(int(conv["B"]), int(conv["seed"]))
    B = 9472
    assert engine.n2_use_pipeline(B, zNo)
    pick = np.r_[idx, np.setdiff1d(np.arange(B + 3), idx)[:B - len(idx)]]       # fixture instances first, then others
    sub = {k: np.ascontiguousarray(np.asarray(v)[pick]) for k, v in full.items()}
    cm = engine.compile_model_n2(mi, B, zNo)
    assert cm.lanes == 0
    r = engine.n2_solve_ensemble(cm, mi, sub, B, zNo=zNo, tNo=5, period=0.5, rtol=1e-8, atol=1e-11, keep_on_device=True)
    assert bool((r.status == 0).all())
    got_all = r.out[..., :len(idx)].cpu().numpy()                       # [tNo][rows][zNo][3]
    for j, i in enumerate(idx):
        for s_ in range(5):
            ref = conv["dataYs"][j][s_]
            rel = np.abs(got_all[s_, :, :, j] - ref)/np.abs(ref)
            assert rel[-1].max() < 1e-6, (i, s_, rel[-1].max())
            assert rel[:, -1].max() < 1e-6, (i, s_, rel[:, -1])
            assert rel.max() < 1e-5, (i, s_, rel.max())


def test_n2_ensemble_matches_single_and_reports_failures(n2_settings):
    from rmt_app_b200 import engine
    mi = cases.ch4_input("N2")
    cm = engine.compile_model(mi)
    B = 70
    rng = np.random.default_rng(1)
    sw = {"temperature": rng.uniform(900, 1000, B), "k0": 7.2e-4*rng.uniform(0.5, 2.0, B)}
    sw["temperature"][5] = np.nan
    r = engine.n2_solve_ensemble(cm, mi, sw, B, zNo=12, tNo=3, period=5.0)
    assert r.status[5] != 0 and (np.delete(r.status, 5) == 0).all()
    assert np.isnan(r.out[:, :, :, 5]).all()
    one = engine.n2_solve_ensemble(cm, mi, {k: v[11:12] for k, v in sw.items()}, 1, zNo=12, tNo=3, period=5.0)
    np.testing.assert_array_equal(one.out[..., 0], r.out[..., 11])


def test_n2_config5_share_properties(n2_settings):
    """BASELINE configs[4] per-GPU share (12 500 reactors x 200 nodes, period 0.5 s, 5 slabs): everything converges,
    the profiles are physical, and a reactor solved inside the ensemble (8 lanes) equals the same reactor solved
    alone (32 lanes) to rounding."""
    from rmt_app_b200 import engine, rmtExeBatchN2
    mi = cases.methanol_readme_input("N2")
    B, zNo = 12500, 200
    sw = cases.config3_sweep(B, 20240613)
    r = rmtExeBatchN2(mi, sw, zNo=zNo, tNo=5)
    assert r["success"].all()
    Y = r["dataYs"]                                            # [B][tNo][nc+1][zNo]
    assert Y.shape == (B, 5, 7, zNo)
    np.testing.assert_allclose(Y[:, :, :6, :].sum(axis=2), 1.0, rtol=1e-12)
    assert (Y[:, :, :6, :] > 0).all()
    T = Y[:, :, 6, :]
    assert T.min() > 440.0 and T.max() < 760.0                                   # reverse water-gas shift cools a few K below the feed
    assert np.median(np.abs(T[:, :, 0] - sw["temperature"][:, None])) < 5.0       # the first node mostly stays near the feed temperature
    st = r["stats"]
    assert 20 < st[0].mean() < 80
    for i in (0, 6789, B - 1):
        one = rmtExeBatchN2(mi, {k: v[i:i + 1] for k, v in sw.items()}, zNo=zNo, tNo=5)
        np.testing.assert_allclose(one["dataYs"][0], Y[i], rtol=1e-10, atol=0)
    assert engine.n2_lanes(B, zNo) == 8 and engine.n2_lanes(1, zNo) == 32


def test_n2_isothermal_parity(n2_settings):
    """N2 with process-type "iso-thermal" (nc unknowns per node): RHS against the reference fixture, rmtExe against
    the converged oracle run (1e-6) and against the reference's default run, several lanes per reactor."""
    from rmt_app_b200 import engine, rmtExe
    g = np.load(os.path.join(GOLDEN, "n2_iso_reference.npz"))
    z = int(g["zNo"])
    mi = cases.ch4_input("N2", "iso-thermal")
    cm = engine.compile_model(mi)
    assert cm.spec.iso and cm.spec.n == 3
    Y, F = g["rhs_Y"], g["rhs_F"]
    Fg = engine.n2_rhs_batch(cm, mi, Y, z)
    Fr, Fq = F.reshape(len(F), 3, z), Fg.reshape(len(F), 3, z)
    scale = np.max(np.abs(Fr), axis=1, keepdims=True)
    assert np.max(np.abs(Fq - Fr)/scale) < 2e-9
    n2_settings["N2"]["zNo"] = z
    O.solverSetting["N2"]["zNo"] = z
    want = O.rmtExe(mi, method="LSODA", rtol=1e-11, atol=1e-13)["resModel"]["dataPack"]
    tight = dict(mi); tight["solver-config"] = dict(mi["solver-config"], rtol=1e-9, atol=1e-12)
    ours = rmtExe(tight)["resModel"]["dataPack"]
    assert len(ours) == 5
    for a, b in zip(ours, want):
        assert a["dataYs"].shape == b["dataYs"].shape
        np.testing.assert_allclose(a["dataYs"], b["dataYs"], rtol=1e-6)
        np.testing.assert_allclose(a["dataYCons2"], b["dataYCons2"], rtol=1e-6)
    dflt = rmtExe(mi)["resModel"]["dataPack"]
    for i, d in enumerate(dflt):
        np.testing.assert_allclose(d["dataYs"], g["default__dataYs"][i], rtol=5e-3)
    ref = None
    for lanes, block in ((1, 64), (8, 64), (32, 32)):
        c2 = engine.compile_model(mi, block=block, lanes=lanes)
        r = engine.n2_solve_ensemble(c2, mi, None, 1, zNo=z, tNo=5, period=10.0, rtol=1e-8, atol=1e-11)
        assert r.status[0] == 0
        ref = r if ref is None else ref
        np.testing.assert_allclose(r.out, ref.out, rtol=1e-11)
    # ensemble form: the temperature row of the iso-thermal dataYs is each reactor's own feed temperature
    from rmt_app_b200 import rmtExeBatchN2
    sw = {"temperature": np.array([960.0, 973.0, 985.0])}
    rb = rmtExeBatchN2(mi, sw, zNo=z, tNo=5)
    assert rb["dataYs"].shape == (3, 5, 4, z) and rb["success"].all()
    np.testing.assert_array_equal(rb["dataYs"][:, :, 3, :], np.broadcast_to(sw["temperature"].reshape(3, 1, 1), (3, 5, z)))
    np.testing.assert_allclose(rb["dataYs"][1, :, :3, :], np.array([d["dataYs"][:3] for d in dflt]), rtol=1e-12)


def _slab_errors(dp, ref):
    rel = np.abs(dp["dataYs"] - ref)/np.abs(ref)
    return rel[-1].max(), rel[:, -1].max(), rel.max()


def test_n2_config2_z50_tight_against_converged_oracle(n2_settings):
    """BASELINE config 2 (README inputs, 50 axial nodes) at the parity-grade tolerance: outlet mole fractions and the
    temperature profile of all five slabs within 1e-6 of the converged oracle run (BDF rtol 1e-9 / atol 1e-12,
    tests/golden/n2_sol_m50_oracle_tight.npz; the oracle RHS is pinned to modelEquationN2 at 50 nodes to 1e-12)."""
    from rmt_app_b200 import rmtExe
    conv = np.load(os.path.join(GOLDEN, "n2_sol_m50_oracle_tight.npz"))
    assert int(conv["zNo"]) == 50
    n2_settings["N2"]["zNo"] = 50
    mi = cases.methanol_readme_input("N2")
    mi["solver-config"].update(rtol=1e-8, atol=1e-11)
    res = rmtExe(mi)["resModel"]
    assert len(res["dataPack"]) == 5
    for i, dp in enumerate(res["dataPack"]):
        eT, eOut, eAll = _slab_errors(dp, conv["dataYs"][i])
        assert eT < 1e-6, (i, eT)                # temperature profile, all 50 nodes
        assert eOut < 1e-6, (i, eOut)            # outlet mole fractions + outlet temperature
        assert eAll < 1e-5, (i, eAll)            # every species at every node (trace species near 1e-5 mole fraction)
        np.testing.assert_allclose(dp["dataTime"], conv["dataTime"][i])


def test_n2_ch4_tight_against_converged_oracle(n2_settings):
    """Methane coupling (nc = 3, one reaction), 20 nodes: all slabs within 1e-6 of the converged ORACLE run
    (n2_sol_ch4_oracle_tight.npz) — the twin of the check against the reference's own tight run above."""
    from rmt_app_b200 import rmtExe
    conv = np.load(os.path.join(GOLDEN, "n2_sol_ch4_oracle_tight.npz"))
    n2_settings["N2"]["zNo"] = int(conv["zNo"])
    mi = cases.ch4_input("N2")
    mi["solver-config"].update(rtol=1e-9, atol=1e-12)
    res = rmtExe(mi)["resModel"]
    for i, dp in enumerate(res["dataPack"]):
        np.testing.assert_allclose(dp["dataYs"], conv["dataYs"][i], rtol=1e-6)


def test_n2_config5_z200_against_converged_oracle(n2_settings):
    """BASELINE config 5's grid (200 nodes, period 0.5 s, 5 slabs): three instances of the per-GPU share (seed 20240613,
    indices 0 / 6789 / 12499) against converged oracle runs (tests/golden/n2_sol_config5_z200_oracle_tight.npz, made by
    make_golden_oracle.py config5).  The instances are solved INSIDE an ensemble large enough for the 8-lane launch shape
    config 5 uses (25 node groups of 8, hand-over between groups through shared memory), and alone (32 lanes)."""
    from rmt_app_b200 import engine, rmtExeBatchN2
    conv = np.load(os.path.join(GOLDEN, "n2_sol_config5_z200_oracle_tight.npz"))
    zNo, idx = int(conv["zNo"]), [int(i) for i in conv["index"]]
    assert zNo == 200
    mi = cases.methanol_readme_input("N2")
    full = cases.config3_sweep(int(conv["B"]), int(conv["seed"]))
    # 12 500-reactor launch shape (8 lanes per reactor) with the three fixture instances planted among other reactors
    B = 12500
    assert engine.n2_lanes(B, zNo) == 8
    r = rmtExeBatchN2(mi, full, zNo=zNo, tNo=5, rtol=1e-8, atol=1e-11)
    assert r["success"].all()
    for j, i in enumerate(idx):
        got = r["dataYs"][i]                                          # [tNo][nc+1][zNo]
        for s in range(5):
            ref = conv["dataYs"][j][s]
            rel = np.abs(got[s] - ref)/np.abs(ref)
            assert rel[-1].max() < 1e-6, (i, s, rel[-1].max())        # temperature profile, all 200 nodes
            assert rel[:, -1].max() < 1e-6, (i, s, rel[:, -1])        # outlet
            assert rel.max() < 1e-5, (i, s, rel.max())
    # the same reactors alone (32 lanes per reactor): same solution to rounding
    one = rmtExeBatchN2(mi, {k: v[idx] for k, v in full.items()}, zNo=zNo, tNo=5, rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(one["dataYs"], r["dataYs"][idx], rtol=1e-9)
