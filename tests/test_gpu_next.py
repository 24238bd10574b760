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


def test_m7_dimensional_twin_parity():
    """Model M7 (pbReactor.runM3): RHS against the reference's modelEquationM3, Jacobian against differences,
    solution against the reference at tight tolerance (1e-6) and at default tolerance (its own scatter)."""
    import os
    import pyremot_oracle as O
    from conftest import GOLDEN
    from rmt_app_b200 import engine, rmtExe, rmtExeBatch
    g = np.load(os.path.join(GOLDEN, "m7_reference.npz"))
    mi = cases.methanol_m7_input()
    cm = engine.compile_model(mi)
    assert cm.spec.n == 8 and cm.spec.model == "M7"
    Y, F = g["rhs_Y"], g["rhs_F"]
    Fg, J, _ = engine.n1_rhs_batch(cm, mi, Y, jac=True)
    o = O.M7Oracle(mi)
    rng = np.random.default_rng(3)
    for y, f, fg in zip(Y, F, Fg):
        dev = 0.0
        for _ in range(6):
            dev = max(dev, np.max(np.abs(np.array(o.rhs(0, y*(1 + 2.2e-16*rng.choice([-1, 0, 1], size=8)))) - f)))
        assert np.max(np.abs(fg - f)) <= 1e-13*np.max(np.abs(f)) + 50*dev
    for y, Jg in zip(Y[[0, 8, 12]], J[[0, 8, 12]]):
        Jfd = np.zeros((8, 8))
        for j in range(8):
            h = 1e-6*max(abs(y[j]), 1e-3)
            yp, ym = y.copy(), y.copy()
            yp[j] += h; ym[j] -= h
            Jfd[:, j] = (np.array(o.rhs(0, yp)) - np.array(o.rhs(0, ym)))/(2*h)
        rowscale = np.max(np.abs(Jfd), axis=1, keepdims=True)
        assert np.max(np.abs(Jg - Jfd)/rowscale) < 5e-6
    tight = dict(mi); tight["solver-config"] = dict(mi["solver-config"], rtol=1e-9, atol=1e-12)
    res = rmtExe(tight)["resModel"]
    assert res["dataYs"].shape == g["tight__dataYs"].shape == (7, 30)
    assert np.max(np.abs(res["dataYs"] - g["tight__dataYs"])/np.abs(g["tight__dataYs"])) < 1e-6
    assert len(res["XYList"]) == 7 and res["dataList"][-1]["leg"] == "Temperature"
    np.testing.assert_allclose(res["dataPressure"], g["tight__soly"][7], rtol=1e-8)
    dflt = rmtExe(mi)["resModel"]["dataYs"][:, -1]
    conv = g["tight__dataYs"][:, -1]
    theirs = np.max(np.abs(g["default__dataYs"][:, -1] - conv)/np.abs(conv))
    assert np.max(np.abs(dflt - conv)/np.abs(conv)) < max(3*theirs, 2e-3)
    # ensemble form
    B = 300
    sw = cases.config3_sweep(B, seed=6)
    r = rmtExeBatch(mi, sw, rtol=1e-8, atol=1e-11)
    assert r["success"].all() and r["labelList"][-2:] == ["Temperature", "Pressure"]
    i = 123
    want = O.rmtExe(cases.instance_input(mi, sw, i), method="LSODA", rtol=1e-10, atol=1e-12)["resModel"]["solY"][:, -1]
    got = r["dataYs"][i]
    np.testing.assert_allclose(got[6:], want[6:], rtol=1e-6)
    np.testing.assert_allclose(got[:6], want[:6]/want[:6].sum(), rtol=1e-6)


def test_m9_dimensional_dynamic_twin_parity():
    """Model M9 (pbReactor.runM5): RHS against the reference's modelEquationM5 (fixture), solution at tight tolerance
    against the converged oracle run (1e-6: a wrong Jacobian block would cost the Rosenbrock method its order and show
    here), default tolerance against the reference's own default run, rmtExe's plot lists, and the ensemble form."""
    import os
    import pyremot_oracle as O
    from conftest import GOLDEN
    from rmt_app_b200 import engine, rmtExe, rmtExeBatchN2, solverSetting
    g = np.load(os.path.join(GOLDEN, "m9_reference.npz"))
    t = np.load(os.path.join(GOLDEN, "m9_sol_oracle_tight.npz"))
    mi = cases.methanol_m9_input()
    cm = engine.compile_model(mi)
    assert cm.spec.model == "M9" and cm.spec.n == 7 and cm.lanes == 1
    zNo, tNo = 12, 3
    Y, F = g["rhs_Y"], g["rhs_F"]
    Fg = engine.n2_rhs_batch(cm, mi, Y, zNo)
    Fr, Fq = F.reshape(len(F), 7, zNo), Fg.reshape(len(F), 7, zNo)
    scale = np.max(np.abs(Fr), axis=1, keepdims=True)
    assert np.max(np.abs(Fq - Fr)/scale) < 2e-9
    assert np.max(np.abs(Fq[0] - Fr[0])/np.maximum(np.abs(Fr[0]), 1e-9*scale[0])) < 1e-10
    old = dict(solverSetting["S2"])
    solverSetting["S2"].update(zNo=zNo, tNo=tNo)
    try:
        tight = dict(mi); tight["solver-config"] = dict(mi["solver-config"], rtol=1e-9, atol=1e-12)
        res = rmtExe(tight)["resModel"]
        ours = np.array([d["dataYs"] for d in res["dataPack"]])
        assert ours.shape == t["dataYs"].shape == (tNo, 7, zNo)
        assert np.max(np.abs(ours - t["dataYs"])/np.abs(t["dataYs"])) < 1e-6
        assert len(res["XYList"]) == tNo and [d["leg"] for d in res["dataList"]] == list(g["default__legends"])
        np.testing.assert_array_equal(res["XYList"][0][0], g["default__x"])
        dflt = rmtExe(mi)["resModel"]
        Td = np.array([xy[1] for xy in dflt["XYList"]])
        theirs = np.max(np.abs(g["default__T_profiles"] - t["dataYs"][:, 6])/t["dataYs"][:, 6])
        assert np.max(np.abs(Td - t["dataYs"][:, 6])/t["dataYs"][:, 6]) < max(3*theirs, 2e-3)
        # ensemble form: a few inlet temperatures, one checked against the oracle
        B = 6
        sw = {"temperature": 523.0 + np.arange(B)*2.0}
        r = rmtExeBatchN2(mi, sw, rtol=1e-8, atol=1e-11)
        assert r["success"].all() and r["dataYs"].shape == (B, tNo, 7, zNo)
        O.solverSetting["S2"].update(zNo=zNo, tNo=tNo)
        one = dict(mi); one["operating-conditions"] = dict(mi["operating-conditions"], temperature=float(sw["temperature"][4]))
        want = O.rmtExe(one, method="BDF", rtol=1e-10, atol=1e-13)["resModel"]["dataPack"]
        for i in range(tNo):
            np.testing.assert_allclose(r["dataYs"][4, i], want[i]["dataYs"], rtol=1e-6)
    finally:
        solverSetting["S2"].update(old)
        O.solverSetting["S2"].update(tNo=10, zNo=100)


def test_m9_stage_pipeline_equals_one_lane_per_reactor():
    """M9 through the stage-pipelined kernel (four threads per reactor: the velocity march and its linearisation are
    chains private to a thread, rmt_kernels.cu "stage pipeline") against the lanes kernel with one lane per reactor: same
    step sequence, states equal to <= 1e-9, for node counts below and above the ring depth; the single reactor of the
    fixture stays within 1e-6 of the converged oracle run; ensembles of a few hundred reactors take the pipeline."""
    import os
    from conftest import GOLDEN
    from rmt_app_b200 import engine
    mi = cases.methanol_m9_input()
    B = 100
    rng = np.random.default_rng(4)
    sw = {"temperature": 523.0 + rng.uniform(-8.0, 8.0, B), "pressure": mi["operating-conditions"]["pressure"]*rng.uniform(0.9, 1.1, B)}
    period = float(mi["operating-conditions"]["period"])
    one = engine.compile_model(mi, block=64, lanes=1)
    pipe = engine.compile_model(mi, block=256, lanes=0)
    assert pipe.load(0).info.lanes == 0 and pipe.load(0).info.model == 9
    for zNo in (12, 37):
        a = engine.n2_solve_ensemble(one, mi, sw, B, zNo=zNo, tNo=3, period=period, rtol=1e-6, atol=1e-9)
        b = engine.n2_solve_ensemble(pipe, mi, sw, B, zNo=zNo, tNo=3, period=period, rtol=1e-6, atol=1e-9)
        assert (a.status == 0).all() and (b.status == 0).all()
        np.testing.assert_array_equal(a.stats, b.stats)
        np.testing.assert_allclose(b.out, a.out, rtol=1e-9, atol=0)
    t = np.load(os.path.join(GOLDEN, "m9_sol_oracle_tight.npz"))
    r = engine.n2_solve_ensemble(pipe, mi, None, 1, zNo=12, tNo=3, period=period, rtol=1e-9, atol=1e-12)
    assert r.status[0] == 0
    assert np.max(np.abs(r.out[..., 0] - t["dataYs"])/np.abs(t["dataYs"])) < 1e-6
    assert engine.compile_model_n2(mi, 1000, 50).lanes == 0 and engine.compile_model_n2(mi, 6, 12).lanes == 1


def test_full_state_integrator_when_reactions_do_not_outnumber_species():
    """nr >= nc: no reaction-extent form (codegen.use_extents), the integrator works on the full state; the solution
    must match the oracle on the same synthetic three-reaction methane case, for the outlet-only (Ros4) and the
    profile (Rodas4) paths."""
    import pyremot_oracle as O
    from rmt_app_b200 import engine, rmtExe, rmtExeBatch
    mi = cases.ch4_three_reaction_input("N1")
    cm = engine.compile_model(mi)
    assert not cm.reduced and cm.m == cm.spec.n == 4 and "#define RMT_REDUCED 0" in cm.header
    ref = O.rmtExe(mi, method="LSODA", rtol=1e-11, atol=1e-13)["resModel"][0]["dataYs"]
    tight = dict(mi); tight["solver-config"] = dict(mi["solver-config"], rtol=1e-9, atol=1e-12)
    ours = rmtExe(tight)["resModel"][0]["dataYs"]
    assert np.max(np.abs(ours - ref)/np.abs(ref)) < 1e-6
    B = 64
    sw = {"temperature": np.linspace(940, 1000, B)}
    r = rmtExeBatch(mi, sw)                                  # default tolerance, outlet only -> Ros4
    assert r["success"].all()
    i = 40
    one = dict(mi); one["operating-conditions"] = dict(mi["operating-conditions"], temperature=float(sw["temperature"][i]))
    want = O.rmtExe(one, method="LSODA", rtol=1e-11, atol=1e-13)["resModel"][0]["dataYs"][:, -1]
    np.testing.assert_allclose(r["dataYs"][i], want, rtol=5e-3)
    r9 = rmtExeBatch(mi, sw, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(r9["dataYs"][i], want, rtol=1e-6)


def test_integration_stub_with_raw_ctypes():
    """The binding INTEGRATION.md shows a maintainer (raw ctypes against include/rmt_b200.h, no rmt_app_b200.capi):
    compile, load, rmt_setup, rmt_n1_solve — and the result equals rmtExe's."""
    import ctypes as C
    import os
    import torch
    from rmt_app_b200 import engine, rmtExe
    from rmt_app_b200.model import ModelSpec
    from rmt_app_b200.codegen import generate_model_header
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rmt_app_b200")
    lib = C.CDLL(os.path.join(root, "librmtb200.so"))
    lib.rmt_last_error.restype = C.c_char_p

    def ck(rc):
        if rc:
            raise RuntimeError(lib.rmt_last_error().decode())
    ck(lib.rmt_init(torch.cuda.current_device()))
    mi = cases.methanol_readme_input("N1")
    spec = ModelSpec(mi)
    blob, mod = C.c_uint64(), C.c_uint64()
    src = open(os.path.join(root, "csrc", "rmt_kernels.cu")).read()
    ck(lib.rmt_nvrtc_compile(generate_model_header(spec).encode(), src.encode(), b"sm_100a", 256, None, 0, C.byref(blob)))
    data, size = C.c_void_p(), C.c_size_t()
    ck(lib.rmt_blob_data(blob, C.byref(data), C.byref(size)))
    ck(lib.rmt_module_load(data, size, C.byref(mod)))
    nin = spec.nin
    row_map = (C.c_int32*nin)(*([-1]*nin))
    uniform = (C.c_double*nin)(*engine.uniform_inputs(spec, mi))
    consts = torch.empty((29 + spec.nc + spec.nkp, 1), dtype=torch.float64, device="cuda")
    ck(lib.rmt_setup(mod, C.c_int64(1), None, 0, row_map, uniform, C.c_void_p(consts.data_ptr()), None))
    z = np.linspace(0, 1, 101)
    out = torch.empty((101, 2*spec.n + spec.nc, 1), dtype=torch.float64, device="cuda")
    status = torch.empty(1, dtype=torch.int32, device="cuda")
    stats = torch.empty((4, 1), dtype=torch.int32, device="cuda")
    ck(lib.rmt_n1_solve(mod, C.c_int64(1), C.c_void_p(consts.data_ptr()), 101, z.ctypes.data_as(C.POINTER(C.c_double)),
                        C.c_double(1e-3), C.c_double(1e-6), 100000, 1, 2,
                        C.c_void_p(out.data_ptr()), C.c_void_p(status.data_ptr()), C.c_void_p(stats.data_ptr()),
                        None, None, None, None))
    torch.cuda.synchronize()
    assert int(status.cpu()[0]) == 0
    dataYs = out.cpu().numpy()[:, spec.n + spec.nc:, 0].T                           # rows per point: raw | C_i | dataYs
    ref = rmtExe(mi)["resModel"][0]["dataYs"]
    np.testing.assert_allclose(dataYs, ref, rtol=1e-12)
    ck(lib.rmt_blob_free(blob)); ck(lib.rmt_module_free(mod))
