"""GPU parity tests for the N1 path (run with -m gpu on the B200 box).

Level 1: RHS / constants / Jacobian against the oracle and the reference's golden vectors.
Level 2: tight-tolerance solutions against the reference run at rtol=1e-10/atol=1e-12
         (north_star tolerance: relative 1e-6 on outlet mole fractions and T profiles).
Level 3: default-tolerance solutions within the reference's own default-vs-converged error.
Every call goes through the C ABI (rmt_app_b200.capi -> librmtb200.so)."""
import numpy as np
import pytest

import cases
import pyremot_oracle as O

pytestmark = pytest.mark.gpu

N1_CASES = {
    "methanol_readme": lambda: cases.methanol_readme_input("N1"),
    "methanol_testfile": lambda: cases.methanol_testfile_input("N1"),
    "ch4_noniso": lambda: cases.ch4_input("N1", "non-iso-thermal"),
    "ch4_iso": lambda: cases.ch4_input("N1", "iso-thermal"),
}
TIGHT = dict(rtol=1e-9, atol=1e-12)


def _engine():
    from rmt_app_b200 import engine
    return engine


def _ulp_sensitivity(o, y, f, rng):
    """How much does the reference RHS itself move under 1-ulp input perturbations?
    (near chemical equilibrium the rates cancel and amplify rounding)."""
    dev = 0.0
    for _ in range(6):
        fp = np.array(o.rhs(0.0, y*(1 + 2.2e-16*rng.choice([-1, 0, 1], size=y.size))))
        dev = max(dev, np.max(np.abs(fp - f)))
    return dev


@pytest.mark.parametrize("name", list(N1_CASES))
def test_rhs_parity_with_reference_golden(golden_n1, name):
    eng = _engine()
    mi = N1_CASES[name]()
    cm = eng.compile_model(mi)
    Y, F = golden_n1[name + "__rhs_Y"], golden_n1[name + "__rhs_F"]
    Fg, _, consts = eng.n1_rhs_batch(cm, mi, Y)
    o = O.N1Oracle(mi)
    rng = np.random.default_rng(5)
    for y, f, fg in zip(Y, F, Fg):
        tol = 1e-13*np.max(np.abs(f)) + 50*_ulp_sensitivity(o, y, f, rng)
        assert np.max(np.abs(fg - f)) <= tol, (name, y, fg - f, tol)
    # well-conditioned states (feed + perturbed): plain relative agreement
    assert np.max(np.abs(Fg[0] - F[0])/np.abs(F[0])) < 1e-11
    # setup constants written by rmt_setup vs the reference's paramsSet
    g = lambda k: float(golden_n1[name + "__" + k])
    np.testing.assert_allclose(consts[10, 0], g("const_GaMiVi"), rtol=1e-13)      # Wilke mixture viscosity
    np.testing.assert_allclose(consts[9, 0], g("da_GaHeCoTe0"), rtol=1e-13)
    np.testing.assert_allclose(consts[8, 0], golden_n1[name + "__da_GaMaCoTe0"][0], rtol=1e-13)
    np.testing.assert_allclose(consts[6, 0], g("bc_GaDe0"), rtol=1e-13)
    np.testing.assert_allclose(consts[7, 0], g("bc_GaCpMeanMix0"), rtol=1e-13)
    np.testing.assert_allclose(consts[3, 0], g("bc_SpCo0"), rtol=1e-14)


@pytest.mark.parametrize("name", ["methanol_readme", "ch4_iso", "ch4_noniso"])
def test_analytic_jacobian_against_oracle_differences(golden_n1, name):
    eng = _engine()
    mi = N1_CASES[name]()
    cm = eng.compile_model(mi)
    o = O.N1Oracle(mi)
    Y = golden_n1[name + "__rhs_Y"][[0, 10, 13, 20, 25]]
    if name == "ch4_noniso":
        Y = Y[:3]        # later states sit at T -> 0 K where differences are meaningless
    _, J, _ = eng.n1_rhs_batch(cm, mi, Y, jac=True)
    for y, Jg in zip(Y, J):
        n = y.size
        Jfd = np.zeros((n, n))
        for j in range(n):
            h = 1e-6*max(abs(y[j]), 1e-3)
            yp, ym = y.copy(), y.copy()
            yp[j] += h; ym[j] -= h
            Jfd[:, j] = (np.array(o.rhs(0, yp)) - np.array(o.rhs(0, ym)))/(2*h)
        assert np.max(np.abs(Jg - Jfd)) < 2e-6*np.max(np.abs(Jfd)), name


@pytest.mark.parametrize("name", list(N1_CASES))
def test_reaction_extent_system_is_the_projected_jacobian(golden_n1, name):
    """The integrator solves in reaction extents (nr + 2 unknowns) when nr < nc.  With E = [[nu^T, 0], [0, I]]
    its g and A = dg/dx must satisfy E g == f and E A == J E exactly (up to rounding): then the Rosenbrock
    iterates are those of the full system."""
    eng = _engine()
    mi = N1_CASES[name]()
    cm = eng.compile_model(mi)
    spec = cm.spec
    assert cm.reduced and cm.m == spec.nr + spec.n - spec.nc
    Y = golden_n1[name + "__rhs_Y"][[0, 10, 13, 20]]
    F, J, _ = eng.n1_rhs_batch(cm, mi, Y, jac=True)
    g, A, _ = eng.n1_rhs_batch(cm, mi, Y, system=True)
    nx = spec.n - spec.nc
    E = np.zeros((spec.n, cm.m))
    E[:spec.nc, :spec.nr] = spec.nu.T
    E[spec.nc:, spec.nr:] = np.eye(nx)
    for f, Jf, gi, Ai in zip(F, J, g, A):
        np.testing.assert_allclose(E @ gi, f, rtol=0, atol=1e-13*np.max(np.abs(f)))
        lhs, rhs = E @ Ai, Jf @ E
        assert np.max(np.abs(lhs - rhs)) <= 1e-12*np.max(np.abs(rhs)), (name, lhs - rhs)
    # the full-state build of the same model is still available and agrees to integration tolerance
    cf = eng.compile_model(mi, reduced=False)
    assert not cf.reduced and cf.m == spec.n
    sweep = {"temperature": np.array([mi["operating-conditions"]["temperature"]]*3)*np.array([1.0, 1.01, 0.99])}
    a = eng.n1_solve_ensemble(cm, mi, sweep, B=3, rtol=1e-9, atol=1e-12)
    b = eng.n1_solve_ensemble(cf, mi, sweep, B=3, rtol=1e-9, atol=1e-12)
    assert (a.status == 0).all() and (b.status == 0).all()
    np.testing.assert_allclose(a.out, b.out, rtol=2e-8, atol=1e-12)
    assert np.max(np.abs(a.stats[0] - b.stats[0])) <= 2      # same step sequence, up to rounding at the accept test


@pytest.mark.parametrize("name", list(N1_CASES))
def test_rmtexe_tight_matches_reference_tight(golden_n1, name):
    """Level 2: outlet mole fractions and the 101-point T profile within 1e-6 (north_star)."""
    from rmt_app_b200 import rmtExe
    mi = N1_CASES[name]()
    mi["solver-config"].update(TIGHT)
    dp = rmtExe(mi)["resModel"][0]
    ref = golden_n1[name + "__tight_LSODA__dataYs"]
    assert dp["dataYs"].shape == ref.shape
    rel = np.abs(dp["dataYs"] - ref)/np.abs(ref)
    tol = 1e-6
    assert rel[:, -1].max() < tol, rel[:, -1]
    assert rel.max() < tol, rel.max()
    # result schema of runN1 (pbHomoReactor.py:2991-3007)
    for k in ("dataXs", "dataYCons1", "dataYCons2", "dataYTemp1", "dataYTemp2"):
        assert np.asarray(dp[k]).shape == golden_n1[name + "__" + k].shape, k
    assert dp["labelList"] == list(golden_n1[name + "__labelList"])
    assert dp["indexList"] == list(golden_n1[name + "__indexList"])
    assert dp["modelId"] == "N1" and dp["successStatus"] is True and dp["dataTime"] == []
    soly = golden_n1[name + "__tight_LSODA__soly"]
    nc = len(mi["feed"]["components"]["shell"])
    np.testing.assert_allclose(dp["dataYCons1"], soly[:nc], rtol=2e-6, atol=1e-12)


@pytest.mark.parametrize("name", list(N1_CASES))
def test_rmtexe_default_tolerance_level3(golden_n1, name):
    """Level 3: at SciPy's default tolerances we must be no further from the converged
    solution than the reference's own LSODA/BDF/Radau runs are."""
    from rmt_app_b200 import rmtExe
    dp = rmtExe(N1_CASES[name]())["resModel"][0]
    tight = golden_n1[name + "__tight_LSODA__dataYs"]
    ours = np.abs(dp["dataYs"][:, -1] - tight[:, -1])/np.abs(tight[:, -1])
    theirs = max(np.max(np.abs(golden_n1["%s__%s__dataYs" % (name, m)][:, -1] - tight[:, -1])/np.abs(tight[:, -1]))
                 for m in ("default", "BDF", "Radau"))
    assert ours.max() < max(3*theirs, 2e-3), (ours, theirs)


def test_config3_corners_tight_and_default(golden_corners):
    from rmt_app_b200 import rmtExeBatch
    base = cases.methanol_readme_input("N1")
    sw = cases.config3_corners()
    tight = golden_corners["tight_dataYs"]          # [36][8][101]
    r = rmtExeBatch(base, sw, profile=True, **TIGHT)
    assert r["success"].all()
    rel = np.abs(r["dataYs"] - tight)/np.abs(tight)
    assert rel[:, :, -1].max() < 1e-6               # outlets
    assert rel[:, -1, :].max() < 1e-6               # temperature profiles
    assert rel.max() < 5e-6                         # every species at every point (dense output of trace species)
    # default tolerance: outlet error distribution no worse than the reference's LSODA
    d = rmtExeBatch(base, sw)
    assert d["success"].all()
    ours = (np.abs(d["dataYs"] - tight[:, :, -1])/np.abs(tight[:, :, -1])).max(axis=1)
    theirs = (np.abs(golden_corners["default_dataYs"][:, :, -1] - tight[:, :, -1])/np.abs(tight[:, :, -1])).max(axis=1)
    assert np.median(ours) <= 2*np.median(theirs) and ours.max() <= theirs.max(), (np.median(ours), ours.max(), theirs.max())


def test_random_sweep_against_oracle_port():
    from rmt_app_b200 import rmtExeBatch
    base = cases.methanol_readme_input("N1")
    B = 4096
    sw = cases.config3_sweep(B, seed=99)
    r = rmtExeBatch(base, sw, **TIGHT)
    assert r["success"].all()
    np.testing.assert_allclose(r["dataYs"][:, :6].sum(axis=1), 1.0, rtol=1e-13)
    for i in (0, 1234, 4095):
        want = O.rmtExe(cases.instance_input(base, sw, i), method="LSODA", rtol=1e-10, atol=1e-12)["resModel"][0]["dataYs"][:, -1]
        assert np.max(np.abs(r["dataYs"][i] - want)/np.abs(want)) < 1e-6


def test_kinetic_parameter_sweep_and_objective():
    """Config-4 style: Arrhenius parameters as per-instance VARS slots, fused objective, reduction."""
    from rmt_app_b200 import engine
    base = cases.methanol_readme_input("N1")
    base["reaction-rates"] = cases.methanol_kinetics_param(1171.2)
    B = 2048
    pop = cases.config4_population(B)
    cm = engine.compile_model(base)
    nominal = engine.n1_solve_ensemble(cm, base, None, 1, **TIGHT).out[0, :, 0]
    # nominal parameters through the parametrised kinetics == fixed-constant kinetics
    plain = engine.n1_solve_ensemble(engine.compile_model(cases.methanol_readme_input("N1")),
                                     cases.methanol_readme_input("N1"), None, 1, **TIGHT).out[0, :, 0]
    np.testing.assert_allclose(nominal, plain, rtol=1e-9)
    res = engine.n1_solve_ensemble(cm, base, pop, B, objective_ref=nominal, keep_on_device=True, **TIGHT)
    out = res.out.cpu().numpy()[0]                    # [n][B]
    obj = res.objective.cpu().numpy()
    idx = [0, 1, 2, 3, 4, 5, 7]
    want = (((out[idx] - nominal[idx, None])/nominal[idx, None])**2).sum(axis=0)
    np.testing.assert_allclose(obj, want, rtol=1e-12)
    s, mn, am = cm.module.reduce_objective(B, res.objective, index_offset=1000)
    np.testing.assert_allclose(s, obj.sum(), rtol=1e-12)
    assert mn == obj.min() and am == 1000 + int(obj.argmin())
    # one perturbed instance against the oracle with the same parameters
    i = 77
    want = O.rmtExe(cases.instance_input(base, pop, i), method="LSODA", rtol=1e-10, atol=1e-12)["resModel"][0]["dataYs"][:, -1]
    assert np.max(np.abs(out[:, i] - want)/np.abs(want)) < 1e-6


def test_edge_cases_and_failure_reporting():
    from rmt_app_b200 import rmtExeBatch, engine
    base = cases.methanol_readme_input("N1")
    # ragged sizes around the block size; a single instance; results independent of batch composition
    sw = cases.config3_sweep(131, seed=5)
    full = rmtExeBatch(base, sw)
    for B in (1, 31, 129):
        part = rmtExeBatch(base, {k: v[:B] for k, v in sw.items()})
        np.testing.assert_array_equal(part["dataYs"], full["dataYs"][:B])      # bit-identical
    again = rmtExeBatch(base, sw)
    np.testing.assert_array_equal(again["dataYs"], full["dataYs"])             # deterministic
    # one poisoned instance does not abort its neighbours
    bad = {k: v.copy() for k, v in sw.items()}
    bad["concentration"][7, 0] = -5.0
    bad["temperature"][9] = np.nan
    r = rmtExeBatch(base, bad)
    assert r["status"][7] != 0 and r["status"][9] != 0
    ok = np.ones(131, bool); ok[[7, 9]] = False
    assert r["success"][ok].all()
    np.testing.assert_array_equal(r["dataYs"][ok], full["dataYs"][ok])
    assert np.isnan(r["dataYs"][7]).all()
    # step budget exhausted -> status 1
    r = rmtExeBatch(base, sw, max_steps=5)
    assert (r["status"] == 1).all()
    # no-dense mode lands on the output points: must agree with dense output at tight tolerance
    z = np.linspace(0, 1, 11)
    a = rmtExeBatch(base, {k: v[:8] for k, v in sw.items()}, z_eval=z, dense=True, **TIGHT)
    b = rmtExeBatch(base, {k: v[:8] for k, v in sw.items()}, z_eval=z, dense=False, **TIGHT)
    np.testing.assert_allclose(a["dataYs"], b["dataYs"], rtol=1e-7)
    # argument validation errors come from the C ABI, loudly
    from rmt_app_b200.capi import RmtError
    with pytest.raises(RmtError, match="strictly increasing"):
        rmtExeBatch(base, sw, z_eval=np.array([0.5, 0.5, 1.0]))
    with pytest.raises(KeyError):
        rmtExeBatch(base, {"nonsense": np.ones(4)})


def test_host_buffer_entry_point_matches_device_path():
    from rmt_app_b200 import engine
    base = cases.methanol_readme_input("N1")
    B = 777
    sw = cases.config3_sweep(B, seed=3)
    cm = engine.compile_model(base)
    dev = engine.n1_solve_ensemble(cm, base, sw, B)
    rows, row_map = engine.sweep_rows(cm.spec, sw, B)
    uniform = engine.uniform_inputs(cm.spec, base)
    out = np.empty((1, cm.spec.n, B)); status = np.empty(B, np.int32); stats = np.empty((4, B), np.int32)
    cm.module.n1_solve_host(B, rows, rows.shape[0], row_map, uniform, np.array([1.0]), 1e-3, 1e-6, out, status, stats)
    np.testing.assert_array_equal(out, dev.out)
    np.testing.assert_array_equal(status, dev.status)
    np.testing.assert_array_equal(stats, dev.stats)
    # >= 2^18 reactors: the library pipelines three chunks on two streams; same bits, with profile points and objective
    B = (1 << 18) + 777
    sw = cases.config3_sweep(B, seed=4)
    z = np.array([0.5, 1.0])
    ref = np.array([0.6, 0.2, 0.02, 0.02, 0.15, 1e-4, 5e6, 600.0])
    dev = engine.n1_solve_ensemble(cm, base, sw, B, z_eval=z, objective_ref=ref, pipeline=False)
    rows, row_map = engine.sweep_rows(cm.spec, sw, B)
    out = np.empty((2, cm.spec.n, B)); status = np.empty(B, np.int32); stats = np.empty((4, B), np.int32); obj = np.empty(B)
    cm.module.n1_solve_host(B, rows, rows.shape[0], row_map, uniform, z, 1e-3, 1e-6, out, status, stats, obj_ref=ref, h_obj=obj)
    np.testing.assert_array_equal(out, dev.out)
    np.testing.assert_array_equal(status, dev.status)
    np.testing.assert_array_equal(stats, dev.stats)
    np.testing.assert_array_equal(obj, dev.objective)


def test_full_size_ensemble_properties():
    """BASELINE size (2^20 reactors): everything converges, outputs are physical and the
    statistics match the 36-corner reference box (T_out within its extremes)."""
    from rmt_app_b200 import rmtExeBatch
    base = cases.methanol_readme_input("N1")
    B = 1 << 20
    sw = cases.config3_sweep(B)
    r = rmtExeBatch(base, sw)
    assert r["success"].all()
    y = r["dataYs"]
    np.testing.assert_allclose(y[:, :6].sum(axis=1), 1.0, rtol=1e-12)
    assert (y[:, :6] > 0).all()
    assert (y[:, 6] < sw["pressure"]).all() and (y[:, 6] > 0.99*sw["pressure"]).all()     # small Ergun pressure drop
    assert y[:, 7].min() > 473.0 and y[:, 7].max() < 700.0                                # SURVEY App. B.4 extremes
    st = r["stats"]
    assert 30 < st[0].mean() < 80 and st[0].max() < 400
    # an ensemble of this size with host inputs runs as a three-chunk copy/compute pipeline: same bits as one launch
    from rmt_app_b200 import engine
    cm = engine.compile_model(base, method=engine.choose_method(base, 1e-3, 1))
    one = engine.n1_solve_ensemble(cm, base, sw, B, pipeline=False)
    np.testing.assert_array_equal(one.out[0].T, y)
    np.testing.assert_array_equal(one.status, r["status"])
    np.testing.assert_array_equal(one.stats, st)


def test_pipelined_ensemble_equals_single_launch():
    """Chunked copy/compute pipeline (engine.n1_solve_ensemble, pipeline=True) against the one-launch path:
    NumPy inputs (staged through pinned memory), pinned torch tensors, profiles and the fused objective."""
    import torch
    from rmt_app_b200 import engine
    base = cases.methanol_readme_input("N1")
    B = 7000
    sw = cases.config3_sweep(B, seed=11)
    cm = engine.compile_model(base)
    z = np.linspace(0, 1, 6)
    ref = np.array([0.6, 0.2, 0.02, 0.02, 0.15, 1e-4, 5e6, 600.0])
    a = engine.n1_solve_ensemble(cm, base, sw, B, z_eval=z, objective_ref=ref, pipeline=False)
    for form in ("numpy", "pinned"):
        s2 = sw if form == "numpy" else {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in sw.items()}
        b = engine.n1_solve_ensemble(cm, base, s2, B, z_eval=z, objective_ref=ref, pipeline=True, workspace=engine.Workspace())
        np.testing.assert_array_equal(a.out, b.out)
        np.testing.assert_array_equal(a.status, b.status)
        np.testing.assert_array_equal(a.stats, b.stats)
        np.testing.assert_array_equal(a.objective, b.objective)
        assert b.h2d_bytes == a.h2d_bytes and b.d2h_bytes == a.d2h_bytes


def test_branch_free_device_math_accuracy():
    """rmt_exp / rmt_log / rmt_sqrt / rmt_exp10 / rmt_rcp (csrc/rmt_kernels.cu) against NumPy: <= 2 ulp."""
    import torch
    from rmt_app_b200 import engine
    cm = engine.compile_model(cases.methanol_readme_input("N1"))
    mod = cm.load(torch.cuda.current_device())
    rng = np.random.default_rng(2)
    xs = np.concatenate([rng.uniform(-300, 300, 20000), np.exp(rng.uniform(-40, 40, 20000)),
                         rng.uniform(300, 1200, 5000), [1.0, 2.0, 0.5, 10.0, 1e-300, 1e300, 709.0, -708.0]])
    n = xs.size
    d_x = torch.from_numpy(xs).cuda()
    d_o = torch.empty((5, n), dtype=torch.float64, device="cuda")
    mod.math_probe(n, d_x, d_o, stream=torch.cuda.current_stream().cuda_stream)
    o = d_o.cpu().numpy()
    ulp = 2.220446049250313e-16

    def relerr(got, want, mask):
        return np.max(np.abs(got[mask] - want[mask])/np.abs(want[mask]))
    with np.errstate(all="ignore"):
        m = (xs > -700) & (xs < 700)
        assert relerr(o[0], np.exp(xs), m) < 2*ulp
        pos = xs > 0
        lg = np.log(np.where(pos, xs, 1.0))
        big = pos & (np.abs(lg) > 1e-2)
        assert relerr(o[1], lg, big) < 2*ulp
        assert np.max(np.abs(o[1][pos] - lg[pos])) < 1e-15*np.maximum(1.0, np.abs(lg[pos])).max()
        assert np.isnan(o[1][xs < 0]).all()
        assert relerr(o[2], np.sqrt(np.where(pos, xs, 1.0)), pos) < 2*ulp
        assert np.isnan(o[2][xs < 0]).all()
        m10 = (xs > -300) & (xs < 300)
        assert relerr(o[3], np.power(10.0, xs), m10) < 4*ulp
        nz = (np.abs(xs) > 1e-290) & (np.abs(xs) < 1e290)
        assert relerr(o[4], 1.0/xs, nz) < 2*ulp
    # special values: what the branch-free versions do with them is part of the contract (DESIGN.md section 6;
    # solver-config["exact-math"] restores IEEE semantics).  exp / 10^x saturate at e^+-708 / 10^+-307 and swallow NaN; log
    # is NaN outside the positive normal range (NaN, +-Inf, +-0, subnormals); sqrt keeps +-0 and +-Inf, gives NaN for NaN,
    # negative and subnormal arguments; the refined reciprocal of 0, Inf, NaN or a subnormal is NaN.  A non-finite value
    # makes the error test reject the step and ends in status 3 — never in a silent wrong answer; what the saturating exp
    # can hide is an overflow the reference would have turned into 0 or Inf (test_exact_math_option_gives_ieee_special_values).
    sp = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 5e-324, 1e-310, 800.0, -800.0])
    d_s = torch.from_numpy(sp).cuda()
    d_q = torch.empty((5, sp.size), dtype=torch.float64, device="cuda")
    mod.math_probe(sp.size, d_s, d_q, stream=torch.cuda.current_stream().cuda_stream)
    ex, lg_, sq, e10, rc = d_q.cpu().numpy()
    assert np.isfinite(ex).all() and ex[0] > 1e307 and ex[1] > 1e307 and ex[7] > 1e307 and 0 < ex[2] < 1e-307 and 0 < ex[8] < 1e-307
    assert (ex[3:7] == 1.0).all() and np.isfinite(e10).all() and e10[7] == 1e307 and e10[8] == 1e-307
    assert np.isnan(lg_[:7]).all() and np.isnan(lg_[8]) and abs(lg_[7] - np.log(800.0)) < 1e-14
    assert np.isnan(sq[0]) and sq[1] == np.inf and sq[2] == -np.inf and sq[3] == 0.0 and sq[4] == 0.0
    assert np.isnan(sq[5]) and np.isnan(sq[6]) and np.isnan(sq[8]) and abs(sq[7] - np.sqrt(800.0)) < 1e-14
    assert np.isnan(rc[:7]).all() and rc[7] == 1.0/800.0 and rc[8] == -1.0/800.0


def _with_inert(mi, sym="CO"):
    """ch4 input plus a zero-feed inert species `sym` (never formed by the reaction)."""
    mi["feed"]["components"]["shell"] = list(mi["feed"]["components"]["shell"]) + [sym]
    mi["feed"]["concentration"] = np.append(np.asarray(mi["feed"]["concentration"], float), 0.0)
    return mi


@pytest.mark.parametrize("process", ["iso-thermal", "non-iso-thermal"])
@pytest.mark.parametrize("reduced", [True, False])
def test_zero_feed_species_inside_a_denominator(process, reduced):
    """a species with zero feed that is never produced stays exactly 0; when it only occurs in an LHHW-type
    term 1 + K*y_i there is no pole at 0, the sign analysis does not flag it, and the integration runs like the
    reference's (which has no positivity restriction, pbHomoReactor.py:3148-3175) — in the reaction-extent form and in
    the full-state form."""
    eng = _engine()
    mi = _with_inert(cases.ch4_input("N1", process))
    mi["reaction-rates"] = {
        "VARS": {"k0": 0.0072*1e-1, "K": 40.0, "C_CH4": lambda x: x['SpCoi'][0], "y_CO": lambda x: x['MoFri'][3]},
        "RATES": {"r1": lambda x: x['k0']*(x['C_CH4']**2)/(1 + x['K']*x['y_CO'])},
    }
    cm = eng.compile_model(mi, reduced=reduced)
    assert cm.reduced == reduced
    assert cm.spec.kin.positive_species() == [False, False, False, False]
    z = np.linspace(0, 1, 11)
    for tol in (dict(rtol=1e-3, atol=1e-6), TIGHT):
        r = eng.n1_solve_ensemble(cm, mi, None, 1, z_eval=z, **tol)
        assert r.status[0] == 0, r.status
        assert r.stats[1, 0] <= 0.25*r.stats[0, 0] + 3            # no rejection storm
        # the inert's mole fraction: exactly zero in the extent form (y = y0 + nu^T xi, nu = 0), at the rounding floor
        # of the linear solves in the full-state form
        assert (r.out[:, 3, 0] == 0.0).all() if reduced else (np.abs(r.out[:, 3, 0]) < 1e-15).all()
    want = O.N1Oracle(mi).solve(method="LSODA", rtol=1e-11, atol=1e-13, t_eval=z)
    got = r.out[:, :, 0].T                                        # dataYs rows: y_i..., P, (T)
    ref = O.N1Oracle(mi).pack(want)[0]["dataYs"]
    live = [0, 1, 2] + list(range(4, got.shape[0]))
    assert np.max(np.abs(got[live] - ref[live])/np.abs(ref[live])) < 1e-6
    # the same species under a true pole (division by its mole fraction) IS flagged, and the run fails loudly
    mi2 = _with_inert(cases.ch4_input("N1", process))
    mi2["reaction-rates"] = {
        "VARS": {"k0": 0.0072*1e-1, "C_CH4": lambda x: x['SpCoi'][0], "y_CO": lambda x: x['MoFri'][3]},
        "RATES": {"r1": lambda x: x['k0']*(x['C_CH4']**2)*x['y_CO']/x['y_CO']},
    }
    cm2 = eng.compile_model(mi2, reduced=reduced)
    assert cm2.spec.kin.positive_species() == [False, False, False, True]
    assert eng.n1_solve_ensemble(cm2, mi2, None, 1, **TIGHT).status[0] != 0


@pytest.mark.parametrize("reduced", [True, False])
def test_irreversible_reaction_at_complete_conversion(reduced):
    """a fast irreversible first-order reaction consumes its reactant completely within the first tenth of the
    bed; afterwards the exact solution's zero sits at the rounding floor (y = y0 + nu^T xi in the extent form) with a
    random sign.  A negative value within the error tolerance is that zero: no rejection storm, status 0, and the
    products agree with the oracle."""
    eng = _engine()
    mi = cases.ch4_input("N1", "iso-thermal")
    mi["reaction-rates"] = {"VARS": {"k0": 40.0, "C_CH4": lambda x: x['SpCoi'][0]},
                            "RATES": {"r1": lambda x: x['k0']*x['C_CH4']}}
    cm = eng.compile_model(mi, reduced=reduced)
    assert cm.spec.kin.positive_species() == [False, False, False]
    z = np.linspace(0, 1, 21)
    want = O.N1Oracle(mi).pack(O.N1Oracle(mi).solve(method="LSODA", rtol=1e-12, atol=1e-14, t_eval=z))[0]["dataYs"]
    assert want[0, -1] < 1e-9                                     # the reactant is gone at the outlet
    for tol, bar in ((dict(rtol=1e-3, atol=1e-6), 5e-3), (TIGHT, 1e-6)):
        r = eng.n1_solve_ensemble(cm, mi, None, 1, z_eval=z, **tol)
        assert r.status[0] == 0, r.status
        assert r.stats[1, 0] <= 0.5*r.stats[0, 0] + 3, r.stats[:, 0]
        got = r.out[:, :, 0].T
        assert (got[:3] >= -1e-9).all()                           # interpolated points: zero within the tolerance
        assert np.max(np.abs(got[1:] - want[1:])/np.abs(want[1:])) < bar
        assert np.max(np.abs(got[0] - want[0])) < bar             # the vanishing reactant: absolute (mole fraction)


def test_exact_math_option_gives_ieee_special_values():
    """the branch-free device math trades IEEE special values for speed (exp saturates at e^708 instead of
    overflowing to Inf, x * rcp(Inf) is NaN instead of 0).  Kinetics that rely on them — here a rate switched off by an
    overflowing np.exp in the denominator, which NumPy evaluates to exactly 0 — fail LOUDLY in the default build (status
    3, NaN outputs) and integrate like the reference with solver-config["exact-math"] = True (libdevice + IEEE division)."""
    eng = _engine()

    def make(exact):
        mi = cases.ch4_input("N1", "iso-thermal")
        mi["reaction-rates"] = {"VARS": {"k0": 0.0072*1e-1, "C_CH4": lambda x: x['SpCoi'][0], "y": lambda x: x['MoFri'][0]},
                                "RATES": {"r1": lambda x: x['k0']*(x['C_CH4']**2)/(1 + np.exp(900.0*x['y']))**3}}
        mi["solver-config"] = dict(mi["solver-config"], **({"exact-math": True} if exact else {}))
        return mi
    fast = eng.compile_model(make(False))
    exact = eng.compile_model(make(True))
    assert exact is not fast and exact.exact_math and exact.key() != fast.key()
    r = eng.n1_solve_ensemble(fast, make(False), None, 1)
    assert r.status[0] in (2, 3) and np.isnan(r.out).all()      # every step is rejected on NaN: step underflow or the NaN counter
    r = eng.n1_solve_ensemble(exact, make(True), None, 1, **TIGHT)
    assert r.status[0] == 0
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = O.N1Oracle(make(True)).pack(O.N1Oracle(make(True)).solve(method="LSODA", rtol=1e-11, atol=1e-13))[0]["dataYs"][:, -1]
    np.testing.assert_allclose(r.out[0, :, 0], want, rtol=1e-9)
    np.testing.assert_allclose(r.out[0, :3, 0], [0.9, 0.05, 0.05], rtol=1e-4)          # no reaction: the feed composition
