"""Kinetics tracer / symbolic differentiation / CUDA code generator (host logic, CPU)."""
import math

import numpy as np
import pytest

import cases
import pyremot_oracle as O
from rmt_app_b200 import capi
from rmt_app_b200.codegen import generate_model_header, model_flops
from rmt_app_b200.expr import TraceError
from rmt_app_b200.kinetics import trace_kinetics
from rmt_app_b200.model import ModelSpec, parse_reaction


def _random_points(nc, n, seed=0):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        T = rng.uniform(450, 700)
        P = rng.uniform(1e6, 9e6)
        y = rng.uniform(0.01, 1.0, nc)
        y /= y.sum()
        C = y*P/(O.R_CONST*T)
        yield T, P, y, C


@pytest.mark.parametrize("kin,nc", [(cases.methanol_kinetics(1171.2), 6), (cases.methanol_kinetics_param(1209.02), 6),
                                     (cases.ch4_input()["reaction-rates"], 3)])
def test_traced_rates_equal_python_lambdas(kin, nc):
    ir = trace_kinetics(kin["VARS"], kin["RATES"], nc)
    for T, P, y, C in _random_points(nc, 20):
        want = O.reaction_rate_exe((T, P, y, C), kin["VARS"], kin["RATES"])
        got = ir.evaluate(T, P, y, C)
        np.testing.assert_allclose(got, want, rtol=1e-15, atol=0)


def test_parameter_slots():
    kin = cases.methanol_kinetics_param(1171.2)
    ir = trace_kinetics(kin["VARS"], kin["RATES"], 6)
    assert ir.param_names[0] == "CaBeDe" and set(cases.METHANOL_ARRHENIUS) <= set(ir.param_names)
    T, P, y, C = next(_random_points(6, 1))
    p = list(ir.param_defaults)
    p[ir.param_names.index("k01")] *= 2.0
    a, b = ir.evaluate(T, P, y, C), ir.evaluate(T, P, y, C, p)
    assert b[0] == pytest.approx(2*a[0], rel=1e-14) and b[1] == a[1]


def test_partials_against_central_differences():
    kin = cases.methanol_kinetics(1171.2)
    ir = trace_kinetics(kin["VARS"], kin["RATES"], 6)
    part = ir.differentiate()
    g = ir.g
    for T, P, y, C in _random_points(6, 5, seed=3):
        env = {"T": T, "P": P, "kp0": 1171.2}
        env.update({"y%d" % i: y[i] for i in range(6)})
        env.update({"C%d" % i: C[i] for i in range(6)})
        for w in ["T", "P"] + ["y%d" % i for i in range(6)]:
            an = [g.evaluate([d], env)[0] if d is not None else 0.0 for d in part[w]]
            h = abs(env[w])*1e-6
            ep, em = dict(env), dict(env)
            ep[w] += h; em[w] -= h
            fd = (np.array(g.evaluate(ir.rates, ep)) - np.array(g.evaluate(ir.rates, em)))/(2*h)
            np.testing.assert_allclose(an, fd, rtol=2e-6, atol=1e-9*np.max(np.abs(fd)))
    assert all(d is None for i in range(6) for d in part["C%d" % i])      # methanol rates use MoFri only


def test_positive_species_analysis():
    m = ModelSpec(cases.methanol_readme_input())
    assert m.kin.positive_species() == [True, True, True, True, False, False]
    assert ModelSpec(cases.ch4_input()).kin.positive_species() == [False, False, False]


def test_shadowing_and_numpy_and_closures():
    import numpy as np_
    k = 3.5
    varis = {
        "P": lambda x: x["T"]*4000.0,              # user key shadows the base P, keeps its slot (rmtReaction.py:39)
        "a": 2.0,
        "s": lambda x: np_.sum(x["MoFri"]) + np_.exp(-1000.0/x["T"]) + k,
        "m": lambda x: max(x["SpCoi"][0], 1e-30)**0.5 + abs(x["P"])*0 + math.log(x["P"], 10),
    }
    rates = {"r": lambda x: x["a"]*x["s"]*x["m"]}
    ir = trace_kinetics(varis, rates, 2)
    T, P, y, C = 500.0, 2e6, np.array([0.3, 0.7]), np.array([10.0, 20.0])
    want = O.reaction_rate_exe((T, P, y, C), varis, rates)
    np.testing.assert_allclose(ir.evaluate(T, P, y, C), want, rtol=1e-14)


def test_data_dependent_branch_is_rejected():
    varis = {"k": lambda x: 1.0 if x["T"] > 500 else 2.0}
    with pytest.raises(TraceError, match="branch"):
        trace_kinetics(varis, {"r": lambda x: x["k"]}, 2)


def test_reaction_parser_and_stoichiometry():
    assert parse_reaction("CO2 + 3H2 <=> CH3OH + H2O") == ([("CO2", -1.0), ("H2", -3.0)], [("CH3OH", 1.0), ("H2O", 1.0)])
    assert parse_reaction("0.5C4H10+1.5N2=C3H6 + CH4")[0] == [("C4H10", -0.5), ("N2", -1.5)]
    m = ModelSpec(cases.methanol_readme_input())
    np.testing.assert_array_equal(m.nu, [[-3, -1, 1, 0, 1, 0], [1, 1, -1, -1, 0, 0], [0, 0, 1, 0, -2, 1]])
    np.testing.assert_allclose(m.dH25, [-49010.0, -41160.0, -24520.0], rtol=1e-12)
    # dCp cubic against the oracle's per-reaction evaluation
    o = O.N1Oracle(cases.methanol_readme_input())
    for T in (480.0, 650.0):
        want = np.array(O.enthalpy_change_of_reaction(o.reactionListSorted, T))/(T - O.Tref)
        got = m.dcp[:, 0] + m.dcp[:, 1]*T + m.dcp[:, 2]*T**2 + m.dcp[:, 3]*T**3
        np.testing.assert_allclose(got, want, rtol=1e-12)


def test_rates_count_must_match_reactions():
    mi = cases.ch4_input()
    mi["reactions"]["R2"] = "C2H4 + H2 <=> 2CH4"
    with pytest.raises(ValueError, match="matched by position"):
        ModelSpec(mi)


def test_unsupported_model_is_loud():
    mi = cases.ch4_input()
    mi["model"] = "M2"
    with pytest.raises(NotImplementedError):
        ModelSpec(mi)


def test_flop_counts_match_survey_estimate():
    fl = model_flops(ModelSpec(cases.methanol_readme_input()))
    assert 250 <= fl["rhs_alg"] <= 350           # ~300 flop per RHS
    assert 800 <= fl["rhs_weighted"] <= 1100     # ~850-900 weighted
    assert fl["rates_alg"] == 103


@pytest.mark.parametrize("mk", [lambda: cases.methanol_readme_input("N1"), lambda: cases.ch4_input("N1", "iso-thermal"),
                                lambda: cases.ch4_input("N2")])
def test_generated_translation_unit_compiles_for_sm100a(mk):
    """NVRTC (through the C ABI) needs no GPU: this is the real device compiler."""
    spec = ModelSpec(mk())
    src = generate_model_header(spec)
    assert "rmt_rates_jac" in src and "RMT_ROS_A" in src
    cubin, log, ptx = capi.nvrtc_compile(src, block=64, want_ptx=True)
    assert cubin[:4] == b"\x7fELF" and len(cubin) > 10000
    assert ".target sm_100a" in ptx
    assert "fma.rn.f64" in ptx


def test_nvrtc_error_is_reported():
    with pytest.raises(capi.RmtError, match="NVRTC compilation failed"):
        capi.nvrtc_compile("#define RMT_NC oops\n")
