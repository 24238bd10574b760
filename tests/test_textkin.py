"""Dashboard text kinetics -> VARS/RATES (host logic)."""
import numpy as np
import pytest

import cases
import pyremot_oracle as O
from rmt_app_b200 import parse_reaction_rates
from rmt_app_b200.kinetics import trace_kinetics

VARS = """
"CaBeDe" : CaBeDe;
"RT": x['R_CONST']*x['T'];
"K1": 35.45*math.exp(-1.7069e4/x['RT']);
"K2": 7.3976*math.exp(-2.0436e4/x['RT']);
"K3": 8.2894e4*math.exp(-5.2940e4/x['RT']);
"KH2": 0.249*math.exp(3.4394e4/x['RT']);
"KCO2": 1.02e-7*math.exp(6.74e4/x['RT']);
"KCO": 7.99e-7*math.exp(5.81e4/x['RT']);
"Ln_KP1": 4213/x['T'] - 5.752 * math.log(x['T']) - 1.707e-3*x['T'] + 2.682e-6 * (math.pow(x['T'], 2)) - 7.232e-10*(math.pow(x['T'], 3)) + 17.6;
"KP1": math.exp(x['Ln_KP1']);
"log_KP2": 2167/x['T'] - 0.5194 * math.log10(x['T']) + 1.037e-3*x['T'] - 2.331e-7 * (math.pow(x['T'], 2)) - 1.2777;
"KP2": math.pow(10, x['log_KP2']);
"Ln_KP3": 4019/x['T'] + 3.707 * math.log(x['T']) - 2.783e-3*x['T'] + 3.8e-7 * (math.pow(x['T'], 2)) - 6.56e-4/(math.pow(x['T'], 3)) - 26.64;
"KP3": math.exp(x['Ln_KP3']);
"yi_H2": x['MoFri'][0]; "yi_CO2": x['MoFri'][1]; "yi_H2O": x['MoFri'][2];
"yi_CO": x['MoFri'][3]; "yi_CH3OH": x['MoFri'][4]; "yi_DME": x['MoFri'][5];
"PH2": x['P']*(x['yi_H2'])*1e-5; "PCO2": x['P']*(x['yi_CO2'])*1e-5; "PH2O": x['P']*(x['yi_H2O'])*1e-5;
"PCO": x['P']*(x['yi_CO'])*1e-5; "PCH3OH": x['P']*(x['yi_CH3OH'])*1e-5; "PCH3OCH3": x['P']*(x['yi_DME'])*1e-5;
"ra1": x['PCO2']*x['PH2'];
"ra2": 1 + (x['KCO2']*x['PCO2']) + (x['KCO']*x['PCO']) + math.sqrt(x['KH2']*x['PH2']);
"ra3": (1/x['KP1'])*((x['PH2O']*x['PCH3OH'])/(x['PCO2']*(math.pow(x['PH2'], 3))));
"ra4": x['PH2O'] - (1/x['KP2'])*((x['PCO2']*x['PH2'])/x['PCO']);
"ra5": (math.pow(x['PCH3OH'], 2)/x['PH2O'])-(x['PCH3OCH3']/x['KP3'])
"""
RATES = """
"r1": 1000*x['K1']*(x['ra1']/(math.pow(x['ra2'], 3)))*(1-x['ra3'])*x['CaBeDe'];
"r2": 1000*x['K2']*(1/x['ra2'])*x['ra4']*x['CaBeDe'];
"r3": 1000*x['K3']*x['ra5']*x['CaBeDe']
"""


def test_text_sections_equal_the_lambda_form():
    rr = parse_reaction_rates(VARS, RATES, {"CaBeDe": 1171.2})
    ref = cases.methanol_kinetics(1171.2)
    assert list(rr["VARS"]) == list(ref["VARS"]) and list(rr["RATES"]) == list(ref["RATES"])
    assert rr["VARS"]["CaBeDe"] == 1171.2
    T, P = 560.0, 4.2e6
    y = np.array([0.5, 0.2, 0.02, 0.25, 0.02, 0.01])
    C = y*P/(O.R_CONST*T)
    want = O.reaction_rate_exe((T, P, y, C), ref["VARS"], ref["RATES"])
    got = O.reaction_rate_exe((T, P, y, C), rr["VARS"], rr["RATES"])
    assert got == want
    ir = trace_kinetics(rr["VARS"], rr["RATES"], 6)
    np.testing.assert_allclose(ir.evaluate(T, P, y, C), want, rtol=1e-15)
    assert ir.param_names == ["CaBeDe"]


def test_scalar_expressions_become_parameter_slots():
    rr = parse_reaction_rates('"k0": 0.0072*1e-1; "C": x[\'SpCoi\'][0]', '"r1": x["k0"]*(x["C"]**2)')
    assert rr["VARS"]["k0"] == pytest.approx(7.2e-4) and callable(rr["VARS"]["C"])


def test_malformed_entry():
    with pytest.raises(ValueError):
        parse_reaction_rates("just an expression", '"r": 1.0')


@pytest.mark.gpu
def test_text_kinetics_on_the_device_equal_the_lambda_form(golden_n1):
    """README.md:83-173 text sections -> parse -> trace -> generated CUDA -> rmt_n1_rhs: the lambda form of the same
    kinetics to rounding (the two generated translation units number their temporaries differently, so the compiler
    may fuse multiply-adds differently: near chemical equilibrium that is amplified to ~1e-10 of the row scale), the
    oracle's RHS to 1e-11 on well-conditioned states, and the same converged solution through rmtExe."""
    from rmt_app_b200 import engine, rmtExe
    mi_l = cases.methanol_readme_input("N1")
    mi_t = cases.methanol_readme_input("N1")
    mi_t["reaction-rates"] = parse_reaction_rates(VARS, RATES, {"CaBeDe": 1171.2})
    Y = golden_n1["methanol_readme__rhs_Y"]
    cm_l, cm_t = engine.compile_model(mi_l), engine.compile_model(mi_t)
    Fl, _, _ = engine.n1_rhs_batch(cm_l, mi_l, Y)
    Ft, Jt, _ = engine.n1_rhs_batch(cm_t, mi_t, Y, jac=True)
    scale = np.max(np.abs(Fl), axis=1, keepdims=True)
    assert np.max(np.abs(Ft - Fl)/scale) < 1e-9
    assert np.max(np.abs(Ft[:3] - Fl[:3])/scale[:3]) < 1e-12          # well-conditioned states: plain rounding
    o = O.N1Oracle(mi_t)                          # the oracle evaluates the parsed lambdas themselves
    for k in (0, 1, 2):                           # feed state and small perturbations of it
        f = np.array(o.rhs(0.0, Y[k]))
        # feed state: every entry to 1e-11 of itself; perturbed states: 1e-11 of the state's largest entry (a small row
        # such as CH3OH's r1 - 2 r3 cancels: 1 ulp in the rates is ~3e-11 of it, cf. test_rhs_parity_with_reference_golden)
        den = np.abs(f) if k == 0 else np.max(np.abs(f))
        assert np.max(np.abs(Ft[k] - f)/den) < 1e-11, k
    _, Jl, _ = engine.n1_rhs_batch(cm_l, mi_l, Y, jac=True)
    assert np.max(np.abs(Jt - Jl))/np.max(np.abs(Jl)) < 1e-9
    for mi in (mi_l, mi_t):
        mi["solver-config"] = dict(mi["solver-config"], rtol=1e-9, atol=1e-12)
    a, b = rmtExe(mi_t)["resModel"][0]["dataYs"], rmtExe(mi_l)["resModel"][0]["dataYs"]
    np.testing.assert_allclose(a, b, rtol=1e-9)
    np.testing.assert_allclose(a, golden_n1["methanol_readme__tight_LSODA__dataYs"], rtol=1e-6)
