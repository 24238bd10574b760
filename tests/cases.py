"""Model inputs used by the tests, the golden-vector generator and bench.py.

These are *inputs* in the reference's own `modelInput` format (the dict that
`rmtExe` takes, PyREMOT/rmt.py:21-80).  Three kinetics sets:

* methanol/DME synthesis, 6 species / 3 reactions — the README / notebook
  TEST1 instance (README.md:83-229) and the `tests/test_rmt_N1_DME.py:25-269`
  instance (different bed density, U, Tm and a float32-derived feed);
* methane coupling, 3 species / 1 reaction (`tests/test_rmt_N2_CH4.py:23-250`),
  which exercises `SpCoi`, a scalar VARS entry, `**`, `MeTe == 0` (adiabatic)
  and the iso-thermal mode.

plus the synthetic sweeps of SURVEY.md §8(d) (config 3 / 4 / 5).

The module deliberately imports nothing from the reference or from the
package under test so both sides can consume identical inputs.
"""
import math

import numpy as np

R_CONST = 8.314472  # PyREMOT/core/constants.py:8


# ----------------------------------------------------------------------------
# kinetics (user plug-in sections VARS / RATES)
# ----------------------------------------------------------------------------
def methanol_kinetics(CaBeDe):
    """VARS/RATES of the CO2->methanol/DME case (README.md:83-173)."""
    varis0 = {
        "CaBeDe": CaBeDe,
        "RT": lambda x: x['R_CONST']*x['T'],
        "K1": lambda x: 35.45*math.exp(-1.7069e4/x['RT']),
        "K2": lambda x: 7.3976*math.exp(-2.0436e4/x['RT']),
        "K3": lambda x: 8.2894e4*math.exp(-5.2940e4/x['RT']),
        "KH2": lambda x: 0.249*math.exp(3.4394e4/x['RT']),
        "KCO2": lambda x: 1.02e-7*math.exp(6.74e4/x['RT']),
        "KCO": lambda x: 7.99e-7*math.exp(5.81e4/x['RT']),
        "Ln_KP1": lambda x: 4213/x['T'] - 5.752 *
        math.log(x['T']) - 1.707e-3*x['T'] + 2.682e-6 *
        (math.pow(x['T'], 2)) - 7.232e-10*(math.pow(x['T'], 3)) + 17.6,
        "KP1": lambda x: math.exp(x['Ln_KP1']),
        "log_KP2": lambda x: 2167/x['T'] - 0.5194 *
        math.log10(x['T']) + 1.037e-3*x['T'] - 2.331e-7 *
        (math.pow(x['T'], 2)) - 1.2777,
        "KP2": lambda x: math.pow(10, x['log_KP2']),
        "Ln_KP3": lambda x: 4019/x['T'] + 3.707 *
        math.log(x['T']) - 2.783e-3*x['T'] + 3.8e-7 *
        (math.pow(x['T'], 2)) - 6.56e-4/(math.pow(x['T'], 3)) - 26.64,
        "KP3": lambda x: math.exp(x['Ln_KP3']),
        "yi_H2": lambda x: x['MoFri'][0],
        "yi_CO2": lambda x: x['MoFri'][1],
        "yi_H2O": lambda x: x['MoFri'][2],
        "yi_CO": lambda x: x['MoFri'][3],
        "yi_CH3OH": lambda x: x['MoFri'][4],
        "yi_DME": lambda x: x['MoFri'][5],
        "PH2": lambda x: x['P']*(x['yi_H2'])*1e-5,
        "PCO2": lambda x: x['P']*(x['yi_CO2'])*1e-5,
        "PH2O": lambda x: x['P']*(x['yi_H2O'])*1e-5,
        "PCO": lambda x: x['P']*(x['yi_CO'])*1e-5,
        "PCH3OH": lambda x: x['P']*(x['yi_CH3OH'])*1e-5,
        "PCH3OCH3": lambda x: x['P']*(x['yi_DME'])*1e-5,
        "ra1": lambda x: x['PCO2']*x['PH2'],
        "ra2": lambda x: 1 + (x['KCO2']*x['PCO2']) + (x['KCO']*x['PCO']) + math.sqrt(x['KH2']*x['PH2']),
        "ra3": lambda x: (1/x['KP1'])*((x['PH2O']*x['PCH3OH'])/(x['PCO2']*(math.pow(x['PH2'], 3)))),
        "ra4": lambda x: x['PH2O'] - (1/x['KP2'])*((x['PCO2']*x['PH2'])/x['PCO']),
        "ra5": lambda x: (math.pow(x['PCH3OH'], 2)/x['PH2O'])-(x['PCH3OCH3']/x['KP3']),
    }
    rates0 = {
        "r1": lambda x: 1000*x['K1']*(x['ra1']/(math.pow(x['ra2'], 3)))*(1-x['ra3'])*x['CaBeDe'],
        "r2": lambda x: 1000*x['K2']*(1/x['ra2'])*x['ra4']*x['CaBeDe'],
        "r3": lambda x: 1000*x['K3']*x['ra5']*x['CaBeDe'],
    }
    return {"VARS": varis0, "RATES": rates0}


# Arrhenius pairs of the methanol model lifted into scalar VARS slots
# (SURVEY.md §8(d), config 4): k0* pre-exponentials, E* activation energies.
METHANOL_ARRHENIUS = {
    "k01": 35.45, "E1": -1.7069e4,
    "k02": 7.3976, "E2": -2.0436e4,
    "k03": 8.2894e4, "E3": -5.2940e4,
    "k0H2": 0.249, "EH2": 3.4394e4,
    "k0CO2": 1.02e-7, "ECO2": 6.74e4,
    "k0CO": 7.99e-7, "ECO": 5.81e4,
}


def methanol_kinetics_param(CaBeDe, arrhenius=None):
    """Same kinetics as `methanol_kinetics` but with the six Arrhenius pairs
    exposed as scalar (non-callable) VARS entries — exactly the values
    `reactionRateExe` passes through as constants (rmtReaction.py:46-49) and
    therefore the per-instance kinetic-parameter slots of the ensemble."""
    a = dict(METHANOL_ARRHENIUS)
    if arrhenius:
        a.update(arrhenius)
    base = methanol_kinetics(CaBeDe)
    varis = {"CaBeDe": CaBeDe}
    varis.update(a)
    varis["RT"] = base["VARS"]["RT"]
    varis["K1"] = lambda x: x['k01']*math.exp(x['E1']/x['RT'])
    varis["K2"] = lambda x: x['k02']*math.exp(x['E2']/x['RT'])
    varis["K3"] = lambda x: x['k03']*math.exp(x['E3']/x['RT'])
    varis["KH2"] = lambda x: x['k0H2']*math.exp(x['EH2']/x['RT'])
    varis["KCO2"] = lambda x: x['k0CO2']*math.exp(x['ECO2']/x['RT'])
    varis["KCO"] = lambda x: x['k0CO']*math.exp(x['ECO']/x['RT'])
    for k, v in base["VARS"].items():
        if k not in varis:
            varis[k] = v
    return {"VARS": varis, "RATES": base["RATES"]}


METHANOL_COMPONENTS = ["H2", "CO2", "H2O", "CO", "CH3OH", "DME"]
METHANOL_REACTIONS = {
    "R1": "CO2 + 3H2 <=> CH3OH + H2O",
    "R2": "CO + H2O <=> H2 + CO2",
    "R3": "2CH3OH <=> DME + H2O",
}


def feed_mole_fraction(H2COxRatio, CO2COxRatio, dtype=np.float64):
    """Feed composition formula of PyREMOT/data/initData.py:11-41 (vectorised;
    the reference helper returns float32 — pass dtype=np.float32 to mimic)."""
    H2COxRatio = np.asarray(H2COxRatio, dtype=np.float64)
    CO2COxRatio = np.asarray(CO2COxRatio, dtype=np.float64)
    y_tr = 0.00001
    tmf0 = 1 - (y_tr + y_tr + y_tr)
    COx = tmf0/(H2COxRatio + 1)
    y_H2 = H2COxRatio*COx
    y_CO2 = CO2COxRatio*COx
    y_CO = COx - y_CO2
    tr = np.full_like(y_H2, y_tr)
    return np.stack([y_H2, y_CO2, tr, y_CO, tr, tr], axis=-1).astype(dtype)


def methanol_readme_input(model="N1", ivp="default", process_type="non-iso-thermal"):
    """README / notebook TEST1 canonical instance = BASELINE config 1
    (README.md:175-229)."""
    CaBeDe = 1171.2
    mi = {
        "model": model,
        "operating-conditions": {
            "pressure": 5000000,
            "temperature": 523,
            "process-type": process_type,
        },
        "feed": {
            "volumetric-flowrate": 0.000228,
            "concentration": [574.8978, 287.4489, 1.15e-02, 287.4489, 1.15e-02, 1.15e-02],
            "components": {"shell": list(METHANOL_COMPONENTS)},
        },
        "reactions": dict(METHANOL_REACTIONS),
        "reaction-rates": methanol_kinetics(CaBeDe),
        "external-heat": {"OvHeTrCo": 50, "EfHeTrAr": 4/0.0381, "MeTe": 523},
        "reactor": {
            "ReInDi": 0.0381, "ReLe": 1, "PaDi": 0.002, "BeVoFr": 0.39,
            "CaBeDe": CaBeDe, "CaDe": 1920, "CaSpHeCa": 960,
        },
        "solver-config": {"ivp": ivp, "display-result": "False"},
    }
    if model == "N2":
        mi["operating-conditions"]["period"] = 0.5
    return mi


def methanol_testfile_input(model="N1", ivp="default"):
    """`PyREMOT/tests/test_rmt_N1_DME.py:25-269` instance:
    float32-derived feed rounded to 7 dp in kmol/m^3, bulk density 1982*0.61,
    U = 100, Tm = T-1."""
    P, T = 5*1e6, 523
    bed_por, rea_D, rea_L, cat_d, cat_rho, cat_Cp = 0.39, 0.0381, 1, 0.002, 1982, 960
    bulk_rho = cat_rho*(1 - bed_por)
    y32 = feed_mole_fraction(1, 0.5, dtype=np.float32)
    ct0 = np.round(np.array([(P/(R_CONST*T))*float(v)/1000 for v in y32]), 7)
    ct0_CONV = 1e3*ct0
    InGaVe = 0.2/bed_por
    rea_CSA = bed_por*(math.pi*(rea_D**2)/4)
    VoFlRa = InGaVe*rea_CSA
    mi = {
        "model": model,
        "operating-conditions": {
            "pressure": P, "temperature": T, "period": 0.5,
            "process-type": "non-iso-thermal",
        },
        "feed": {
            "volumetric-flowrate": VoFlRa,
            "concentration": ct0_CONV,
            "components": {"shell": list(METHANOL_COMPONENTS)},
        },
        "reactions": dict(METHANOL_REACTIONS),
        "reaction-rates": methanol_kinetics(bulk_rho),
        "external-heat": {"OvHeTrCo": 100, "EfHeTrAr": 4/rea_D, "MeTe": T - 1},
        "reactor": {
            "ReInDi": rea_D, "ReLe": rea_L, "PaDi": cat_d, "BeVoFr": bed_por,
            "CaBeDe": bulk_rho, "CaDe": cat_rho, "CaSpHeCa": cat_Cp/1000,
        },
        "solver-config": {"ivp": ivp, "display-result": "False"},
    }
    return mi


def methanol_m7_input(ivp="default"):
    """`PyREMOT/tests/test_rmt_DME3.py:20-262` instance: model M7, the dimensional steady-state twin of N1
    (pbReactor.py runM3 :1170-1575).  Needs `feed.mixture-viscosity` and uses `external-heat.EfHeTrAr` as given."""
    P, T = 5*1e6, 523
    bed_por, rea_D, rea_L, cat_d, cat_rho, cat_Cp, cat_por = 0.39, 0.0381, 1, 0.002, 1982, 960, 0.45
    bulk_rho = cat_rho*(1 - bed_por)
    y32 = feed_mole_fraction(1, 0.5, dtype=np.float32)
    ct0 = np.round(np.array([(P/(R_CONST*T))*float(v)/1000 for v in y32]), 7)
    InGaVe = 0.2/bed_por
    rea_CSA = bed_por*(math.pi*(rea_D**2)/4)
    VoFlRa = InGaVe*rea_CSA
    kin = methanol_kinetics(bulk_rho)
    varis = {"CaDe": cat_rho, "CaBeDe": bulk_rho, "CaPo": cat_por}
    for k, v in kin["VARS"].items():
        if k not in varis:
            varis[k] = v
    return {
        "model": "M7",
        "operating-conditions": {"pressure": P, "temperature": T, "period": 50},
        "feed": {
            "volumetric-flowrate": VoFlRa, "concentration": ct0*1000, "mixture-viscosity": 1e-5,
            "components": {"shell": list(METHANOL_COMPONENTS), "tube": [], "medium": []},
        },
        "reactions": dict(METHANOL_REACTIONS),
        "reaction-rates": {"VARS": varis, "RATES": kin["RATES"]},
        "external-heat": {"OvHeTrCo": 50, "EfHeTrAr": 4/rea_D, "MeTe": 523},
        "reactor": {"ReInDi": rea_D, "ReLe": rea_L, "PaDi": cat_d, "BeVoFr": bed_por, "CaBeDe": bulk_rho,
                    "CaDe": cat_rho, "CaSpHeCa": cat_Cp/1000},
        "solver-config": {"ivp": ivp},
    }


def methanol_m9_input(ivp="default", period=2.0):
    """`PyREMOT/tests/test_rmt_DME5.py:20-232` instance: model M9, the dimensional dynamic twin of N2
    (pbReactor.py runM5 :1997-2660).  Concentrations in kmol/m^3 and rates in kmol/(m^3 s) as that script has them
    (its RATES lack the factor 1000 of the M7 script)."""
    mi = methanol_m7_input(ivp)
    kin = mi["reaction-rates"]
    rates = {
        "r1": lambda x: x['K1']*(x['ra1']/(math.pow(x['ra2'], 3)))*(1-x['ra3'])*x['CaBeDe'],
        "r2": lambda x: x['K2']*(1/x['ra2'])*x['ra4']*x['CaBeDe'],
        "r3": lambda x: x['K3']*x['ra5']*x['CaBeDe'],
    }
    mi["model"] = "M9"
    mi["operating-conditions"] = dict(mi["operating-conditions"], period=period)
    mi["feed"] = dict(mi["feed"], concentration=np.array(mi["feed"]["concentration"])/1000, **{"superficial-velocity": 0.2})
    mi["reaction-rates"] = {"VARS": kin["VARS"], "RATES": rates}
    return mi


def ch4_input(model="N1", process_type="non-iso-thermal", ivp="default"):
    """Methane-coupling instance of `PyREMOT/tests/test_rmt_N2_CH4.py:23-250`
    (SURVEY.md App. B.5)."""
    P, T = 3*1e5, 973
    bed_por, rea_dia, cat_d, cat_rho, cat_cp = 0.39, 0.007, 0.002, 1982, 960
    bulk_rho = cat_rho*(1 - bed_por)
    MoFri0 = np.array([1 - (0.05 + 0.05), 0.05, 0.05])
    ct0 = np.round(np.array([(P/(R_CONST*T))*v/1000 for v in MoFri0]), 7)
    ct0_CONV = 1e3*ct0
    InGaVe = 0.01/bed_por
    rea_CSA = bed_por*(math.pi*(rea_dia**2)/4)
    VoFlRa = InGaVe*rea_CSA
    varis0 = {
        "k0": 0.0072*1e-1,
        "y_CH4": lambda x: x['MoFri'][0],
        "C_CH4": lambda x: x['SpCoi'][0],
    }
    rates0 = {"r1": lambda x: x['k0']*(x['C_CH4']**2)}
    mi = {
        "model": model,
        "operating-conditions": {
            "pressure": P, "temperature": T, "period": 10,
            "process-type": process_type,
        },
        "feed": {
            "volumetric-flowrate": VoFlRa,
            "concentration": ct0_CONV,
            "components": {"shell": ["CH4", "C2H4", "H2"], "tube": [], "medium": []},
        },
        "reactions": {"R1": "2CH4 <=> C2H4 + 2H2"},
        "reaction-rates": {"VARS": varis0, "RATES": rates0},
        "external-heat": {"OvHeTrCo": 50, "EfHeTrAr": 4/rea_dia, "MeTe": 0},
        "reactor": {
            "ReInDi": rea_dia, "ReLe": 1, "PaDi": cat_d, "BeVoFr": bed_por,
            "CaBeDe": bulk_rho, "CaDe": cat_rho, "CaSpHeCa": cat_cp/1000,
        },
        "solver-config": {"ivp": ivp, "display-result": "False"},
    }
    return mi


# ----------------------------------------------------------------------------
# synthetic sweeps
# ----------------------------------------------------------------------------
def ch4_three_reaction_input(model="N1", process_type="iso-thermal"):
    """Synthetic variant of the methane-coupling case with as many reactions as species (forward, reverse and a slow
    side path written as separate reactions): nr = nc = 3, so the steady-state integrator keeps the full state
    instead of switching to reaction extents."""
    mi = ch4_input(model, process_type)
    mi["reactions"] = {"R1": "2CH4 <=> C2H4 + 2H2", "R2": "C2H4 + 2H2 <=> 2CH4", "R3": "2CH4 <=> C2H4 + 2H2"}
    varis = dict(mi["reaction-rates"]["VARS"])
    varis["C_C2H4"] = lambda x: x['SpCoi'][1]
    varis["C_H2"] = lambda x: x['SpCoi'][2]
    mi["reaction-rates"] = {"VARS": varis, "RATES": {
        "r1": lambda x: x['k0']*(x['C_CH4']**2),
        "r2": lambda x: 0.05*x['k0']*x['C_C2H4']*x['C_H2'],
        "r3": lambda x: 0.01*x['k0']*x['C_CH4'],
    }}
    return mi


def config3_sweep(B, seed=20240611):
    """Config 3: T0 ~ U[473,573] K, P0 ~ U[2e6,8e6] Pa, H2/COx ~ U[1,3],
    CO2/COx ~ U[0.2,0.8]; float64 feed C0 = y*P0/(R*T0); Tm = T0.
    Returns the per-instance `sweep` dict understood by `rmtExeBatch`."""
    rng = np.random.default_rng(seed)
    T0 = rng.uniform(473.0, 573.0, B)
    P0 = rng.uniform(2e6, 8e6, B)
    r = rng.uniform(1.0, 3.0, B)
    c = rng.uniform(0.2, 0.8, B)
    y = feed_mole_fraction(r, c)
    C0 = y*(P0/(R_CONST*T0))[:, None]
    return {
        "temperature": T0,
        "pressure": P0,
        "concentration": C0,
        "MeTe": T0.copy(),
    }


def config3_corners():
    """The 36 corner instances of the config-3 box."""
    Ts, Ps, rs, cs = [], [], [], []
    for T0 in (473.0, 523.0, 573.0):
        for P0 in (2e6, 5e6, 8e6):
            for r in (1.0, 3.0):
                for c in (0.2, 0.8):
                    Ts.append(T0); Ps.append(P0); rs.append(r); cs.append(c)
    T0 = np.array(Ts); P0 = np.array(Ps)
    y = feed_mole_fraction(np.array(rs), np.array(cs))
    C0 = y*(P0/(R_CONST*T0))[:, None]
    return {"temperature": T0, "pressure": P0, "concentration": C0, "MeTe": T0.copy()}


def config4_population(B, seed=20240612):
    """Config 4: pre-exponentials x lognormal(sigma=0.2), activation energies
    x N(1, 0.02); operating point = config 1."""
    rng = np.random.default_rng(seed)
    sweep = {}
    for k, v in METHANOL_ARRHENIUS.items():
        if k.startswith("k0"):
            sweep[k] = v*rng.lognormal(0.0, 0.2, B)
        else:
            sweep[k] = v*rng.normal(1.0, 0.02, B)
    return sweep


def instance_input(base, sweep, i):
    """Materialise instance `i` of a sweep as a plain single-reactor
    modelInput (what one would hand to the reference's rmtExe)."""
    import copy
    mi = copy.copy(base)
    mi["operating-conditions"] = dict(base["operating-conditions"])
    mi["feed"] = dict(base["feed"])
    mi["external-heat"] = dict(base["external-heat"])
    mi["reactor"] = dict(base["reactor"])
    rr = base["reaction-rates"]
    mi["reaction-rates"] = {"VARS": dict(rr["VARS"]), "RATES": rr["RATES"]}
    for k, v in sweep.items():
        val = np.asarray(v)[i]
        if k == "temperature":
            mi["operating-conditions"]["temperature"] = float(val)
        elif k == "pressure":
            mi["operating-conditions"]["pressure"] = float(val)
        elif k == "concentration":
            mi["feed"]["concentration"] = np.array(val, dtype=np.float64)
        elif k == "volumetric-flowrate":
            mi["feed"]["volumetric-flowrate"] = float(val)
        elif k in ("MeTe", "OvHeTrCo"):
            mi["external-heat"][k] = float(val)
        elif k in mi["reactor"]:
            mi["reactor"][k] = float(val)
        elif k in mi["reaction-rates"]["VARS"]:
            mi["reaction-rates"]["VARS"][k] = float(val)
        else:
            raise KeyError(k)
    return mi
