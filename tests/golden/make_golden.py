#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE.

The reference (PyREMOT, /root/reference) is imported unmodified — with a stub
`matplotlib` because the image has none and every model module imports it at
import time (PyREMOT/library/plot.py:7) — and driven through its own public
entry point `rmtExe` (PyREMOT/rmt.py:21).  `scipy.integrate.solve_ivp` as seen
by `PyREMOT/docs/pbHomoReactor.py` is wrapped so we can (i) capture the RHS
callable + paramsSet the reference builds, (ii) inject rtol/atol/method, and
(iii) record nfev/njev.  Nothing from the reference is copied into the repo:
only numbers it produced.

This script can only run where /root/reference exists (the build container);
the GPU box consumes the committed .npz files.

usage: python tests/golden/make_golden.py <part> [...]
parts: n1 corners n2rhs n2sol_ch4 n2sol_m20_lsoda n2sol_m20_bdf n2sol_m50_bdf
       n2sol_m20_tight all_fast
"""
import io
import os
import sys
import time
import types
import contextlib
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402

REF_ROOT = os.environ.get("RMT_REFERENCE", "/root/reference")


def load_reference():
    for n in ("matplotlib", "matplotlib.pyplot"):
        m = types.ModuleType(n)
        m.__getattr__ = lambda k: (lambda *a, **kw: None)
        sys.modules.setdefault(n, m)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    warnings.simplefilter("ignore")
    import PyREMOT  # noqa
    import PyREMOT.docs.pbHomoReactor as H
    from PyREMOT.solvers import solverSetting
    H.printProgressBar = lambda *a, **k: None
    return PyREMOT, H, solverSetting


class Capture:
    """Wraps the solve_ivp symbol inside pbHomoReactor."""

    def __init__(self, H, **inject):
        self.H = H
        self.inject = inject
        self.calls = []
        self._orig = H.solve_ivp

    def __enter__(self):
        def patched(fun, t_span, y0, **kw):
            kw = dict(kw)
            kw.update(self.inject)
            sol = self._orig(fun, t_span, y0, **kw)
            self.calls.append(dict(fun=fun, t_span=np.array(t_span, float), y0=np.array(y0, float),
                                   args=kw.get("args"), nfev=sol.nfev, njev=sol.njev,
                                   nlu=sol.nlu, t=sol.t, y=sol.y, method=kw.get("method")))
            return sol
        self.H.solve_ivp = patched
        return self

    def __exit__(self, *a):
        self.H.solve_ivp = self._orig


def run_ref(PyREMOT, mi):
    buf = io.StringIO()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(buf):
        res = PyREMOT.rmtExe(mi)
    return res, time.perf_counter() - t0


def consts_from_params(paramsSet):
    """Flatten the numeric constants of the reference's paramsSet tuple."""
    FunParam, DA = paramsSet[2], paramsSet[3]
    out = {}
    for k in ("CrSeAr", "GaMiVi", "varNo"):
        out["const_" + k] = np.array(FunParam["const"][k], dtype=float)
    out["const_StHeRe25"] = np.array(FunParam["const"]["StHeRe25"], float)
    out["const_MoWei"] = np.array(FunParam["const"]["MoWei"], float)
    for k in ("SpCo0", "GaDe0", "GaCpMeanMix0", "P0", "T0", "VoFlRa0"):
        out["bc_" + k] = np.array(FunParam["constBC1"][k], float)
    out["bc_SpCoi0"] = np.array(FunParam["constBC1"]["SpCoi0"], float)
    for k in ("Cf", "Tf", "Pf", "vf", "zf", "Cpf", "GaHeCoTe0"):
        out["da_" + k] = np.array(DA[k], float)
    for k in ("Cif", "Cpif", "GaMaCoTe0"):
        out["da_" + k] = np.array(DA[k], float)
    out["exhe_EfHeTrAr"] = np.array(FunParam["ExHe"]["EfHeTrAr"], float)
    return out


def n1_case_inputs():
    return {
        "methanol_readme": cases.methanol_readme_input("N1"),
        "methanol_testfile": cases.methanol_testfile_input("N1"),
        "ch4_noniso": cases.ch4_input("N1", "non-iso-thermal"),
        "ch4_iso": cases.ch4_input("N1", "iso-thermal"),
    }


def part_n1():
    PyREMOT, H, _ = load_reference()
    rng = np.random.default_rng(7)
    out = {}
    for name, mi in n1_case_inputs().items():
        # default tolerance, each solver
        for meth in ("default", "BDF", "Radau"):
            mi["solver-config"]["ivp"] = meth
            with Capture(H) as cap:
                res, wall = run_ref(PyREMOT, mi)
            c = cap.calls[0]
            dp = res["resModel"][0]
            tag = f"{name}__{meth}"
            out[tag + "__dataYs"] = np.array(dp["dataYs"])
            out[tag + "__soly"] = np.array(c["y"])
            out[tag + "__nfev_njev_wall"] = np.array([c["nfev"], c["njev"], wall])
            if meth == "default":
                out[name + "__dataXs"] = np.array(dp["dataXs"])
                out[name + "__dataYCons1"] = np.array(dp["dataYCons1"])
                out[name + "__dataYCons2"] = np.array(dp["dataYCons2"])
                out[name + "__dataYTemp1"] = np.array(dp["dataYTemp1"])
                out[name + "__dataYTemp2"] = np.array(dp["dataYTemp2"])
                out[name + "__labelList"] = np.array(dp["labelList"])
                out[name + "__indexList"] = np.array(dp["indexList"])
                fun, ps, y0, traj = c["fun"], c["args"][0], c["y0"], c["y"]
        mi["solver-config"]["ivp"] = "default"
        # tight tolerance
        for meth in ("LSODA", "BDF"):
            with Capture(H, rtol=1e-10, atol=1e-12, method=meth) as cap:
                res, wall = run_ref(PyREMOT, mi)
            c = cap.calls[0]
            tag = f"{name}__tight_{meth}"
            out[tag + "__dataYs"] = np.array(res["resModel"][0]["dataYs"])
            out[tag + "__soly"] = np.array(c["y"])
            out[tag + "__nfev_njev_wall"] = np.array([c["nfev"], c["njev"], wall])
        # RHS known answers: feed state, trajectory states, random perturbations
        idx = [0, 1, 2, 3, 5, 10, 25, 50, 75, 100]
        Y = [y0] + [traj[:, i] for i in idx[1:]]
        for i in idx:
            for _ in range(3):
                Y.append(traj[:, i]*(1 + 0.05*rng.uniform(-1, 1, traj.shape[0])))
        Y = np.array(Y)
        F = np.array([fun(0.0, list(y), ps) for y in Y])
        out[name + "__rhs_Y"] = Y
        out[name + "__rhs_F"] = F
        for k, v in consts_from_params(ps).items():
            out[name + "__" + k] = v
        print(name, "done", flush=True)
    # rates known answer (SURVEY App. B.2 last entry) through the reference's reactionRateExe
    from PyREMOT.docs.rmtReaction import reactionRateExe
    y = np.array([0.45, 0.26, 0.024, 0.24, 0.007, 0.016]); y = y/y.sum()
    kin = cases.methanol_kinetics(1982*(1 - 0.39))
    Pt, Tt = 4.99e6, 620.0
    C = y*Pt/(cases.R_CONST*Tt)
    out["rates_ka__inputs"] = np.concatenate([[Tt, Pt], y, C])
    out["rates_ka__R"] = np.array(reactionRateExe((Tt, Pt, y, C), kin["VARS"], kin["RATES"]))
    import scipy
    out["versions"] = np.array([np.__version__, scipy.__version__, sys.version.split()[0]])
    np.savez_compressed(os.path.join(HERE, "n1_reference.npz"), **out)


def part_corners():
    PyREMOT, H, _ = load_reference()
    base = cases.methanol_readme_input("N1")
    sweep = cases.config3_corners()
    B = len(sweep["temperature"])
    dflt, tight, stats = [], [], []
    for i in range(B):
        mi = cases.instance_input(base, sweep, i)
        with Capture(H) as cap:
            res, wall = run_ref(PyREMOT, mi)
        dflt.append(np.array(res["resModel"][0]["dataYs"]))
        c0 = cap.calls[0]
        with Capture(H, rtol=1e-10, atol=1e-12, method="LSODA") as cap:
            res, wall_t = run_ref(PyREMOT, mi)
        tight.append(np.array(res["resModel"][0]["dataYs"]))
        stats.append([c0["nfev"], c0["njev"], wall, cap.calls[0]["nfev"], wall_t])
        print("corner", i, stats[-1], flush=True)
    np.savez_compressed(os.path.join(HERE, "n1_corners_reference.npz"),
                        default_dataYs=np.array(dflt), tight_dataYs=np.array(tight),
                        stats=np.array(stats), **{"sweep_" + k: v for k, v in sweep.items()})


def _capture_n2(PyREMOT, H, mi, **inject):
    with Capture(H, **inject) as cap:
        res, wall = run_ref(PyREMOT, mi)
    return res, wall, cap.calls


def part_n2rhs():
    from scipy.integrate import solve_ivp
    PyREMOT, H, solverSetting = load_reference()
    rng = np.random.default_rng(11)
    out = {}
    # capture by letting the reference run with a 1e-9 s period (cheap) so runN2 builds paramsSet
    for name, mi, zNo in (("methanol_testfile_z20", cases.methanol_testfile_input("N2"), 20),
                          ("methanol_readme_z50", cases.methanol_readme_input("N2"), 50),
                          ("ch4_z20", cases.ch4_input("N2"), 20)):
        solverSetting['N2']['zNo'] = zNo
        mi["operating-conditions"]["period"] = 1e-9
        res, wall, calls = _capture_n2(PyREMOT, H, mi)
        fun, ps, y0 = calls[0]["fun"], calls[0]["args"][0], calls[0]["y0"]
        # a real mid-transient state: integrate the captured reference RHS for a short time
        tmid = 0.05 if "ch4" not in name else 2.0
        sol = solve_ivp(fun, [0, tmid], y0, method="BDF", args=(ps,), rtol=1e-5, atol=1e-8)
        ymid = sol.y[:, -1]
        Y = [y0, ymid, sol.y[:, len(sol.t)//2]]
        for base in (y0, ymid):
            for _ in range(3):
                Y.append(base*(1 + 0.05*rng.uniform(-1, 1, base.size)) + 1e-4*rng.uniform(0, 1, base.size))
        # clamp exercise: a few non-positive concentrations
        yneg = ymid.copy(); yneg[1*zNo + 3] = -1e-7; yneg[2*zNo + 5] = 0.0
        Y.append(yneg)
        Y = np.array(Y)
        F = np.array([fun(0.0, y, ps) for y in Y])
        out[name + "__rhs_Y"] = Y
        out[name + "__rhs_F"] = F
        out[name + "__zNo"] = np.array(zNo)
        print(name, "rhs done", sol.nfev, flush=True)
    solverSetting['N2']['zNo'] = 20
    np.savez_compressed(os.path.join(HERE, "n2_rhs_reference.npz"), **out)


def part_n2rhs_iso():
    """N2 with process-type "iso-thermal" (nc unknowns per node, T frozen at the feed temperature): RHS known answers
    and the reference's own default-tolerance run (methane case, 20 nodes)."""
    from scipy.integrate import solve_ivp
    PyREMOT, H, solverSetting = load_reference()
    rng = np.random.default_rng(19)
    out = {}
    zNo = 20
    solverSetting['N2']['zNo'] = zNo
    mi = cases.ch4_input("N2", "iso-thermal")
    mi["operating-conditions"]["period"] = 1e-9
    res, wall, calls = _capture_n2(PyREMOT, H, mi)
    fun, ps, y0 = calls[0]["fun"], calls[0]["args"][0], calls[0]["y0"]
    sol = solve_ivp(fun, [0, 2.0], y0, method="BDF", args=(ps,), rtol=1e-5, atol=1e-8)
    ymid = sol.y[:, -1]
    Y = [y0, ymid, sol.y[:, len(sol.t)//2]]
    for base in (y0, ymid):
        for _ in range(3):
            Y.append(base*(1 + 0.05*rng.uniform(-1, 1, base.size)) + 1e-4*rng.uniform(0, 1, base.size))
    yneg = ymid.copy(); yneg[1*zNo + 3] = -1e-7; yneg[2*zNo + 5] = 0.0
    Y.append(yneg)
    Y = np.array(Y)
    out["rhs_Y"] = Y
    out["rhs_F"] = np.array([fun(0.0, y, ps) for y in Y])
    out["zNo"] = np.array(zNo)
    res, wall, calls = _capture_n2(PyREMOT, H, cases.ch4_input("N2", "iso-thermal"))
    out.update({"default__" + k: v for k, v in _n2_pack(res, calls, wall).items()})
    np.savez_compressed(os.path.join(HERE, "n2_iso_reference.npz"), **out)
    print("n2 iso done: %d RHS states, default run nfev %s wall %.1f" % (len(Y), out["default__nfev"], wall))


def _n2_pack(res, calls, wall):
    dps = res["resModel"]["dataPack"]
    return dict(dataYs=np.array([d["dataYs"] for d in dps]),
                dataTime=np.array([d["dataTime"] for d in dps]),
                dataXs=np.array(dps[0]["dataXs"]),
                soly_last=np.array([c["y"][:, -1] for c in calls]),
                nfev=np.array([c["nfev"] for c in calls]), njev=np.array([c["njev"] for c in calls]),
                wall=np.array(wall))


def part_n2sol(which):
    PyREMOT, H, solverSetting = load_reference()
    cfg = {
        "ch4": ("ch4", 20, {}, "n2_sol_ch4_reference.npz"),
        "ch4_tight": ("ch4", 20, dict(rtol=1e-10, atol=1e-12, method="LSODA"), "n2_sol_ch4_tight_reference.npz"),
        "m20_lsoda": ("methanol_testfile", 20, {}, "n2_sol_m20_lsoda_reference.npz"),
        "m20_bdf": ("methanol_testfile", 20, dict(method="BDF"), "n2_sol_m20_bdf_reference.npz"),
        "m50_bdf": ("methanol_readme", 50, dict(method="BDF"), "n2_sol_m50_bdf_reference.npz"),
        "m20_tight": ("methanol_testfile", 20, dict(method="BDF", rtol=1e-8, atol=1e-10), "n2_sol_m20_tight_reference.npz"),
        "mr20_lsoda": ("methanol_readme", 20, {}, "n2_sol_mr20_lsoda_reference.npz"),
    }[which]
    name, zNo, inject, fname = cfg
    mi = {"ch4": cases.ch4_input, "methanol_testfile": cases.methanol_testfile_input,
          "methanol_readme": cases.methanol_readme_input}[name]("N2")
    solverSetting['N2']['zNo'] = zNo
    res, wall, calls = _capture_n2(PyREMOT, H, mi, **inject)
    solverSetting['N2']['zNo'] = 20
    out = _n2_pack(res, calls, wall)
    out["zNo"] = np.array(zNo)
    print(which, "wall", wall, "nfev", out["nfev"], flush=True)
    np.savez_compressed(os.path.join(HERE, fname), **out)


def part_m7():
    """Model M7 (pbReactor.runM3, the dimensional twin of N1): RHS known answers and solutions."""
    PyREMOT, H, _ = load_reference()
    import PyREMOT.docs.pbReactor as PB
    PB.pltc.plots2DSub = staticmethod(lambda *a, **k: None)      # runM3 always plots (pbReactor.py:1350-1356)
    rng = np.random.default_rng(13)
    out = {}
    orig = PB.solve_ivp

    def run(inject):
        calls = []

        def patched(fun, t_span, y0, **kw):
            kw = dict(kw); kw.update(inject)
            sol = orig(fun, t_span, y0, **kw)
            calls.append(dict(fun=fun, y0=np.array(y0, float), args=kw.get("args"), nfev=sol.nfev, y=sol.y, t=sol.t))
            return sol
        PB.solve_ivp = patched
        try:
            res, wall = run_ref(PyREMOT, cases.methanol_m7_input())
        finally:
            PB.solve_ivp = orig
        return res, calls[0], wall
    res, c, wall = run({})
    out["default__dataYs"] = np.array(res["resModel"]["dataYs"])
    out["default__soly"] = c["y"]; out["default__t"] = c["t"]
    out["default__nfev_wall"] = np.array([c["nfev"], wall])
    res_t, ct, wall_t = run(dict(rtol=1e-10, atol=1e-12, method="LSODA"))
    out["tight__dataYs"] = np.array(res_t["resModel"]["dataYs"])
    out["tight__soly"] = ct["y"]
    out["tight__nfev_wall"] = np.array([ct["nfev"], wall_t])
    fun, args, traj = c["fun"], c["args"], c["y"]
    Y = [c["y0"]] + [traj[:, i] for i in (1, 2, 3, 5, 10, 20, 29)]
    for i in (0, 1, 3, 10, 29):
        for _ in range(3):
            Y.append(traj[:, i]*(1 + 0.05*rng.uniform(-1, 1, traj.shape[0])))
    Y = np.array(Y)
    out["rhs_Y"] = Y
    out["rhs_F"] = np.array([fun(0.0, y, *args) for y in Y])
    np.savez_compressed(os.path.join(HERE, "m7_reference.npz"), **out)
    print("m7 done: default nfev %d wall %.2f, tight nfev %d" % (c["nfev"], wall, ct["nfev"]))


def part_m9():
    """Model M9 (pbReactor.runM5, the dimensional dynamic twin of N2): RHS known answers and the slab-end states of
    a default-tolerance run, on a 12-node grid (solverSetting['S2'] is the reference's module-level setting)."""
    PyREMOT, H, solverSetting = load_reference()
    import PyREMOT.docs.pbReactor as PB
    PB.pltc.plots2DSub = staticmethod(lambda *a, **k: None)      # runM5 always plots (pbReactor.py:2243-2246)
    old = dict(solverSetting["S2"])
    solverSetting["S2"].update(zNo=12, tNo=3)
    rng = np.random.default_rng(17)
    out = {}
    orig = PB.solve_ivp
    calls = []

    def patched(fun, t_span, y0, **kw):
        sol = orig(fun, t_span, y0, **kw)
        calls.append(dict(fun=fun, y0=np.array(y0, float), args=kw.get("args"), nfev=sol.nfev, y=sol.y, t=sol.t))
        return sol
    PB.solve_ivp = patched
    try:
        res, wall = run_ref(PyREMOT, cases.methanol_m9_input())
    finally:
        PB.solve_ivp = orig
        solverSetting["S2"].update(old)
    rm = res["resModel"]
    out["default__T_profiles"] = np.array([xy[1] for xy in rm["XYList"]])            # temperature at the end of each slab
    out["default__x"] = np.array(rm["XYList"][0][0])
    out["default__legends"] = np.array([d["leg"] for d in rm["dataList"]])
    out["default__slab_end_states"] = np.array([c["y"][:, -1] for c in calls])       # [tNo][(nc+1)*zNo]
    out["default__nfev_wall"] = np.array([sum(c["nfev"] for c in calls), wall])
    fun, args = calls[0]["fun"], calls[0]["args"]
    Y = [calls[0]["y0"]] + [c["y"][:, k] for c in calls for k in (1, -1)]
    for c in calls:
        for _ in range(2):
            Y.append(c["y"][:, -1]*(1 + 0.05*rng.uniform(-1, 1, c["y"].shape[0])))
    neg = np.array(calls[0]["y"][:, 2]); neg[[62, 64, 70]] = -1e-9                    # clamped entries (:2487-2491): DME at three nodes (a clamped reactant makes the rates 1e31)
    Y.append(neg)
    Y = np.array(Y)
    out["rhs_Y"] = Y
    out["rhs_F"] = np.array([fun(0.0, y, *args) for y in Y])
    np.savez_compressed(os.path.join(HERE, "m9_reference.npz"), **out)
    print("m9 done: nfev %d wall %.1f s, %d RHS states" % (out["default__nfev_wall"][0], wall, len(Y)))


def part_props():
    """Component-table known answers: Cp_i(T), viscosity_i(T), dHf25, MW for all 12 species,
    Wilke mixture viscosity and reaction parsing, straight from the reference's helpers."""
    PyREMOT, H, _ = load_reference()
    from PyREMOT.docs.rmtThermo import calHeatCapacityAtConstantPressure, calStandardEnthalpyOfReaction
    from PyREMOT.docs.gasTransPor import calGasViscosity, calMixturePropertyM1
    from PyREMOT.data import componentSymbolList, componentDataStore
    from PyREMOT.docs.rmtUtility import rmtUtilityClass as U
    syms = list(componentSymbolList)
    Ts = np.array([298.15, 473.0, 523.0, 620.5, 973.0])
    cp = np.array([calHeatCapacityAtConstantPressure(syms, T) for T in Ts])
    mu = np.array([calGasViscosity(syms, T) for T in Ts])
    MW = np.array([c["MW"] for c in componentDataStore["payload"]], float)
    dHf = np.array([c["dHf25"]["val"] for c in componentDataStore["payload"]], float)
    y = np.arange(1, 13, dtype=float); y /= y.sum()
    wilke = np.array([calMixturePropertyM1(12, m, y, MW) for m in mu])
    reactions = {"R1": "CO2 + 3H2 <=> CH3OH + H2O", "R2": "CO + H2O <=> H2 + CO2", "R3": "2CH3OH <=> DME + H2O",
                 "R4": "2CH4 <=> C2H4 + 2H2", "R5": "C3H8 => C3H6 + H2", "R6": "0.5C4H10+1.5N2=C3H6 + CH4"}
    dH = np.array([calStandardEnthalpyOfReaction(r) for r in reactions.values()])
    vec = U.buildReactionCoeffVector(U.buildReactionCoefficient(reactions))
    flat = np.array([[j, syms.index(s), v] for j, r in enumerate(vec) for s, v in r], float)
    np.savez_compressed(os.path.join(HERE, "component_props_reference.npz"), symbols=np.array(syms), Ts=Ts, cp=cp, mu=mu,
                        MW=MW, dHf25=dHf, wilke_y=y, wilke=wilke, reactions=np.array(list(reactions.values())),
                        dH25=dH, stoich=flat, rmtCom=np.array(PyREMOT.rmtCom()))
    print("props done")


class PlotRecorder:
    """Stand-in for matplotlib.pyplot that records what is drawn."""

    def __init__(self):
        self.calls = []

    def install(self, module):
        for name in ("plot", "title", "xlabel", "ylabel", "legend", "show", "figure", "subplots"):
            setattr(module, name, self._make(name))

    def _make(self, name):
        def f(*a, **kw):
            if name == "plot":
                x, y = np.asarray(a[0], float), np.asarray(a[1], float)
                self.calls.append(["plot", str(kw.get("label")), int(x.size), float(x[0]), float(x[-1]), float(y[0]), float(y[-1])])
            elif name in ("title", "xlabel", "ylabel"):
                self.calls.append([name, str(a[0])])
            else:
                self.calls.append([name])
        return f


def synthetic_packs():
    """Result dictionaries with the reference's schema (pbHomoReactor.py:2991-3007, :3664-3696) and made-up numbers."""
    xs = np.linspace(0, 1, 5)
    ys = np.arange(25, dtype=float).reshape(5, 5)/7.0
    steady = [{"modelId": "N1", "processType": "non-iso-thermal", "successStatus": True, "computation-time": 0.123,
               "dataShape": xs.shape, "labelList": ["A", "B", "C", "Pressure", "Temperature"], "indexList": [3, 3, 4],
               "dataTime": [], "dataXs": xs, "dataYs": ys}]
    steady_iso = [dict(steady[0], processType="iso-thermal", labelList=["A", "B", "C", "Pressure"], dataYs=ys[:4])]
    dyn = {"computation-time": 4.5, "dataPack": [
        {"modelId": "N2", "processType": "non-iso-thermal", "successStatus": True, "labelList": ["A", "B", "C", "Temperature"],
         "indexList": [3, 4, 3], "dataTime": 0.1*(i + 1), "dataXs": xs, "dataYs": ys[:4] + i} for i in range(6)]}
    return steady, steady_iso, dyn


def part_plots():
    """What the reference's plot hooks draw (solResultAnalysis.py:307-459) for synthetic result dictionaries ->
    tests/golden/plot_calls_reference.json (compared call by call with rmt_app_b200/plotting.py)."""
    import json
    load_reference()
    import PyREMOT.solvers.solResultAnalysis as SRA
    import types
    import PyREMOT.library.plot as PL
    rec = PlotRecorder()
    PL.plt = types.SimpleNamespace()              # the symbol plots2D draws through (library/plot.py:7)
    rec.install(PL.plt)
    steady, steady_iso, dyn = synthetic_packs()
    out = {}
    SRA.plotResultsSteadyState(steady); out["steady"] = rec.calls; rec.calls = []
    SRA.plotResultsSteadyState(steady_iso); out["steady_iso"] = rec.calls; rec.calls = []
    for seed in (7, 8):
        np.random.seed(seed)
        SRA.plotResultsDynamic(dyn, 6); out["dynamic_seed%d" % seed] = rec.calls; rec.calls = []
    with open(os.path.join(HERE, "plot_calls_reference.json"), "w") as f:
        json.dump(out, f, indent=0)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    for part in sys.argv[1:]:
        if part == "n1":
            part_n1()
        elif part == "corners":
            part_corners()
        elif part == "m7":
            part_m7()
        elif part == "m9":
            part_m9()
        elif part == "n2rhs_iso":
            part_n2rhs_iso()
        elif part == "props":
            part_props()
        elif part == "plots":
            part_plots()
        elif part == "n2rhs":
            part_n2rhs()
        elif part.startswith("n2sol_"):
            part_n2sol(part[len("n2sol_"):])
        else:
            raise SystemExit("unknown part " + part)
