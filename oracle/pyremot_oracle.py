"""CPU oracle for the PyREMOT N1 / N2 hot path  —  TEST INFRASTRUCTURE ONLY.

This file is a plain NumPy/SciPy restatement of what the reference computes on
the path named by BASELINE.json (`rmtExe` with model "N1" or "N2", plus their
dimensional twins "M7" and "M9" from the next-row list of SURVEY 8(f)).  It exists
so that the CUDA path can be checked on machines where /root/reference is not
available (the GPU box).  It is NOT part of the product: only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product (`rmt_app_b200`) never does and fails
loudly without its CUDA library.

Parity status: PINNED.  `tests/test_oracle_golden*.py` check every function
below against numbers produced by running the unmodified reference in the
build container (`tests/golden/make_golden.py` -> `tests/golden/*.npz`):
RHS values to ~1e-14 relative, solutions through the same SciPy integrators
to the integrator's own reproducibility.

The integrator itself is third-party in the reference too:
`scipy.integrate.solve_ivp` (unpinned in the reference's setup.py:26-27;
SciPy 1.18.1 in this image), called at PyREMOT/docs/pbHomoReactor.py:2931 and
:3609 with method LSODA unless `solver-config.ivp` says otherwise and with
SciPy's default rtol=1e-3 / atol=1e-6.  The oracle calls the same function.

Each function cites the reference file:line it follows (paths relative to
/root/reference/PyREMOT/).
"""
import math
import re
import types

import numpy as np
from scipy.integrate import solve_ivp

R_CONST = 8.314472          # core/constants.py:8
EPS_CONST = 1e-30           # core/constants.py:11
PI_CONST = math.pi          # core/constants.py:14
Tref = 273.15 + 25.00       # core/constants.py:17-23

# N1/N2 grid sizes, solvers/solSetting.py:30-39 (module-level mutable dict in
# the reference; same here so tests can override zNo exactly as callers do).
solverSetting = {"N1": {"zNo": 100}, "N2": {"zNo": 20, "rNo": 5, "tNo": 5, "timesNo": 5}, "M9": {"zNo": 30},
                 "S2": {"tNo": 10, "zNo": 100, "rNo": 7, "timesNo": 5}}

# ----------------------------------------------------------------------------
# component data — data/componentData.py:11-22 (MW), :72-86 (dHf25),
# :119-405 (Cp polynomial a0 + a1*T + a2*T**2 + a3*T**3 [J/mol/K]),
# data/dataGasViscosity.py:8-141 (viscosity eq.1 parameters; DME uses eq.2).
# ----------------------------------------------------------------------------
_DB = {
    #          MW      dHf25     Cp a0     a1          a2           a3          viscosity [A, B, C, D]
    "CO2":   (44.01, -393.51, (22.243, 5.98E-02, -3.50E-05, 7.46E-09), (4.719875, 0.373279, 512.686300, -6119.961)),
    "H2":    (2.0,      0.0,  (26.879, 4.35E-03, -3.30E-07, None),     (0.169104, 0.692485, -7.634394, 467.120)),
    "CH3OH": (32.04, -200.7,  (19.038, 9.15E-02, -1.22E-05, -8.03E-09), (0.477915, 0.641076, 284.838034, -3230.713)),
    "H2O":   (18.01, -241.820, (29.163, 1.45E-02, -2.02E-06, None),    (0.501246, 0.709247, 869.465599, -90063.891)),
    "CO":    (28.01, -110.53, (27.113, 6.55E-03, -1.00E-06, None),     (0.734306, 0.588574, 52.318660, 1018.822)),
    "DME":   (46.07, -184.1,  (19.8, 0.17, -5.66e-5, None),            None),
    "N2":    (28,       0,    (28.883, -1.57E-03, 8.08E-06, -2.87E-09), (0.847662, 0.574033, 75.437536, 56.771)),
    "CH4":   (16.04,  -74.90, (19.875, 5.021E-02, 1.268E-05, -11.004E-09), (1.119178, 0.493234, 214.627200, -3952.087)),
    "C2H4":  (28.05,   52.32, (3.950, 15.628E-02, -8.339E-05, 17.657E-09), (1.503552, 0.456140, 288.342422, 73.362)),
    "C3H6":  (42.08,   20.4,  (3.151, 23.812E-02, -12.176E-05, 24.603E-09), (0.876767, 0.520871, 293.618650, -182.857)),
    "C3H8":  (44.1,  -103.9,  (-4.042, 30.456E-02, -15.711E-05, 31.716E-09), (0.173966, 0.734798, 143.207060, -7147.859)),
    "C4H10": (58.12, -126.2,  (-7.908, 41.573E-02, -22.992E-05, 49.875E-09), (0.075828, 0.837082, 67618677, -2141.762)),
}
componentSymbolList = tuple(_DB.keys())


def cp_component(sym, T):
    """Cp_i(T) [J/mol/K]: rmtThermo.py:16-49 evaluates the DB string
    "a0 + a1*T + a2*(T**2) + a3*(T**3)" left to right (componentData.py:119...)."""
    a0, a1, a2, a3 = _DB[sym][2]
    v = a0 + a1*T + a2*(T**2)
    if a3 is not None:
        v = v + a3*(T**3)
    return v


def cp_mean_list(comList, T):
    """rmtThermo.py:52-75: (Cp(Tref) + Cp(T))*0.5 per component."""
    return np.array([(cp_component(s, Tref) + cp_component(s, T))*0.50 for s in comList])


def gas_viscosity(comList, T):
    """gasTransPor.py:137-154 (eq.1), dataGasViscosity.py:133 (DME, eq.2)."""
    out = []
    for s in comList:
        p = _DB[s][3]
        if p is None:
            if s != "DME":
                raise KeyError(s)
            out.append(2.68e-7*(T**0.3975)/(1+(534/T)))
        else:
            A, B, C, D = p
            out.append(A*1e-6*(T**B)/(1+C*(1/T)+D*(T**-2)))
    return np.array(out)


def wilke_mixture(compNo, Xi, MoFri, MWi):
    """gasTransPor.py:229-274 (method of Wilke)."""
    w = np.zeros((compNo, compNo))
    for i in range(compNo):
        for j in range(compNo):
            if i == j:
                w[i, j] = 1
            elif i < j:
                A = 1 + math.sqrt(Xi[i]/Xi[j])*((MWi[j]/MWi[i])**(1/4))
                w[i, j] = (A**2)/math.sqrt(8*(1+(MWi[i]/MWi[j])))
            else:
                w[i, j] = (Xi[i]/Xi[j])*(MWi[j]/MWi[i])*w[j, i]
    mix = np.zeros(compNo)
    for i in range(compNo):
        mix[i] = (Xi[i]*MoFri[i])/np.sum(MoFri*w[i, :])
    return np.sum(mix)


_TERM = re.compile(r"([0-9.]*)([a-zA-Z0-9.]+)")


def parse_reactions(reactionDict):
    """rmtUtility.py:172-249: reaction strings -> sorted reactant/product
    lists with signed coefficients, and the [[symbol, nu], ...] vectors."""
    rs, vec = [], []
    for reaction in reactionDict.values():
        lhs, rhs = reaction.replace("<", "").replace(">", "").replace(" ", "").split("=")
        reac = [{"symbol": s, "coeff": -1*float(c) if len(c) else -1.0} for c, s in _TERM.findall(lhs)]
        prod = [{"symbol": s, "coeff": float(c) if len(c) else 1.0} for c, s in _TERM.findall(rhs)]
        rs.append({"reactants": reac, "products": prod})
        vec.append([[d["symbol"], float(d["coeff"])] for d in reac + prod])
    return rs, vec


def standard_enthalpy_of_reaction(reaExpr):
    """rmtThermo.py:129-198: 1000*(sum_prod nu*dHf25 - sum_react nu*dHf25) [J/mol]."""
    lhs, rhs = reaExpr.replace("<", "").replace(">", "").replace(" ", "").split("=")
    def side(txt):
        return np.sum(np.array([_DB[s][1]*float(c) if len(c) else _DB[s][1]*1 for c, s in _TERM.findall(txt)]))
    return (side(rhs) - side(lhs))*1000.00


def enthalpy_change_of_reaction(reactionListSorted, T):
    """rmtThermo.py:258-312: (sum_i nu_i*CpMean_i)*(T - Tref) per reaction,
    products' dot + reactants' dot."""
    out = []
    for item in reactionListSorted:
        r = np.dot(cp_mean_list([i['symbol'] for i in item['reactants']], T),
                   np.array([i['coeff'] for i in item['reactants']]))
        p = np.dot(cp_mean_list([i['symbol'] for i in item['products']], T),
                   np.array([i['coeff'] for i in item['products']]))
        out.append((p + r)*(T - Tref))
    return out


def reaction_rate_exe(loopVars, varDict, rateDict):
    """rmtReaction.py:11-61: ordered evaluation of VARS (callables get the
    dict built so far) then RATES; result is a list matched to reactions by
    position."""
    T, P, MoFri, SpCoi = loopVars
    merged = {"R_CONST": R_CONST, "T": T, "P": P, "MoFri": MoFri, "SpCoi": SpCoi}
    merged.update(varDict)
    exe = {}
    for k, v in merged.items():
        exe[k] = v(exe) if isinstance(v, types.FunctionType) else v
    return [f(exe) for f in rateDict.values()]


def component_formation_rate(compNo, comList, reactionStochCoeff, Ri):
    """rmtReaction.py:64-97: r_i = sum_j nu_ij R_j with string matching,
    accumulated in reaction order then term order."""
    ri = np.zeros(compNo)
    for k in range(compNo):
        acc = 0
        for m in range(len(reactionStochCoeff)):
            for sym, nu in reactionStochCoeff[m]:
                if comList[k] == sym:
                    acc += nu*Ri[m]
        ri[k] = acc
    return ri


class HomoReactorSetup:
    """Per-solve constants of runN1 (docs/pbHomoReactor.py:2708-2852) /
    runN2 (:3332-3507); identical arithmetic in both."""

    def __init__(self, modelInput):
        mi = modelInput
        self.modelInput = mi
        self.modelId = mi['model']
        oc = mi['operating-conditions']
        self.P, self.T = oc['pressure'], oc['temperature']
        self.processType = oc['process-type']
        self.iso = self.processType == "iso-thermal"      # docs/modelSetting.py:21-23
        self.reactionList = list(mi['reactions'].values())
        self.reactionListSorted, self.reactionStochCoeff = parse_reactions(mi['reactions'])
        self.varis = mi['reaction-rates']['VARS']
        self.rates = mi['reaction-rates']['RATES']
        self.compList = list(mi['feed']['components']['shell'])
        for c in self.compList:                             # rmt.py:55-57
            if c not in componentSymbolList:
                raise Exception("Component database is not up to date!")
        self.compNo = nc = len(self.compList)
        rs = mi['reactor']
        self.ReSpec = rs
        self.ReInDi, self.ReLe, self.PaDi = rs['ReInDi'], rs['ReLe'], rs['PaDi']
        self.BeVoFr, self.CaBeDe = rs['BeVoFr'], rs['CaBeDe']
        self.CrSeAr = PI_CONST*(self.ReInDi ** 2)/4                     # :2751
        self.VoFlRa0 = mi['feed']['volumetric-flowrate']
        self.SpCoi0 = 1*np.array(mi['feed']['concentration'], dtype=float)
        self.SpCo0 = np.sum(self.SpCoi0)
        self.SuGaVe0_run = self.VoFlRa0/self.CrSeAr                     # :2763
        self.MoFri0 = self.SpCoi0/np.sum(self.SpCoi0)
        self.MoWei = [_DB[s][0] for s in self.compList]
        eh = mi['external-heat']
        self.Tm, self.U = eh['MeTe'], eh['OvHeTrCo']
        self.a = 4/self.ReInDi                                          # :2778 (EfHeTrAr ignored)
        self.GaVii0 = gas_viscosity(self.compList, self.T)
        self.GaMiVi = wilke_mixture(nc, self.GaVii0, self.MoFri0, self.MoWei)   # :2782-2783
        self.GaCpMeanList0 = cp_mean_list(self.compList, self.T)
        self.GaCpMeanMix0 = np.dot(self.MoFri0, self.GaCpMeanList0)     # :2787-2790
        self.MiMoWe0 = np.dot(np.copy(self.MoFri0), np.array(self.MoWei))*1e-3   # rmtUtility.py:57-95
        self.GaDe0 = self.MiMoWe0*self.SpCo0                            # :2796
        self.StHeRe25 = np.array([standard_enthalpy_of_reaction(r) for r in self.reactionList])

    def scales(self, vf):
        """Dimensionless-analysis block (:2799-2823 / :3442-3466)."""
        self.Cif = np.copy(self.SpCoi0)
        self.Cf, self.Tf, self.Pf = self.SpCo0, self.T, self.P
        self.vf, self.zf = vf, self.ReLe
        self.Cpf = self.GaCpMeanMix0
        self.GaMaCoTe0 = (vf/self.zf)*np.repeat(np.max(self.Cif), self.compNo)   # GaMaCoTe0 == "MAX"
        self.GaHeCoTe0 = (self.GaDe0*vf*self.Tf*(self.Cpf/self.MiMoWe0)/self.zf)

    # shared per-point physics -------------------------------------------------
    def heat_exchange(self, T):
        """rmtUtility.py:424-452 (Tm == 0 => adiabatic)."""
        return 0 if self.Tm == 0 else self.U*self.a*(self.Tm - T)


class N1Oracle(HomoReactorSetup):
    """Steady-state dimensionless model: runN1 + modelEquationN1
    (docs/pbHomoReactor.py:2694-3314)."""

    def __init__(self, modelInput):
        super().__init__(modelInput)
        self.scales(self.SuGaVe0_run)
        nc = self.compNo
        self.varNo = nc + 1 if self.iso else nc + 2
        IV = np.zeros(self.varNo)
        IV[0:nc] = self.SpCoi0/np.max(self.SpCoi0)        # :2833
        IV[nc] = self.P/self.Pf
        if not self.iso:
            IV[nc+1] = (self.T - self.Tf)/self.Tf
        self.IV = IV
        self.times = np.linspace(0, 1, solverSetting['N1']['zNo']+1)

    def rhs(self, t, y):
        """modelEquationN1, :3017-3314."""
        nc = self.compNo
        BeVoFr, PaDi = self.BeVoFr, self.PaDi
        InGaVe0 = self.VoFlRa0/(self.CrSeAr*BeVoFr)            # :3137
        SuGaVe0 = InGaVe0*BeVoFr
        yLoop = np.array(y)
        CoSpi = yLoop[0:nc]
        P = yLoop[nc]
        T = yLoop[nc+1] if not self.iso else 0
        CoSpi_ReVa = CoSpi*np.max(self.SpCoi0)                 # :3159-3162
        CoSp_ReVa = np.sum(CoSpi_ReVa)
        T_ReVa = T*self.Tf + self.Tf
        P_ReVa = P*self.Pf
        MoFri = CoSpi_ReVa/np.sum(CoSpi_ReVa)
        InGaVe = InGaVe0*(CoSp_ReVa/self.SpCo0)*(self.P/P_ReVa)   # rmtUtility.py:405-421
        InGaVe_DiLeVa = InGaVe/InGaVe0
        SuGaVe = InGaVe*BeVoFr
        SuGaVe_DiLeVa = SuGaVe/SuGaVe0
        MiMoWe = np.dot(MoFri, np.array(self.MoWei))*1e-3
        GaDeEOS = P_ReVa/((R_CONST/MiMoWe)*T_ReVa)             # rmtThermo.py:353-369
        GaDe_DiLeVa = GaDeEOS/self.GaDe0
        ergA = 150*self.GaMiVi*SuGaVe/(PaDi**2)                # :3214-3220
        ergB = ((1-BeVoFr)**2)/(BeVoFr**3)
        ergC = 1.75*GaDeEOS*(SuGaVe**2)/PaDi
        ergD = (1-BeVoFr)/(BeVoFr**3)
        RHS_ergun = -1*(ergA*ergB + ergC*ergD)/(self.Pf/self.zf)
        Ri = np.array(reaction_rate_exe((T_ReVa, P_ReVa, MoFri, CoSpi_ReVa), self.varis, self.rates))
        ri = component_formation_rate(nc, self.compList, self.reactionStochCoeff, Ri)
        CpMeanList = cp_mean_list(self.compList, T_ReVa)
        GaCpMeanMix = np.dot(MoFri, CpMeanList)
        GaCpMeanMixEff_DiLeVa = (GaCpMeanMix/self.GaCpMeanMix0)*BeVoFr
        HeReT = np.array(np.array(enthalpy_change_of_reaction(self.reactionListSorted, T_ReVa)) + self.StHeRe25)
        OvHeReT = np.dot(Ri, HeReT)
        Qm = self.heat_exchange(T_ReVa)
        dxdt = np.zeros(self.varNo)
        constC1 = 1/SuGaVe_DiLeVa
        constT1 = 1/(GaDe_DiLeVa*GaCpMeanMixEff_DiLeVa*InGaVe_DiLeVa)
        for i in range(nc):
            dxdt[i] = constC1*(ri[i]/self.GaMaCoTe0[i])
        dxdt[nc] = RHS_ergun
        if not self.iso:
            dxdt[nc+1] = constT1*((-OvHeReT + Qm)/self.GaHeCoTe0)
        return dxdt.tolist()

    def solve(self, method=None, rtol=None, atol=None, t_eval=None):
        ivp = self.modelInput['solver-config']['ivp']
        method = method or ("LSODA" if ivp == 'default' else ivp)      # :2918
        kw = {}
        if rtol is not None:
            kw["rtol"] = rtol
        if atol is not None:
            kw["atol"] = atol
        return solve_ivp(lambda t, y: self.rhs(t, y), np.array([0, 1]), self.IV, method=method,
                         t_eval=self.times if t_eval is None else t_eval, **kw)

    def pack(self, sol):
        """Post-processing of runN1 :2949-3007 incl. sortResult4
        (solvers/solResultAnalysis.py:191-249)."""
        nc = self.compNo
        dataYs = sol.y
        ncol = dataYs.shape[1]
        cons = dataYs[0:nc, :]
        Pd = dataYs[nc, :]
        Td = dataYs[nc+1, :] if not self.iso else np.repeat(0, ncol).reshape(ncol)
        C = cons*np.max(self.Cif)
        Pr = (Pd*self.Pf).reshape(1, ncol)
        Tr = (Td*self.Tf + self.Tf).reshape(1, ncol)
        y = C/np.sum(C, axis=0)
        allv = np.concatenate((y, Pr, Tr), axis=0) if not self.iso else np.concatenate((y, Pr), axis=0)
        labelList = self.compList.copy() + ["Pressure"] + ([] if self.iso else ["Temperature"])
        return [{
            "modelId": self.modelId, "processType": self.processType, "successStatus": sol.success,
            "computation-time": 0.0, "dataShape": np.array(sol.t).shape, "labelList": labelList,
            "indexList": [nc, nc, nc + 1], "dataTime": [], "dataXs": sol.t,
            "dataYCons1": cons, "dataYCons2": C, "dataYTemp1": Td, "dataYTemp2": Tr, "dataYs": allv,
        }]


class N2Oracle(HomoReactorSetup):
    """Dynamic method-of-lines model: runN2 + modelEquationN2
    (docs/pbHomoReactor.py:3319-4134)."""

    def __init__(self, modelInput):
        super().__init__(modelInput)
        BeVoFr = self.BeVoFr
        self.InGaVe0 = self.VoFlRa0/(self.CrSeAr*BeVoFr)           # :3391
        self.SuGaVe0 = self.InGaVe0*BeVoFr
        self.scales(self.SuGaVe0)
        self.opT = modelInput['operating-conditions']['period']
        self.zNo = zNo = solverSetting['N2']['zNo']
        self.dataXs = np.linspace(0, 1, zNo)
        self.dz = 1/(zNo-1)
        nc = self.compNo
        self.varNo = nc if self.iso else nc + 1
        IV2D = np.zeros((self.varNo, zNo))
        for i in range(nc):
            IV2D[i, :] = self.SpCoi0[i]/np.max(self.SpCoi0)
        self.IV = IV2D.flatten()
        self.tNo = solverSetting['N2']['tNo']
        self.timesNo = solverSetting['N2']['timesNo']

    def rhs(self, t, y):
        """modelEquationN2, :3706-4134."""
        nc, zNo, dz = self.compNo, self.zNo, self.dz
        BeVoFr, PaDi = self.BeVoFr, self.PaDi
        InGaVe0 = self.VoFlRa0/(self.CrSeAr*BeVoFr)
        SuGaVe0 = InGaVe0*BeVoFr
        vf, zf, Tf = self.vf, self.zf, self.Tf
        Cmax = np.max(self.SpCoi0)
        yLoop = np.reshape(y, (self.varNo, zNo))
        SpCoi_z = yLoop[0:nc, :]
        T_z = yLoop[nc, :] if not self.iso else np.repeat(0, zNo)
        P_z = np.zeros(zNo + 1); P_z[0] = self.P
        v_z = np.zeros(zNo + 1); v_z[0] = SuGaVe0
        dxdtMat = np.zeros((self.varNo, zNo))
        MoWei = np.array(self.MoWei)
        CoSpi = np.zeros(nc)
        for z in range(zNo):
            for i in range(nc):
                CoSpi[i] = max(SpCoi_z[i][z], EPS_CONST)          # :3897-3904
            CoSpi_ReVa = CoSpi*Cmax
            T = T_z[z]
            T_ReVa = T*Tf + Tf
            P = P_z[z]
            v = v_z[z]
            v_DiLeVa = v/vf
            MoFri = CoSpi_ReVa/np.sum(CoSpi_ReVa)
            SuGaVe = v
            InGaVe = SuGaVe/BeVoFr
            InGaVe_DiLeVa = InGaVe/InGaVe0
            MiMoWe = np.dot(MoFri, MoWei)*1e-3
            GaDeEOS = P/((R_CONST/MiMoWe)*T_ReVa)
            GaDe_DiLeVa = GaDeEOS/self.GaDe0
            ergA = 150*self.GaMiVi*SuGaVe/(PaDi**2)               # :3970-3979
            ergB = ((1-BeVoFr)**2)/(BeVoFr**3)
            ergC = 1.75*GaDeEOS*(SuGaVe**2)/PaDi
            ergD = (1-BeVoFr)/(BeVoFr**3)
            dxdt_P = -1*(ergA*ergB + ergC*ergD)
            P_z[z+1] = dxdt_P*dz + P_z[z]
            Ri = np.array(reaction_rate_exe((T_ReVa, P_z[z], MoFri, CoSpi_ReVa), self.varis, self.rates))
            ri = component_formation_rate(nc, self.compList, self.reactionStochCoeff, Ri)
            CpMeanList = cp_mean_list(self.compList, T_ReVa)
            GaCpMeanMix = np.dot(MoFri, CpMeanList)
            GaCpMeanMix_DiLeVa = GaCpMeanMix/self.GaCpMeanMix0
            GaCpMeanMixEff_DiLeVa = GaCpMeanMix_DiLeVa*BeVoFr
            HeReT = np.array(np.array(enthalpy_change_of_reaction(self.reactionListSorted, T_ReVa)) + self.StHeRe25)
            OvHeReT = np.dot(Ri, HeReT)
            Qm = self.heat_exchange(T_ReVa)
            v_z[z+1] = v_z[z]                                     # :4066
            const_F1 = 1/(BeVoFr*(zf/vf))
            const_T2 = 1/(GaDe_DiLeVa*GaCpMeanMix_DiLeVa*BeVoFr*(zf/vf))
            for i in range(nc):
                Ci_c = SpCoi_z[i][z]
                Ci_b = self.SpCoi0[i]/Cmax if z == 0 else max(SpCoi_z[i][z - 1], EPS_CONST)
                dCdz = (Ci_c - Ci_b)/dz
                dxdtMat[i][z] = const_F1*(-v_DiLeVa*dCdz + (ri[i]/self.GaMaCoTe0[i]))
            if not self.iso:
                T_c = T_z[z]
                T_b = (self.T - Tf)/Tf if z == 0 else T_z[z - 1]
                dTdz = (T_c - T_b)/dz
                conv = -1*InGaVe_DiLeVa*GaDe_DiLeVa*GaCpMeanMixEff_DiLeVa*dTdz
                hform = (1/self.GaHeCoTe0)*(-OvHeReT)
                hexch = (1/self.GaHeCoTe0)*Qm
                dxdtMat[nc][z] = const_T2*(conv + hform + hexch)
        return dxdtMat.flatten().tolist()

    def solve(self, method=None, rtol=None, atol=None, **ivp_kw):
        """Slab loop of runN2 :3589-3685: tNo restarted solve_ivp calls, only
        the last column of each is kept.  `ivp_kw` goes to solve_ivp unchanged (fixture generation passes
        `jac_sparsity` to the implicit methods so that a 200-node run finishes in minutes; the sparsity only
        shapes the finite-difference Jacobian of the Newton iteration, not the converged solution)."""
        ivp = self.modelInput['solver-config']['ivp']
        method = method or ("LSODA" if ivp == 'default' else ivp)
        kw = dict(ivp_kw)
        if rtol is not None:
            kw["rtol"] = rtol
        if atol is not None:
            kw["atol"] = atol
        opTSpan = np.linspace(0, self.opT, self.tNo + 1)
        IV = self.IV
        nc, zNo = self.compNo, self.zNo
        dataPack, nfev = [], 0
        for i in range(self.tNo):
            t = np.array([opTSpan[i], opTSpan[i+1]])
            times = np.linspace(t[0], t[1], self.timesNo)
            sol = solve_ivp(lambda tt, yy: self.rhs(tt, yy), t, IV, method=method, t_eval=times, **kw)
            if sol.success is False:
                raise RuntimeError("ODE Error")
            nfev += sol.nfev
            last = sol.y[:, -1]
            R = np.reshape(last, (self.varNo, zNo))
            cons = R[:-1]
            Td = R[-1] if not self.iso else np.repeat(0, zNo).reshape(zNo)
            C = (R[:-1] if not self.iso else R[:])[:nc]*np.max(self.Cif)      # sortResult5, solResultAnalysis.py:252-301
            Tr = (Td*self.Tf + self.Tf).reshape(1, zNo)
            yv = C/np.sum(C, axis=0)
            dataPack.append({
                "modelId": self.modelId, "processType": self.processType, "successStatus": sol.success,
                "dataShape": np.array(sol.t[-1]).shape,
                "labelList": self.compList.copy() + ["Temperature"],
                "indexList": [nc, nc + 1, nc], "dataTime": sol.t[-1], "dataXs": self.dataXs,
                "dataYCons1": cons, "dataYCons2": C, "dataYTemp1": Td, "dataYTemp2": Tr,
                "dataYs": np.concatenate((yv, Tr), axis=0), "solY": last,
            })
            IV = sol.y[:, -1]
        self.nfev = nfev
        return {"computation-time": 0.0, "dataPack": dataPack}


class M7Oracle:
    """Model M7 = PackedBedReactorClass.runM3 + modelEquationM3 (docs/pbReactor.py:1170-1575): the
    DIMENSIONAL steady-state twin of N1, unknowns [C_i [mol/m^3]..., T [K], P [Pa]] over z in [0, ReLe]."""

    def __init__(self, modelInput):
        mi = self.modelInput = modelInput
        self.P, self.T = mi['operating-conditions']['pressure'], mi['operating-conditions']['temperature']
        self.reactionList = list(mi['reactions'].values())
        self.reactionListSorted, self.reactionStochCoeff = parse_reactions(mi['reactions'])
        self.varis, self.rates = mi['reaction-rates']['VARS'], mi['reaction-rates']['RATES']
        self.compList = list(mi['feed']['components']['shell'])
        for c in self.compList:
            if c not in componentSymbolList:
                raise Exception("Component database is not up to date!")
        self.compNo = nc = len(self.compList)
        rs = mi['reactor']
        self.ReLe, self.PaDi, self.BeVoFr = rs['ReLe'], rs['PaDi'], rs['BeVoFr']
        self.CrSeAr = PI_CONST*(rs['ReInDi'] ** 2)/4                      # :1218
        self.VoFlRa0 = mi['feed']['volumetric-flowrate']
        self.SpCoi0 = 1*np.array(mi['feed']['concentration'], dtype=float)
        self.SpCo0 = np.sum(self.SpCoi0)
        self.MoWei = [_DB[s][0] for s in self.compList]
        self.ExHe = mi['external-heat']                                   # EfHeTrAr used as given (:1508)
        self.GaMiVi = mi['feed']['mixture-viscosity']                     # :1235 (input, not Wilke)
        self.StHeRe25 = np.array([standard_enthalpy_of_reaction(r) for r in self.reactionList])
        self.IV = np.concatenate([self.SpCoi0, [self.T, self.P]])         # :1240-1246
        self.times = np.linspace(0, self.ReLe, solverSetting['M9']['zNo'])   # :1283-1287

    def rhs(self, t, y):
        """modelEquationM3, :1371-1575."""
        nc, BeVoFr, PaDi = self.compNo, self.BeVoFr, self.PaDi
        InGaVe0 = self.VoFlRa0/(self.CrSeAr*BeVoFr)
        y = np.asarray(y, dtype=float)
        CoSpi, T, P = y[0:nc], y[nc], y[nc+1]
        CoSp = np.sum(CoSpi)
        MoFri = CoSpi/np.sum(CoSpi)
        InGaVe = InGaVe0*(CoSp/self.SpCo0)*(self.P/P)
        SuGaVe = InGaVe*BeVoFr
        MoFl = (CoSp*SuGaVe*self.CrSeAr)/self.CrSeAr
        MiMoWe = np.dot(MoFri, np.array(self.MoWei))*1e-3
        GaDe = MiMoWe*CoSp                                              # calDensityIG, not the EOS one (:1471)
        ergA = 150*self.GaMiVi*SuGaVe/(PaDi**2)
        ergB = ((1-BeVoFr)**2)/(BeVoFr**3)
        ergC = 1.75*GaDe*(SuGaVe**2)/PaDi
        ergD = (1-BeVoFr)/(BeVoFr**3)
        RHS_ergun = -1*(ergA*ergB + ergC*ergD)
        Ri = np.array(reaction_rate_exe((T, P, MoFri, CoSpi), self.varis, self.rates))
        ri = component_formation_rate(nc, self.compList, self.reactionStochCoeff, Ri)
        CpMeanMixture = np.dot(MoFri, cp_mean_list(self.compList, T))
        HeReT = np.array(np.array(enthalpy_change_of_reaction(self.reactionListSorted, T)) + self.StHeRe25)
        OvHeReT = np.dot(Ri, HeReT)
        Qm = (self.ExHe['OvHeTrCo']*self.ExHe['EfHeTrAr'])*(self.ExHe['MeTe'] - T)     # no adiabatic switch here
        const_C1 = 1/SuGaVe
        const_T1 = 1/(MoFl*CpMeanMixture)
        return [const_C1*ri[i] for i in range(nc)] + [const_T1*(-OvHeReT + Qm), RHS_ergun]

    def solve(self, method=None, rtol=None, atol=None):
        ivp = self.modelInput['solver-config']['ivp']
        method = method or ("LSODA" if ivp == 'default' else ivp)
        kw = {}
        if rtol is not None:
            kw["rtol"] = rtol
        if atol is not None:
            kw["atol"] = atol
        return solve_ivp(lambda t, y: self.rhs(t, y), np.array([0, self.ReLe]), self.IV, method=method,
                         t_eval=self.times, **kw)

    def pack(self, sol):
        """:1301-1368 — mole fractions + temperature rows, plot-helper lists (library/plot.py:85-115)."""
        nc = self.compNo
        C = sol.y[0:nc, :]
        dataYs = np.concatenate((C/np.sum(C, axis=0), [sol.y[nc, :]]), axis=0)
        labelList = self.compList.copy() + ["Temperature", "Pressure"]
        XYList = [[sol.t, item] for item in dataYs]
        dataList = [{"x": XYList[i][0], "y": XYList[i][1], "leg": labelList[i]} for i in range(len(XYList))]
        return {"dataYs": dataYs, "XYList": XYList, "dataList": dataList, "nfev": sol.nfev, "solY": sol.y}


class M9Oracle:
    """Model M9 = PackedBedReactorClass.runM5 + modelEquationM5 (docs/pbReactor.py:1997-2660): the DIMENSIONAL
    dynamic twin of N2.  Unknowns [C_i..., T] at zNo nodes (variable-major); pressure AND superficial velocity are
    marched node by node inside the RHS (:2546-2612).  Grid/time settings come from solverSetting['S2']."""

    def __init__(self, modelInput):
        mi = self.modelInput = modelInput
        self.P, self.T = mi['operating-conditions']['pressure'], mi['operating-conditions']['temperature']
        self.opT = mi['operating-conditions']['period']
        self.reactionList = list(mi['reactions'].values())
        self.reactionListSorted, self.reactionStochCoeff = parse_reactions(mi['reactions'])
        self.varis, self.rates = mi['reaction-rates']['VARS'], mi['reaction-rates']['RATES']
        self.compList = list(mi['feed']['components']['shell'])
        for c in self.compList:
            if c not in componentSymbolList:
                raise Exception("Component database is not up to date!")
        self.compNo = nc = len(self.compList)
        rs = self.ReSpec = mi['reactor']
        self.ReLe, self.PaDi, self.BeVoFr = rs['ReLe'], rs['PaDi'], rs['BeVoFr']
        self.CrSeAr = PI_CONST*(rs['ReInDi'] ** 2)/4                      # :2043
        self.VoFlRa0 = mi['feed']['volumetric-flowrate']
        self.SpCoi0 = np.array(mi['feed']['concentration'], dtype=float)
        self.SpCo0 = np.sum(self.SpCoi0)
        self.MoWei = [_DB[s][0] for s in self.compList]
        self.ExHe = mi['external-heat']
        self.GaMiVi = mi['feed']['mixture-viscosity']                     # :2069
        s2 = solverSetting['S2']
        self.zNo, self.tNo, self.timesNo = s2['zNo'], s2['tNo'], s2['timesNo']
        self.dataXs = np.linspace(0, self.ReLe, self.zNo)
        self.dz = self.ReLe/(self.zNo - 1)                                # :2076
        self.varNo = nc + 1
        IV2D = np.zeros((self.varNo, self.zNo))
        for i in range(nc):
            IV2D[i, :] = self.SpCoi0[i]
        IV2D[nc, :] = self.T
        self.IV = IV2D.flatten()                                          # :2090-2103
        self.StHeRe25 = np.array([standard_enthalpy_of_reaction(r) for r in self.reactionList])

    def rhs(self, t, y):
        """modelEquationM5, :2296-2660."""
        nc, zNo, dz = self.compNo, self.zNo, self.dz
        BeVoFr, PaDi = self.BeVoFr, self.PaDi
        CaDe, CaSpHeCa = self.ReSpec['CaDe'], self.ReSpec['CaSpHeCa']
        InGaVe0 = self.VoFlRa0/(self.CrSeAr*BeVoFr)
        SuGaVe0 = InGaVe0*BeVoFr                                          # :2421 (the feed's superficial-velocity is not used)
        P_z = np.zeros(zNo + 1); P_z[0] = self.P
        v_z = np.zeros(zNo + 1); v_z[0] = SuGaVe0
        yLoop = np.reshape(y, (self.varNo, zNo))
        SpCoi_z = yLoop[0:nc, :]
        T_z = yLoop[nc, :]
        dxdtMat = np.zeros((self.varNo, zNo))
        MoWei = np.array(self.MoWei)
        CoSpi = np.zeros(nc)
        for z in range(zNo):
            for i in range(nc):
                CoSpi[i] = max(SpCoi_z[i][z], EPS_CONST)                  # :2487-2491
            CoSp = np.sum(CoSpi)
            T, P, v = T_z[z], P_z[z], v_z[z]
            MoFri = CoSpi/np.sum(CoSpi)
            SuGaVe = v
            MoFlRa = CoSp*SuGaVe*self.CrSeAr
            MoFl = MoFlRa/self.CrSeAr
            MiMoWe = np.dot(MoFri, MoWei)*1e-3
            GaDe = MiMoWe*CoSp                                            # calDensityIG (:2528)
            ergA = 150*self.GaMiVi*SuGaVe/(PaDi**2)
            ergB = ((1-BeVoFr)**2)/(BeVoFr**3)
            ergC = 1.75*GaDe*(SuGaVe**2)/PaDi
            ergD = (1-BeVoFr)/(BeVoFr**3)
            dxdt_P = -1*(ergA*ergB + ergC*ergD)
            P_z[z+1] = dxdt_P*dz + P_z[z]                                 # :2546
            Ri = np.array(reaction_rate_exe((T_z[z], P_z[z], MoFri, CoSpi), self.varis, self.rates))
            ri = component_formation_rate(nc, self.compList, self.reactionStochCoeff, Ri)
            OvR = np.sum(ri)
            CpMeanMixture = np.dot(MoFri, cp_mean_list(self.compList, T))
            HeReT = np.array(np.array(enthalpy_change_of_reaction(self.reactionListSorted, T)) + self.StHeRe25)
            OvHeReT = np.dot(Ri, HeReT)
            Tm, U, a = self.ExHe['MeTe'], self.ExHe['OvHeTrCo'], self.ExHe['EfHeTrAr']
            Qm = 0 if Tm == 0 else U*a*(Tm - T)                           # rmtUtility.py:424-452 with unit 'kJ/m^3.s'
            if Qm != 0:
                Qm = Qm*1e-3
            T_b = self.T if z == 0 else T_z[z - 1]
            dxdt_v_T = (T_z[z] - T_b)/dz
            dxdt_v = (1/(CoSp*1000))*((-SuGaVe/R_CONST)*((1/T)*dxdt_P - (P/T**2)*dxdt_v_T) + OvR*1000)    # :2606-2608
            v_z[z+1] = dxdt_v*dz + v_z[z]
            const_F1 = 1/BeVoFr
            const_T1 = MoFl*CpMeanMixture
            const_T2 = 1/(CoSp*CpMeanMixture*BeVoFr + (1-BeVoFr)*CaDe*CaSpHeCa)
            for i in range(nc):
                Ci_c = SpCoi_z[i][z]
                Ci_b = self.SpCoi0[i] if z == 0 else max(SpCoi_z[i][z - 1], EPS_CONST)
                dCdz = (Ci_c - Ci_b)/dz
                dxdtMat[i][z] = const_F1*(-v_z[z]*dCdz - Ci_c*dxdt_v + ri[i])
            dTdz = (T_z[z] - T_b)/dz
            dxdtMat[nc][z] = const_T2*(-const_T1*dTdz + (-OvHeReT + Qm))
        return dxdtMat.flatten().tolist()

    def solve(self, method=None, rtol=None, atol=None):
        """Slab loop of runM5 :2161-2216 and the plot lists it returns (:2262-2294): the (x, y) series of the LAST
        variable (temperature) at the end of every slab."""
        ivp = self.modelInput['solver-config']['ivp']
        method = method or ("LSODA" if ivp == 'default' else ivp)
        kw = {}
        if rtol is not None:
            kw["rtol"] = rtol
        if atol is not None:
            kw["atol"] = atol
        opTSpan = np.linspace(0, self.opT, self.tNo + 1)
        IV = self.IV
        nc, zNo = self.compNo, self.zNo
        dataPack, nfev = [], 0
        dataPacktime = np.zeros((self.varNo, self.tNo, zNo))
        for i in range(self.tNo):
            t = np.array([opTSpan[i], opTSpan[i+1]])
            sol = solve_ivp(lambda tt, yy: self.rhs(tt, yy), t, IV, method=method,
                            t_eval=np.linspace(t[0], t[1], self.timesNo), **kw)
            if sol.success is False:
                raise RuntimeError("ODE Error")
            nfev += sol.nfev
            C = np.reshape(sol.y[0:nc*zNo, -1], (nc, zNo))
            Tr = np.array([sol.y[nc*zNo:(nc + 1)*zNo, -1]])
            dataYs = np.concatenate((C/np.sum(C, axis=0), Tr), axis=0)
            dataPack.append({"successStatus": sol.success, "dataTime": sol.t[-1], "dataYCons": C, "dataYTemp": Tr,
                             "dataYs": dataYs, "solY": sol.y[:, -1]})
            for m in range(self.varNo):
                dataPacktime[m][i, :] = dataYs[m, :]
            IV = sol.y[:, -1]
        self.nfev = nfev
        labelList = self.compList.copy() + ["Temperature"]
        XYList = [[self.dataXs, item] for item in dataPacktime[self.varNo - 1]]
        names = [labelList[self.varNo - 1] + " at t=" + str(opTSpan[t + 1]) for t in range(self.tNo)]
        dataList = [{"x": XYList[i][0], "y": XYList[i][1], "leg": names[i]} for i in range(len(XYList))]
        return {"XYList": XYList, "dataList": dataList, "dataPack": dataPack}


def rmtExe(modelInput, method=None, rtol=None, atol=None):
    """rmt.py:21-80 restricted to the N1/N2 branch of rmtCore.py:63-127."""
    if modelInput['model'] == "N1":
        o = N1Oracle(modelInput)
        sol = o.solve(method=method, rtol=rtol, atol=atol)
        if sol.success is False:
            raise RuntimeError("ODE Error")
        res = o.pack(sol)
        res[0]["nfev"] = sol.nfev
    elif modelInput['model'] == "N2":
        o = N2Oracle(modelInput)
        res = o.solve(method=method, rtol=rtol, atol=atol)
    elif modelInput['model'] == "M9":
        res = M9Oracle(modelInput).solve(method=method, rtol=rtol, atol=atol)
    elif modelInput['model'] == "M7":
        o = M7Oracle(modelInput)
        sol = o.solve(method=method, rtol=rtol, atol=atol)
        if sol.success is False:
            raise RuntimeError("ODE Error")
        res = o.pack(sol)
    else:
        raise NotImplementedError(modelInput['model'])
    return {"resModel": res, "comTime": 0.0}


def rmtCom():
    """rmt.py:83-92."""
    return ",".join(componentSymbolList)
