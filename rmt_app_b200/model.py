"""Model specification: what `rmtCoreClass.modExe` derives from a modelInput
before it dispatches to runN1/runN2 (PyREMOT/docs/rmtCore.py:63-183), restated
as a compile-time description for the CUDA code generator.

* component rows in `compList` order          (rmtCore.py:129-164)
* reaction strings -> stoichiometry            (rmtUtility.py:172-249)
* standard heats of reaction at 25 C           (rmtThermo.py:129-198)
* traced kinetics                              (rmtReaction.py:11-61)
"""
import hashlib
import re

import numpy as np

from .componentdb import COMPONENTS, Tref, componentSymbolList
from .kinetics import trace_kinetics

# same token grammar as the reference: optional number, then a symbol
_TERM = re.compile(r"([0-9.]*)([a-zA-Z0-9.]+)")

# per-instance primary inputs of the device setup kernel, in row order.
# "concentration" expands to nc rows; kinetic parameter slots follow.
SCALAR_INPUTS = ("temperature", "pressure", "volumetric-flowrate", "ReInDi", "ReLe", "PaDi", "BeVoFr",
                 "OvHeTrCo", "MeTe", "mixture-viscosity", "EfHeTrAr", "CaDe", "CaSpHeCa")


def parse_reaction(expr):
    """"A + 3B <=> C + D" -> ([(sym, -nu)...], [(sym, +nu)...]) exactly as
    rmtUtility.buildReactionCoefficient (:172-220) tokenises it."""
    sides = expr.replace("<", "").replace(">", "").replace(" ", "").split("=")
    if len(sides) != 2:
        raise ValueError("reaction %r must contain exactly one '='" % expr)
    reac = [(s, -1*float(c) if len(c) else -1.0) for c, s in _TERM.findall(sides[0])]
    prod = [(s, float(c) if len(c) else 1.0) for c, s in _TERM.findall(sides[1])]
    return reac, prod


class ModelSpec:
    """Everything the code generator needs; hashable into a module key."""

    def __init__(self, modelInput):
        mi = modelInput
        self.model = mi["model"]
        if self.model not in ("N1", "N2", "M7", "M9"):
            raise NotImplementedError(
                "rmt_app_b200 implements the pseudo-homogeneous packed-bed models N1, N2, M7 and M9 only "
                "(got model=%r)" % (self.model,))
        self.compList = list(mi["feed"]["components"]["shell"])
        for c in self.compList:
            if c not in componentSymbolList:                          # rmt.py:55-57
                raise Exception("Component database is not up to date!")
        self.nc = len(self.compList)
        # modelSetting.py:21-23; M7 (pbReactor.runM3) has no process-type switch: always with the energy balance
        self.iso = self.model not in ("M7", "M9") and mi["operating-conditions"]["process-type"] == "iso-thermal"
        self.reactions = list(mi["reactions"].values())
        self.nr = len(self.reactions)
        self.components = [COMPONENTS[c] for c in self.compList]

        nu = np.zeros((self.nr, self.nc))
        dH25 = np.zeros(self.nr)
        dcp = np.zeros((self.nr, 4))        # dCp_j(T) = sum nu*CpMean_i(T) as a cubic in T
        for j, expr in enumerate(self.reactions):
            reac, prod = parse_reaction(expr)
            for sym, v in reac + prod:
                if sym in self.compList:                               # rmtReaction.py:88-90 (string match)
                    nu[j, self.compList.index(sym)] += v
                if sym not in COMPONENTS:
                    raise Exception("Component database is not up to date!")
            # rmtThermo.py:129-198: (sum_prod - sum_react)*1000, |nu| on each side
            hp = np.sum(np.array([COMPONENTS[s].dHf25*v for s, v in prod]))
            hr = np.sum(np.array([COMPONENTS[s].dHf25*(-v) for s, v in reac]))
            dH25[j] = (hp - hr)*1000.00
            # rmtThermo.py:258-312 with CpMean = (Cp(Tref)+Cp(T))/2 folded into one cubic
            for sym, v in reac + prod:
                comp = COMPONENTS[sym]
                a = comp.cp
                dcp[j, 0] += v*0.5*(comp.cp_at(Tref) + a[0])
                dcp[j, 1] += v*0.5*a[1]
                dcp[j, 2] += v*0.5*a[2]
                dcp[j, 3] += v*0.5*a[3]
        self.nu, self.dH25, self.dcp = nu, dH25, dcp

        rr = mi["reaction-rates"]
        self.kin = trace_kinetics(rr["VARS"], rr["RATES"], self.nc)
        if self.kin.nr != self.nr:
            raise ValueError("RATES has %d entries for %d reactions (matched by position, rmtReaction.py:56-58)"
                             % (self.kin.nr, self.nr))
        self.kin.differentiate()
        self.nkp = len(self.kin.param_names)

    # -- input row bookkeeping --------------------------------------------------
    @property
    def n(self):
        """unknowns per axial point: N1 = nc + P (+ T); M7 = nc + T + P; N2 = nc (+ T)."""
        if self.model == "M7":
            return self.nc + 2
        if self.model == "N1":
            return self.nc + (1 if self.iso else 2)
        return self.nc + (0 if self.iso else 1)

    def input_names(self):
        names = ["temperature", "pressure"] + ["concentration[%d]" % i for i in range(self.nc)]
        names += list(SCALAR_INPUTS[2:])
        names += ["VARS:" + k for k in self.kin.param_names]
        return names

    @property
    def nin(self):
        return 2 + self.nc + len(SCALAR_INPUTS) - 2 + self.nkp

    def key(self, extra=""):
        h = hashlib.sha256()
        h.update(repr((self.model, self.compList, self.iso, self.reactions, self.kin.param_names)).encode())
        g = self.kin.g
        h.update(repr([(n.op, tuple(a.id for a in n.args), n.value, n.name) for n in g.topo(self.kin.rates)]).encode())
        h.update(extra.encode())
        return h.hexdigest()[:16]
