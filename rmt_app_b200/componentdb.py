"""Component property table of the reference, as data.

Facts restated from PyREMOT/data/componentData.py:11-22 (MW [g/mol]), :72-86
(heat of formation at 25 C [kJ/mol]), :119-405 (Cp(T) = a0 + a1*T + a2*T^2 +
a3*T^3 [J/mol/K]; the reference stores these as strings and `eval`s them per
call, rmtThermo.py:37) and PyREMOT/data/dataGasViscosity.py:8-141 (low-pressure
gas viscosity, eq.1: A*1e-6*T^B/(1 + C/T + D/T^2) [Pa.s], gasTransPor.py:137-154;
DME uses the closed form of dataGasViscosity.py:133, "eq.2").

Order matters: `rmtCom()` (rmt.py:83-92) joins the symbols in table order.
"""
from collections import OrderedDict

Tref = 273.15 + 25.00       # PyREMOT/core/constants.py:17-23
R_CONST = 8.314472          # PyREMOT/core/constants.py:8
PI_CONST = 3.141592653589793


class Component:
    __slots__ = ("symbol", "MW", "dHf25", "cp", "cp_terms", "visc_eq", "visc")

    def __init__(self, symbol, MW, dHf25, cp, visc_eq, visc):
        self.symbol, self.MW, self.dHf25 = symbol, float(MW), float(dHf25)
        self.cp = tuple(float(c) for c in cp) + (0.0,)*(4 - len(cp))
        self.cp_terms = len(cp)
        self.visc_eq, self.visc = visc_eq, tuple(float(v) for v in visc)

    def cp_at(self, T):
        """Same left-to-right evaluation as the reference's eval'd string."""
        a0, a1, a2, a3 = self.cp
        v = a0 + a1*T + a2*(T**2)
        if self.cp_terms == 4:
            v = v + a3*(T**3)
        return v


def _c(sym, MW, dHf, cp, visc, eq=1):
    return sym, Component(sym, MW, dHf, cp, eq, visc)


# eq.2 (DME): mu = p0 * T^p1 / (1 + p2/T)
COMPONENTS = OrderedDict([
    _c("CO2", 44.01, -393.51, (22.243, 5.98E-02, -3.50E-05, 7.46E-09), (4.719875, 0.373279, 512.686300, -6119.961)),
    _c("H2", 2.0, 0.0, (26.879, 4.35E-03, -3.30E-07), (0.169104, 0.692485, -7.634394, 467.120)),
    _c("CH3OH", 32.04, -200.7, (19.038, 9.15E-02, -1.22E-05, -8.03E-09), (0.477915, 0.641076, 284.838034, -3230.713)),
    _c("H2O", 18.01, -241.820, (29.163, 1.45E-02, -2.02E-06), (0.501246, 0.709247, 869.465599, -90063.891)),
    _c("CO", 28.01, -110.53, (27.113, 6.55E-03, -1.00E-06), (0.734306, 0.588574, 52.318660, 1018.822)),
    _c("DME", 46.07, -184.1, (19.8, 0.17, -5.66e-5), (2.68e-7, 0.3975, 534.0, 0.0), eq=2),
    _c("N2", 28, 0, (28.883, -1.57E-03, 8.08E-06, -2.87E-09), (0.847662, 0.574033, 75.437536, 56.771)),
    _c("CH4", 16.04, -74.90, (19.875, 5.021E-02, 1.268E-05, -11.004E-09), (1.119178, 0.493234, 214.627200, -3952.087)),
    _c("C2H4", 28.05, 52.32, (3.950, 15.628E-02, -8.339E-05, 17.657E-09), (1.503552, 0.456140, 288.342422, 73.362)),
    _c("C3H6", 42.08, 20.4, (3.151, 23.812E-02, -12.176E-05, 24.603E-09), (0.876767, 0.520871, 293.618650, -182.857)),
    _c("C3H8", 44.1, -103.9, (-4.042, 30.456E-02, -15.711E-05, 31.716E-09), (0.173966, 0.734798, 143.207060, -7147.859)),
    _c("C4H10", 58.12, -126.2, (-7.908, 41.573E-02, -22.992E-05, 49.875E-09), (0.075828, 0.837082, 67618677, -2141.762)),
])

componentSymbolList = tuple(COMPONENTS.keys())
