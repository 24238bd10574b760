"""Plot hooks behind `solver-config.display-result == "True"`: what `plotResultsSteadyState` /
`plotResultsDynamic` (PyREMOT/solvers/solResultAnalysis.py:307-459) put on screen through
`plotClass.plots2D` (PyREMOT/library/plot.py:36-82) — one figure per quantity group, each with the
same lines, legends, axis labels and title.  Presentation only; needs matplotlib (absent from the
B200 image, in which case a notice is printed and nothing is drawn).

Steady state (:307-371): three figures — the nc species rows of dataYs, the pressure row, the
temperature row (no temperature figure when iso-thermal) — titled
"Steady-State Modeling {modelId}, computation-time {t}", x label "Reactor Length (m)", y labels
"Concentration (mol/$m^3$)" / "Pressure (bar)" / "Temperature (K)" (the reference's labels, kept although the
rows hold mole fractions and Pa).

Dynamic (:373-459): the slabs to draw are the first, the last and two interior ones drawn without
replacement from NumPy's global generator (`selectRandomForList`, core/utilities.py:88-110 — seed with
`np.random.seed`); per slab two figures (species rows; temperature row), title as above + " at t={dataTime}".
"""
import numpy as np

X_LABEL = "Reactor Length (m)"
Y_LABELS_STEADY = ("Concentration (mol/$m^3$)", "Pressure (bar)", "Temperature (K)")
Y_LABELS_DYNAMIC = ("Concentration (mol/$m^3$)", "Temperature (K)")


def _plt():
    try:
        import matplotlib.pyplot as plt
        return plt
    except Exception:
        print("display-result: matplotlib is not installed; skipping the figures")
        return None


def _figure(plt, lines, y_label, title):
    """plots2D (library/plot.py:36-82): a list of {x, y, leg} or a single one."""
    for ln in (lines if isinstance(lines, list) else [lines]):
        plt.plot(ln["x"], ln["y"], label=ln.get("leg", "line"))
    if len(title) > 0:
        plt.title(title)
    plt.xlabel(X_LABEL)
    plt.ylabel(y_label)
    plt.legend()
    plt.show()


def _data_list(xs, ys, labels):
    """plots2DSetXYList + plots2DSetDataList (library/plot.py:85-115)."""
    return [{"x": xs, "y": row, "leg": labels[i]} for i, row in enumerate(ys)]


def select_slabs(tNo, no=2):
    """selectRandomForList(range(tNo), no): [first, `no` sorted interior indices drawn at random, last]."""
    idx = list(range(tNo))
    inner = np.sort(np.random.choice(idx[1:-1], no, replace=False))
    return [idx[0], *[int(i) for i in inner], idx[-1]]


def plotResultsSteadyState(dataPack):
    plt = _plt()
    dp = dataPack[0]
    if plt is None or dp.get("successStatus", True) is not True:
        return
    nc, i_p, i_t = dp["indexList"]
    title = "Steady-State Modeling %s, computation-time %s" % (dp["modelId"], dp["computation-time"])
    lines = _data_list(dp["dataXs"], dp["dataYs"], dp["labelList"])
    groups = [lines[0:nc], lines[i_p]]
    if dp.get("processType") != "iso-thermal":
        groups.append(lines[i_t])
    for f, g in enumerate(groups):
        _figure(plt, g, Y_LABELS_STEADY[f], title)


def plotResultsDynamic(resPack, tNo):
    plt = _plt()
    if plt is None:
        return
    packs = resPack["dataPack"]
    d0 = packs[0]
    nc, _, i_t = d0["indexList"]
    title = "Steady-State Modeling %s, computation-time %s" % (d0["modelId"], resPack["computation-time"])     # the reference's wording (:410)
    for i in select_slabs(tNo, 2):
        dp = packs[i]
        if dp.get("successStatus", True) is not True:
            continue
        lines = _data_list(dp["dataXs"], dp["dataYs"], d0["labelList"])
        groups = [lines[0:nc]]
        if d0.get("processType") != "iso-thermal":
            groups.append(lines[i_t])
        for f, g in enumerate(groups):
            _figure(plt, g, Y_LABELS_DYNAMIC[f], title + " at t=%s" % (dp["dataTime"],))
