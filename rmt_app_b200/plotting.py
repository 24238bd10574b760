"""Plot hooks behind `solver-config.display-result == "True"` — the thin equivalent of
`plotResultsSteadyState` / `plotResultsDynamic` (PyREMOT/solvers/solResultAnalysis.py:307-459).
Presentation only; needs matplotlib (absent from the B200 image, in which case a notice is printed)."""


def _plt():
    try:
        import matplotlib.pyplot as plt
        return plt
    except Exception:
        print("display-result: matplotlib is not installed; skipping the figures")
        return None


def plotResultsSteadyState(dataPack):
    plt = _plt()
    if plt is None:
        return
    dp = dataPack[0]
    xs, ys, labels = dp["dataXs"], dp["dataYs"], dp["labelList"]
    nc = dp["indexList"][0]
    fig, ax = plt.subplots(1, 3, figsize=(14, 4))
    for i in range(nc):
        ax[0].plot(xs, ys[i], label=labels[i])
    ax[0].set_xlabel("dimensionless length"); ax[0].set_ylabel("mole fraction"); ax[0].legend()
    ax[1].plot(xs, ys[nc]); ax[1].set_xlabel("dimensionless length"); ax[1].set_ylabel("pressure [Pa]")
    if len(ys) > nc + 1:
        ax[2].plot(xs, ys[nc + 1]); ax[2].set_xlabel("dimensionless length"); ax[2].set_ylabel("temperature [K]")
    fig.suptitle("Steady-State Modeling %s, computation-time %s" % (dp["modelId"], dp["computation-time"]))
    plt.show()


def plotResultsDynamic(resPack, tNo):
    plt = _plt()
    if plt is None:
        return
    packs = resPack["dataPack"]
    fig, ax = plt.subplots(1, 2, figsize=(11, 4))
    for dp in packs:                      # the reference draws two random slabs (:421-422); all are drawn here
        nc = dp["indexList"][0]
        for i in range(nc):
            ax[0].plot(dp["dataXs"], dp["dataYs"][i])
        ax[1].plot(dp["dataXs"], dp["dataYs"][-1], label="t=%.3g" % dp["dataTime"])
    ax[0].set_xlabel("dimensionless length"); ax[0].set_ylabel("mole fraction")
    ax[1].set_xlabel("dimensionless length"); ax[1].set_ylabel("temperature [K]"); ax[1].legend()
    fig.suptitle("Dynamic Modeling %s, computation-time %s" % (packs[0]["modelId"], resPack["computation-time"]))
    plt.show()
