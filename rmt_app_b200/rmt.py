"""Public API — same names, arguments and result layout as the reference's
`PyREMOT/rmt.py` for the N1 / N2 branch of `rmtCoreClass.modExe`
(PyREMOT/docs/rmtCore.py:124-127).

    rmtExe(modelInput) -> {"resModel": dataPack | resPack, "comTime": float}
    rmtCom()           -> "CO2,H2,CH3OH,..."

`rmtExeBatch` is the ensemble extension: the same modelInput plus per-instance
arrays for any operating / feed / reactor input or scalar VARS entry.
"""
from timeit import default_timer as timer

import numpy as np

from . import engine
from .componentdb import componentSymbolList
from .engine import solverSetting

STATUS_TEXT = {0: "ok", 1: "max_steps reached", 2: "step size underflow", 3: "non-finite state"}


def _component_list(FeCom):
    """rmtUtility.buildComponentList (docs/rmtUtility.py:313-340)."""
    comp = []
    for part in ("shell", "tube", "medium"):
        v = FeCom.get(part)
        if v:
            comp.extend(v)
    return list(dict.fromkeys(comp))


def _check_components(modelInput):
    for c in _component_list(modelInput['feed']['components']):
        if c not in componentSymbolList:
            raise Exception("Component database is not up to date!")      # rmt.py:55-57


def rmtExe(modelInput):
    """Drop-in for PyREMOT.rmtExe (rmt.py:21-80) on models "N1" and "N2"."""
    try:
        tic = timer()
        modelType = modelInput['model']
        _check_components(modelInput)
        if modelType == "N1":
            resModel = _runN1(modelInput)
        elif modelType == "N2":
            resModel = _runN2(modelInput)
        elif modelType == "M7":
            resModel = _runM7(modelInput)
        elif modelType == "M9":
            resModel = _runM9(modelInput)
        else:
            raise NotImplementedError(
                "model %r: this build accelerates the pseudo-homogeneous packed-bed models N1, N2, M7 and M9 only"
                % (modelType,))
        return {"resModel": resModel, "comTime": (timer() - tic)*1000}
    except Exception as e:
        print(e)
        raise


def rmtCom():
    """rmt.py:83-92."""
    return ",".join(componentSymbolList)


def _display(modelInput):
    return modelInput.get('solver-config', {}).get('display-result', "False") == "True"


def _runN1(modelInput):
    """runN1 (pbHomoReactor.py:2694-3015): one reactor, 101 output points."""
    start = timer()
    cm = engine.compile_model(modelInput)
    spec = cm.spec
    nc, n = spec.nc, spec.n
    times = np.linspace(0, 1, solverSetting['N1']['zNo'] + 1)                # :2858-2862
    res = engine.n1_solve_ensemble(cm, modelInput, None, 1, z_eval=times, out_mode=2)
    if int(res.status[0]) != 0:
        print('ODE Error')                                                    # :2944-2947
        raise RuntimeError("ODE Error: integrator status %d (%s)" % (res.status[0], STATUS_TEXT.get(int(res.status[0]))))
    o = res.out[:, :, 0].T                       # [rows][n_eval]
    raw, C, allv = o[:n], o[n:n + nc], o[n + nc:]
    ncol = times.size
    processType = modelInput['operating-conditions']['process-type']
    labelList = list(spec.compList) + ["Pressure"] + ([] if spec.iso else ["Temperature"])
    Td = raw[nc + 1, :] if not spec.iso else np.repeat(0, ncol).reshape(ncol)
    Tr = (allv[nc + 1, :] if not spec.iso else np.repeat(float(modelInput['operating-conditions']['temperature']), ncol)
          ).reshape(1, ncol)
    dataPack = [{
        "modelId": modelInput['model'],
        "processType": processType,
        "successStatus": True,
        "computation-time": np.round(timer() - start, 3),
        "dataShape": times.shape,
        "labelList": labelList,
        "indexList": [nc, nc, nc + 1],
        "dataTime": [],
        "dataXs": times,
        "dataYCons1": raw[0:nc, :],
        "dataYCons2": C,
        "dataYTemp1": Td,
        "dataYTemp2": Tr,
        "dataYs": allv,
        "solverStats": {k: int(v) for k, v in zip(("accepted", "rejected", "nfev", "njev"), res.stats[:, 0])},
    }]
    if _display(modelInput):
        from .plotting import plotResultsSteadyState
        plotResultsSteadyState(dataPack)
    return dataPack


def _runM7(modelInput):
    """runM3 (PyREMOT/docs/pbReactor.py:1170-1368), model id "M7": the dimensional twin of N1.
    Result: {"dataYs": (nc+1) x zNo [mole fractions..., T], "XYList", "dataList"} (:1301-1368); the
    number of output points comes from solverSetting['M9']['zNo'] (:1283).  The reference always draws
    a figure here; that only happens with solver-config.display-result == "True"."""
    cm = engine.compile_model(modelInput)
    spec = cm.spec
    nc = spec.nc
    times = np.linspace(0, modelInput['reactor']['ReLe'], solverSetting['M9']['zNo'])
    res = engine.n1_solve_ensemble(cm, modelInput, None, 1, z_eval=times, out_mode=1)
    if int(res.status[0]) != 0:
        raise RuntimeError("ODE Error: integrator status %d (%s)" % (res.status[0], STATUS_TEXT.get(int(res.status[0]))))
    rows = res.out[:, :, 0].T                         # y_i..., T, P
    dataYs = np.ascontiguousarray(rows[:nc + 1])
    labelList = list(spec.compList) + ["Temperature", "Pressure"]
    XYList = [[times, item] for item in dataYs]       # library/plot.py:85-115
    dataList = [{"x": XYList[i][0], "y": XYList[i][1], "leg": labelList[i]} for i in range(len(XYList))]
    out = {"dataYs": dataYs, "XYList": XYList, "dataList": dataList, "dataPressure": rows[nc + 1].copy()}
    if _display(modelInput):
        from .plotting import plotResultsSteadyState
        plotResultsSteadyState([{"dataXs": times, "dataYs": rows, "labelList": labelList[:nc] + ["Pressure", "Temperature"],
                                 "indexList": [nc, nc + 1, nc], "modelId": "M7", "computation-time": 0.0}])
    return out


def _runM9(modelInput):
    """runM5 (PyREMOT/docs/pbReactor.py:1997-2294), model id "M9": the dimensional twin of N2.  Grid and slabs come
    from solverSetting['S2'] (:2072, :2145).  The reference returns only the plot lists of the LAST variable it looped
    over — the temperature profile at the end of every slab (:2262-2294); the per-slab records it builds on the way
    (:2203-2209) are returned here as well under "dataPack"."""
    from .engine import n2_solve_ensemble
    zNo, tNo = solverSetting['S2']['zNo'], solverSetting['S2']['tNo']
    cm = engine.compile_model(modelInput, block=engine.n2_block(1))
    spec = cm.spec
    nc, n = spec.nc, spec.n
    opT = modelInput['operating-conditions']['period']
    res = n2_solve_ensemble(cm, modelInput, None, 1, zNo=zNo, tNo=tNo, period=opT, out_mode=2)
    if int(res.status[0]) != 0:
        raise RuntimeError("ODE Error: integrator status %d (%s)" % (res.status[0], STATUS_TEXT.get(int(res.status[0]))))
    opTSpan = np.linspace(0, opT, tNo + 1)
    dataXs = np.linspace(0, float(modelInput['reactor']['ReLe']), zNo)
    labelList = list(spec.compList) + ["Temperature"]
    dataPack = []
    for i in range(tNo):
        o = res.out[i, :, :, 0]                 # [rows][zNo]: raw state | C_i | (y_i, T)
        raw, allv = o[:n], o[n + nc:]
        dataPack.append({"successStatus": True, "dataTime": opTSpan[i + 1], "dataYCons": raw[:nc].copy(),
                         "dataYTemp": raw[nc:nc + 1].copy(), "dataYs": allv.copy()})
    Tt = np.array([d["dataYs"][nc] for d in dataPack])                      # dataPacktime[indexTemp], :2211-2212
    XYList = [[dataXs, item] for item in Tt]                                # library/plot.py:85-90
    names = [labelList[nc] + " at t=" + str(opTSpan[t + 1]) for t in range(tNo)]
    dataList = [{"x": XYList[i][0], "y": XYList[i][1], "leg": names[i]} for i in range(len(XYList))]
    out = {"XYList": XYList, "dataList": dataList, "dataPack": dataPack}
    if _display(modelInput):
        from .plotting import plotResultsDynamic
        plotResultsDynamic({"computation-time": 0.0,
                            "dataPack": [dict(d, dataXs=dataXs, labelList=labelList, indexList=[nc, nc + 1, nc],
                                              modelId="M9") for d in dataPack]}, tNo)
    return out


def _runN2(modelInput):
    from .engine import n2_solve_ensemble
    start = timer()
    zNo, tNo = solverSetting['N2']['zNo'], solverSetting['N2']['tNo']
    cm = engine.compile_model_n2(modelInput, 1, zNo)
    spec = cm.spec
    nc, n = spec.nc, spec.n
    opT = modelInput['operating-conditions']['period']
    res = n2_solve_ensemble(cm, modelInput, None, 1, zNo=zNo, tNo=tNo, period=opT, out_mode=2)
    if int(res.status[0]) != 0:
        raise RuntimeError("ODE Error: integrator status %d (%s)" % (res.status[0], STATUS_TEXT.get(int(res.status[0]))))
    opTSpan = np.linspace(0, opT, tNo + 1)
    dataXs = np.linspace(0, 1, zNo)
    processType = modelInput['operating-conditions']['process-type']
    dataPack = []
    for i in range(tNo):
        o = res.out[i, :, :, 0]                 # [rows][zNo]
        raw, C, allv = o[:n], o[n:n + nc], o[n + nc:]
        if spec.iso:
            # the reference still appends a temperature row: T-hat = 0 everywhere, i.e. the feed temperature (:3641-3661)
            allv = np.concatenate((allv, np.full((1, zNo), float(modelInput['operating-conditions']['temperature']))), axis=0)
        dataPack.append({
            "modelId": modelInput['model'],
            "processType": processType,
            "successStatus": True,
            "dataShape": np.array(opTSpan[i + 1]).shape,
            "labelList": list(spec.compList) + ["Temperature"],
            "indexList": [nc, nc + 1, nc],
            "dataTime": opTSpan[i + 1],
            "dataXs": dataXs,
            "dataYCons1": raw[:-1] if not spec.iso else raw[:-1],       # :3638 (kept, incl. the iso slicing quirk)
            "dataYCons2": C,
            "dataYTemp1": raw[-1] if not spec.iso else np.repeat(0, zNo).reshape(zNo),
            "dataYTemp2": allv[-1].reshape(1, zNo) if not spec.iso else np.repeat(
                float(modelInput['operating-conditions']['temperature']), zNo).reshape(1, zNo),
            "dataYs": allv,
        })
    resPack = {"computation-time": np.round(timer() - start, 3), "dataPack": dataPack}
    if _display(modelInput):
        from .plotting import plotResultsDynamic
        plotResultsDynamic(resPack, tNo)
    return resPack


def rmtExeBatch(modelInput, sweep=None, B=None, *, rtol=None, atol=None, profile=False, z_eval=None,
                objective_ref=None, dense=True, max_steps=100000, keep_on_device=False, workspace=None,
                return_stats=True):
    """Ensemble form of rmtExe for model "N1": B independent reactors that share
    `modelInput` except for the per-instance arrays in `sweep` (keys:
    "temperature", "pressure", "concentration" [B, nc], "volumetric-flowrate",
    "ReInDi", "ReLe", "PaDi", "BeVoFr", "OvHeTrCo", "MeTe", or any scalar VARS name).

    Returns {"dataYs": [B, n] outlet (or [B, n, n_eval] with profile/z_eval),
             "status": [B], "success": [B] bool, "stats": [4, B], "dataXs": z_eval,
             "objective": [B] or None, "comTime": ms}.
    Failed instances are flagged in `status` and never abort the ensemble.
    `workspace` (engine.Workspace) reuses pinned/device buffers across calls; the
    returned arrays are then views valid until the next call with that workspace.
    Sweep values may be NumPy arrays (staged through pinned memory), pinned torch CPU
    tensors (copied directly) or torch CUDA tensors (no host transfer at all)."""
    tic = timer()
    _check_components(modelInput)
    if modelInput['model'] not in ("N1", "M7"):
        raise NotImplementedError("rmtExeBatch covers the steady-state models N1 and M7; use rmtExeBatchN2 for N2")
    if B is None:
        if not sweep:
            raise ValueError("give B or a non-empty sweep")
        first = next(iter(sweep.values()))
        B = int(first.shape[0]) if hasattr(first, "shape") else len(first)
    if z_eval is None:
        if modelInput['model'] == "M7":       # dimensional axial coordinate; ReLe must be the same for the whole batch
            L = float(modelInput['reactor']['ReLe'])
            if sweep and "ReLe" in sweep:
                raise ValueError("model M7 integrates over [0, ReLe]: ReLe cannot be swept")
            z_eval = np.linspace(0, L, solverSetting['M9']['zNo']) if profile else np.array([L])
        else:
            z_eval = np.linspace(0, 1, solverSetting['N1']['zNo'] + 1) if profile else np.array([1.0])
    rt = modelInput.get('solver-config', {}).get('rtol', engine.DEFAULT_RTOL) if rtol is None else rtol
    cm = engine.compile_model(modelInput, method=engine.choose_method(modelInput, rt, len(z_eval), dense))
    res = engine.n1_solve_ensemble(cm, modelInput, sweep, B, z_eval=z_eval, rtol=rtol, atol=atol, out_mode=1,
                                   dense=dense, max_steps=max_steps, objective_ref=objective_ref,
                                   keep_on_device=keep_on_device, workspace=workspace, want_stats=return_stats)
    if keep_on_device:
        out = res.out.permute(2, 1, 0)
        data = out[:, :, 0] if out.shape[2] == 1 else out
        success = res.status == 0
    else:
        out = np.transpose(res.out, (2, 1, 0))           # [B][n][n_eval]
        data = out[:, :, 0] if out.shape[2] == 1 else out
        success = res.status == 0
    return {"dataYs": data, "status": res.status, "success": success, "stats": res.stats, "dataXs": res.z_eval,
            "objective": res.objective, "h2d_bytes": res.h2d_bytes, "d2h_bytes": res.d2h_bytes,
            "labelList": list(cm.spec.compList) + (["Temperature", "Pressure"] if cm.spec.model == "M7" else
                                                   ["Pressure"] + ([] if cm.spec.iso else ["Temperature"])),
            "comTime": (timer() - tic)*1000}


def rmtExeBatchN2(modelInput, sweep=None, B=None, *, zNo=None, tNo=None, rtol=None, atol=None,
                  keep_on_device=False, workspace=None):
    """Ensemble form of rmtExe for the dynamic model "N2": B independent reactors integrated over
    [0, period]; returns the profiles at the end of each of the tNo slabs.

    Returns {"dataYs": [B, tNo, nc+1, zNo] (rows y_i..., T [K] — each [b, i] is one dataPack
    entry's dataYs, pbHomoReactor.py:3660-3677), "dataTime": [tNo], "dataXs": [zNo], "status": [B], ...}."""
    tic = timer()
    _check_components(modelInput)
    if modelInput['model'] not in ("N2", "M9"):
        raise NotImplementedError("rmtExeBatchN2 covers the dynamic models N2 and M9")
    if B is None:
        if not sweep:
            raise ValueError("give B or a non-empty sweep")
        first = next(iter(sweep.values()))
        B = int(first.shape[0]) if hasattr(first, "shape") else len(first)
    grid = solverSetting['N2' if modelInput['model'] == "N2" else 'S2']
    zNo = int(grid['zNo'] if zNo is None else zNo)
    tNo = int(grid['tNo'] if tNo is None else tNo)
    cm = engine.compile_model_n2(modelInput, B, zNo)
    res = engine.n2_solve_ensemble(cm, modelInput, sweep, B, zNo=zNo, tNo=tNo, rtol=rtol, atol=atol, out_mode=1,
                                   keep_on_device=keep_on_device, workspace=workspace)
    out = res.out.permute(3, 0, 1, 2) if keep_on_device else np.transpose(res.out, (3, 0, 1, 2))
    if cm.spec.iso:
        # iso-thermal: the reference's dataYs still carries a temperature row (the feed temperature, per reactor)
        T0 = sweep["temperature"] if sweep and "temperature" in sweep else np.full(B, float(modelInput['operating-conditions']['temperature']))
        if keep_on_device:
            import torch
            Tr = torch.as_tensor(np.asarray(T0, dtype=np.float64), device=out.device).view(B, 1, 1, 1).expand(B, tNo, 1, zNo)
            out = torch.cat((out, Tr), dim=2)
        else:
            Tr = np.broadcast_to(np.asarray(T0, dtype=np.float64).reshape(B, 1, 1, 1), (B, tNo, 1, zNo))
            out = np.concatenate((out, Tr), axis=2)
    period = float(modelInput['operating-conditions']['period'])
    return {"dataYs": out, "dataTime": np.linspace(0, period, tNo + 1)[1:], "dataXs": np.linspace(0, 1, zNo),
            "status": res.status, "success": res.status == 0, "stats": res.stats,
            "labelList": list(cm.spec.compList) + ([] if cm.spec.iso else ["Temperature"]),
            "comTime": (timer() - tic)*1000}
