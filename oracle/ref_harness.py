"""Drive the UNMODIFIED reference (staged by oracle/stage_reference.py, or /root/reference in the build container).

TEST / BENCH INFRASTRUCTURE ONLY — never imported by rmt_app_b200/.  Used by bench.py's CPU legs (`--impl reference`,
`cpu_baseline`, `--cpu-baselines`) and by tests that cross-check the oracle port against the live reference.

What is NOT the reference's code here (all of it outside the timed arithmetic):
* a stub `matplotlib` in sys.modules — the image has no matplotlib and every model module imports it at import time
  (PyREMOT/library/plot.py:7); figures are never drawn because the inputs carry display-result = "False";
* the console progress bar is silenced (PyREMOT/docs/pbHomoReactor.py prints one line per RHS call to stdout);
* `Capture` wraps the `solve_ivp` symbol the model module calls, to read nfev/njev and (optionally) to inject a
  method or tolerances — the default run passes everything through untouched.
"""
import contextlib
import io
import os
import sys
import time
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
_loaded = None


def reference_root():
    """Where an importable `PyREMOT` package lives: oracle/_ref (staged copy) first, then $RMT_REFERENCE or
    /root/reference.  None when neither exists."""
    for root in (os.path.join(HERE, "_ref"), os.environ.get("RMT_REFERENCE", "/root/reference")):
        if root and os.path.isfile(os.path.join(root, "PyREMOT", "rmt.py")):
            return root
    return None


def available():
    return reference_root() is not None


def load():
    """(PyREMOT module, pbHomoReactor module, solverSetting dict, root) of the unmodified reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference is neither staged under oracle/_ref nor present at /root/reference")
    for n in ("matplotlib", "matplotlib.pyplot"):
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__getattr__ = lambda k: (lambda *a, **kw: None)
            sys.modules[n] = m
    if root not in sys.path:
        sys.path.insert(0, root)
    warnings.simplefilter("ignore")
    import PyREMOT
    import PyREMOT.docs.pbHomoReactor as H
    from PyREMOT.solvers import solverSetting
    H.printProgressBar = lambda *a, **k: None
    _loaded = (PyREMOT, H, solverSetting, root)
    return _loaded


class Capture:
    """Wraps `solve_ivp` as seen by PyREMOT/docs/pbHomoReactor.py (call sites :2931, :3609)."""

    def __init__(self, inject=None, keep_fun=False, stop_after=None, capture_only=False):
        self.H = load()[1]
        self.inject = dict(inject or {})
        self.calls = []
        self.keep_fun = keep_fun
        self.stop_after = stop_after
        self.capture_only = capture_only        # record (fun, y0, args) of the first call and stop without integrating
        self._orig = self.H.solve_ivp

    def __enter__(self):
        def patched(fun, t_span, y0, **kw):
            kw = dict(kw)
            kw.update(self.inject)
            if self.capture_only:
                self.calls.append(dict(fun=fun, args=kw.get("args"), y0=y0, t_span=tuple(float(v) for v in t_span)))
                raise StopAfter()
            t0 = time.perf_counter()
            sol = self._orig(fun, t_span, y0, **kw)
            rec = dict(nfev=sol.nfev, njev=sol.njev, nlu=sol.nlu, wall=time.perf_counter() - t0,
                       t_span=tuple(float(v) for v in t_span))
            if self.keep_fun:
                rec.update(fun=fun, args=kw.get("args"), y0=y0, y=sol.y, t=sol.t)
            self.calls.append(rec)
            if self.stop_after is not None and len(self.calls) >= self.stop_after:
                raise StopAfter()
            return sol
        self.H.solve_ivp = patched
        return self

    def __exit__(self, *a):
        self.H.solve_ivp = self._orig


class StopAfter(Exception):
    pass


def rmtExe(modelInput, quiet=True):
    """The reference's public entry point (PyREMOT/rmt.py:21-80) on `modelInput`; returns (result, wall seconds)."""
    P = load()[0]
    mi = dict(modelInput)
    mi["solver-config"] = dict(mi.get("solver-config", {}), **{"display-result": "False"})
    t0 = time.perf_counter()
    if quiet:
        with contextlib.redirect_stdout(io.StringIO()):
            res = P.rmtExe(mi)
    else:
        res = P.rmtExe(mi)
    return res, time.perf_counter() - t0


def set_grid(model, **kw):
    """Mutate the reference's own module-level grid dict (solvers/solSetting.py:30-39), the only way to change zNo."""
    load()[2][model].update(kw)
