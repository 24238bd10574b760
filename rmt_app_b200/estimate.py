"""Population-based kinetic-parameter estimation on top of the ensemble kernel
(the use-case the reference's README.md:5 claims but never implements).

The inner loop — "integrate one steady-state reactor per parameter set and score its outlet
against data" — is `rmt_n1_solve` with the fused objective; this module only proposes
parameter sets (differential evolution, rand/1/bin) and keeps the best.  With several ranks
(torch.distributed initialised) the population is sharded with `rmtExeBatchSharded` and the
per-instance objectives are all-gathered, so every rank evolves the same population.
"""
import numpy as np


def differential_evolution(modelInput, bounds, outlet_ref, popsize=4096, generations=30, F=0.6, CR=0.9, seed=0,
                           rtol=None, atol=None, callback=None, workspace=None):
    """Minimise sum_k ((out_k - ref_k)/ref_k)^2 over the scalar VARS entries named in `bounds`
    ({name: (lo, hi)}); `outlet_ref` = [y_i..., P, T] like a dataYs column.

    Returns {"x": best parameters (dict), "fun": best objective, "history": best per generation,
             "population": [popsize, d], "objective": [popsize], "nsolves": total reactor solves}."""
    from . import engine
    from .ensemble import rmtExeBatchSharded, world_info
    names = list(bounds)
    lo = np.array([bounds[k][0] for k in names], float)
    hi = np.array([bounds[k][1] for k in names], float)
    if not np.all(hi > lo):
        raise ValueError("every bound needs hi > lo")
    rng = np.random.default_rng(seed)                 # same seed on every rank -> same proposals
    d = len(names)
    ws = workspace if workspace is not None else engine.Workspace()
    ref = np.asarray(outlet_ref, float)
    multi = world_info()[1] > 1
    rt = modelInput.get('solver-config', {}).get('rtol', engine.DEFAULT_RTOL) if rtol is None else rtol
    cm = engine.compile_model(modelInput, method=engine.choose_method(modelInput, rt, 1))

    def score(P):
        sweep = {k: np.ascontiguousarray(P[:, j]) for j, k in enumerate(names)}
        if multi:
            r = rmtExeBatchSharded(modelInput, sweep, len(P), rtol=rtol, atol=atol, objective_ref=ref, workspace=ws)
            obj = r["objective"]
        else:
            r = engine.n1_solve_ensemble(cm, modelInput, sweep, len(P), rtol=rtol, atol=atol, objective_ref=ref,
                                         workspace=ws, want_stats=False)
            obj = np.array(r.objective)
        return np.where(np.isfinite(obj), obj, np.inf)

    pop = lo + rng.random((popsize, d))*(hi - lo)
    fit = score(pop)
    nsolves = popsize
    history = [float(fit.min())]
    idx = np.arange(popsize)
    for g in range(generations):
        a, b, c = (rng.permutation(popsize) for _ in range(3))
        mutant = np.clip(pop[a] + F*(pop[b] - pop[c]), lo, hi)
        cross = rng.random((popsize, d)) < CR
        cross[idx, rng.integers(0, d, popsize)] = True
        trial = np.where(cross, mutant, pop)
        tf = score(trial)
        nsolves += popsize
        better = tf <= fit
        pop[better], fit[better] = trial[better], tf[better]
        history.append(float(fit.min()))
        if callback is not None and callback(g, pop, fit):
            break
    best = int(np.argmin(fit))
    return {"x": {k: float(pop[best, j]) for j, k in enumerate(names)}, "fun": float(fit[best]), "history": history,
            "population": pop, "objective": fit, "nsolves": nsolves}
