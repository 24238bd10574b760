"""Rosenbrock tableau: order of convergence, embedded estimate, dense output (NumPy statement)."""
import numpy as np
from scipy.integrate import solve_ivp

from rmt_app_b200.tableau import RODAS4, dense_eval, reference_step


def _vdp():
    mu = 5.0
    f = lambda y: np.array([y[1], mu*(1 - y[0]**2)*y[1] - y[0]])
    J = lambda y: np.array([[0, 1], [-2*mu*y[0]*y[1] - 1, mu*(1 - y[0]**2)]])
    return f, J, np.array([2.0, 0.0])


def test_rodas4_orders():
    f, J, y0 = _vdp()
    ref = solve_ivp(lambda t, y: f(y), [0, 1], y0, rtol=1e-13, atol=1e-14, method="Radau", dense_output=True)
    errs, derr, est = [], [], []
    for N in (40, 80, 160):
        h, y, dmax, emax = 1.0/N, y0.copy(), 0.0, 0.0
        for i in range(N):
            yn, e, K = reference_step(RODAS4, f, J, y, h)
            dmax = max(dmax, np.max(np.abs(dense_eval(RODAS4, y, yn, K, 0.5) - ref.sol((i + 0.5)*h))))
            emax = max(emax, np.max(np.abs(e)))
            y = yn
        errs.append(np.max(np.abs(y - ref.y[:, -1]))); derr.append(dmax); est.append(emax)
    order = np.log2(np.array(errs[:-1])/np.array(errs[1:]))
    assert np.all(order > 3.8), order                    # 4th-order propagated solution
    assert np.all(np.log2(np.array(derr[:-1])/np.array(derr[1:])) > 3.3)     # 3rd-order dense output (+1 global)
    assert np.all(np.log2(np.array(est[:-1])/np.array(est[1:])) > 3.3)       # embedded 3rd order: local estimate O(h^4)


def test_other_tableaux_orders():
    """Rodas3 (order 3) and the L-stable Ros4 (order 4): convergence order and re-used function values."""
    from rmt_app_b200.tableau import RODAS3, ROS4, new_function_flags
    f, J, y0 = _vdp()
    ref = solve_ivp(lambda t, y: f(y), [0, 1], y0, rtol=1e-13, atol=1e-14, method="Radau")
    for tab, want in ((RODAS3, 2.9), (ROS4, 3.75)):
        errs = []
        for N in (80, 160, 320):
            h, y = 1.0/N, y0.copy()
            for _ in range(N):
                y, _, _ = reference_step(tab, f, J, y, h)
            errs.append(np.max(np.abs(y - ref.y[:, -1])))
        order = np.log2(np.array(errs[:-1])/np.array(errs[1:]))
        assert np.all(order > want), (tab["name"], order)
        lam = -1e9
        yn, _, _ = reference_step(tab, lambda y: lam*(y - 1.0), lambda y: np.array([[lam]]), np.array([5.0]), 0.1)
        assert abs(yn[0] - 1.0) < 1e-3                     # L-stable: R(inf) = 0
    assert new_function_flags(RODAS3) == [1, 0, 1, 1] and new_function_flags(ROS4) == [1, 1, 1, 0]
    assert new_function_flags(RODAS4) == [1]*6


def test_rodas4_is_stiffly_accurate_and_l_stable():
    # y' = lam*(y - 1): one huge step must land on the slow manifold (R(inf) = 0)
    lam = -1e9
    f = lambda y: lam*(y - 1.0)
    J = lambda y: np.array([[lam]])
    yn, e, _ = reference_step(RODAS4, f, J, np.array([5.0]), 0.1)
    assert abs(yn[0] - 1.0) < 1e-6
    assert sum(RODAS4["e"]) == 1.0 and RODAS4["m"][-1] == 1.0 and RODAS4["gamma"] == 0.25
