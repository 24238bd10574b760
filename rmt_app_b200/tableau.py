"""Rosenbrock tableaux used by the device integrators.

Single source of truth: `codegen.py` emits these numbers as `constexpr`
arrays into the CUDA translation unit, `tests/test_tableau.py` checks order of
convergence, stiff accuracy and the dense-output polynomial numerically.

Formulation (Hairer & Wanner, "Solving ODEs II", IV.7, the form implemented in
their RODAS code), autonomous y' = f(y), J = f'(y_n):

    (I/(h*gamma) - J) k_i = f(y_n + sum_{j<i} a_ij k_j) + sum_{j<i} (c_ij/h) k_j
    y_{n+1} = y_n + sum_i m_i k_i ,   err = sum_i e_i k_i

The reactor models are autonomous in the integration variable (no explicit z
or t in modelEquationN1/N2, PyREMOT/docs/pbHomoReactor.py:3017, :3706), so the
c_i / d_i time-derivative coefficients of the non-autonomous form are not
needed.

Dense output (RODAS form), s = (x - x_n)/h:
    y(x) = y_n (1-s) + s ( y_{n+1} + (1-s) ( D2 + s D3 ) ),  Dq = sum_i d_qi k_i
"""

RODAS4 = {
    "name": "rodas4",
    "stages": 6,
    "order": 4,
    "gamma": 0.25,
    # a[i][j], i = 1..5 (0-based row i = stage i+1), strictly lower
    "a": [
        [],
        [0.1544000000000000e+01],
        [0.9466785280815826e+00, 0.2557011698983284e+00],
        [0.3314825187068521e+01, 0.2896124015972201e+01, 0.9986419139977817e+00],
        [0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950e+00],
        # stage 6 argument = stage-5 argument + k5
        [0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950e+00, 1.0],
    ],
    "c": [
        [],
        [-0.5668800000000000e+01],
        [-0.2430093356833875e+01, -0.2063599157091915e+00],
        [-0.1073529058151375e+00, -0.9594562251023355e+01, -0.2047028614809616e+02],
        [0.7496443313967647e+01, -0.1024680431464352e+02, -0.3399990352819905e+02, 0.1170890893206160e+02],
        [0.8083246795921522e+01, -0.7981132988064893e+01, -0.3152159432874371e+02, 0.1631930543123136e+02,
         -0.6058818238834054e+01],
    ],
    # y_{n+1} = (stage-6 argument) + k6
    "m": [0.1221224509226641e+01, 0.6019134481288629e+01, 0.1253708332932087e+02, -0.6878860361058950e+00, 1.0, 1.0],
    "e": [0.0, 0.0, 0.0, 0.0, 0.0, 1.0],
    "dense": [
        [0.1012623508344586e+02, -0.7487995877610167e+01, -0.3480091861555747e+02, -0.7992771707568823e+01,
         0.1025137723295662e+01, 0.0],
        [-0.6762803392801253e+00, 0.6087714651680015e+01, 0.1643084320892478e+02, 0.2476722511418386e+02,
         -0.6594389125716872e+01, 0.0],
    ],
}

# Rodas3 (Sandu, Verwer et al. 1997, the KPP "Rodas3" integrator): 4 stages, order 3(2), L-stable,
# stiffly accurate; stage 2 is evaluated at y_n (a21 = 0) so it re-uses f(y_n): 3 RHS evaluations
# (one of them together with the Jacobian) per step.  No dense output: output points are hit by
# stepping onto them.
RODAS3 = {
    "name": "rodas3",
    "stages": 4,
    "order": 3,
    "gamma": 0.5,
    "a": [[], [0.0], [2.0, 0.0], [2.0, 0.0, 1.0]],
    "c": [[], [4.0], [1.0, -1.0], [1.0, -1.0, -8.0/3.0]],
    "m": [2.0, 0.0, 1.0, 1.0],
    "e": [0.0, 0.0, 0.0, 1.0],
    "dense": None,
}

# Ros4 (Hairer & Wanner IV.7, the "L-stable method of order 4"; KPP's Ros4): 4 stages, order 4(3), L-stable
# but not stiffly accurate; stage 4 has the same argument as stage 3 and re-uses its function value:
# 3 RHS evaluations (one with the Jacobian) per step.
ROS4 = {
    "name": "ros4",
    "stages": 4,
    "order": 4,
    "gamma": 0.5728200000000000,
    "a": [[], [2.0], [1.867943637803922, 0.2344449711399156], [1.867943637803922, 0.2344449711399156, 0.0]],
    "c": [[], [-7.137615036412310], [2.580708087951457, 0.6515950076447975],
          [-2.137148994382534, -0.3214669691237626, -0.6949742501781779]],
    "m": [2.255570073418735, 0.2870493262186792, 0.4353179431840180, 1.093502252409163],
    "e": [-0.2815431932141155, -0.07276199124938920, -0.1082196201495311, -1.093502252409163],
    "dense": None,
}

TABLEAUX = {"rodas4": RODAS4, "rodas3": RODAS3, "ros4": ROS4}


def new_function_flags(tab):
    """flag[i] = stage i evaluates f at a new argument; 0 = its argument equals the previous stage's
    (or y_n for the second stage), whose function value is re-used."""
    flags = [1]
    for i in range(1, tab["stages"]):
        row = list(tab["a"][i]) + [0.0]*(tab["stages"] - len(tab["a"][i]))
        prev = (list(tab["a"][i - 1]) + [0.0]*(tab["stages"] - len(tab["a"][i - 1]))) if i > 1 else [0.0]*tab["stages"]
        flags.append(0 if row == prev else 1)
    return flags


def reference_step(tab, f, jac, y, h):
    """NumPy statement of one step (used by tests and by the host-side
    documentation of the kernel; the device code is generated from the same
    coefficient tables).  Returns (y_new, err, K)."""
    import numpy as np
    n = len(y)
    s = tab["stages"]
    W = np.eye(n)/(h*tab["gamma"]) - jac(y)
    K = []
    for i in range(s):
        yi = np.array(y, dtype=float)
        for j in range(len(tab["a"][i])):
            yi = yi + tab["a"][i][j]*K[j]
        rhs = np.array(f(yi), dtype=float)
        for j in range(len(tab["c"][i])):
            rhs = rhs + (tab["c"][i][j]/h)*K[j]
        K.append(np.linalg.solve(W, rhs))
    ynew = np.array(y, dtype=float)
    err = np.zeros(n)
    for i in range(s):
        ynew = ynew + tab["m"][i]*K[i]
        err = err + tab["e"][i]*K[i]
    return ynew, err, K


def dense_eval(tab, y0, y1, K, s):
    import numpy as np
    D2 = sum(d*k for d, k in zip(tab["dense"][0], K))
    D3 = sum(d*k for d, k in zip(tab["dense"][1], K))
    return np.asarray(y0)*(1 - s) + s*(np.asarray(y1) + (1 - s)*(D2 + s*D3))
