"""Host-side (offline) model setup — the Python mirror of the reference's setup functions.

These run once per script on the host in the reference too (MATLAB Symbolic/Control toolboxes,
MPT3); they are inputs of the hot path, not part of it:

  mgcmDLTI()    functions/mgcmDLTI.m:1-43    linearise + exact discretisation (Ts = 0.01)
  matOCP()      functions/matOCP.m:1-33      K by pole placement, steady-state map, Q,R,P,T
  getCONS()     functions/getCONS.m:1-60     boxes + precomputed LMPC terminal set (term_set.mat)
  getCONSPOLY() functions/getCONSPOLY.m:1-74 boxes, X(-)D and the robust terminal set.  MPT3 is not
                available here: the robust sets for the reference's default uncertainty bound
                state_uncert=[0.02;5e-4;0;0] (LBMPC_RunExample.m:38) ship as data
                (data/moore_greitzer_sets.npz, taken from the reference's saved workspace).
"""
import os

import numpy as np
import scipy.linalg as sla
import scipy.signal as ssig

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "moore_greitzer_sets.npz")

# working point of the Moore-Greitzer compressor (LBMPC_RunExample.m:52-56)
X_WP = np.array([0.5, 1.6875, 1.1547, 0.0])
U_WP = 1.1547


def mgcmDLTI():
    """[A,B,C,D,Ts] of mgcmDLTI.m: Jacobian of the MGCM ODE at the working point, then
    Ad = expm(A Ts), Bd = (Ad - I) A^-1 B (mgcmDLTI.m:24-39)."""
    wn, zeta, beta = np.sqrt(1000.0), 1.0 / np.sqrt(2.0), 1.0
    x1, x2, x3 = X_WP[0], X_WP[1], X_WP[2]
    A = np.array([[1.5 - 1.5 * x1 ** 2, -1.0, 0.0, 0.0],
                  [1.0 / beta ** 2, -x3 / (2.0 * np.sqrt(x2)) / beta ** 2, -np.sqrt(x2) / beta ** 2, 0.0],
                  [0.0, 0.0, 0.0, 1.0],
                  [0.0, 0.0, -wn ** 2, -2.0 * zeta * wn]])
    B = np.array([[0.0], [0.0], [0.0], [wn ** 2]])
    Ts = 0.01
    Ad = sla.expm(A * Ts)
    Bd = (Ad - np.eye(4)) @ np.linalg.solve(A, B)
    return Ad, Bd, np.eye(4), np.zeros((4, 1)), Ts


def matOCP(A, B, C=None):
    """[Ks,Klqr,Q,R,P,T,Mtheta,LAMBDA,PSI,LAMBDA_0,PSI_0] of matOCP.m."""
    n, m = B.shape
    C = np.eye(n) if C is None else C
    o = C.shape[0]
    K = ssig.place_poles(A, B, [0.75, 0.78, 0.98, 0.99][:n]).gain_matrix   # matOCP.m:7-8
    Ks = -K
    M = np.block([[A - np.eye(n), B, np.zeros((n, o))], [C, np.zeros((o, m)), -np.eye(o)]])
    Mtheta = sla.null_space(M)                                             # matOCP.m:12-16
    if Mtheta[0, 0] < 0:                                                   # MATLAB's null() sign
        Mtheta = -Mtheta
    LAMBDA, PSI = Mtheta[:n, :], Mtheta[n:n + m, :]
    Q, R = np.eye(n), np.eye(m)
    Klqr = -_dlqr(A, B, Q, R)
    P = sla.solve_discrete_are(A + B @ Ks, B, Q, R)                        # matOCP.m:30
    T = 1000.0
    return Ks, Klqr, Q, R, P, T, Mtheta, LAMBDA, PSI, np.zeros((n, 1)), np.zeros((m, 1))


def _dlqr(A, B, Q, R):
    X = sla.solve_discrete_are(A, B, Q, R)
    return np.linalg.solve(R + B.T @ X @ B, B.T @ X @ A)


def _boxes(xmax, xmin, umax, umin, x_wp, u_wp):
    xmax, xmin = np.atleast_1d(xmax).astype(float), np.atleast_1d(xmin).astype(float)
    umax, umin = np.atleast_1d(umax).astype(float), np.atleast_1d(umin).astype(float)
    n, m = xmax.size, umax.size
    F_u = np.vstack([np.eye(m), -np.eye(m)]); h_u = np.concatenate([umax - u_wp, -umin + u_wp])
    F_x = np.vstack([np.eye(n), -np.eye(n)]); h_x = np.concatenate([xmax - x_wp, -xmin + x_wp])
    return F_x, h_x, F_u, h_u


# default physical limits (LBMPC_RunExample.m:26-34)
XMAX = np.array([1.0, 2.1875, 2.1547, 20.0])
XMIN = np.array([0.0, 1.1875, 0.1547, -20.0])
UMAX, UMIN = 2.1547, 0.1547


def getCONS(xmax=XMAX, xmin=XMIN, umax=UMAX, umin=UMIN, x_wp=X_WP, u_wp=U_WP, recompute=False):
    """[F_x,h_x,F_u,h_u,F_w_N,h_w_N] of getCONS.m: boxes (:15-16) + precomputed terminal set (:57-58, term_set.mat).
    Other boxes (or recompute=True) run the invariant-set iteration the reference has commented out (:30-50) through
    lbmpc_b200.sets — about a minute of linear programmes for the 616-row set, which it reproduces row for row."""
    F_x, h_x, F_u, h_u = _boxes(xmax, xmin, umax, umin, x_wp, u_wp)
    default = (np.allclose(xmax, XMAX) and np.allclose(xmin, XMIN) and np.allclose(umax, UMAX) and np.allclose(umin, UMIN)
               and np.allclose(x_wp, X_WP) and np.allclose(u_wp, U_WP))
    if default and not recompute:
        d = np.load(_DATA)
        return F_x, h_x, F_u, h_u, d["term_set_F_w_N"].copy(), d["term_set_h_w_N"].copy()
    from .sets import lmpc_terminal_set
    A, B, C, _, _ = mgcmDLTI()
    Ks, _, _, _, _, _, _, LAMBDA, PSI, L0, P0 = matOCP(A, B, C)
    F_w_N, h_w_N = lmpc_terminal_set(A, B, Ks, LAMBDA, PSI, F_x, h_x, F_u, h_u, L0, P0)
    return F_x, h_x, F_u, h_u, F_w_N, h_w_N


def pdiff(F_u, h_u, F_v, h_v):
    """Pontryagin difference {x: F_u x <= h_u} (-) {v: F_v v <= h_v} in H-representation, as utilities/pdiff.m:8-17: one
    LP per row (support function of the subtrahend in the direction of the row).  Host-side, scipy linprog (HiGHS)."""
    from scipy.optimize import linprog
    F_u, h_u = np.atleast_2d(np.asarray(F_u, float)), np.asarray(h_u, float).ravel()
    sup = np.empty(len(h_u))
    for i, row in enumerate(F_u):
        res = linprog(-row, A_ub=np.atleast_2d(F_v), b_ub=np.asarray(h_v, float).ravel(), bounds=[(None, None)] * F_u.shape[1],
                      method="highs")
        if res.status != 0:
            raise ValueError(f"pdiff: support LP of row {i} failed ({res.message})")
        sup[i] = -res.fun
    return F_u.copy(), h_u - sup


def tightened_state_set(F_x, h_x, state_uncert):
    """X (-) D of getCONSPOLY.m:17-30 with the box D = {|d_j| <= state_uncert_j} (:17-19): [F_x_d, h_x_d].  (MPT's
    minHRep only reorders / normalises the rows of this box-minus-box; the rows here keep the order of F_x.)"""
    w = np.asarray(state_uncert, float).ravel()
    n = w.size
    return pdiff(F_x, h_x, np.vstack([np.eye(n), -np.eye(n)]), np.concatenate([w, w]))


def getCONSPOLY(xmax=XMAX, xmin=XMIN, umax=UMAX, umin=UMIN, state_uncert=(0.02, 5e-4, 0.0, 0.0), x_wp=X_WP,
                u_wp=U_WP):
    """[F_x,h_x,F_u,h_u,F_w_N,h_w_N,F_x_d,h_x_d] of getCONSPOLY.m.  The reference's default uncertainty bound ships as
    data (the values of its own MPT run, examples/DSS_NMPC.m); any other bound or box is computed here by the MPT-free
    restatement in lbmpc_b200.sets (which reproduces the shipped sets, tests/test_sets.py)."""
    F_x, h_x, F_u, h_u = _boxes(xmax, xmin, umax, umin, x_wp, u_wp)
    default = (np.allclose(state_uncert, (0.02, 5e-4, 0.0, 0.0)) and np.allclose(xmax, XMAX) and np.allclose(xmin, XMIN)
               and np.allclose(umax, UMAX) and np.allclose(umin, UMIN) and np.allclose(x_wp, X_WP) and np.allclose(u_wp, U_WP))
    if default:
        d = np.load(_DATA)
        return (F_x, h_x, F_u, h_u, d["lbmpc_F_w_N"].copy(), d["lbmpc_h_w_N"].copy(), d["lbmpc_F_x_d"].copy(),
                d["lbmpc_h_x_d"].copy())
    from .sets import lbmpc_terminal_set
    A, B, C, _, _ = mgcmDLTI()
    _, _, Q, R, _, _, _, LAMBDA, PSI, L0, P0 = matOCP(A, B, C)
    F_w_N, h_w_N, F_x_d, h_x_d = lbmpc_terminal_set(A, B, Q, R, LAMBDA, PSI, F_x, h_x, F_u, h_u, state_uncert, L0, P0)
    return F_x, h_x, F_u, h_u, F_w_N, h_w_N, F_x_d, h_x_d


def moore_greitzer_model(variant="LMPC"):
    """Convenience: dict with every matrix of the ocpLBMPC.m:1-6 argument list."""
    A, B, C, D, Ts = mgcmDLTI()
    Ks, Klqr, Q, R, P, T, Mtheta, LAMBDA, PSI, L0, P0 = matOCP(A, B, C)
    mdl = dict(A=A, B=B, K=Ks, Klqr=Klqr, Q=Q, R=R, P=P, T=T, Mtheta=Mtheta, LAMBDA=LAMBDA, PSI=PSI, Ts=Ts,
               x_wp=X_WP.copy(), u_wp=U_WP)
    if variant == "LMPC":
        F_x, h_x, F_u, h_u, F_w_N, h_w_N = getCONS()
    else:
        F_x, h_x, F_u, h_u, F_w_N, h_w_N, F_x_d, h_x_d = getCONSPOLY()
        mdl.update(F_x_d=F_x_d, h_x_d=h_x_d)
    mdl.update(F_x=F_x, h_x=h_x, F_u=F_u, h_u=h_u, F_w_N=F_w_N, h_w_N=h_w_N)
    return mdl


def double_integrator_model(lam=0.99, invariant=False):
    """The second problem shape of the reference (matlab/trackingMPC/RunExample.m): sampled double integrator with two
    inputs (:20-28), LQR gain, P = dare(A+BK,B,Q,R), T = 100 P (:57-62), boxes |x| <= 5, |u| <= 0.3 (:64-67), steady-state
    parametrisation null space (:40-45) and the extended admissible set X_ext of (x, theta) written out in closed form at
    :84-93.  invariant=True replaces it by the maximal positively invariant subset the reference then computes with
    MPT (:105-108, compute_MPIS.m) — here by lbmpc_b200.sets.compute_mpis.  nx = nu = nt = 2."""
    A = np.array([[1.0, 1.0], [0.0, 1.0]])
    B = np.array([[0.0, 0.5], [1.0, 0.5]])
    Cm = np.array([[1.0, 0.0]])
    n, m, o = 2, 2, 1
    Q, R = np.eye(n), np.eye(m)
    M = np.block([[A - np.eye(n), B, np.zeros((n, o))], [Cm, np.zeros((o, m)), -np.eye(o)]])
    Mtheta = sla.null_space(M)
    LAMBDA, PSI = Mtheta[:n, :], Mtheta[n:n + m, :]
    K = -_dlqr(A, B, Q, R)
    P = sla.solve_discrete_are(A + B @ K, B, Q, R)
    T = 100.0 * P
    F_u = np.vstack([np.eye(m), -np.eye(m)]); h_u = np.full(2 * m, 0.3)
    F_x = np.vstack([np.eye(n), -np.eye(n)]); h_x = np.full(2 * n, 5.0)
    L = PSI - K @ LAMBDA
    F_w = np.block([[F_x, np.zeros((2 * n, m))], [np.zeros((2 * n, n)), F_x @ LAMBDA], [F_u @ K, F_u @ L],
                    [np.zeros((2 * m, n)), F_u @ PSI]])
    h_w = np.concatenate([h_x, lam * h_x, h_u, lam * h_u])
    if invariant:
        from .sets import compute_mpis, min_hrep
        Ak = np.block([[A + B @ K, B @ L], [np.zeros((m, n)), np.eye(m)]])
        F_w, h_w = min_hrep(*compute_mpis(F_w, h_w, Ak)[:2])
    return dict(A=A, B=B, K=K, Q=Q, R=R, P=P, T=T, Mtheta=Mtheta, LAMBDA=LAMBDA, PSI=PSI, F_x=F_x, h_x=h_x, F_u=F_u, h_u=h_u,
                F_w_N=F_w, h_w_N=h_w, x_wp=np.zeros(n), u_wp=np.zeros(m))
