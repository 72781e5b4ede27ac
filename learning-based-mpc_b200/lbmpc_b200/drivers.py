"""Python mirrors of the reference's closed-loop drivers, argument for argument:

    ocpLBMPC  <-  matlab/LBMPC/functions/ocpLBMPC.m:1-47      ocpLMPC  <-  matlab/LBMPC/functions/ocpLMPC.m:1-40

The loop bodies follow the .m files line by line (data acquisition, `update_data` window, first move to the plant,
histories w.r.t. the working point).  The one thing replaced is the solver call `fmincon(COSTFUN, opt_var, ..., CONSFUN,
options)` (ocpLBMPC.m:27-31, ocpLMPC.m:20-24): it becomes one call into the GPU engine through the C ABI
(`Solver.solve_batch` / `Solver.solve_sqp`, include/lbmpc.h).  Nothing here solves anything on the CPU.

F-form LBMPC is stated exactly as the reference states it: costLBMPC.m:27 rolls the LEARNED model (nominal + L2NW oracle,
u = K x + c on the learned state), constraintsLBMPC.m:23 rolls the NOMINAL model.  The non-convex problem is solved as a
sequence of QPs (lbmpc_solve_sqp_ex, twin = 1): by default first order — oracle value and Jacobian along the previous
solution give an LTV QP on the learned sequence whose rows follow the nominal sequence — which reproduces the reference's
saved LBMPC_N50_sys_full.mat history to 5e-6 on the inputs over 40 steps; sqp_order=0 freezes the oracle value only.

The plant: trueModel.m integrates the continuous Moore-Greitzer model with ode23 over one sampling period; MATLAB's ode23 is
not reproducible here, so the mirror takes one RK4 step of length Ts (the CasADi scripts' `dynamic`,
DMS_tracking_LMPC_casadi.m:297-304) — step 1 of a run is independent of the plant, later steps agree with the reference's
saved histories to the accuracy of the two integrators (tests compare at 1e-3).
"""
import numpy as np

from .capi import Solver


def mgcm_model(x, u):
    """Continuous Moore-Greitzer compressor model, trueModel.m:21-41."""
    wn2, zeta_wn2 = 1000.0, 2.0 * (1.0 / np.sqrt(2.0)) * np.sqrt(1000.0)
    return np.array([-x[1] + 1.0 + 3.0 * (x[0] / 2.0) - (x[0] ** 3 / 2.0),
                     x[0] + 1.0 - x[2] * np.sqrt(x[1]),
                     x[3],
                     -wn2 * x[2] - zeta_wn2 * x[3] + wn2 * u])


def trueModel(xk, uk, Ts):
    """trueModel.m:1-18 with one RK4 step instead of ode23 (see module docstring)."""
    xk = np.asarray(xk, float).reshape(4)
    k1 = mgcm_model(xk, uk)
    k2 = mgcm_model(xk + Ts / 2 * k1, uk)
    k3 = mgcm_model(xk + Ts / 2 * k2, uk)
    k4 = mgcm_model(xk + Ts * k3, uk)
    return xk + Ts / 6 * (k1 + 2 * k2 + 2 * k3 + k4), uk


def transitionTrue(xk, ck, xw, r0, K, dT):
    """transitionTrue.m:11-12: u = K (x - x_wp) + c + u_wp, then the plant."""
    uk = float(np.asarray(K, float).reshape(-1) @ (np.asarray(xk, float).reshape(-1) - np.asarray(xw, float).reshape(-1))
               + float(np.ravel(ck)[0]) + float(np.ravel(r0)[0]))
    xk1, _ = trueModel(xk, uk, dT)
    return xk1, uk


def update_data(X, Y, q, it, data_old):
    """utilities/update_data.m:3-10: moving window of q data points (columns)."""
    X, Y = np.asarray(X, float).reshape(-1, 1), np.asarray(Y, float).reshape(-1, 1)
    if it < q:
        return {"X": np.hstack([data_old["X"], X]), "Y": np.hstack([data_old["Y"], Y])}
    return {"X": np.hstack([data_old["X"][:, 1:], X]), "Y": np.hstack([data_old["Y"][:, 1:], Y])}


def _model(A, B, K, Q, R, P, T, LAMBDA, PSI, F_x, h_x, F_u, h_u, F_w_N, h_w_N, F_x_d=None, h_x_d=None):
    return dict(A=A, B=B, K=K, Q=Q, R=R, P=P, T=T, LAMBDA=LAMBDA, PSI=PSI, F_x=F_x, h_x=h_x, F_u=F_u, h_u=h_u, F_w_N=F_w_N,
                h_w_N=h_w_N, F_x_d=F_x_d, h_x_d=h_x_d)


def ocpLBMPC(x, x_wp, dx_init, dx_ref, u_wp, N, Ts, iterations, options, opt_var, data, A, B, Kstabil, Q, R, P, T, Mtheta,
             LAMBDA, PSI, m, F_x, h_x, F_u, h_u, F_w_N, h_w_N, F_x_d, h_x_d, sysHistory, art_refHistory, true_refHistory,
             sqp_iters=2, sqp_order=1, q=100, device=0, info=None):
    """ocpLBMPC.m:1-47 with the fmincon call replaced by the GPU engine.  `options` (fmincon options) is accepted and
    ignored.  Histories are returned as (5, 1 + iterations), (m, 1 + iterations), (n, 1 + iterations) arrays like the
    reference's horizontally concatenated matrices.  `info` (optional dict) receives per-step iterations / status."""
    x, x_wp = np.asarray(x, float).reshape(-1), np.asarray(x_wp, float).reshape(-1)
    dx_ref = np.asarray(dx_ref, float).reshape(-1)
    n = x.size
    sol = Solver(_model(A, B, Kstabil, Q, R, P, T, LAMBDA, PSI, F_x, h_x, F_u, h_u, F_w_N, h_w_N, F_x_d, h_x_d),
                 "F", "LBMPC", int(N), device=device, max_batch=1)
    opt_var = np.asarray(opt_var, float).reshape(-1).copy()
    data = {"X": np.asarray(data["X"], float).reshape(3, -1), "Y": np.asarray(data["Y"], float).reshape(n, -1)}
    sysH = [np.asarray(sysHistory, float).reshape(n + m, -1)]
    artH = [np.asarray(art_refHistory, float).reshape(m, -1)]
    refH = [np.asarray(true_refHistory, float).reshape(n, -1)]
    Mtheta = np.asarray(Mtheta, float)
    A_, B_ = np.asarray(A, float), np.asarray(B, float).reshape(n, m)
    its, sts = [], []
    u = x_k1 = None
    try:
        for k in range(1, int(iterations) + 1):
            if k > 1:
                X = np.concatenate([x[:2] - x_wp[:2], [u - float(np.ravel(u_wp)[0])]])            # ocpLBMPC.m:14
                Y = (x_k1 - x_wp) - (A_ @ (x - x_wp) + B_[:, 0] * (u - float(np.ravel(u_wp)[0])))   # ocpLBMPC.m:15
                x = x_k1
                data = update_data(X, Y, q, k, data)                                               # ocpLBMPC.m:19
                dx = x - x_wp
            else:
                dx = np.asarray(dx_init, float).reshape(-1)
            # ---- replaces fmincon(COSTFUN, opt_var, ..., CONSFUN, options), ocpLBMPC.m:27-31 ----
            out = sol.solve_sqp(dx[None, :], data["X"].T[None], data["Y"].T[None], sqp_iters=sqp_iters, dx_ref=dx_ref[None, :],
                                warm=opt_var[None, :], twin=True, order=sqp_order, want_x=False)
            opt_var = np.concatenate([out["uc"].reshape(-1), out["theta"].reshape(-1)])
            its.append(int(out["iters"][0])); sts.append(int(out["status"][0]))
            theta_opt = opt_var[-m:]
            c = opt_var[:m]
            art_ref = Mtheta @ theta_opt
            x_k1, u = transitionTrue(x, c, x_wp, u_wp, Kstabil, Ts)                                 # ocpLBMPC.m:37
            sysH.append(np.concatenate([x - x_wp, [u - float(np.ravel(u_wp)[0])]]).reshape(-1, 1))   # ocpLBMPC.m:40
            artH.append(np.asarray(art_ref[:m], float).reshape(m, 1))
            refH.append(dx_ref.reshape(-1, 1))
    finally:
        sol.close()
    if info is not None:
        info.update(iters=np.array(its), status=np.array(sts), data=data, opt_var=opt_var)
    return np.hstack(sysH), np.hstack(artH), np.hstack(refH)


def ocpLMPC(x, x_wp_init, x_wp, x_wp_ref, u_wp, N, Ts, iterations, options, opt_var, Kstabil, Q, R, P, T, Mtheta, LAMBDA, PSI,
            m, F_x, h_x, F_u, h_u, F_w_N, h_w_N, sysHistory, art_refHistory, true_refHistory, A=None, B=None, device=0,
            info=None):
    """ocpLMPC.m:1-40 with the fmincon call replaced by the GPU engine.  The reference hard-codes A, B inside
    nominalModel.m:14-21; pass them (or leave None for the Moore-Greitzer values of mgcmDLTI)."""
    from .model import mgcmDLTI
    if A is None or B is None:
        A, B = mgcmDLTI()[:2]
    x, x_wp = np.asarray(x, float).reshape(-1), np.asarray(x_wp, float).reshape(-1)
    x_wp_ref = np.asarray(x_wp_ref, float).reshape(-1)
    n = x.size
    sol = Solver(_model(A, B, Kstabil, Q, R, P, T, LAMBDA, PSI, F_x, h_x, F_u, h_u, F_w_N, h_w_N), "F", "LMPC", int(N),
                 device=device, max_batch=1)
    opt_var = np.asarray(opt_var, float).reshape(-1).copy()
    sysH = [np.asarray(sysHistory, float).reshape(n + m, -1)]
    artH = [np.asarray(art_refHistory, float).reshape(m, -1)]
    refH = [np.asarray(true_refHistory, float).reshape(n, -1)]
    Mtheta = np.asarray(Mtheta, float)
    its, sts = [], []
    try:
        for k in range(1, int(iterations) + 1):
            dx = x - x_wp if k > 1 else np.asarray(x_wp_init, float).reshape(-1)                    # ocpLMPC.m:13-17
            out = sol.solve_batch(dx[None, :], x_wp_ref[None, :], None, opt_var[None, :], want_x=False)   # ocpLMPC.m:20-24
            opt_var = np.concatenate([out["uc"].reshape(-1), out["theta"].reshape(-1)])
            its.append(int(out["iters"][0])); sts.append(int(out["status"][0]))
            art_ref = Mtheta @ opt_var[-m:]
            x, u = transitionTrue(x, opt_var[:m], x_wp, u_wp, Kstabil, Ts)                          # ocpLMPC.m:30
            sysH.append(np.concatenate([x - x_wp, [u - float(np.ravel(u_wp)[0])]]).reshape(-1, 1))   # ocpLMPC.m:33 (the NEW state)
            artH.append(np.asarray(art_ref[:m], float).reshape(m, 1))
            refH.append(x_wp_ref.reshape(-1, 1))
    finally:
        sol.close()
    if info is not None:
        info.update(iters=np.array(its), status=np.array(sts), opt_var=opt_var)
    return np.hstack(sysH), np.hstack(artH), np.hstack(refH)
