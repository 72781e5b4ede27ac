"""Independent dense cross-check of the LBMPC/LMPC QPs (TEST INFRASTRUCTURE ONLY).

This file is *not* the oracle the GPU path is held to (that is the C restatement in
`oracle/lbmpc_oracle.c`); it is a second, structurally different statement of the same
optimisation problems used to validate the C oracle: the problem is *condensed* to the
(N*m + m) decision variables the reference hands to fmincon (`opt_var=[c;theta]`,
LBMPC_RunExample.m:69-71) / the input+theta part of CasADi's `y` (DMS_tracking_LMPC_casadi.m:122),
all inequality rows are stacked into one dense `G y <= h` exactly in the reference's
row order, and the QP is solved with a dense Mehrotra predictor-corrector.

Problem definitions follow
  F-form: costLMPC.m:20-45, costLBMPC.m:20-45, constraintsLMPC.m:18-41, constraintsLBMPC.m:18-45,
          transitionNominal.m:12 (u = K x + c)
  C-form: DMS_tracking_LMPC_casadi.m:223-291, LBMPC_casadi.m:240-305
Only tests/ may import this module.
"""
import numpy as np


def condensed_qp(form, variant, mdl, N, dx0, dx_ref=None, delta=0.01, d_off=None, cost_shift=None):
    """Return (H, g, c0, G, h, meta) with J(y) = 0.5 y'Hy + g'y + c0 and rows G y <= h.

    form    'F' (fmincon scripts, y=[c;theta]) or 'C' (CasADi scripts, y=[du;theta])
    variant 'LMPC' (terminal set on last state) or 'LBMPC' (robust rows on x_1)
    mdl     dict with A,B,K,Q,R,P,T,LAMBDA,PSI,F_x,h_x,F_u,h_u,F_w_N,h_w_N[,F_x_d,h_x_d]
    cost_shift (N+1, n): the objective sees x_k + cost_shift_k, the rows see x_k (twin sequences, DMS_LBMPC_casadi.m:252-319)
    """
    A, B, K = mdl["A"], mdl["B"], mdl["K"]
    n, m = B.shape
    Q, R, P = mdl["Q"], mdl["R"], mdl["P"]
    T = np.atleast_2d(mdl["T"])
    if T.shape == (1, 1):
        T = T[0, 0] * np.eye(n)
    Lam, Psi = mdl["LAMBDA"], mdl["PSI"]
    nt = Lam.shape[1]
    nv = N * m + nt
    dx_ref = np.zeros(n) if dx_ref is None else np.asarray(dx_ref, float)
    # affine maps: x_k = Phi_k dx0 + Gam_k y + off_k ;  u_k = Uk y + ku_k
    Abar = A + B @ K if form == "F" else A
    Xmap = [np.zeros((n, nv))]
    Xoff = [np.asarray(dx0, float).copy()]
    Umap, Uoff = [], []
    for k in range(N):
        E = np.zeros((m, nv))
        E[:, k * m:(k + 1) * m] = np.eye(m)
        if form == "F":
            Umap.append(K @ Xmap[k] + E)
            Uoff.append(K @ Xoff[k])
        else:
            Umap.append(E)
            Uoff.append(np.zeros(m))
        dk = np.zeros(n) if d_off is None else d_off[:, k]
        Xmap.append(A @ Xmap[k] + B @ Umap[k])
        Xoff.append(A @ Xoff[k] + B @ Uoff[k] + dk)
    Th = np.zeros((nt, nv))
    Th[:, N * m:] = np.eye(nt)
    H = np.zeros((nv, nv))
    g = np.zeros(nv)
    c0 = 0.0

    def add_quad(M, off, W, scale):
        nonlocal H, g, c0
        H += 2 * scale * M.T @ W @ M
        g += 2 * scale * M.T @ W @ off
        c0 += scale * off @ W @ off

    if form == "C":
        stages, sc = range(N), delta                 # k=1..N, delta-scaled (…casadi.m:233-237)
    else:
        stages, sc = range(N - 2), 1.0               # "if k < N-1" (costLMPC.m:30)
    ek = (lambda k: 0.0) if cost_shift is None else (lambda k: np.asarray(cost_shift, float)[k])
    for k in stages:
        add_quad(Xmap[k] - Lam @ Th, Xoff[k] + ek(k), Q, sc)
        add_quad(Umap[k] - Psi @ Th, Uoff[k], R, sc)
    add_quad(Xmap[N] - Lam @ Th, Xoff[N] + ek(N), P, 1.0)    # terminal cost on x_N
    add_quad(Lam @ Th, -dx_ref, T, 1.0)              # (LAMBDA*theta - xs)'T(.)
    rows, rhs = [], []

    def add_rows(F, M, off, hh):
        rows.append(F @ M)
        rhs.append(np.asarray(hh, float).reshape(-1) - F @ off)

    Fx, hx, Fu, hu = mdl["F_x"], mdl["h_x"], mdl["F_u"], mdl["h_u"]
    Fw, hw = mdl["F_w_N"], mdl["h_w_N"]

    def XT(k):
        return np.vstack([Xmap[k], Th]), np.concatenate([Xoff[k], np.zeros(nt)])

    if form == "C":
        if variant == "LBMPC":                       # LBMPC_casadi.m:286-290 (k==1 block first)
            add_rows(mdl["F_x_d"], Xmap[1], Xoff[1], mdl["h_x_d"])
            add_rows(Fw, *XT(1), hw)
        for k in range(N):                           # …casadi.m:278-280
            add_rows(Fx, Xmap[k + 1], Xoff[k + 1], hx)
            add_rows(Fu, Umap[k], Uoff[k], hu)
        if variant == "LMPC":
            add_rows(Fw, *XT(N), hw)                 # …casadi.m:285-286
    else:
        for k in range(N - 1):                       # "if (k<N)" constraintsLMPC.m:21
            if variant == "LBMPC" and k == 0:        # constraintsLBMPC.m:26-31
                add_rows(mdl["F_x_d"], Xmap[1], Xoff[1], mdl["h_x_d"])
                add_rows(Fw, *XT(1), hw)
            add_rows(Fx, Xmap[k + 1], Xoff[k + 1], hx)
            add_rows(Fu, Umap[k], Uoff[k], hu)
        if variant == "LMPC":                        # else-branch reuses x_{N-1} (constraintsLMPC.m:36-38)
            add_rows(Fw, *XT(N - 1), hw)
    G = np.vstack(rows)
    h = np.concatenate(rhs)
    meta = dict(Xmap=Xmap, Xoff=Xoff, Umap=Umap, Uoff=Uoff, n=n, m=m, nt=nt)
    return H, g, c0, G, h, meta


def mehrotra_dense(H, g, G, h, y0=None, tol=1e-9, tol_mu=1e-10, max_iter=60, verbose=False):
    """Dense Mehrotra predictor-corrector, same initial point / sigma / step rule as the
    C oracle (oracle/lbmpc_oracle.c) but on the condensed problem with a Cholesky solve."""
    nv, mrow = H.shape[0], G.shape[0]
    y = np.zeros(nv) if y0 is None else np.asarray(y0, float).copy()
    s = np.maximum(h - G @ y, 1.0)
    lam = np.ones(mrow)
    info = dict(status=1, iters=max_iter)
    for it in range(max_iter):
        rd = H @ y + g + G.T @ lam
        rp = G @ y + s - h
        mu = s @ lam / mrow
        if verbose:
            print(it, np.abs(rd).max(), np.abs(rp).max(), mu)
        if np.abs(rd).max() < tol * max(1.0, 100.0 * lam.max()) and np.abs(rp).max() < tol and mu < tol_mu:
            info = dict(status=0, iters=it)
            break
        w = lam / s
        Lc = np.linalg.cholesky(H + G.T @ (w[:, None] * G))

        def solve(rc):
            rhs = -rd - G.T @ ((-rc + lam * rp) / s)
            dy = np.linalg.solve(Lc.T, np.linalg.solve(Lc, rhs))
            ds = -rp - G @ dy
            dl = (-rc - lam * ds) / s
            return dy, ds, dl

        def amax(ds, dl):
            a = np.inf
            neg = ds < 0
            if neg.any():
                a = min(a, (-s[neg] / ds[neg]).min())
            neg = dl < 0
            if neg.any():
                a = min(a, (-lam[neg] / dl[neg]).min())
            return a

        dya, dsa, dla = solve(s * lam)
        aa = min(1.0, amax(dsa, dla))
        mu_aff = (s + aa * dsa) @ (lam + aa * dla) / mrow
        sigmu = max((mu_aff / mu) ** 2 * mu, 0.1 * tol_mu)      # same floor as lbmpc_oracle.c: no centring below the target gap
        dy, ds, dl = solve(s * lam + dsa * dla - sigmu)
        a = min(1.0, 0.99 * amax(ds, dl))
        y += a * dy
        s += a * ds
        lam += a * dl
    info.update(rd=np.abs(H @ y + g + G.T @ lam).max(), rp=np.abs(G @ y + s - h).max(), mu=s @ lam / mrow)
    return y, s, lam, info
