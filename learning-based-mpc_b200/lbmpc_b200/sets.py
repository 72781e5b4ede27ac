"""Host-side polytope computations of the reference WITHOUT the MPT toolbox.

The reference computes its constraint sets offline with MPT3 (`Polyhedron`, `minHRep`, set intersection and equality
inside `compute_MPIS.m`) and ships one of them as data (`saved_data+plots/data/term_set.mat`, 616 rows).  Everything
here is H-representation arithmetic on {x : F x <= h} with one linear programme per row (scipy / HiGHS):

  min_hrep           Polyhedron.minHRep()                      getCONSPOLY.m:27,68 ; getCONS.m:50
  contains / equal   the `==` of compute_MPIS.m:17
  compute_mpis       compute_MPIS.m:1-24  (Gilbert & Tan: intersect {F Aw^i x <= 1} until nothing is added)
  lmpc_terminal_set  getCONS.m:25-50      (the commented-out code that produced term_set.mat)
  lbmpc_terminal_set getCONSPOLY.m:32-69  (robust one-step set: extended constraints (-) disturbance, pdiff.m)

The sets stay offline host work (BASELINE.json north_star); the engine only consumes F_w_N, h_w_N.
"""
import numpy as np
from scipy.optimize import linprog

from .model import _dlqr, pdiff


def _support(c, F, h):
    """max c'x s.t. F x <= h  (inf when unbounded, -inf when the set is empty)."""
    res = linprog(-np.asarray(c, float), A_ub=F, b_ub=h, bounds=[(None, None)] * F.shape[1], method="highs")
    if res.status == 3:
        return np.inf
    if res.status == 2:
        return -np.inf
    if res.status != 0:
        raise RuntimeError(f"linprog failed: {res.message}")
    return -res.fun


def _normalise(F, h):
    """rows scaled to unit 2-norm (keeps the set, makes tolerances comparable)"""
    F, h = np.asarray(F, float), np.asarray(h, float).ravel()
    nrm = np.linalg.norm(F, axis=1)
    keep = nrm > 0
    return F[keep] / nrm[keep, None], h[keep] / nrm[keep]


def min_hrep(F, h, tol=1e-9):
    """Irredundant H-representation of {F x <= h}: duplicate rows merged, then every row whose support over the
    OTHER rows does not exceed its right-hand side is dropped (one LP per row).  Row order is kept."""
    Fn, hn = _normalise(F, h)
    # merge duplicates (same direction): keep the tighter bound
    order = np.lexsort(np.round(Fn, 10).T[::-1])
    keep = np.ones(len(hn), bool)
    for a, b in zip(order[:-1], order[1:]):
        if keep[a] and np.allclose(Fn[a], Fn[b], atol=1e-10, rtol=0):
            drop = a if hn[a] >= hn[b] else b
            keep[drop] = False
    idx = np.flatnonzero(keep)
    alive = np.zeros(len(hn), bool)
    alive[idx] = True
    for i in idx:
        alive[i] = False
        if not alive.any() or _support(Fn[i], Fn[alive], hn[alive]) > hn[i] + tol:
            alive[i] = True
    F, h = np.asarray(F, float), np.asarray(h, float).ravel()
    nz = np.flatnonzero(np.linalg.norm(F, axis=1) > 0)
    sel = nz[alive]
    return F[sel].copy(), h[sel].copy()


def contains(F_out, h_out, F_in, h_in, tol=1e-8):
    """{F_in x <= h_in} is a subset of {F_out x <= h_out} (up to tol on unit-norm rows)."""
    Fo, ho = _normalise(F_out, h_out)
    Fi, hi = np.asarray(F_in, float), np.asarray(h_in, float).ravel()
    return all(_support(Fo[i], Fi, hi) <= ho[i] + tol for i in range(len(ho)))


def equal(F1, h1, F2, h2, tol=1e-8):
    return contains(F1, h1, F2, h2, tol) and contains(F2, h2, F1, h1, tol)


def compute_mpis(F, h, Aw, max_iter=1000, tol=1e-9, verbose=False):
    """Maximal positively invariant set of x+ = Aw x inside {F x <= h} (compute_MPIS.m): rows F Aw^i, right-hand side h,
    added level by level while a level still cuts the set; returns (F_mpi, h_mpi, levels)."""
    F, h = _normalise(F, h)
    Fs, hs = F.copy(), h.copy()
    Fi = F.copy()
    for level in range(1, max_iter + 1):
        Fi = Fi @ Aw
        nrm = np.linalg.norm(Fi, axis=1)
        added = 0
        for r in range(len(h)):
            if nrm[r] < 1e-14:
                continue
            row, rhs = Fi[r] / nrm[r], h[r] / nrm[r]
            if _support(row, Fs, hs) > rhs + tol:
                Fs = np.vstack([Fs, row])
                hs = np.append(hs, rhs)
                added += 1
        if verbose:
            print(f"level {level}: +{added} rows ({len(hs)})")
        if added == 0:
            return Fs, hs, level
    raise RuntimeError("compute_mpis: no fixed point within max_iter levels")


def _extended_rows(F_x, h_x, F_u, h_u, K, LAMBDA, PSI, LAMBDA_0, PSI_0, lam):
    n, m = F_x.shape[1], F_u.shape[1]
    L = PSI - K @ LAMBDA
    L0 = PSI_0 - K @ LAMBDA_0
    F_w = np.block([[F_x, np.zeros((len(h_x), m))],
                    [np.zeros((len(h_x), n)), F_x @ LAMBDA],
                    [F_u @ K, F_u @ L],
                    [np.zeros((len(h_u), n)), F_u @ PSI]])
    h_w = np.concatenate([h_x, lam * (h_x - (F_x @ LAMBDA_0).ravel()), h_u - (F_u @ L0).ravel(),
                          lam * (h_u - (F_u @ PSI_0).ravel())])
    return F_w, h_w, L, L0


def lmpc_terminal_set(A, B, K, LAMBDA, PSI, F_x, h_x, F_u, h_u, LAMBDA_0=None, PSI_0=None, lam=0.99, unit_rhs=True):
    """Terminal set of tracking LMPC in the extended state w = [x; theta] (getCONS.m:25-50): MPIS of
    w+ = [[A+BK, B L], [0, I]] w inside the extended admissible set, irredundant.  With unit_rhs the rows are scaled
    to h = 1 like the shipped term_set.mat (compute_MPIS.m:6-8)."""
    n, m = B.shape
    LAMBDA_0 = np.zeros((n, 1)) if LAMBDA_0 is None else LAMBDA_0
    PSI_0 = np.zeros((m, 1)) if PSI_0 is None else PSI_0
    F_w, h_w, L, _ = _extended_rows(F_x, h_x, F_u, h_u, K, LAMBDA, PSI, LAMBDA_0, PSI_0, lam)
    Aw = np.block([[A + B @ K, B @ L], [np.zeros((m, n)), np.eye(m)]])
    Fs, hs, _ = compute_mpis(F_w, h_w, Aw)
    Fs, hs = min_hrep(Fs, hs)
    if unit_rhs:
        Fs, hs = Fs / hs[:, None], np.ones_like(hs)
    return Fs, hs


def lbmpc_terminal_set(A, B, Q, R, LAMBDA, PSI, F_x, h_x, F_u, h_u, state_uncert, LAMBDA_0=None, PSI_0=None, lam=0.99,
                       control_weight=10.0):
    """[F_w_N, h_w_N, F_x_d, h_x_d] of getCONSPOLY.m:25-69 for any uncertainty box |d_j| <= state_uncert_j."""
    n, m = B.shape
    LAMBDA_0 = np.zeros((n, 1)) if LAMBDA_0 is None else LAMBDA_0
    PSI_0 = np.zeros((m, 1)) if PSI_0 is None else PSI_0
    w = np.asarray(state_uncert, float).ravel()
    F_d, h_d = np.vstack([np.eye(n), -np.eye(n)]), np.concatenate([w, w])
    F_x_d, h_x_d = min_hrep(*pdiff(F_x, h_x, F_d, h_d))                      # X (-) D  (:25-28)
    K_t = -_dlqr(A, B, Q, control_weight * R)                                # :35-36
    F_w, h_w, L, L0 = _extended_rows(F_x, h_x, F_u, h_u, K_t, LAMBDA, PSI, LAMBDA_0, PSI_0, lam)
    F_w = np.vstack([F_w, np.hstack([F_x_d @ (A + B @ K_t), F_x_d @ B @ L])])            # :45-49
    h_w = np.concatenate([h_w, h_x_d - (F_x_d @ B @ L0).ravel()])                         # :50-54
    F_d_w = np.block([[F_d, np.zeros((2 * n, m))], [np.zeros((m, n)), np.eye(m)], [np.zeros((m, n)), -np.eye(m)]])
    h_d_w = np.concatenate([h_d, np.zeros(2 * m)])                                        # :57-61
    F0, h0 = pdiff(F_w, h_w, F_d_w, h_d_w)                                                # :64
    F_w_N, h_w_N = min_hrep(F0, h0)                                                       # :66-68
    return F_w_N, h_w_N, F_x_d, h_x_d
