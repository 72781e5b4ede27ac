"""Host-side set computations without MPT (lbmpc_b200.sets) against the sets the reference ships:
the 16-row robust terminal set of its LBMPC run (examples/DSS_NMPC.m dump) and the 616-row tracking terminal set
(saved_data+plots/data/term_set.mat), both committed as data in lbmpc_b200/ (tests/golden/make_fixtures.py)."""
import numpy as np

import lbmpc_b200
from lbmpc_b200 import sets
from lbmpc_b200.model import getCONS, getCONSPOLY
from oracle_py import OracleProblem


def _rows_match(F1, h1, F2, h2, tol):
    """same irredundant H-representation up to row order and scaling"""
    a = np.hstack([F1, h1[:, None]]) / np.linalg.norm(F1, axis=1)[:, None]
    b = np.hstack([F2, h2[:, None]]) / np.linalg.norm(F2, axis=1)[:, None]
    if a.shape != b.shape:
        return False
    d = np.abs(a[:, None, :] - b[None, :, :]).max(axis=2)
    return d.min(axis=1).max() < tol and d.min(axis=0).max() < tol


def test_min_hrep_and_containment():
    F = np.array([[1.0, 0], [-1, 0], [0, 1], [0, -1], [1, 1], [2, 0], [1, 0]])
    h = np.array([1.0, 1, 1, 1, 3, 2, 5])                       # rows 4..6: redundant / duplicate directions
    Fm, hm = sets.min_hrep(F, h)
    assert Fm.shape == (4, 2) and sets.equal(Fm, hm, F[:4], h[:4])
    assert sets.contains(F[:4], 2 * h[:4], Fm, hm) and not sets.contains(Fm, hm, F[:4], 2 * h[:4])
    Fc, hc = sets.min_hrep(np.vstack([F[:4], [[1.0, 1.0]]]), np.append(h[:4], 1.5))   # a cutting row stays
    assert Fc.shape == (5, 2)


def test_lbmpc_robust_sets_regenerated_and_other_uncertainty_bounds():
    ref = getCONSPOLY()                                          # shipped = the reference's own MPT output
    F_x, h_x, F_u, h_u = ref[:4]
    A, B, C, _, _ = lbmpc_b200.mgcmDLTI()
    _, _, Q, R, _, _, _, LAMBDA, PSI, L0, P0 = lbmpc_b200.matOCP(A, B, C)
    F_w, h_w, F_xd, h_xd = sets.lbmpc_terminal_set(A, B, Q, R, LAMBDA, PSI, F_x, h_x, F_u, h_u, (0.02, 5e-4, 0.0, 0.0), L0, P0)
    assert F_w.shape == ref[4].shape == (16, 5)
    assert sets.equal(F_w, h_w, ref[4], ref[5], 1e-6) and _rows_match(F_w, h_w, ref[4], ref[5], 1e-5)
    assert _rows_match(F_xd, h_xd, ref[6], ref[7], 1e-12)
    small = getCONSPOLY(state_uncert=(0.01, 2.5e-4, 0.0, 0.0))   # half the uncertainty: a larger robust set
    assert sets.contains(small[4], small[5], ref[4], ref[5]) and not sets.contains(ref[4], ref[5], small[4], small[5])
    mdl = lbmpc_b200.moore_greitzer_model("LBMPC")
    mdl.update(F_w_N=small[4], h_w_N=small[5], F_x_d=small[6], h_x_d=small[7])
    out = OracleProblem("C", "LBMPC", mdl, 20).solve(np.array([-0.35, -0.4, 0.0, 0.0]))
    assert out["status"] == 0


def test_double_integrator_invariant_terminal_set():
    """trackingMPC/RunExample.m:99-108: compute_MPIS of the extended admissible set under [A+BK, BL; 0, I]."""
    m0, m1 = lbmpc_b200.double_integrator_model(), lbmpc_b200.double_integrator_model(invariant=True)
    F, h = m1["F_w_N"], m1["h_w_N"]
    assert sets.contains(m0["F_w_N"], m0["h_w_N"], F, h) and not sets.contains(F, h, m0["F_w_N"], m0["h_w_N"])
    K, L = m1["K"], m1["PSI"] - m1["K"] @ m1["LAMBDA"]
    Ak = np.block([[m1["A"] + m1["B"] @ K, m1["B"] @ L], [np.zeros((2, 2)), np.eye(2)]])
    assert sets.contains(F, h, F @ np.linalg.inv(Ak), h, 1e-7)  # image of the set under Ak stays inside: invariant
    out = OracleProblem("C", "LMPC", m1, 5).solve(np.array([1.0, -0.2]), np.array([0.5, 0.0]))
    assert out["status"] == 0


def test_tracking_terminal_set_616_rows_regenerated():
    """getCONS.m:30-50 (commented out in the reference, result shipped as term_set.mat): the invariant-set iteration
    reproduces the shipped 616-row set row for row."""
    ref = getCONS()
    new = getCONS(recompute=True)
    assert new[4].shape == ref[4].shape == (616, 5)
    assert _rows_match(new[4], new[5], ref[4], ref[5], 1e-6)
