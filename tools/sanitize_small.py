"""Smallest batch through every thread mapping (compute-sanitizer target): 9 QPs, N=20, C-form LBMPC + F-form LMPC(616 rows)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "learning-based-mpc_b200"))
import numpy as np
import lbmpc_b200
from lbmpc_b200.dist import sample_initial_states

X0 = sample_initial_states(9, 4)
for variant, form in (("LBMPC", "C"), ("LMPC", "F")):
    res = {}
    for kernel in ("warp", "cta", "stream"):
        s = lbmpc_b200.Solver(lbmpc_b200.moore_greitzer_model(variant), form, variant, 20, max_batch=9, kernel=kernel)
        res[kernel] = s.solve_batch(X0)
        s.close()
        print(variant, form, kernel, res[kernel]["status"].tolist(), res[kernel]["iters"].tolist(), flush=True)
    for k in ("cta", "stream"):
        assert np.array_equal(res[k]["status"], res["warp"]["status"]) and np.abs(res[k]["uc"] - res["warp"]["uc"]).max() < 1e-7
print("sanitize_small ok")
