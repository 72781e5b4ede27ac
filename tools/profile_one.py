"""One solve call of a named configuration (used under ncu).  usage: profile_one.py FORM VARIANT N BATCH [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "learning-based-mpc_b200"))
import numpy as np
import lbmpc_b200
form, variant, N, nb = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
rng = np.random.default_rng(0)
lo, hi = np.array([-0.40, -0.45, -0.05, -1.0]), np.array([0.10, 0.10, 0.05, 1.0])
dx0 = lo + (hi - lo) * rng.random((nb, 4)); dx0[0] = [-0.35, -0.4, 0, 0]
sol = lbmpc_b200.Solver(lbmpc_b200.moore_greitzer_model(variant), form, variant, N, max_batch=nb)
for _ in range(reps):
    out = sol.solve_batch(dx0)
if os.environ.get("LBMPC_PHASES"):
    sol.phase_cycles(True)
    out = sol.solve_batch(dx0)
    ph = sol.phase_cycles(False)
    it = max(ph.pop("iterations"), 1)
    print("phase cycles/iter (", it, "it):", {k: v // it for k, v in ph.items()})
print(form, variant, N, nb, "kernel_ms", sol.last_kernel_ms, "iters_mean", out["iters"].mean(), "status", np.bincount(out["status"]))
