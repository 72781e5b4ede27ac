"""Scratch driver for early GPU bring-up: parity + timing of a few configurations."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "learning-based-mpc_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import lbmpc_b200
from oracle_py import OracleProblem

rng = np.random.default_rng(0)
lo, hi = np.array([-0.40, -0.45, -0.05, -1.0]), np.array([0.10, 0.10, 0.05, 1.0])
for form, variant, N, nb in (("C", "LBMPC", 50, 1024), ("C", "LMPC", 50, 1024), ("F", "LMPC", 50, 256),
                             ("F", "LBMPC", 50, 256), ("C", "LBMPC", 200, 512), ("C", "LBMPC", 50, 16384)):
    mdl = lbmpc_b200.moore_greitzer_model(variant)
    dx0 = lo + (hi - lo) * rng.random((nb, 4)); dx0[0] = [-0.35, -0.4, 0, 0]
    sol = lbmpc_b200.Solver(mdl, form, variant, N, max_batch=nb)
    t = time.time(); got = sol.solve_batch(dx0); t_gpu = time.time() - t
    ms = sol.last_kernel_ms
    for _ in range(3):
        sol.solve_batch(dx0)
    ms2 = sol.last_kernel_ms
    nref = min(nb, 1024)
    t = time.time(); ref = OracleProblem(form, variant, mdl, N).solve_batch(dx0[:nref], nthreads=os.cpu_count()); t_cpu = time.time() - t
    st_eq = (got["status"][:nref] == ref["status"]).mean()
    dit = np.abs(got["iters"][:nref] - ref["iters"]).max()
    ok = (ref["status"] == 0) & (got["status"][:nref] == 0)
    err = np.abs(got["uc"][:nref][ok] - ref["uc"][ok]).max() / np.abs(ref["uc"][ok]).max()
    ej = (np.abs(got["obj"][:nref][ok] - ref["obj"][ok]) / np.maximum(1, np.abs(ref["obj"][ok]))).max()
    print(f"{form}-{variant} N={N} batch={nb} slots={sol.slots_per_cta}: status_eq={st_eq:.4f} d_iters={dit} "
          f"rel_err_u={err:.2e} rel_err_J={ej:.2e} iters_mean={got['iters'].mean():.2f} "
          f"kernel_ms first={ms:.3f} warm={ms2:.3f} -> {nb/ms2*1e3:.0f} QP/s ; host call {t_gpu*1e3:.1f} ms ; "
          f"oracle {nref} QPs {t_cpu*1e3:.0f} ms ({os.cpu_count()} threads)", flush=True)
    sol.close()
