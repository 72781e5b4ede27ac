"""One-off robustness sweep: GPU engine vs CPU oracle on large random batches of all problem kinds.
Prints, per case, the verdict mismatches, iteration-count differences and the error quantiles.  usage: stress_parity.py [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "learning-based-mpc_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import lbmpc_b200
from lbmpc_b200.dist import sample_initial_states
from oracle_py import OracleProblem

res = {}
cases = [("C", "LBMPC", 50, 65536, 101), ("C", "LMPC", 50, 32768, 102), ("F", "LBMPC", 50, 16384, 103), ("F", "LMPC", 50, 16384, 104),
         ("C", "LBMPC", 200, 4096, 105), ("C", "LMPC", 20, 16384, 106), ("C", "LBMPC", 3, 16384, 107)]
for form, variant, N, nb, seed in cases:
    mdl = lbmpc_b200.moore_greitzer_model(variant)
    X0 = sample_initial_states(nb, seed)
    got = lbmpc_b200.Solver(mdl, form, variant, N, max_batch=nb).solve_batch(X0)
    ref = OracleProblem(form, variant, mdl, N).solve_batch(X0, nthreads=os.cpu_count())
    ok = (got["status"] == 0) & (ref["status"] == 0)
    scale = np.maximum(np.abs(ref["uc"]).reshape(nb, -1).max(1), 1e-3)
    err = np.abs(got["uc"] - ref["uc"]).reshape(nb, -1).max(1) / scale
    eo = np.abs(got["obj"] - ref["obj"]) / np.maximum(1.0, np.abs(ref["obj"]))
    r = {"batch": nb, "status_counts_gpu": np.bincount(got["status"], minlength=4).tolist(),
         "status_mismatch": int((got["status"] != ref["status"]).sum()),
         "iters_diff_gt1": int((np.abs(got["iters"] - ref["iters"]) > 1).sum()),
         "iters_diff_eq1": int((np.abs(got["iters"] - ref["iters"]) == 1).sum()),
         "iters_max": int(got["iters"].max()),
         "u_relerr_p50": float(np.median(err[ok])), "u_relerr_p999": float(np.quantile(err[ok], 0.999)), "u_relerr_max": float(err[ok].max()),
         "frac_u_within_1e-8": float((err[ok] < 1e-8).mean()), "obj_relerr_max": float(eo[ok].max())}
    res[f"{form}_{variant}_N{N}"] = r
    print(form, variant, N, json.dumps(r), flush=True)
out = os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "stress_parity.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(res, open(out, "w"), indent=1)
