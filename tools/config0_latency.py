"""BASELINE configs[0] / SURVEY 8d config 1: the reference's own closed-loop run (LBMPC_RunExample.m defaults: F-form LBMPC, N = 50,
dx_init = [-0.35;-0.4;0;0], q = 100, one QP per control step) through the Python mirror of ocpLBMPC.m on the GPU engine — per-step
latency and the agreement with the saved history LBMPC_N50_sys_full.mat (fixture copy; nothing here reads /root/reference).
usage: config0_latency.py [steps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "learning-based-mpc_b200"))
import numpy as np
import lbmpc_b200
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
fx = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))
X_EQ, U_EQ, DX0 = np.array([0.5, 1.6875, 1.1547, 0.0]), 1.1547, np.array([-0.35, -0.4, 0.0, 0.0])
out = {}
for variant in ("LBMPC", "LMPC"):
    mdl = lbmpc_b200.moore_greitzer_model(variant)
    mats = [mdl[k] for k in ("K", "Q", "R", "P", "T")] + [np.vstack([mdl["LAMBDA"], mdl["PSI"]]), mdl["LAMBDA"], mdl["PSI"], 1]
    rows = [mdl[k] for k in ("F_x", "h_x", "F_u", "h_u", "F_w_N", "h_w_N")]
    hist0 = [np.concatenate([DX0, [0.0]]).reshape(5, 1), np.zeros((1, 1)), np.zeros((4, 1))]
    info = {}
    for rep in range(2):   # first run warms the context up
        t0 = time.perf_counter()
        if variant == "LBMPC":
            sysH, artH, _ = lbmpc_b200.ocpLBMPC(X_EQ + DX0, X_EQ, DX0, np.zeros(4), U_EQ, 50, 0.01, steps, None, np.zeros(51),
                                                {"X": np.zeros((3, 1)), "Y": np.zeros((4, 1))}, mdl["A"], mdl["B"], *mats, *rows,
                                                mdl["F_x_d"], mdl["h_x_d"], *hist0, info=info)
        else:
            sysH, artH, _ = lbmpc_b200.ocpLMPC(X_EQ + DX0, DX0, X_EQ, np.zeros(4), U_EQ, 50, 0.01, steps, None, np.zeros(51), *mats, *rows,
                                               *hist0, A=mdl["A"], B=mdl["B"], info=info)
        dt = time.perf_counter() - t0
    ref = fx[f"{variant}_N50__sysH"]
    n = min(steps + 1, ref.shape[1], 40)
    err = np.abs(sysH[:, :n] - ref[:, :n])
    out[variant] = {"steps": steps, "ms_per_control_step": 1e3 * dt / steps, "status_all_optimal": bool((info["status"] == 0).all()),
                    "ipm_iterations_mean": float(info["iters"].mean()), "first_input_abs_err": float(abs(sysH[4, 1] - ref[4, 1])),
                    "input_abs_err_first_40_steps": float(err[4].max()), "state_abs_err_first_40_steps": float(err[:4].max()),
                    "reference_solver_s_per_step": "fmincon SQP / IPOPT: 0.096 s median per solve (BASELINE.md)"}
    print(variant, json.dumps(out[variant]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_config0_latency.json"), "w"), indent=1)
