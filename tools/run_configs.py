"""Measure the five BASELINE.json configs on ONE B200 (per-GPU share of the multi-GPU ones) and write
profiles/<out>.json.  Timing: CUDA events around the solve kernel (lbmpc_last_kernel_ms), device-resident
inputs, median of `reps` launches after 2 warm-ups.  usage: python tools/run_configs.py [out.json]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "learning-based-mpc_b200"))
import numpy as np
import torch
import lbmpc_b200
from lbmpc_b200.dist import sample_initial_states

dev = torch.device("cuda", 0)
X_EQ, U_EQ = lbmpc_b200.X_WP, float(lbmpc_b200.U_WP)
peak = lbmpc_b200.measure_fp64_peak(0)
res = {"fp64_peak_tflops_measured": peak}


def flops_per_iter(N, n_g):
    return 1355.0 * N + 108.0 * n_g


def timed_solve(sol, args, reps=7, **kw):
    for _ in range(2):
        out = sol.solve_batch(*args, want_x=False, **kw)
    ms = []
    for _ in range(reps):
        out = sol.solve_batch(*args, want_x=False, out=out, **kw)
        torch.cuda.synchronize()
        ms.append(sol.last_kernel_ms)
    return out, float(np.median(ms))


def summarise(name, out, ms, nb, N, n_g, extra=None):
    it = out["iters"].cpu().numpy()
    st = out["status"].cpu().numpy()
    fl = float(it.sum()) * flops_per_iter(N, n_g)
    r = {"batch": nb, "horizon": N, "kernel_ms": ms, "qp_per_s": nb / ms * 1e3, "us_per_qp": 1e3 * ms / nb,
         "iters_mean": float(it.mean()), "iters_max": int(it.max()), "status_counts": np.bincount(st, minlength=4).tolist(),
         "fp64_tflops": fl / (ms * 1e-3) * 1e-12, "roofline_frac": fl / (ms * 1e-3) * 1e-12 / peak}
    if extra:
        r.update(extra)
    res[name] = r
    print(name, json.dumps(r), flush=True)


def dev_solver(variant, form, N):
    return lbmpc_b200.Solver(lbmpc_b200.moore_greitzer_model(variant), form, variant, N, device_pointers=True)


# ---- config 1: the reference closed loop, one QP per step (LBMPC_RunExample.m defaults; RK4 plant) ----
mdl = lbmpc_b200.moore_greitzer_model("LBMPC")
hs = lbmpc_b200.Solver(mdl, "F", "LBMPC", 50, max_batch=1)
dx = np.array([[-0.35, -0.4, 0.0, 0.0]])
warm = None
t0 = time.perf_counter()
lat = []
for k in range(200):
    t1 = time.perf_counter()
    o = hs.solve_batch(dx, warm=warm)
    lat.append(time.perf_counter() - t1)
    warm = np.concatenate([o["uc"].reshape(1, -1), o["theta"].reshape(1, -1)], axis=1)
    u = float(mdl["K"].reshape(-1) @ dx[0] + o["uc"].reshape(-1)[0])
    dx = (dx @ mdl["A"].T + u * mdl["B"].reshape(1, -1))           # nominal plant keeps the loop self-contained
res["config1_fform_lbmpc_closed_loop_batch1"] = {
    "steps": 200, "host_call_us_p50": 1e6 * float(np.median(lat)), "host_call_us_p95": 1e6 * float(np.percentile(lat, 95)),
    "kernel_us_last": 1e3 * hs.last_kernel_ms, "iters_last": int(o["iters"][0]), "note": "F-form LBMPC N=50, warm start = previous opt_var (ocpLBMPC.m:31), host pointers: one synchronous C-ABI call per control step"}
print("config1", json.dumps(res["config1_fform_lbmpc_closed_loop_batch1"]), flush=True)
hs.close()
cl = lbmpc_b200.Solver(mdl, "C", "LBMPC", 50, max_batch=1)
t1 = time.perf_counter()
h = cl.closed_loop(np.array([[0.15, 1.2875, 1.1547, 0.0]]), 500, X_EQ, U_EQ, q=100, use_oracle=False)
dt = time.perf_counter() - t1
res["config1_cform_lbmpc_closed_loop_batch1"] = {"steps": 500, "wall_s": dt, "us_per_step": 1e6 * dt / 500,
                                                 "final_state_minus_eq": (h["x"][0, -1] - X_EQ).tolist(),
                                                 "note": "LBMPC_casadi.m loop (500 steps, RK4 plant, warm-start shift) in one lbmpc_closed_loop call"}
print("config1c", json.dumps(res["config1_cform_lbmpc_closed_loop_batch1"]), flush=True)
cl.close()

# ---- config 2: batch 1024, both polytope sets ----
for variant, n_g in (("LBMPC", 24), ("LMPC", 616)):
    s = dev_solver(variant, "C", 50)
    x = torch.from_numpy(sample_initial_states(1024, 0)).to(dev)
    out, ms = timed_solve(s, (x,))
    summarise(f"config2_cform_{variant.lower()}_b1024", out, ms, 1024, 50, n_g)
    s.close()

# ---- config 3: tracking LMPC, terminal invariant set, batch 16384 ----
mdl = lbmpc_b200.moore_greitzer_model("LMPC")
s = dev_solver("LMPC", "C", 50)
nb = 16384
rng = np.random.default_rng(1)
xref = mdl["LAMBDA"][:, 0][None, :] * rng.uniform(-0.1, 0.1, (nb, 1)) * (rng.random((nb, 1)) < 0.5)
x = torch.from_numpy(sample_initial_states(nb, 1)).to(dev)
out, ms = timed_solve(s, (x, torch.from_numpy(xref).to(dev)))
summarise("config3_tracking_lmpc_b16384", out, ms, nb, 50, 616)
s.close()

# ---- config 4: long horizon N=200 with learned-oracle offsets, batch 65536 (whole batch on one GPU) ----
fx = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))
data = fx["casadi_train_data__data"]
s = dev_solver("LBMPC", "C", 200)
nb, q = 65536, 100
rng = np.random.default_rng(2)
offs = rng.integers(0, data.shape[1] - q, nb)
idx = offs[:, None] + np.arange(q)[None, :]
Xw = torch.from_numpy(np.ascontiguousarray(data[:3][:, idx].transpose(1, 2, 0))).to(dev)   # (nb, q, 3)
Yw = torch.from_numpy(np.ascontiguousarray(data[3:7][:, idx].transpose(1, 2, 0))).to(dev)
x = torch.from_numpy(sample_initial_states(nb, 2)).to(dev)
du = torch.zeros((nb, 200, 1), dtype=torch.float64, device=dev)                              # u = u_eq along the rollout
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
d_off = s.oracle_apply(x, du, Xw, Yw)
e0.record()
d_off = s.oracle_apply(x, du, Xw, Yw)
e1.record()
torch.cuda.synchronize()
oracle_ms = e0.elapsed_time(e1)
out, ms = timed_solve(s, (x,), reps=3, d_off=d_off)
summarise("config4_lbmpc_n200_oracle_b65536", out, ms, nb, 200, 24,
          {"oracle_apply_ms": oracle_ms, "oracle_gexp_per_s": nb * 200 * q / (oracle_ms * 1e-3) * 1e-9,
           "d_off_absmax": float(d_off.abs().max())})
s.close()
del Xw, Yw, d_off

# ---- config 5: Monte-Carlo closed loop, per-GPU share of 1 M scenarios (125 k), T = 20 steps ----
nb, T = 125000, 100
s = lbmpc_b200.Solver(lbmpc_b200.moore_greitzer_model("LBMPC"), "C", "LBMPC", 50, max_batch=nb)
x_init = X_EQ[None, :] + sample_initial_states(nb, 3)
wbar = np.array([0.02, 5e-4, 0.0, 0.0])
s.closed_loop(x_init[:1024], 2, X_EQ, U_EQ, q=100, use_oracle=True, wbar=wbar, seed=7)      # warm-up
t1 = time.perf_counter()
h = s.closed_loop(x_init, T, X_EQ, U_EQ, q=100, use_oracle=True, wbar=wbar, seed=7)
dt = time.perf_counter() - t1
it = h["iters"]
st = h["status"]
res["config5_montecarlo_closed_loop_125k_per_gpu"] = {
    "scenarios": nb, "steps": T, "wall_s": dt, "qp_per_s": nb * T / dt, "iters_mean": float(it.mean()),
    "status_counts": np.bincount(st.ravel(), minlength=4).tolist(),
    "note": "solve + L2NW oracle (q=100) + RK4 plant + disturbance + window update per step, all on the GPU; host call incl. H2D of x_init and D2H of the histories"}
print("config5", json.dumps(res["config5_montecarlo_closed_loop_125k_per_gpu"]), flush=True)
s.close()

out_path = os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "configs.json")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
json.dump(res, open(out_path, "w"), indent=1)
