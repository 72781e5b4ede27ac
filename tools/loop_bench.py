"""Monte-Carlo closed loop throughput: fused (stream) vs per-step launches.  usage: loop_bench.py <scenarios> <steps> [kernels]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "learning-based-mpc_b200"))
import numpy as np
import torch
import lbmpc_b200
from lbmpc_b200.dist import sample_initial_states
nb, T = int(sys.argv[1]), int(sys.argv[2])
kernels = sys.argv[3].split(",") if len(sys.argv) > 3 else ["auto", "stream"]
x_init = torch.from_numpy(lbmpc_b200.X_WP[None, :] + sample_initial_states(nb, 3)).cuda()
wbar = np.array([0.02, 5e-4, 0.0, 0.0])
for k in kernels:
    s = lbmpc_b200.Solver(lbmpc_b200.moore_greitzer_model("LBMPC"), "C", "LBMPC", 50, device_pointers=True, max_batch=nb, kernel=k)
    s.closed_loop(x_init[:2048], 2, lbmpc_b200.X_WP, float(lbmpc_b200.U_WP), q=100, use_oracle=True, wbar=wbar, seed=7)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    h = s.closed_loop(x_init, T, lbmpc_b200.X_WP, float(lbmpc_b200.U_WP), q=100, use_oracle=True, wbar=wbar, seed=7)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = h["status"].cpu().numpy()
    print(k, s.last_kernel, "scenarios", nb, "steps", T, "wall_s %.3f" % dt, "QP/s %.0f" % (nb * T / dt), "iters mean %.2f" % float(h["iters"].float().mean()),
          "status", np.bincount(st.ravel(), minlength=4).tolist(), "chunk", os.environ.get("LBMPC_LOOP_CHUNK"), flush=True)
    s.close()
