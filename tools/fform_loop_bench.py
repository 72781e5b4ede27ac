"""F-form closed loop on the device (ocpLBMPC.m / ocpLMPC.m for a batch of scenarios): throughput.  usage: fform_loop_bench.py <scenarios> <steps>"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "learning-based-mpc_b200"))
import numpy as np
import torch
import lbmpc_b200
from lbmpc_b200.dist import sample_initial_states
nb, T = int(sys.argv[1]), int(sys.argv[2])
x_init = torch.from_numpy(lbmpc_b200.X_WP[None, :] + 0.3 * sample_initial_states(nb, 3)).cuda()
for variant in ("LBMPC", "LMPC"):
    s = lbmpc_b200.Solver(lbmpc_b200.moore_greitzer_model(variant), "F", variant, 50, device_pointers=True, max_batch=nb)
    s.closed_loop(x_init, 2, lbmpc_b200.X_WP, float(lbmpc_b200.U_WP), q=100, use_oracle=True)   # same width: every scratch buffer is sized before the timed call
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    h = s.closed_loop(x_init, T, lbmpc_b200.X_WP, float(lbmpc_b200.U_WP), q=100, use_oracle=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = h["status"].cpu().numpy()
    print("F-form", variant, "scenarios", nb, "steps", T, "wall_s %.3f" % dt, "control steps/s %.0f" % (nb * T / dt), "last kernel", s.last_kernel,
          "IPM iterations of a step's last QP, mean %.2f" % float(h["iters"].float().mean()),
          "status", np.bincount(st.ravel(), minlength=4).tolist(), flush=True)
    s.close()
