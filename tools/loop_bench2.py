import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "learning-based-mpc_b200"))
import numpy as np, torch, lbmpc_b200
from lbmpc_b200.dist import sample_initial_states
nb, T = int(sys.argv[1]), int(sys.argv[2])
x_init = torch.from_numpy(lbmpc_b200.X_WP[None, :] + sample_initial_states(nb, 3)).cuda()
wbar = np.array([0.02, 5e-4, 0.0, 0.0])
for k, orc, w in (("stream", False, None), ("stream", False, wbar), ("stream", True, wbar), ("auto", False, None), ("auto", True, wbar)):
    s = lbmpc_b200.Solver(lbmpc_b200.moore_greitzer_model("LBMPC"), "C", "LBMPC", 50, device_pointers=True, max_batch=nb, kernel=k)
    s.closed_loop(x_init[:2048], 2, lbmpc_b200.X_WP, float(lbmpc_b200.U_WP), q=100, use_oracle=orc, wbar=w, seed=7)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h = s.closed_loop(x_init, T, lbmpc_b200.X_WP, float(lbmpc_b200.U_WP), q=100, use_oracle=orc, wbar=w, seed=7)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    it = h["iters"].float()
    print(k, "oracle", orc, "w", w is not None, "QP/s %.0f" % (nb * T / dt), "QP-iter/s %.0f" % (float(it.sum()) / dt), "iters mean %.2f max %d" % (float(it.mean()), int(it.max())),
          "status", np.bincount(h["status"].cpu().numpy().ravel(), minlength=4).tolist(), flush=True)
    s.close()
x = torch.from_numpy(sample_initial_states(nb, 3)).cuda()
for k in ("stream", "warp"):
    s = lbmpc_b200.Solver(lbmpc_b200.moore_greitzer_model("LBMPC"), "C", "LBMPC", 50, device_pointers=True, max_batch=nb, kernel=k)
    o = s.solve_batch(x, want_x=False); o = s.solve_batch(x, want_x=False, out=o); torch.cuda.synchronize()
    print("solve only", k, nb, "ms %.2f" % s.last_kernel_ms, "QP/s %.0f" % (nb / s.last_kernel_ms * 1e3), "QP-iter/s %.0f" % (float(o["iters"].sum()) / s.last_kernel_ms * 1e3))
