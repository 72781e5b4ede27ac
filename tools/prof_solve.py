"""One device-pointer solve call for profiling (ncu): prof_solve.py <kernel> <variant> <N> <batch> [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "learning-based-mpc_b200"))
import torch
import lbmpc_b200
from lbmpc_b200.dist import sample_initial_states

kernel, variant, N, nb = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
s = lbmpc_b200.Solver(lbmpc_b200.moore_greitzer_model(variant), "C", variant, N, device_pointers=True, kernel=kernel, max_batch=nb)
x = torch.from_numpy(sample_initial_states(nb, 0)).cuda()
out = None
for _ in range(reps):
    out = s.solve_batch(x, want_x=False, out=out)
    torch.cuda.synchronize()
print(kernel, s.last_kernel, variant, N, nb, "kernel_ms", s.last_kernel_ms, "iters mean", float(out["iters"].float().mean()))
