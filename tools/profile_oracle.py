"""One lbmpc_oracle_apply call of BASELINE configs[3] shape (N=200, q=100) for ncu.  usage: profile_oracle.py [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "learning-based-mpc_b200"))
import numpy as np
import torch
import lbmpc_b200
from lbmpc_b200.dist import sample_initial_states
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N, q = 200, 100
dev = torch.device("cuda", 0)
data = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))["casadi_train_data__data"]
rng = np.random.default_rng(2)
idx = rng.integers(0, data.shape[1] - q, nb)[:, None] + np.arange(q)[None, :]
Xw = torch.from_numpy(np.ascontiguousarray(data[:3][:, idx].transpose(1, 2, 0))).to(dev)
Yw = torch.from_numpy(np.ascontiguousarray(data[3:7][:, idx].transpose(1, 2, 0))).to(dev)
x = torch.from_numpy(sample_initial_states(nb, 2)).to(dev)
du = torch.zeros((nb, N, 1), dtype=torch.float64, device=dev)
s = lbmpc_b200.Solver(lbmpc_b200.moore_greitzer_model("LBMPC"), "C", "LBMPC", N, device_pointers=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e30
for _ in range(4):
    e0.record(); d = s.oracle_apply(x, du, Xw, Yw); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
ms = best
print(f"oracle_apply batch {nb} N {N} q {q}: {ms:.3f} ms, {nb*N*q/ms*1e-6:.1f} G kernel evaluations/s, |d|max {float(d.abs().max()):.3e}")
