"""Throughput vs batch size with the engine's own kernel choice (device-resident I/O, median of 5 launches).
usage: batch_sweep.py [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "learning-based-mpc_b200"))
import numpy as np
import torch
import lbmpc_b200
from lbmpc_b200.dist import sample_initial_states

dev = torch.device("cuda", 0)
res = {}
for variant, N in (("LBMPC", 50), ("LMPC", 50), ("LBMPC", 200)):
    s = lbmpc_b200.Solver(lbmpc_b200.moore_greitzer_model(variant), "C", variant, N, device_pointers=True)
    rows = []
    for nb in (1, 8, 64, 148, 296, 592, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072):
        if N == 200 and nb > 65536:
            continue
        x = torch.from_numpy(sample_initial_states(nb, 0)).to(dev)
        out = s.solve_batch(x, want_x=False)
        ms = []
        for _ in range(5):
            out = s.solve_batch(x, want_x=False, out=out)
            torch.cuda.synchronize()
            ms.append(s.last_kernel_ms)
        m = float(np.median(ms))
        it = out["iters"].cpu().numpy()
        rows.append({"batch": nb, "kernel_ms": m, "qp_per_s": nb / m * 1e3, "iters_mean": float(it.mean()), "iters_max": int(it.max())})
        print(variant, N, rows[-1], flush=True)
    res[f"C_{variant}_N{N}"] = rows
    s.close()
out = os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "batch_sweep.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(res, open(out, "w"), indent=1)
