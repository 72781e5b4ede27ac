"""Throughput of each thread mapping (warp / cta / stream / mixed) vs batch size, device-resident I/O, median of 5
launches (CUDA events around the solve kernel).  usage: kernel_sweep.py [out.json] [cases]
cases: comma list of variant:N, default LBMPC:50,LMPC:50,LBMPC:200"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "learning-based-mpc_b200"))
import numpy as np
import torch
import lbmpc_b200
from lbmpc_b200.dist import sample_initial_states

dev = torch.device("cuda", 0)
peak = lbmpc_b200.measure_fp64_peak(0)
res = {"fp64_peak_tflops": peak, "stream_ctas_env": os.environ.get("LBMPC_STREAM_CTAS")}
cases = (sys.argv[2] if len(sys.argv) > 2 else "LBMPC:50,LMPC:50,LBMPC:200").split(",")
for case in cases:
    variant, N = case.split(":")
    N = int(N)
    ng = 24 if variant == "LBMPC" else 616
    rows = []
    for kernel in ("auto", "warp", "cta", "stream", "mixed"):
        for nb in (1024, 4096, 16384, 65536, 262144):
            if kernel in ("warp", "cta") and nb > 65536:
                continue
            if kernel == "cta" and nb > 16384:
                continue
            if N == 200 and nb > 65536:
                continue
            s = lbmpc_b200.Solver(lbmpc_b200.moore_greitzer_model(variant), "C", variant, N, device_pointers=True, kernel=kernel,
                                  max_batch=nb)
            x = torch.from_numpy(sample_initial_states(nb, 0)).to(dev)
            out = s.solve_batch(x, want_x=False)
            ms = []
            for _ in range(5):
                out = s.solve_batch(x, want_x=False, out=out)
                torch.cuda.synchronize()
                ms.append(s.last_kernel_ms)
            m = float(np.median(ms))
            it = out["iters"].cpu().numpy()
            fl = float(it.sum()) * (1355.0 * N + 108.0 * ng)
            rows.append({"kernel": kernel, "used": s.last_kernel, "batch": nb, "kernel_ms": m, "qp_per_s": nb / m * 1e3,
                         "iters_mean": float(it.mean()), "iters_max": int(it.max()),
                         "roofline_frac": fl / (m * 1e-3) * 1e-12 / peak})
            print(variant, N, rows[-1], flush=True)
            s.close()
            del x, out
    res[f"C_{variant}_N{N}"] = rows
out = os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "kernel_sweep.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(res, open(out, "w"), indent=1)
