"""Single-QP latency through the C ABI (host pointers): time of the lbmpc_solve_batch call itself vs the kernel time, for pageable
and page-locked caller arrays.  usage: python tools/latency_one.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'learning-based-mpc_b200'))
import lbmpc_b200
from lbmpc_b200.capi import _ptr
mdl = lbmpc_b200.moore_greitzer_model("LBMPC")
for pinned in (False, True):
    hs = lbmpc_b200.Solver(mdl, "C", "LBMPC", 50, max_batch=1)
    if pinned:
        import torch
        mk = lambda shape, dt: torch.zeros(shape, dtype=torch.float64 if dt is np.float64 else torch.int32).pin_memory().numpy()
    else:
        mk = lambda shape, dt: np.zeros(shape, dt)
    dx = mk((1,4), np.float64); dx[:] = [-0.35,-0.4,0,0]; warm = mk((1,51), np.float64)
    uc, th, obj = mk((1,50), np.float64), mk((1,1), np.float64), mk((1,), np.float64)
    it, st = mk((1,), np.int32), mk((1,), np.int32)
    args = (hs.h, 1, _ptr(dx), None, None, _ptr(warm), _ptr(uc), _ptr(th), None, _ptr(obj), _ptr(it), _ptr(st), None)
    lat=[]; ker=[]
    for k in range(400):
        t1=time.perf_counter(); rc = hs.lib.lbmpc_solve_batch(*args); lat.append(time.perf_counter()-t1)
        ker.append(hs.last_kernel_ms)
        warm[0,:50]=uc[0]; warm[0,50]=th[0,0]
    print("pinned" if pinned else "pageable", "C-ABI call us p50 %.1f p95 %.1f ; kernel us p50 %.1f ; overhead %.1f" % (1e6*np.median(lat[50:]), 1e6*np.percentile(lat[50:],95), 1e3*np.median(ker[50:]), 1e6*np.median(lat[50:])-1e3*np.median(ker[50:])))
