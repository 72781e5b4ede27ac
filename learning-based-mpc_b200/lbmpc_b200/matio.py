"""`.mat` files in the layout the reference's plotting scripts load (saved_data+plots/compare.m:4-6, :28-29,
LMPCvsLBMPC.m:4-5, horizonsLBMPC.m:4-6), written from closed-loop histories of the engine.

    F-form runs (ocpLBMPC / ocpLMPC):     <name>_sys_full.mat -> sysH (n+m) x (steps+1) = [x - x_wp ; u - u_wp]
                                          <name>_art_full.mat -> art_refH m x (steps+1)
    C-form runs (Solver.closed_loop):     <name>.mat          -> xlo (LBMPC) / xl (LMPC): n x steps, ABSOLUTE states
                                                                 (the CasADi scripts save `x`, DMS_tracking_LMPC_casadi.m:178)
Host-side convenience only (scipy.io.savemat); nothing here is on the compute path.
"""
import os

import numpy as np


def save_fform_run(directory, name, sysH, art_refH):
    """LBMPC_N50 -> LBMPC_N50_sys_full.mat {sysH}, LBMPC_N50_art_full.mat {art_refH} (what LMPCvsLBMPC.m:4-5 loads)."""
    from scipy.io import savemat
    os.makedirs(directory, exist_ok=True)
    p1, p2 = os.path.join(directory, f"{name}_sys_full.mat"), os.path.join(directory, f"{name}_art_full.mat")
    savemat(p1, {"sysH": np.asarray(sysH, float)})
    savemat(p2, {"art_refH": np.atleast_2d(np.asarray(art_refH, float))})
    return p1, p2


def save_cform_run(path, x_hist, variant="LBMPC", scenario=0):
    """One scenario of Solver.closed_loop (x_hist: (batch, steps+1, n) or (steps+1, n), absolute) -> `xlo` (LBMPC) or `xl`
    (LMPC), n x steps as compare.m:8-11 indexes it (the CasADi loops store the state BEFORE each step: `steps` columns)."""
    from scipy.io import savemat
    x = np.asarray(x_hist, float)
    if x.ndim == 3:
        x = x[scenario]
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    savemat(path, {"xlo" if variant == "LBMPC" else "xl": np.ascontiguousarray(x[:-1].T)})
    return path
