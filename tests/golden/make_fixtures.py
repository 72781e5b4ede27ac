#!/usr/bin/env python3
"""Generate tests/golden/reference_fixtures.npz from the read-only reference tree.

Run ONCE in the build container (the GPU box has no /root/reference):

    python tests/golden/make_fixtures.py [/root/reference]

Only *data* are extracted (saved .mat trajectories, the terminal set, the
numeric constants of the MATLAB workspace dump `examples/DSS_NMPC.m`); no
reference source code is copied.  Every array name below records the
reference file it came from; the table is repeated in tests/golden/README.md.
"""
import os
import re
import sys

import numpy as np
import scipy.io as sio

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
LB = os.path.join(REF, "matlab", "LBMPC")
DATA = os.path.join(LB, "saved_data+plots", "data")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fixtures.npz")

NCOL = 64  # closed-loop columns kept per trajectory (first step tight, rest loose)


def parse_dump(path):
    """Parse the MATLAB-generated workspace dump (examples/DSS_NMPC.m).

    Joins `...` continuations, then matches `name = [ ... ];` / scalars.
    """
    txt = open(path).read()
    txt = re.sub(r"\.\.\.[^\n]*\n", " ", txt)
    out = {}
    pat = re.compile(r"^([A-Za-z_]\w*) =\s*(\[.*?\]|[-+0-9.Ee]+|-?Inf);", re.S | re.M)
    for m in pat.finditer(txt):
        name, body = m.group(1), m.group(2)
        if body.startswith("["):
            rows = [r for r in body[1:-1].split(";") if r.strip()]
            try:
                mat = [[float(t.replace("Inf", "inf")) for t in r.split()] for r in rows]
            except ValueError:
                continue
            if len({len(r) for r in mat}) != 1:
                continue
            out[name] = np.array(mat, dtype=np.float64)
        else:
            out[name] = np.array(float(body.replace("Inf", "inf")))
    return out


def main():
    fx = {}
    # --- terminal set of the tracking LMPC (getCONS.m:57-58) -------------------
    ts = sio.loadmat(os.path.join(DATA, "term_set.mat"))
    fx["term_set__F_w_N"] = ts["F_w_N"].astype(np.float64)
    fx["term_set__h_w_N"] = ts["h_w_N"].astype(np.float64).reshape(-1)

    # --- workspace dump constants (examples/DSS_NMPC.m) ------------------------
    dump = parse_dump(os.path.join(LB, "examples", "DSS_NMPC.m"))
    keep = ["A", "B", "F_w_N", "F_x", "F_x_d", "F_u", "K_loc", "Klqr", "Kstabil", "LAMBDA",
            "Mtheta", "P", "PSI", "Q", "R", "T", "h_w_N", "h_x", "h_x_d", "h_u", "N",
            "data", "x", "u", "theta", "y_init", "y_OL", "solve_times", "x_eq", "u_eq",
            "x_init", "delta", "con_lb", "q"]
    for k in keep:
        if k in dump:
            fx["dump__" + k] = dump[k]
    # --- F-form closed-loop histories (LMPC_RunExample.m / LBMPC_RunExample.m) --
    for kind in ("LMPC", "LBMPC"):
        for N in (20, 40, 50, 60, 80):
            f = os.path.join(DATA, f"{kind}_N{N}_sys_full.mat")
            if not os.path.exists(f):
                continue
            fx[f"{kind}_N{N}__sysH"] = sio.loadmat(f)["sysH"][:, :NCOL].copy()
            fa = os.path.join(DATA, f"{kind}_N{N}_art_full.mat")
            fx[f"{kind}_N{N}__art_refH"] = sio.loadmat(fa)["art_refH"][:, :NCOL].copy()
    # --- C-form closed-loop absolute states (*_casadi.m scripts) ---------------
    for name, var in (("DMS_N50_tLMPC", "xl"), ("tLMPC", "xl"), ("DMS_tLMPC_K", "xl"),
                      ("DSS_tLMPC", "xl"), ("DMS_N50_tLBMPC_q100", "xlo"),
                      ("DMS_N50_tLBMPC_q10", "xlo"), ("DMS_tLBMPC_q100", "xlo"),
                      ("DMS_tLBMPC", "xlo"), ("tLBMPC", "xlo")):
        f = os.path.join(DATA, "casadi", name + ".mat")
        fx[f"casadi_{name}__{var}"] = sio.loadmat(f)[var][:, :NCOL].copy()
    # --- oracle training data (7 x 500 = [X;Y]) --------------------------------
    fx["casadi_train_data__data"] = sio.loadmat(os.path.join(DATA, "casadi", "train_data.mat"))["data"]
    # --- saved IPOPT per-solve wall times (solve_stats.m) -> BASELINE numbers ---
    st = sio.loadmat(os.path.join(DATA, "casadi", "intelCPU_solve_sample_fullLMPC.mat"))
    fx["casadi_intelCPU_solve_sample_fullLMPC__times"] = np.stack(
        [st[f"solve_times_{i}"].reshape(-1) for i in range(1, 6)])
    np.savez_compressed(OUT, **fx)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(fx), "arrays")
    for k, v in sorted(fx.items()):
        print(f"  {k:50s} {v.shape}")


if __name__ == "__main__":
    main()
