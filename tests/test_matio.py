"""`.mat` writer: files have the variable names, shapes and dtype of the reference's saved runs (what compare.m / LMPCvsLBMPC.m
load), checked against the fixture copies of those runs."""
import numpy as np

import lbmpc_b200
from lbmpc_b200 import matio


def test_mat_files_have_the_reference_layout(fx, tmp_path):
    from scipy.io import loadmat
    sysH, art = fx["LBMPC_N50__sysH"], fx["LBMPC_N50__art_refH"]
    p1, p2 = matio.save_fform_run(str(tmp_path), "LBMPC_N50", sysH, art)
    a, b = loadmat(p1), loadmat(p2)
    assert p1.endswith("LBMPC_N50_sys_full.mat") and a["sysH"].shape == sysH.shape and a["sysH"].dtype == np.float64
    assert np.array_equal(a["sysH"], sysH) and b["art_refH"].shape == (1, sysH.shape[1])
    x_abs = (lbmpc_b200.X_WP[:, None] + sysH[:4]).T                       # (steps+1, 4) absolute states, like closed_loop["x"][0]
    for variant, key in (("LBMPC", "xlo"), ("LMPC", "xl")):
        p = matio.save_cform_run(str(tmp_path / f"DMS_N50_t{variant}.mat"), x_abs[None], variant)
        m = loadmat(p)
        ref = fx["casadi_DMS_N50_tLBMPC_q100__xlo"]
        assert m[key].shape[0] == ref.shape[0] == 4 and m[key].shape[1] == x_abs.shape[0] - 1 and m[key].dtype == ref.dtype
        assert np.array_equal(m[key][:, 0], x_abs[0])
