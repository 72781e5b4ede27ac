import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "learning-based-mpc_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fx():
    """Golden data extracted from the reference tree (tests/golden/make_fixtures.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))


@pytest.fixture(scope="session")
def models():
    import lbmpc_b200
    return {v: lbmpc_b200.moore_greitzer_model(v) for v in ("LMPC", "LBMPC")}


@pytest.fixture(scope="session")
def emul_lib():
    """tests/emul/emul_core.cpp: the product's per-QP device functions compiled for the host (test only)."""
    import ctypes
    src = os.path.join(ROOT, "tests", "emul", "emul_core.cpp")
    out_dir = os.path.join(ROOT, "tests", "emul", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libemul.so")
    deps = [src] + [os.path.join(ROOT, "learning-based-mpc_b200", "csrc", f) for f in ("lbmpc_core.cuh", "lbmpc_problem.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-x", "c++", src, "-o", so])
    lib = ctypes.CDLL(so)
    lib.emul_last_error.restype = ctypes.c_char_p
    return lib


def sample_ics(n, seed=0):
    from lbmpc_b200.dist import sample_initial_states
    return sample_initial_states(n, seed)


def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


def assert_parity(got, ref, tol=1e-8):
    """The parity rule of SURVEY.md §8c: identical status, iterations within +-1, 1e-8 relative on the
    optimal input sequence / theta / objective of the QPs both sides call optimal."""
    assert (got["status"] == ref["status"]).all(), np.nonzero(got["status"] != ref["status"])
    assert np.abs(got["iters"].astype(int) - ref["iters"].astype(int)).max() <= 1
    ok = ref["status"] == 0
    if ok.any():
        scale = max(1.0, np.abs(ref["uc"][ok]).max())
        assert np.abs(got["uc"][ok] - ref["uc"][ok]).max() / scale < tol
        assert np.abs(got["theta"][ok] - ref["theta"][ok]).max() < tol
        assert (np.abs(got["obj"][ok] - ref["obj"][ok]) / np.maximum(1.0, np.abs(ref["obj"][ok]))).max() < tol
        if got.get("xtraj") is not None and ref.get("xtraj") is not None:
            assert np.abs(got["xtraj"][ok] - ref["xtraj"][ok]).max() / max(1.0, np.abs(ref["xtraj"][ok]).max()) < tol
