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
    deps = [src] + [os.path.join(ROOT, "learning-based-mpc_b200", "csrc", f) for f in ("lbmpc_core.cuh", "lbmpc_problem.hpp", "lbmpc_stream.cuh")]
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


def assert_parity(got, ref, tol=1e-8, frac_tight=0.999, max_dit=1, caps=(1e-7, 1e-6)):
    """The parity rule of SURVEY.md §8c / BASELINE.json north_star, FP64:
      * identical feasibility verdicts (status) for every QP,
      * iteration counts within +-1,
      * objective within 1e-8 relative for every QP both sides call optimal,
      * optimal input sequence / theta / predicted states within 1e-8 relative (inf-norm, per QP).
    The last line holds for every well-conditioned QP.  A handful of initial states on the boundary of the
    feasible set give QPs whose minimiser is only weakly determined (multipliers ~1e3..1e4, slacks that reach
    the ulp of the states in the last interior-point iterations): there the objective still agrees to 1e-12
    but the primal iterate carries round-off noise ~ eps*lambda/mu that ANY FP64 interior-point implementation
    shows (the oracle and the independent dense formulation differ by the same amount,
    tests/test_oracle_golden.py).  So: >= 99.9 % of the optimal QPs must meet 1e-8, none may exceed 1e-7
    (1e-6 when the two sides stopped one iteration apart, where a degenerate vertex is only reached to
    O(sqrt(mu))) — the levels a 184 k-QP sweep measured (profiles/r1_stress_parity.json: 99.97 % within 1e-8, worst
    4e-8), so that a 10x regression fails.  Batches below 1000 QPs: at most ONE QP may sit between 1e-8 and the cap."""
    assert (got["status"] == ref["status"]).all(), np.nonzero(got["status"] != ref["status"])
    dit = np.abs(got["iters"].astype(int) - ref["iters"].astype(int))
    assert dit.max(initial=0) <= max_dit
    ok = ref["status"] == 0
    if not ok.any():
        return
    n = int(ok.sum())
    assert (np.abs(got["obj"][ok] - ref["obj"][ok]) / np.maximum(1.0, np.abs(ref["obj"][ok]))).max() < tol
    errs = []
    for key in ("uc", "theta", "xtraj"):
        if got.get(key) is None or ref.get(key) is None:
            continue
        g, r = got[key][ok].reshape(n, -1), ref[key][ok].reshape(n, -1)
        errs.append(np.abs(g - r).max(1) / np.maximum(1.0, np.abs(r).max(1)))
    if not errs:   # verdict / iterations / objective only (problems whose minimiser is not unique)
        return
    e = np.max(np.stack(errs), axis=0)
    cap = np.where(dit[ok] == 0, caps[0], caps[1])
    assert (e < cap).all(), (e.max(), int(np.argmax(e)))
    n_loose = int((e >= tol).sum())
    assert n_loose <= max(1, int((1.0 - frac_tight) * e.size)), ((e < tol).mean(), e.max())


def double_integrator_cases(n, seed=0):
    """Initial states / references of the second problem shape (matlab/trackingMPC/RunExample.m:11-16: x in the +-5 box,
    reference on the steady-state manifold LAMBDA theta)."""
    import lbmpc_b200
    mdl = lbmpc_b200.double_integrator_model()
    rng = np.random.default_rng(seed)
    X0 = rng.uniform([-5.0, -1.5], [5.0, 1.5], (n, 2))
    X0[0] = [0.0, -2.0]                                           # RunExample.m:11-12
    xref = (mdl["LAMBDA"] @ rng.uniform(-1.0, 1.0, (2, n))).T
    return mdl, X0, xref


def drop_numerical(got, ref, max_frac=0.03):
    """Second problem shape only: a few QPs of the double integrator end at degenerate vertices (more active rows than
    variables) where the last interior-point steps are decided by round-off; either implementation may then stop with
    status 3 (numerical) one iteration before the other converges — the outcome even depends on the compiler's FMA
    contraction.  Those QPs (at most max_frac of the batch) are compared on neither side; everything else obeys assert_parity
    (with iteration counts within +-2 there: the same round-off decides when the gap test mu < tol_mu is met)."""
    bad = (got["status"] == 3) | (ref["status"] == 3)
    assert bad.mean() <= max_frac, bad.mean()
    keep = ~bad
    f = lambda d: {k: (v[keep] if v is not None else None) for k, v in d.items()}
    return f(got), f(ref)


def assert_other_minimiser(mdl, N, X0, got, ref, tol=1e-7, feas=1e-8):
    """Second shape, C-form: where the minimiser is not unique (flat directions of a 0.01-weighted input cost at a vertex) the
    two sides may return DIFFERENT minimisers.  The explanation is checked per QP instead of being asserted in prose: every
    optimal QP whose inputs differ by more than `tol` must (a) reproduce its own predicted states from the dynamics
    x+ = A x + B u (trackingMPC/constraintsFunction.m:24-27), (b) satisfy the input box, the state box on x_1..x_N and the terminal
    set on [x_N; theta] (:28-38) to `feas`, and (c) attain the oracle's objective to 1e-8 — i.e. it is a feasible point with the
    optimal value, hence a minimiser of the same QP.  Returns the number of such QPs."""
    A, B = np.asarray(mdl["A"], float), np.asarray(mdl["B"], float)
    ok = np.nonzero((ref["status"] == 0) & (got["status"] == 0))[0]
    n_other = 0
    for b in ok:
        e = np.abs(got["uc"][b] - ref["uc"][b]).max() / max(1.0, np.abs(ref["uc"][b]).max())
        if e < tol:
            continue
        n_other += 1
        u, x, th = got["uc"][b].reshape(N, -1), got["xtraj"][b].reshape(N + 1, -1), got["theta"][b].reshape(-1)
        assert np.abs(x[0] - X0[b]).max() < 1e-12
        for k in range(N):
            assert np.abs(A @ x[k] + B @ u[k] - x[k + 1]).max() < 1e-9, (b, k)
            assert (mdl["F_u"] @ u[k] <= mdl["h_u"] + feas).all(), (b, k)
            assert (mdl["F_x"] @ x[k + 1] <= mdl["h_x"] + feas).all(), (b, k)
        assert (mdl["F_w_N"] @ np.concatenate([x[N], th]) <= mdl["h_w_N"] + feas).all(), b
        assert abs(got["obj"][b] - ref["obj"][b]) <= 1e-8 * max(1.0, abs(ref["obj"][b])), b
    return n_other
