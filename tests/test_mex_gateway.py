"""The MATLAB MEX gateway (learning-based-mpc_b200/matlab/lbmpc_mex.c) compiled against the stub mex.h and driven
from Python: packing logic, error path without a GPU, and (on the GPU box) the same answers as the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "learning-based-mpc_b200")
STUB = os.path.join(ROOT, "tests", "mex_stub")


@pytest.fixture(scope="module")
def gw():
    subprocess.check_call(["make", "-C", PKG, "-s"])
    out = os.path.join(STUB, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "lbmpc_mex_stub.so")
    srcs = [os.path.join(PKG, "matlab", "lbmpc_mex.c"), os.path.join(STUB, "mex_stub.c")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs + [os.path.join(STUB, "mex.h")]):
        subprocess.check_call(["gcc", "-O1", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", "-I" + STUB,
                               "-I" + os.path.join(ROOT, "include"), *srcs, "-L" + PKG, "-llbmpc_b200",
                               "-Wl,-rpath," + PKG, "-o", so])
    lib = C.CDLL(so)
    lib.mexstub_double.restype = C.c_void_p
    lib.mexstub_double.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
    lib.mexstub_struct.restype = C.c_void_p
    lib.mexstub_struct.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    lib.mxCreateString.restype = C.c_void_p
    lib.mxCreateString.argtypes = [C.c_char_p]
    lib.mxSetField.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_void_p]
    lib.mxGetField.restype = C.c_void_p
    lib.mxGetField.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p]
    lib.mxGetM.restype = lib.mxGetN.restype = C.c_size_t
    lib.mxGetM.argtypes = lib.mxGetN.argtypes = [C.c_void_p]
    lib.mexstub_copy_out.argtypes = [C.c_void_p, C.c_void_p]
    lib.mexstub_call.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]
    lib.mexstub_last_error.restype = C.c_char_p
    lib.mxGetString.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    return lib


def _dbl(lib, a):
    a = np.asfortranarray(np.atleast_2d(np.asarray(a, float)))
    return lib.mexstub_double(a.ctypes.data, a.shape[0], int(np.prod(a.shape[1:])))


def _struct(lib, fields):
    names = (C.c_char_p * len(fields))(*[k.encode() for k in fields])
    s = lib.mexstub_struct(len(fields), names)
    for k, v in fields.items():
        lib.mxSetField(s, 0, k.encode(), lib.mxCreateString(v.encode()) if isinstance(v, str) else _dbl(lib, v))
    return s


def _call(lib, nlhs, *args):
    plhs = (C.c_void_p * max(nlhs, 1))()
    prhs = (C.c_void_p * len(args))(*args)
    rc = lib.mexstub_call(nlhs, plhs, len(args), prhs)
    return rc, plhs[0], lib.mexstub_last_error().decode()


def _model_struct(lib, mdl):
    f = {"A": mdl["A"], "B": mdl["B"], "K": mdl["K"], "Q": mdl["Q"], "R": mdl["R"], "P": mdl["P"],
         "T": np.atleast_2d(mdl["T"]), "LAMBDA": mdl["LAMBDA"], "PSI": mdl["PSI"], "F_x": mdl["F_x"],
         "h_x": np.asarray(mdl["h_x"]).reshape(-1, 1), "F_u": mdl["F_u"], "h_u": np.asarray(mdl["h_u"]).reshape(-1, 1),
         "F_w_N": mdl["F_w_N"], "h_w_N": np.asarray(mdl["h_w_N"]).reshape(-1, 1)}
    if mdl.get("F_x_d") is not None:
        f["F_x_d"] = mdl["F_x_d"]
        f["h_x_d"] = np.asarray(mdl["h_x_d"]).reshape(-1, 1)
    return _struct(lib, f)


def test_gateway_version_and_bad_command(gw):
    rc, out, _ = _call(gw, 1, gw.mxCreateString(b"version"))
    assert rc == 0
    buf = C.create_string_buffer(64)
    assert gw.mxGetString(out, buf, 64) == 0 and buf.value.startswith(b"lbmpc_b200")
    rc, _, msg = _call(gw, 0, gw.mxCreateString(b"nonsense"))
    assert rc == 1 and "unknown command" in msg


def test_gateway_fails_loudly_without_gpu(gw, models):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    cfg = _struct(gw, {"form": "C", "variant": "LBMPC", "N": 50.0, "max_batch": 4.0})
    rc, _, msg = _call(gw, 1, gw.mxCreateString(b"create"), _model_struct(gw, models["LBMPC"]), cfg)
    assert rc == 1 and "lbmpc_create failed" in msg and "no CUDA device" in msg


@pytest.mark.gpu
def test_gateway_matches_c_abi(gw, models):
    import lbmpc_b200
    from conftest import sample_ics
    mdl = models["LBMPC"]
    dx0 = sample_ics(16, seed=5)
    cfg = _struct(gw, {"form": "C", "variant": "LBMPC", "N": 50.0, "max_batch": 16.0})
    rc, h, msg = _call(gw, 1, gw.mxCreateString(b"create"), _model_struct(gw, mdl), cfg)
    assert rc == 0, msg
    empty = gw.mexstub_double(None, 0, 0)
    rc, out, msg = _call(gw, 1, gw.mxCreateString(b"solve"), h, _dbl(gw, dx0.T), empty, empty, empty)
    assert rc == 0, msg

    def get(name, dtype, shape):
        a = np.empty(shape, dtype=dtype, order="F")
        gw.mexstub_copy_out(gw.mxGetField(out, 0, name.encode()), a.ctypes.data)
        return a
    uc = get("u_or_c", np.float64, (50, 16)).T
    th = get("theta", np.float64, (1, 16)).T
    st = get("status", np.int32, (1, 16)).ravel()
    f = get("f", np.float64, (1, 16)).ravel()
    x = get("x", np.float64, (4, 51, 16))
    ref = lbmpc_b200.Solver(mdl, "C", "LBMPC", 50, max_batch=16).solve_batch(dx0)
    assert (st == ref["status"]).all()
    assert np.array_equal(uc, ref["uc"].reshape(16, 50)) and np.array_equal(th.ravel(), ref["theta"].ravel())
    assert np.array_equal(f, ref["obj"])
    assert np.array_equal(np.transpose(x, (2, 1, 0)), ref["xtraj"].reshape(16, 51, 4))
    # optional 7th argument: per-stage cost shift (nx x (N+1) x batch in MATLAB's column-major layout)
    e = 2e-3 * np.random.default_rng(1).standard_normal((16, 51, 4)).cumsum(axis=1)
    e[:, 0] = 0.0
    e_mx = gw.mexstub_double(np.ascontiguousarray(e).ctypes.data, 4 * 51, 16)
    rc, out, msg = _call(gw, 1, gw.mxCreateString(b"solve"), h, _dbl(gw, dx0.T), empty, empty, empty, e_mx)
    assert rc == 0, msg
    ref_s = lbmpc_b200.Solver(mdl, "C", "LBMPC", 50, max_batch=16).solve_batch(dx0, cost_shift=e)
    assert np.array_equal(get("u_or_c", np.float64, (50, 16)).T, ref_s["uc"].reshape(16, 50))
    assert not np.array_equal(ref_s["uc"], ref["uc"])
    rc, _, msg = _call(gw, 0, gw.mxCreateString(b"destroy"), h)
    assert rc == 0, msg
    rc, _, msg = _call(gw, 1, gw.mxCreateString(b"solve"), h, _dbl(gw, dx0.T))
    assert rc == 1 and "destroyed" in msg
