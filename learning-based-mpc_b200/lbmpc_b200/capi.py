"""ctypes binding of the C ABI in include/lbmpc.h (liblbmpc_b200.so).

This is the only way the Python host code reaches the engine: plain pointers and sizes, no
torch types in the signatures.  There is no CPU fallback — if the shared library (built for
sm_100a by `__graft_entry__.build()` / `make -C learning-based-mpc_b200`) is missing, or no CUDA
device is usable, every call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "liblbmpc_b200.so")

FORM = {"F": 0, "C": 1}
VARIANT = {"LMPC": 0, "LBMPC": 1}
ST_OPTIMAL, ST_MAXITER, ST_INFEASIBLE, ST_NUMERICAL = 0, 1, 2, 3
KERNEL = {"auto": 0, "warp": 1, "cta": 2, "stream": 3, "mixed": 4}     # LBMPC_KERNEL_* (include/lbmpc.h)
KERNEL_NAME = {v: k for k, v in KERNEL.items()}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class LbmpcModel(C.Structure):
    """struct lbmpc_model (include/lbmpc.h)."""
    _fields_ = [("nx", C.c_int32), ("nu", C.c_int32), ("nt", C.c_int32),
                ("A", _dp), ("B", _dp), ("K", _dp), ("Q", _dp), ("R", _dp), ("P", _dp), ("T", _dp),
                ("T_is_scalar", C.c_int32), ("LAMBDA", _dp), ("PSI", _dp),
                ("F_x", _dp), ("h_x", _dp), ("n_Fx", C.c_int32),
                ("F_u", _dp), ("h_u", _dp), ("n_Fu", C.c_int32),
                ("F_w_N", _dp), ("h_w_N", _dp), ("n_Fw", C.c_int32),
                ("F_x_d", _dp), ("h_x_d", _dp), ("n_Fxd", C.c_int32)]


class LbmpcConfig(C.Structure):
    """struct lbmpc_config (include/lbmpc.h)."""
    _fields_ = [("form", C.c_int32), ("variant", C.c_int32), ("N", C.c_int32), ("delta", C.c_double),
                ("tol_res", C.c_double), ("tol_mu", C.c_double), ("inf_radius", C.c_double),
                ("max_iter", C.c_int32), ("max_batch", C.c_int64), ("pointers_on_device", C.c_int32)]


def _f(a, shape=None):
    """column-major float64 copy (MATLAB layout)"""
    a = np.asarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return np.asfortranarray(a)


def pack_model(mdl):
    """dict with the ocpLBMPC.m:1-6 matrices -> (LbmpcModel, keepalive list)."""
    B = _f(np.atleast_2d(mdl["B"]))
    nx, nu = B.shape
    LAM = _f(np.asarray(mdl["LAMBDA"], float).reshape(nx, -1))
    nt = LAM.shape[1]
    T = np.atleast_2d(np.asarray(mdl["T"], float))
    t_scalar = int(T.shape == (1, 1) and nx != 1)
    arrs = dict(A=_f(mdl["A"], (nx, nx)), B=B, K=_f(mdl["K"], (nu, nx)), Q=_f(mdl["Q"], (nx, nx)),
                R=_f(np.atleast_2d(mdl["R"]), (nu, nu)), P=_f(mdl["P"], (nx, nx)), T=_f(T), LAMBDA=LAM,
                PSI=_f(np.asarray(mdl["PSI"], float).reshape(nu, nt)),
                F_x=_f(mdl["F_x"]), h_x=_f(np.ravel(mdl["h_x"])), F_u=_f(np.atleast_2d(mdl["F_u"])),
                h_u=_f(np.ravel(mdl["h_u"])), F_w_N=_f(mdl["F_w_N"]), h_w_N=_f(np.ravel(mdl["h_w_N"])))
    if arrs["F_u"].shape[1] != nu:
        arrs["F_u"] = _f(arrs["F_u"].reshape(-1, nu))
    m = LbmpcModel()
    m.nx, m.nu, m.nt, m.T_is_scalar = nx, nu, nt, t_scalar
    if mdl.get("F_x_d") is not None:
        arrs["F_x_d"] = _f(mdl["F_x_d"])
        arrs["h_x_d"] = _f(np.ravel(mdl["h_x_d"]))
        m.n_Fxd = arrs["F_x_d"].shape[0]
    for k, v in arrs.items():
        setattr(m, k, v.ctypes.data_as(_dp))
    m.n_Fx, m.n_Fu, m.n_Fw = arrs["F_x"].shape[0], arrs["F_u"].shape[0], arrs["F_w_N"].shape[0]
    return m, list(arrs.values())


def make_config(form, variant, N, delta=0.01, tol_res=0.0, tol_mu=0.0, inf_radius=0.0, max_iter=0, max_batch=1024,
                pointers_on_device=False):
    c = LbmpcConfig()
    c.form, c.variant, c.N, c.delta = FORM[form], VARIANT[variant], int(N), float(delta)
    c.tol_res, c.tol_mu, c.inf_radius, c.max_iter = tol_res, tol_mu, inf_radius, int(max_iter)
    c.max_batch, c.pointers_on_device = int(max_batch), int(bool(pointers_on_device))
    return c


_lib = None


def load_library(path=None):
    """Load liblbmpc_b200.so and declare the prototypes of include/lbmpc.h.  Raises if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("LBMPC_B200_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build the CUDA extension first "
                           "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.lbmpc_create.argtypes = [C.POINTER(LbmpcModel), C.POINTER(LbmpcConfig), C.c_int, C.POINTER(vp)]
    lib.lbmpc_create.restype = C.c_int
    lib.lbmpc_solve_batch.argtypes = [vp, C.c_int64] + [vp] * 10 + [vp]
    lib.lbmpc_solve_batch.restype = C.c_int
    lib.lbmpc_solve_batch_shifted.argtypes = [vp, C.c_int64] + [vp] * 11 + [vp]
    lib.lbmpc_solve_batch_shifted.restype = C.c_int
    lib.lbmpc_solve_batch_shifted_xu.argtypes = [vp, C.c_int64] + [vp] * 11 + [vp]
    lib.lbmpc_solve_batch_shifted_xu.restype = C.c_int
    lib.lbmpc_oracle_apply.argtypes = [vp, C.c_int64, C.c_int32, C.c_double, C.c_double] + [vp] * 6 + [vp]
    lib.lbmpc_oracle_apply.restype = C.c_int
    lib.lbmpc_solve_sqp.argtypes = [vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double] + [vp] * 13 + [vp]
    lib.lbmpc_solve_sqp.restype = C.c_int
    lib.lbmpc_solve_sqp_ex.argtypes = [vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double] + [vp] * 13 + [vp]
    lib.lbmpc_solve_sqp_ex.restype = C.c_int
    lib.lbmpc_closed_loop.argtypes = [vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, C.c_double,
                                      vp, vp, C.c_uint64, C.c_uint64, vp, vp, vp, vp, vp, vp]
    lib.lbmpc_closed_loop.restype = C.c_int
    lib.lbmpc_set_kernel.argtypes = [vp, C.c_int32, C.c_int32]
    lib.lbmpc_set_kernel.restype = C.c_int
    lib.lbmpc_last_kernel.argtypes = [vp]
    lib.lbmpc_last_kernel.restype = C.c_int
    lib.lbmpc_num_rows.argtypes = [vp]
    lib.lbmpc_slots_per_cta.argtypes = [vp]
    lib.lbmpc_kernel_launches.argtypes = [vp]
    lib.lbmpc_kernel_launches.restype = C.c_int64
    lib.lbmpc_last_kernel_ms.argtypes = [vp]
    lib.lbmpc_last_kernel_ms.restype = C.c_float
    lib.lbmpc_debug_phase_cycles.argtypes = [vp, C.c_int, vp]
    lib.lbmpc_debug_phase_cycles.restype = C.c_int
    lib.lbmpc_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.lbmpc_measure_fp64_peak.restype = C.c_int
    lib.lbmpc_destroy.argtypes = [vp]
    lib.lbmpc_destroy.restype = None
    lib.lbmpc_last_error.restype = C.c_char_p
    lib.lbmpc_version.restype = C.c_char_p
    if path == os.environ.get("LBMPC_B200_LIB", LIB_PATH):
        _lib = lib
    return lib


class LbmpcError(RuntimeError):
    pass


def measure_fp64_peak(device=0):
    """Measured FP64-FMA peak of the device in TFLOP/s (lbmpc_measure_fp64_peak)."""
    lib = load_library()
    v = C.c_double()
    rc = lib.lbmpc_measure_fp64_peak(int(device), C.byref(v))
    if rc != 0:
        raise LbmpcError(f"lbmpc_measure_fp64_peak failed ({rc}): {lib.lbmpc_last_error().decode()}")
    return v.value


def _ptr(a):
    """numpy array / torch tensor / None -> void* address"""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


class Solver:
    """One engine handle = one (model, form, variant, N) on one GPU (lbmpc_create ... lbmpc_destroy).

    Host mode (default): numpy arrays in / out, the call copies and synchronises.
    Device mode (`device_pointers=True`): torch CUDA float64 tensors in / out, asynchronous on the
    given stream.  Array shapes are "one column per QP" as in the header, expressed in numpy as
    (batch, ...) C-contiguous: dx0 (batch,nx), warm (batch,N*nu+nt), d_off (batch,N,nx),
    u_or_c (batch,N,nu), x_traj (batch,N+1,nx).
    """

    def __init__(self, mdl, form, variant, N, delta=0.01, device=0, max_batch=1024, device_pointers=False,
                 tol_res=0.0, tol_mu=0.0, inf_radius=0.0, max_iter=0, lib=None, kernel=None, lockstep=None):
        self.lib = lib or load_library()
        self._model, self._keep = pack_model(mdl)
        self._cfg = make_config(form, variant, N, delta, tol_res, tol_mu, inf_radius, max_iter, max_batch,
                                device_pointers)
        self.nx, self.nu, self.nt, self.N = self._model.nx, self._model.nu, self._model.nt, int(N)
        self.form, self.variant, self.device, self.device_pointers = form, variant, device, device_pointers
        self.max_batch = int(max_batch)
        h = C.c_void_p()
        rc = self.lib.lbmpc_create(C.byref(self._model), C.byref(self._cfg), int(device), C.byref(h))
        if rc != 0:
            raise LbmpcError(f"lbmpc_create failed ({rc}): {self.lib.lbmpc_last_error().decode()}")
        self.h = h
        if kernel is not None or lockstep is not None:
            self.set_kernel(kernel or "auto", lockstep)

    def set_kernel(self, kernel="auto", lockstep=None):
        """lbmpc_set_kernel: force a thread mapping ("auto" | "warp" | "cta" | "stream" | "mixed") and the warp kernel's
        lock-step tick (None = auto)."""
        self._check(self.lib.lbmpc_set_kernel(self.h, KERNEL[kernel], -1 if lockstep is None else int(bool(lockstep))),
                    "lbmpc_set_kernel")

    @property
    def last_kernel(self):
        """mapping the last solve call used: "warp" | "cta" | "stream" | "mixed" ("auto" before the first call)"""
        return KERNEL_NAME[int(self.lib.lbmpc_last_kernel(self.h))]

    # -- introspection --------------------------------------------------------------------------
    @property
    def num_rows(self):
        return self.lib.lbmpc_num_rows(self.h)

    @property
    def slots_per_cta(self):
        return self.lib.lbmpc_slots_per_cta(self.h)

    @property
    def kernel_launches(self):
        return int(self.lib.lbmpc_kernel_launches(self.h))

    @property
    def last_kernel_ms(self):
        return float(self.lib.lbmpc_last_kernel_ms(self.h))

    def phase_cycles(self, enable=True):
        """Read and clear the per-phase SM-cycle counters of CTA 0 (lbmpc_debug_phase_cycles), then switch the
        instrumentation on/off for the following solve calls."""
        out = np.zeros(16, np.uint64)
        self._check(self.lib.lbmpc_debug_phase_cycles(self.h, int(enable), _ptr(out)), "lbmpc_debug_phase_cycles")
        names = ("load_rollout", "EA_update_assemble", "B_factor_adjoint", "B2_affine_sweeps", "C_affine_step",
                 "D_corrector_sweeps", "E_final_step", "iterations", "store", "D_bwd_p1", "D_bwd_p2", "D_bwd_p3_fwd_p1", "D_fwd_p2",
                 "D_fwd_p3", "EA_gen", "EA_reduce")
        return dict(zip(names, (int(v) for v in out[:16])))

    def _check(self, rc, what):
        if rc != 0:
            raise LbmpcError(f"{what} failed ({rc}): {self.lib.lbmpc_last_error().decode()}")

    # -- host-pointer API -----------------------------------------------------------------------
    def solve_batch(self, dx0, dx_ref=None, d_off=None, warm=None, want_x=True, stream=None, out=None, cost_shift=None):
        """lbmpc_solve_batch; with cost_shift (batch,N+1,nx) lbmpc_solve_batch_shifted: objective at x_k + cost_shift_k;
        with cost_shift (batch,N+1,nx+nu) lbmpc_solve_batch_shifted_xu: objective at [x_k + ex_k; u_k + eu_k]."""
        if self.device_pointers:
            return self._solve_device(dx0, dx_ref, d_off, warm, want_x, stream, out, cost_shift)
        nx, nu, nt, N = self.nx, self.nu, self.nt, self.N
        dx0 = np.ascontiguousarray(dx0, np.float64).reshape(-1, nx)
        nb = dx0.shape[0]
        c = lambda a, shp: None if a is None else np.ascontiguousarray(a, np.float64).reshape(shp)
        dx_ref, d_off, warm = c(dx_ref, (nb, nx)), c(d_off, (nb, N, nx)), c(warm, (nb, N * nu + nt))
        xu = cost_shift is not None and np.shape(cost_shift)[-1] == nx + nu
        cost_shift = c(cost_shift, (nb, N + 1, nx + nu if xu else nx))
        fn = self.lib.lbmpc_solve_batch_shifted_xu if xu else self.lib.lbmpc_solve_batch_shifted
        o = dict(uc=np.empty((nb, N, nu)), theta=np.empty((nb, nt)),
                 xtraj=np.empty((nb, N + 1, nx)) if want_x else None, obj=np.empty(nb),
                 iters=np.empty(nb, np.int32), status=np.empty(nb, np.int32))
        rc = fn(self.h, nb, _ptr(dx0), _ptr(dx_ref), _ptr(d_off), _ptr(cost_shift), _ptr(warm),
                _ptr(o["uc"]), _ptr(o["theta"]), _ptr(o["xtraj"]), _ptr(o["obj"]), _ptr(o["iters"]), _ptr(o["status"]), None)
        self._check(rc, "lbmpc_solve_batch")
        return o

    # -- device-pointer API (torch tensors; torch is plumbing for device memory and streams) ----
    def _solve_device(self, dx0, dx_ref, d_off, warm, want_x, stream, out, cost_shift=None):
        import torch
        nx, nu, nt, N = self.nx, self.nu, self.nt, self.N
        nb = dx0.shape[0]
        dev = dx0.device
        if out is None:
            out = dict(uc=torch.empty((nb, N, nu), dtype=torch.float64, device=dev),
                       theta=torch.empty((nb, nt), dtype=torch.float64, device=dev),
                       xtraj=torch.empty((nb, N + 1, nx), dtype=torch.float64, device=dev) if want_x else None,
                       obj=torch.empty(nb, dtype=torch.float64, device=dev),
                       iters=torch.empty(nb, dtype=torch.int32, device=dev),
                       status=torch.empty(nb, dtype=torch.int32, device=dev))
        st = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        for name, t in (("dx0", dx0), ("dx_ref", dx_ref), ("d_off", d_off), ("warm", warm), ("cost_shift", cost_shift)):
            if t is not None and (t.dtype != torch.float64 or not t.is_contiguous() or t.device != dev or not t.is_cuda):
                raise LbmpcError(f"{name}: device-pointer mode takes contiguous float64 CUDA tensors on one device "
                                 f"(got {t.dtype}, contiguous={t.is_contiguous()}, {t.device})")
        xu = cost_shift is not None and cost_shift.shape[-1] == nx + nu
        fn = self.lib.lbmpc_solve_batch_shifted_xu if xu else self.lib.lbmpc_solve_batch_shifted
        rc = fn(self.h, nb, _ptr(dx0), _ptr(dx_ref), _ptr(d_off), _ptr(cost_shift), _ptr(warm),
                                                _ptr(out["uc"]), _ptr(out["theta"]), _ptr(out.get("xtraj")), _ptr(out["obj"]),
                                                _ptr(out["iters"]), _ptr(out["status"]), C.c_void_p(st))
        self._check(rc, "lbmpc_solve_batch")
        return out

    def oracle_apply(self, dx0, du, X, Y, valid=None, bandwidth=0.5, lam=0.001, stream=None):
        """d_off (batch,N,nx) for input sequences du (batch,N,nu) and data windows X (batch,q,3), Y (batch,q,nx)
        — "one column per sample" in MATLAB terms (data.X is 3 x q, oracleL2NW.m:2)."""
        nx, N = self.nx, self.N
        if self.device_pointers:
            import torch
            nb, q = dx0.shape[0], X.shape[1]
            d = torch.empty((nb, N, nx), dtype=torch.float64, device=dx0.device)
            st = stream if stream is not None else torch.cuda.current_stream(dx0.device).cuda_stream
            rc = self.lib.lbmpc_oracle_apply(self.h, nb, q, bandwidth, lam, _ptr(dx0), _ptr(du), _ptr(X), _ptr(Y),
                                             _ptr(valid), _ptr(d), C.c_void_p(st))
        else:
            dx0 = np.ascontiguousarray(dx0, np.float64).reshape(-1, nx)
            nb = dx0.shape[0]
            du = np.ascontiguousarray(du, np.float64).reshape(nb, N, self.nu)
            X = np.ascontiguousarray(X, np.float64)
            Y = np.ascontiguousarray(Y, np.float64)
            q = X.shape[1]
            valid = None if valid is None else np.ascontiguousarray(valid, np.float64)
            d = np.empty((nb, N, nx))
            rc = self.lib.lbmpc_oracle_apply(self.h, nb, q, bandwidth, lam, _ptr(dx0), _ptr(du), _ptr(X), _ptr(Y),
                                             _ptr(valid), _ptr(d), None)
        self._check(rc, "lbmpc_oracle_apply")
        return d

    def solve_sqp(self, dx0, X, Y, valid=None, sqp_iters=3, dx_ref=None, warm=None, bandwidth=0.5, lam=0.001, want_x=True,
                  twin=False, order=0):
        """Learned-oracle problem as a sequence of QPs (lbmpc_solve_sqp): X (batch,q,3), Y (batch,q,nx) data windows.
        Host arrays.  twin: cost on the learned state sequence, rows on the nominal one (DMS_LBMPC_casadi.m).  Returns the last
        QP's solution plus du_step (batch, sqp_iters).  order=1: oracle value AND Jacobian per outer iteration (LTV QP on the
        learned sequence, lbmpc_solve_sqp_ex)."""
        if self.device_pointers:
            raise LbmpcError("solve_sqp is exposed for host-pointer handles")
        nx, nu, nt, N = self.nx, self.nu, self.nt, self.N
        dx0 = np.ascontiguousarray(dx0, np.float64).reshape(-1, nx)
        nb = dx0.shape[0]
        c = lambda a, shp: None if a is None else np.ascontiguousarray(a, np.float64).reshape(shp)
        X, Y = np.ascontiguousarray(X, np.float64), np.ascontiguousarray(Y, np.float64)
        q = X.shape[1]
        valid, dx_ref, warm = c(valid, (nb, q)), c(dx_ref, (nb, nx)), c(warm, (nb, N * nu + nt))
        o = dict(uc=np.empty((nb, N, nu)), theta=np.empty((nb, nt)), xtraj=np.empty((nb, N + 1, nx)) if want_x else None,
                 obj=np.empty(nb), iters=np.empty(nb, np.int32), status=np.empty(nb, np.int32),
                 du_step=np.empty((nb, sqp_iters)))
        rc = self.lib.lbmpc_solve_sqp_ex(self.h, nb, int(sqp_iters), int(bool(twin)), int(order), int(q), float(bandwidth), float(lam),
                                         _ptr(dx0), _ptr(dx_ref), _ptr(X), _ptr(Y), _ptr(valid), _ptr(warm), _ptr(o["uc"]),
                                         _ptr(o["theta"]), _ptr(o["xtraj"]), _ptr(o["obj"]), _ptr(o["iters"]), _ptr(o["status"]),
                                         _ptr(o["du_step"]), None)
        self._check(rc, "lbmpc_solve_sqp")
        return o

    def closed_loop(self, x_init, steps, x_eq, u_eq, q=100, use_oracle=False, warm_shift=True, wbar=None, seed=0,
                    scenario0=0):
        """Batch of closed-loop scenarios.  Returns dict x (batch,steps+1,nx), u, theta, iters, status (numpy arrays for a
        host-pointer handle; torch tensors on x_init's device, filled asynchronously, for a device-pointer handle).
        C-form handles: LBMPC_casadi.m:160-223; F-form handles: ocpLBMPC.m:10-47 / ocpLMPC.m:11-40 (u = K dx + c to the plant,
        unshifted opt_var, update_data window; LBMPC + use_oracle: two first-order SQP iterations per step) — include/lbmpc.h."""
        if self.device_pointers:
            import torch
            nb, dev = x_init.shape[0], x_init.device
            x_eq = np.ascontiguousarray(x_eq, np.float64)                      # tiny parameter vectors: always host
            wbar = None if wbar is None else np.ascontiguousarray(wbar, np.float64)
            o = dict(x=torch.empty((nb, steps + 1, self.nx), dtype=torch.float64, device=dev),
                     u=torch.empty((nb, steps), dtype=torch.float64, device=dev),
                     theta=torch.empty((nb, steps), dtype=torch.float64, device=dev),
                     iters=torch.empty((nb, steps), dtype=torch.int32, device=dev),
                     status=torch.empty((nb, steps), dtype=torch.int32, device=dev))
            st = torch.cuda.current_stream(dev).cuda_stream
            rc = self.lib.lbmpc_closed_loop(self.h, nb, steps, q, int(use_oracle), int(warm_shift), _ptr(x_eq), float(u_eq),
                                            _ptr(x_init), _ptr(wbar), seed, scenario0, _ptr(o["x"]), _ptr(o["u"]),
                                            _ptr(o["theta"]), _ptr(o["iters"]), _ptr(o["status"]), C.c_void_p(st))
            self._check(rc, "lbmpc_closed_loop")
            return o
        nx = self.nx
        x_init = np.ascontiguousarray(x_init, np.float64).reshape(-1, nx)
        nb = x_init.shape[0]
        x_eq = np.ascontiguousarray(x_eq, np.float64)
        wbar = None if wbar is None else np.ascontiguousarray(wbar, np.float64)
        o = dict(x=np.empty((nb, steps + 1, nx)), u=np.empty((nb, steps)), theta=np.empty((nb, steps)),
                 iters=np.empty((nb, steps), np.int32), status=np.empty((nb, steps), np.int32))
        rc = self.lib.lbmpc_closed_loop(self.h, nb, steps, q, int(use_oracle), int(warm_shift), _ptr(x_eq),
                                        float(u_eq), _ptr(x_init), _ptr(wbar), seed, scenario0, _ptr(o["x"]),
                                        _ptr(o["u"]), _ptr(o["theta"]), _ptr(o["iters"]), _ptr(o["status"]), None)
        self._check(rc, "lbmpc_closed_loop")
        return o

    def close(self):
        if getattr(self, "h", None):
            self.lib.lbmpc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
