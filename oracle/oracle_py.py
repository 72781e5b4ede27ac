"""ctypes front-end of the CPU oracle (oracle/lbmpc_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liblbmpc_oracle.so")
FORM = {"F": 0, "C": 1}
VARIANT = {"LMPC": 0, "LBMPC": 1}
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("lbmpc_oracle.c", "lbmpc_oracle.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.lbo_create.restype = C.c_void_p
        _lib.lbo_create_custom.restype = C.c_void_p
        _lib.lbo_last_error.restype = C.c_char_p
        _lib.lbo_destroy.argtypes = [C.c_void_p]
        _lib.lbo_set_options.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double]
        _lib.lbo_num_rows.argtypes = [C.c_void_p]
    return _lib


def _d(a):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


class OracleProblem:
    """One (form, variant, N) problem built from the reference-style model dict
    (keys A,B,K,Q,R,P,T,LAMBDA,PSI,F_x,h_x,F_u,h_u,F_w_N,h_w_N[,F_x_d,h_x_d])."""

    def __init__(self, form, variant, mdl, N, delta=0.01):
        L = lib()
        A, B = _d(mdl["A"]), _d(mdl["B"])
        nx, nu = B.shape
        Lam = _d(np.atleast_2d(mdl["LAMBDA"]).reshape(nx, -1))
        nt = Lam.shape[1]
        T = np.atleast_2d(np.asarray(mdl["T"], float))
        if T.shape == (1, 1):
            T = T[0, 0] * np.eye(nx)
        args = [_d(mdl["K"]).reshape(nu, nx), _d(mdl["Q"]), _d(np.atleast_2d(mdl["R"])), _d(mdl["P"]), _d(T), Lam,
                _d(np.atleast_2d(mdl["PSI"]).reshape(nu, nt))]
        Fx, hx, Fu, hu = _d(mdl["F_x"]), _d(mdl["h_x"]).reshape(-1), _d(mdl["F_u"]), _d(mdl["h_u"]).reshape(-1)
        Fw, hw = _d(mdl["F_w_N"]), _d(mdl["h_w_N"]).reshape(-1)
        if variant == "LBMPC":
            Fxd, hxd = _d(mdl["F_x_d"]), _d(mdl["h_x_d"]).reshape(-1)
            nFxd = Fxd.shape[0]
        else:
            Fxd = hxd = None
            nFxd = 0
        self._keep = [A, B, Fx, hx, Fu, hu, Fw, hw, Fxd, hxd] + args
        self.h = L.lbo_create(C.c_int(FORM[form]), C.c_int(VARIANT[variant]), C.c_int(nx), C.c_int(nu), C.c_int(nt),
                              C.c_int(N), C.c_double(delta), _p(A), _p(B), *[_p(a) for a in args], _p(Fx), _p(hx),
                              C.c_int(Fx.shape[0]), _p(Fu), _p(hu), C.c_int(Fu.shape[0]), _p(Fw), _p(hw),
                              C.c_int(Fw.shape[0]), _p(Fxd), _p(hxd), C.c_int(nFxd))
        if not self.h:
            raise ValueError(L.lbo_last_error().decode())
        self.h = C.c_void_p(self.h)
        self.nx, self.nu, self.nt, self.N = nx, nu, nt, N
        self.form, self.variant = form, variant

    @classmethod
    def custom(cls, nx, nu, nt, N, A, B, Kinit, Q, R, P, T, Lam, Psi, wq, wr, kP, lo_x, hi_x, kx0, kx1, lo_u, hi_u,
               ku0, ku1, G, hg, kg):
        L = lib()
        self = cls.__new__(cls)
        arrs = [_d(a) for a in (A, B, Kinit, Q, R, P, T, Lam, Psi, wq, wr)]
        b1 = [_d(lo_x), _d(hi_x)]
        b2 = [_d(lo_u), _d(hi_u)]
        G, hg = _d(G), _d(hg)
        ng = 0 if G is None else G.shape[0]
        self._keep = arrs + b1 + b2 + [G, hg]
        h = L.lbo_create_custom(C.c_int(nx), C.c_int(nu), C.c_int(nt), C.c_int(N), *[_p(a) for a in arrs], C.c_int(kP),
                                _p(b1[0]), _p(b1[1]), C.c_int(kx0), C.c_int(kx1), _p(b2[0]), _p(b2[1]), C.c_int(ku0),
                                C.c_int(ku1), _p(G), _p(hg), C.c_int(ng), C.c_int(kg))
        if not h:
            raise ValueError(L.lbo_last_error().decode())
        self.h = C.c_void_p(h)
        self.nx, self.nu, self.nt, self.N = nx, nu, nt, N
        self.form = self.variant = "custom"
        return self

    def set_options(self, tol_res=0.0, tol_mu=0.0, max_iter=0, inf_radius=0.0):
        lib().lbo_set_options(self.h, tol_res, tol_mu, max_iter, inf_radius)

    @property
    def num_rows(self):
        return lib().lbo_num_rows(self.h)

    def solve(self, dx0, dx_ref=None, d_off=None, warm=None, cost_shift=None):
        out = self.solve_batch(np.asarray(dx0, float).reshape(1, -1),
                               None if dx_ref is None else np.asarray(dx_ref, float).reshape(1, -1),
                               None if d_off is None else np.asarray(d_off, float).reshape(1, self.N, self.nx),
                               None if warm is None else np.asarray(warm, float).reshape(1, -1), nthreads=1, stats=True,
                               cost_shift=None if cost_shift is None else np.asarray(cost_shift, float).reshape(1, self.N + 1, self.nx))
        return {k: v[0] for k, v in out.items()}

    def solve_batch(self, dx0, dx_ref=None, d_off=None, warm=None, nthreads=1, stats=False, cost_shift=None, row_shift=None,
                    jac=None):
        """dx0 (batch,nx); d_off (batch,N,nx); warm (batch,N*nu+nt); cost_shift (batch,N+1,nx or nx+nu): the objective is
        evaluated at [x_k + ex_k; u_k + eu_k]; row_shift (same shape): the ROWS act on x_k - ex_k, u_k - eu_k instead; jac
        (batch,N,nx,3): LTV dynamics A_k = A + [J 0 0], B_k = B + J(:,3) (lbo_solve_batch_ex).  Returns dict of arrays."""
        assert cost_shift is None or row_shift is None
        rows = row_shift is not None
        if rows:
            cost_shift = row_shift
        jac = _d(jac)
        L = lib()
        dx0 = _d(dx0)
        nb = dx0.shape[0]
        nx, nu, nt, N = self.nx, self.nu, self.nt, self.N
        dx_ref, d_off, warm, cost_shift = _d(dx_ref), _d(d_off), _d(warm), _d(cost_shift)
        cs_stride = nx if cost_shift is None else int(cost_shift.shape[-1])      # nx: state shift; nx+nu: [ex | eu]
        uc = np.empty((nb, N, nu)); th = np.empty((nb, nt)); xt = np.empty((nb, N + 1, nx)); obj = np.empty(nb)
        it = np.empty(nb, np.int32); st = np.empty(nb, np.int32)
        out = dict(uc=uc, theta=th, xtraj=xt, obj=obj, iters=it, status=st)
        if stats and nb == 1 and cs_stride == nx and not rows and jac is None:
            s = np.zeros(6)
            L.lbo_solve_shifted(self.h, _p(dx0), _p(dx_ref), _p(d_off), _p(cost_shift), _p(warm), _p(uc), _p(th), _p(xt), _p(obj),
                                it.ctypes.data_as(_ip), st.ctypes.data_as(_ip), _p(s))
            out["stats"] = s.reshape(1, 6)
        else:
            rc = L.lbo_solve_batch_ex(self.h, C.c_long(nb), _p(dx0), _p(dx_ref), _p(d_off), _p(cost_shift), C.c_int(cs_stride),
                                      C.c_int(int(rows)), _p(jac), _p(warm), _p(uc), _p(th), _p(xt), _p(obj),
                                      it.ctypes.data_as(_ip), st.ctypes.data_as(_ip), C.c_int(nthreads))
            if rc:
                raise ValueError(L.lbo_last_error().decode())
        return out

    def oracle_offsets(self, dx0, du, X, Y, valid=None, bandwidth=0.5, lam=0.001):
        X, Y, valid = _d(X), _d(Y), _d(valid)
        d = np.empty((self.N, self.nx))
        dx0, du = _d(dx0), _d(du)
        lib().lbo_oracle_offsets(self.h, _p(dx0), _p(du), _p(X), _p(Y), _p(valid), C.c_int(X.shape[1]),
                                 C.c_double(bandwidth), C.c_double(lam), _p(d))
        return d

    def oracle_offsets_jac(self, dx0, du, X, Y, valid=None, bandwidth=0.5, lam=0.001):
        """(d_off (N,nx) = g - J xi, jac (N,nx,3), g (N,nx)) along the learned rollout (lbo_oracle_offsets_jac)."""
        X, Y, valid = _d(X), _d(Y), _d(valid)
        d, J, g = np.empty((self.N, self.nx)), np.empty((self.N, self.nx, 3)), np.empty((self.N, self.nx))
        dx0, du = _d(dx0), _d(du)
        lib().lbo_oracle_offsets_jac(self.h, _p(dx0), _p(du), _p(X), _p(Y), _p(valid), C.c_int(X.shape[1]),
                                     C.c_double(bandwidth), C.c_double(lam), _p(d), _p(J), _p(g))
        return d, J, g

    def solve_sqp1(self, dx0, X, Y, valid=None, sqp_iters=3, dx_ref=None, warm=None, bandwidth=0.5, lam=0.001, twin=False,
                   A=None, B=None, K=None):
        """CPU mirror of lbmpc_solve_sqp with order = 1 (first-order SQP): every outer iteration linearises the learned model
        x+ = A x + B u + g(xi) along the learned rollout of the previous solution (oracle value AND Jacobian) and solves one LTV
        QP on the LEARNED state sequence.  twin: the rows follow the NOMINAL sequence (constraintsLBMPC.m:23,
        DMS_LBMPC_casadi.m:283-319): they act on x_k - e_k, u_k - K e_k with the gap e_k = learned - nominal rollout of the
        previous solution, frozen during the QP.  Outputs: uc (c or du), theta, xtraj = the LEARNED sequence."""
        dx0 = np.ascontiguousarray(dx0, float).reshape(-1, self.nx)
        nb, N, nx, nu = dx0.shape[0], self.N, self.nx, self.nu
        Kf = np.zeros((nu, nx)) if K is None else np.asarray(K, float).reshape(nu, nx)
        A_ = None if A is None else np.asarray(A, float)
        B_ = None if B is None else np.asarray(B, float).reshape(nx, nu)
        outs, steps = [], np.empty((nb, sqp_iters))
        for b in range(nb):
            ulin = np.zeros(N) if warm is None else np.array(warm[b][:N], float)
            w = None if warm is None else np.array(warm[b], float)[None, :]
            xr = None if dx_ref is None else np.asarray(dx_ref[b], float)[None, :]
            Xb, Yb = np.ascontiguousarray(X[b].T), np.ascontiguousarray(Y[b].T)
            vb = None if valid is None else valid[b]
            for j in range(sqp_iters):
                d, J, g = self.oracle_offsets_jac(dx0[b], ulin, Xb, Yb, vb, bandwidth, lam)
                rs = None
                if twin:   # gap between the learned and the nominal rollout of the same decision variables
                    xl, xn_ = dx0[b].copy(), dx0[b].copy()
                    rs = np.zeros((N + 1, nx + nu))
                    for k in range(N):
                        rs[k, :nx] = xl - xn_
                        rs[k, nx:] = Kf @ (xl - xn_)
                        ul, un = Kf @ xl + ulin[k], Kf @ xn_ + ulin[k]
                        xl = A_ @ xl + B_ @ ul + g[k]
                        xn_ = A_ @ xn_ + B_ @ un
                    rs[N, :nx] = xl - xn_
                    rs = rs[None]
                o = self.solve_batch(dx0[b:b + 1], xr, d[None], w, row_shift=rs, jac=J[None])
                u = o["uc"][0, :, 0]
                steps[b, j] = np.abs(u - ulin).max()
                ulin = u.copy()
                w = np.concatenate([u, o["theta"][0]])[None, :]
            outs.append(o)
        res = {k: np.concatenate([o[k] for o in outs]) for k in outs[0] if outs[0][k] is not None}
        res["du_step"] = steps
        return res

    def solve_sqp(self, dx0, X, Y, valid=None, sqp_iters=3, dx_ref=None, warm=None, bandwidth=0.5, lam=0.001, twin=False, A=None,
                  B=None, K=None):
        """CPU mirror of lbmpc_solve_sqp (include/lbmpc.h): oracle offsets along the previous inputs, then one QP, repeated.
        X (batch,q,3), Y (batch,q,nx), valid (batch,q) or None.  twin: cost on the learned sequence, rows on the nominal one;
        with K (F-form: u = K x + c) the gap obeys e+ = (A + B K) e + d and the input gap is K e."""
        dx0 = np.ascontiguousarray(dx0, float).reshape(-1, self.nx)
        nb, N = dx0.shape[0], self.N
        outs, steps = [], np.empty((nb, sqp_iters))
        for b in range(nb):
            ulin = np.zeros(N) if warm is None else np.array(warm[b][:N], float)
            w = None if warm is None else np.array(warm[b], float)[None, :]
            xr = None if dx_ref is None else np.asarray(dx_ref[b], float)[None, :]
            for j in range(sqp_iters):
                d = self.oracle_offsets(dx0[b], ulin, np.ascontiguousarray(X[b].T), np.ascontiguousarray(Y[b].T),
                                        None if valid is None else valid[b], bandwidth, lam)
                if twin:   # cost on x + e with e_{k+1} = (A + B K) e_k + d_k (K = 0: C-form), rows and dynamics on x
                    Kf = np.zeros((self.nu, self.nx)) if K is None else np.asarray(K, float).reshape(self.nu, self.nx)
                    Acl = np.asarray(A, float) + (0.0 if K is None else np.asarray(B, float).reshape(self.nx, self.nu) @ Kf)
                    e = np.zeros((N + 1, self.nx + self.nu))
                    for k in range(N + 1):
                        e[k, self.nx:] = Kf @ e[k, :self.nx]
                        if k < N:
                            e[k + 1, :self.nx] = Acl @ e[k, :self.nx] + d[k]
                    o = self.solve_batch(dx0[b:b + 1], xr, None, w, cost_shift=e[None])
                else:
                    o = self.solve_batch(dx0[b:b + 1], xr, d[None], w)
                u = o["uc"][0, :, 0]
                steps[b, j] = np.abs(u - ulin).max()
                ulin = u.copy()
                w = np.concatenate([u, o["theta"][0]])[None, :]
            outs.append(o)
        res = {k: np.concatenate([o[k] for o in outs]) for k in outs[0] if outs[0][k] is not None}
        res["du_step"] = steps
        return res

    def closed_loop(self, x_eq, u_eq, x_init, steps, q=100, use_oracle=False, warm_shift=True, wbar=None, seed=0,
                    scenario=0):
        xe, xi, wb = _d(x_eq), _d(x_init), _d(wbar)
        xh = np.empty((steps + 1, 4)); uh = np.empty(steps); th = np.empty(steps)
        ih = np.empty(steps, np.int32); sh = np.empty(steps, np.int32)
        rc = lib().lbo_closed_loop(self.h, _p(xe), C.c_double(u_eq), _p(xi), C.c_int(steps), C.c_int(q),
                                   C.c_int(int(use_oracle)), C.c_int(int(warm_shift)), _p(wb),
                                   C.c_ulonglong(seed), C.c_ulonglong(scenario), _p(xh), _p(uh), _p(th),
                                   ih.ctypes.data_as(_ip), sh.ctypes.data_as(_ip))
        if rc:
            raise RuntimeError(lib().lbo_last_error().decode())
        return dict(x=xh, u=uh, theta=th, iters=ih, status=sh)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().lbo_destroy(self.h)
                self.h = None
        except Exception:
            pass


def oracle_l2nw(X, Y, xi, valid=None, bandwidth=0.5, lam=0.001):
    X, Y, xi, valid = _d(X), _d(Y), _d(xi), _d(valid)
    g = np.empty(Y.shape[0])
    lib().lbo_oracle_l2nw(_p(X), _p(Y), _p(valid), C.c_int(X.shape[1]), C.c_int(X.shape[0]), C.c_int(Y.shape[0]),
                          _p(xi), C.c_double(bandwidth), C.c_double(lam), _p(g))
    return g


def plant_rk4(x, u, delta=0.01):
    x = _d(x)
    xn = np.empty(4)
    lib().lbo_plant_rk4(_p(x), C.c_double(u), C.c_double(delta), _p(xn))
    return xn
