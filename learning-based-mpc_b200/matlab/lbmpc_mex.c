/*
 * lbmpc_mex.c — MEX gateway: MATLAB <-> the C ABI of include/lbmpc.h (liblbmpc_b200.so).
 *
 * This is the only piece a MATLAB user of bevanda/Learning-Based-MPC has to build:
 *     mex -R2018a lbmpc_mex.c -I../../include -L.. -llbmpc_b200
 * It holds NO numerics: every command packs mxArrays (column-major doubles, exactly the layout
 * the C ABI takes) and forwards to one lbmpc_* entry point.  A non-zero return code becomes a
 * MATLAB error with the library's message (mexErrMsgIdAndTxt), so a missing GPU fails loudly.
 *
 *   h        = lbmpc_mex('create', model, cfg)          model/cfg: structs, fields below
 *   out      = lbmpc_mex('solve', h, dx0, dx_ref, d_off, warm)    [] for unused inputs
 *              out.u_or_c (nu*N x batch), out.theta (nt x batch), out.x (nx x (N+1) x batch),
 *              out.f (1 x batch), out.iters, out.status (int32 1 x batch)
 *   d_off    = lbmpc_mex('oracle', h, q, bandwidth, lambda, dx0, du, X, Y, valid)
 *   out      = lbmpc_mex('solve_sqp', h, sqp_iters, q, bandwidth, lambda, dx0, dx_ref, X, Y, valid, warm, twin, order)   out.du_step added
 *              (twin, order optional, default 0: lbmpc_solve_sqp_ex — order 1 = oracle value AND Jacobian, a first-order SQP step)
 *   hist     = lbmpc_mex('closed_loop', h, steps, q, use_oracle, warm_shift, x_eq, u_eq, x_init, wbar, seed)
 *   lbmpc_mex('destroy', h)        v = lbmpc_mex('version')
 *
 * model fields (the argument list of functions/ocpLBMPC.m:1-6): A B K Q R P T LAMBDA PSI
 *   F_x h_x F_u h_u F_w_N h_w_N [F_x_d h_x_d]
 * cfg fields: form ('F'|'C'), variant ('LMPC'|'LBMPC'), N, [delta tol_res tol_mu inf_radius max_iter max_batch]
 *
 * The handle travels through MATLAB as a uint64 scalar.  The container this repo is built in has
 * no MATLAB: tests/test_mex_gateway.py compiles this file against a stub mex.h
 * (tests/mex_stub/mex.h) and drives mexFunction from C, which checks the packing logic end to end.
 */
#include <stdint.h>
#include <string.h>

#include "mex.h"
#include "lbmpc.h"

#define MAX_CMD 32

static void fail_rc(const char *what, int rc) {
    mexErrMsgIdAndTxt("lbmpc:call", "%s failed (%d): %s", what, rc, lbmpc_last_error());
}

static const mxArray *field(const mxArray *s, const char *name, int required) {
    const mxArray *f = mxIsStruct(s) ? mxGetField(s, 0, name) : NULL;
    if (!f && required) mexErrMsgIdAndTxt("lbmpc:args", "missing struct field '%s'", name);
    return f;
}
static const double *dptr(const mxArray *a) {
    if (!a || mxIsEmpty(a)) return NULL;
    if (!mxIsDouble(a) || mxIsComplex(a)) mexErrMsgIdAndTxt("lbmpc:args", "arrays must be real double");
    return mxGetPr(a);
}
static double scalar_or(const mxArray *s, const char *name, double dflt) {
    const mxArray *f = field(s, name, 0);
    return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;
}
static lbmpc_handle *get_handle(const mxArray *a) {
    if (!a || !mxIsUint64(a) || mxGetNumberOfElements(a) != 1)
        mexErrMsgIdAndTxt("lbmpc:args", "handle must be the uint64 scalar returned by 'create'");
    return (lbmpc_handle *)(uintptr_t)(*(const uint64_t *)mxGetData(a));
}
/* dimensions cached next to the handle (needed to size outputs) */
typedef struct { lbmpc_handle *h; int nx, nu, nt, N; } entry_t;
#define MAX_HANDLES 64
static entry_t g_tab[MAX_HANDLES];
static entry_t *lookup(lbmpc_handle *h) {
    int i;
    for (i = 0; i < MAX_HANDLES; ++i)
        if (g_tab[i].h == h && h) return &g_tab[i];
    mexErrMsgIdAndTxt("lbmpc:args", "unknown or destroyed handle");
    return NULL;
}
static void at_exit(void) {
    int i;
    for (i = 0; i < MAX_HANDLES; ++i)
        if (g_tab[i].h) { lbmpc_destroy(g_tab[i].h); g_tab[i].h = NULL; }
}

static void cmd_create(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    lbmpc_model m;
    lbmpc_config c;
    const mxArray *ms, *cs, *f;
    char buf[16];
    lbmpc_handle *h = NULL;
    int rc, i;
    if (nrhs < 3) mexErrMsgIdAndTxt("lbmpc:args", "usage: h = lbmpc_mex('create', model, cfg)");
    ms = prhs[1]; cs = prhs[2];
    memset(&m, 0, sizeof m); memset(&c, 0, sizeof c);
    f = field(ms, "A", 1);      m.A = dptr(f);  m.nx = (int32_t)mxGetM(f);
    f = field(ms, "B", 1);      m.B = dptr(f);  m.nu = (int32_t)mxGetN(f);
    f = field(ms, "LAMBDA", 1); m.LAMBDA = dptr(f); m.nt = (int32_t)mxGetN(f);
    m.K = dptr(field(ms, "K", 1));   m.Q = dptr(field(ms, "Q", 1));  m.R = dptr(field(ms, "R", 1));
    m.P = dptr(field(ms, "P", 1));   m.PSI = dptr(field(ms, "PSI", 1));
    f = field(ms, "T", 1);      m.T = dptr(f);  m.T_is_scalar = mxGetNumberOfElements(f) == 1;
    f = field(ms, "F_x", 1);    m.F_x = dptr(f); m.n_Fx = (int32_t)mxGetM(f);   m.h_x = dptr(field(ms, "h_x", 1));
    f = field(ms, "F_u", 1);    m.F_u = dptr(f); m.n_Fu = (int32_t)mxGetM(f);   m.h_u = dptr(field(ms, "h_u", 1));
    f = field(ms, "F_w_N", 1);  m.F_w_N = dptr(f); m.n_Fw = (int32_t)mxGetM(f); m.h_w_N = dptr(field(ms, "h_w_N", 1));
    f = field(ms, "F_x_d", 0);
    if (f && !mxIsEmpty(f)) { m.F_x_d = dptr(f); m.n_Fxd = (int32_t)mxGetM(f); m.h_x_d = dptr(field(ms, "h_x_d", 1)); }
    f = field(cs, "form", 1);
    if (mxGetString(f, buf, sizeof buf)) mexErrMsgIdAndTxt("lbmpc:args", "cfg.form must be 'F' or 'C'");
    c.form = (buf[0] == 'F' || buf[0] == 'f') ? LBMPC_FORM_F : LBMPC_FORM_C;
    f = field(cs, "variant", 1);
    if (mxGetString(f, buf, sizeof buf)) mexErrMsgIdAndTxt("lbmpc:args", "cfg.variant must be 'LMPC' or 'LBMPC'");
    c.variant = (strcmp(buf, "LBMPC") == 0 || strcmp(buf, "lbmpc") == 0) ? LBMPC_VARIANT_LBMPC : LBMPC_VARIANT_LMPC;
    c.N = (int32_t)mxGetScalar(field(cs, "N", 1));
    c.delta = scalar_or(cs, "delta", 0.01);
    c.tol_res = scalar_or(cs, "tol_res", 0.0);
    c.tol_mu = scalar_or(cs, "tol_mu", 0.0);
    c.inf_radius = scalar_or(cs, "inf_radius", 0.0);
    c.max_iter = (int32_t)scalar_or(cs, "max_iter", 0.0);
    c.max_batch = (int64_t)scalar_or(cs, "max_batch", 1.0);
    c.pointers_on_device = 0; /* mxArrays live in host memory */
    rc = lbmpc_create(&m, &c, (int)scalar_or(cs, "device", 0.0), &h);
    if (rc != LBMPC_OK) fail_rc("lbmpc_create", rc);
    for (i = 0; i < MAX_HANDLES && g_tab[i].h; ++i) {}
    if (i == MAX_HANDLES) { lbmpc_destroy(h); mexErrMsgIdAndTxt("lbmpc:args", "too many live handles"); }
    g_tab[i].h = h; g_tab[i].nx = m.nx; g_tab[i].nu = m.nu; g_tab[i].nt = m.nt; g_tab[i].N = c.N;
    mexAtExit(at_exit);
    plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
    *(uint64_t *)mxGetData(plhs[0]) = (uint64_t)(uintptr_t)h;
    (void)nlhs;
}

static void cmd_solve(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    static const char *names[] = {"u_or_c", "theta", "x", "f", "iters", "status"};
    entry_t *e;
    int64_t batch;
    mwSize d3[3];
    mxArray *uc, *th, *x, *f, *it, *st;
    int rc;
    if (nrhs < 3) mexErrMsgIdAndTxt("lbmpc:args", "usage: out = lbmpc_mex('solve', h, dx0, dx_ref, d_off, warm, cost_shift)");
    e = lookup(get_handle(prhs[1]));
    if ((int)mxGetM(prhs[2]) != e->nx) mexErrMsgIdAndTxt("lbmpc:args", "dx0 must be nx x batch");
    batch = (int64_t)mxGetN(prhs[2]);
    uc = mxCreateDoubleMatrix((mwSize)(e->nu * e->N), (mwSize)batch, mxREAL);
    th = mxCreateDoubleMatrix((mwSize)e->nt, (mwSize)batch, mxREAL);
    d3[0] = (mwSize)e->nx; d3[1] = (mwSize)(e->N + 1); d3[2] = (mwSize)batch;
    x = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
    f = mxCreateDoubleMatrix(1, (mwSize)batch, mxREAL);
    it = mxCreateNumericMatrix(1, (mwSize)batch, mxINT32_CLASS, mxREAL);
    st = mxCreateNumericMatrix(1, (mwSize)batch, mxINT32_CLASS, mxREAL);
    /* cost_shift (nx x (N+1) x batch, optional): objective at x_k + cost_shift_k — twin sequences, DMS_LBMPC_casadi.m */
    rc = lbmpc_solve_batch_shifted(e->h, batch, dptr(prhs[2]), nrhs > 3 ? dptr(prhs[3]) : NULL,
                                   nrhs > 4 ? dptr(prhs[4]) : NULL, nrhs > 6 ? dptr(prhs[6]) : NULL,
                                   nrhs > 5 ? dptr(prhs[5]) : NULL, mxGetPr(uc), mxGetPr(th), mxGetPr(x), mxGetPr(f),
                                   (int32_t *)mxGetData(it), (int32_t *)mxGetData(st), NULL);
    if (rc != LBMPC_OK) fail_rc("lbmpc_solve_batch", rc);
    plhs[0] = mxCreateStructMatrix(1, 1, 6, names);
    mxSetField(plhs[0], 0, "u_or_c", uc); mxSetField(plhs[0], 0, "theta", th); mxSetField(plhs[0], 0, "x", x);
    mxSetField(plhs[0], 0, "f", f); mxSetField(plhs[0], 0, "iters", it); mxSetField(plhs[0], 0, "status", st);
    (void)nlhs;
}

static void cmd_oracle(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    entry_t *e;
    int64_t batch;
    int q, rc;
    mwSize d3[3];
    if (nrhs < 9) mexErrMsgIdAndTxt("lbmpc:args", "usage: d = lbmpc_mex('oracle', h, q, bandwidth, lambda, dx0, du, X, Y, valid)");
    e = lookup(get_handle(prhs[1]));
    q = (int)mxGetScalar(prhs[2]);
    batch = (int64_t)mxGetN(prhs[5]);
    d3[0] = (mwSize)e->nx; d3[1] = (mwSize)e->N; d3[2] = (mwSize)batch;
    plhs[0] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
    rc = lbmpc_oracle_apply(e->h, batch, q, mxGetScalar(prhs[3]), mxGetScalar(prhs[4]), dptr(prhs[5]), dptr(prhs[6]),
                            dptr(prhs[7]), dptr(prhs[8]), nrhs > 9 ? dptr(prhs[9]) : NULL, mxGetPr(plhs[0]), NULL);
    if (rc != LBMPC_OK) fail_rc("lbmpc_oracle_apply", rc);
    (void)nlhs;
}

static void cmd_solve_sqp(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    static const char *names[] = {"u_or_c", "theta", "x", "f", "iters", "status", "du_step"};
    entry_t *e;
    int64_t batch;
    int its, rc;
    mwSize d3[3];
    mxArray *uc, *th, *x, *f, *it, *st, *ds;
    if (nrhs < 10)
        mexErrMsgIdAndTxt("lbmpc:args", "usage: out = lbmpc_mex('solve_sqp', h, sqp_iters, q, bandwidth, lambda, dx0, dx_ref, X, Y, valid, warm, twin, order)");
    e = lookup(get_handle(prhs[1]));
    its = (int)mxGetScalar(prhs[2]);
    if ((int)mxGetM(prhs[6]) != e->nx) mexErrMsgIdAndTxt("lbmpc:args", "dx0 must be nx x batch");
    batch = (int64_t)mxGetN(prhs[6]);
    uc = mxCreateDoubleMatrix((mwSize)(e->nu * e->N), (mwSize)batch, mxREAL);
    th = mxCreateDoubleMatrix((mwSize)e->nt, (mwSize)batch, mxREAL);
    d3[0] = (mwSize)e->nx; d3[1] = (mwSize)(e->N + 1); d3[2] = (mwSize)batch;
    x = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
    f = mxCreateDoubleMatrix(1, (mwSize)batch, mxREAL);
    it = mxCreateNumericMatrix(1, (mwSize)batch, mxINT32_CLASS, mxREAL);
    st = mxCreateNumericMatrix(1, (mwSize)batch, mxINT32_CLASS, mxREAL);
    ds = mxCreateDoubleMatrix((mwSize)(its > 0 ? its : 1), (mwSize)batch, mxREAL);
    rc = lbmpc_solve_sqp_ex(e->h, batch, its, nrhs > 12 ? (int)mxGetScalar(prhs[12]) : 0, nrhs > 13 ? (int)mxGetScalar(prhs[13]) : 0,
                            (int)mxGetScalar(prhs[3]), mxGetScalar(prhs[4]), mxGetScalar(prhs[5]),
                            dptr(prhs[6]), dptr(prhs[7]), dptr(prhs[8]), dptr(prhs[9]), nrhs > 10 ? dptr(prhs[10]) : NULL,
                            nrhs > 11 ? dptr(prhs[11]) : NULL, mxGetPr(uc), mxGetPr(th), mxGetPr(x), mxGetPr(f),
                            (int32_t *)mxGetData(it), (int32_t *)mxGetData(st), mxGetPr(ds), NULL);
    if (rc != LBMPC_OK) fail_rc("lbmpc_solve_sqp_ex", rc);
    plhs[0] = mxCreateStructMatrix(1, 1, 7, names);
    mxSetField(plhs[0], 0, "u_or_c", uc); mxSetField(plhs[0], 0, "theta", th); mxSetField(plhs[0], 0, "x", x);
    mxSetField(plhs[0], 0, "f", f); mxSetField(plhs[0], 0, "iters", it); mxSetField(plhs[0], 0, "status", st);
    mxSetField(plhs[0], 0, "du_step", ds);
    (void)nlhs;
}

static void cmd_closed_loop(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    static const char *names[] = {"x", "u", "theta", "iters", "status"};
    entry_t *e;
    int64_t batch;
    int steps, q, rc;
    mwSize d3[3];
    mxArray *x, *u, *th, *it, *st;
    if (nrhs < 9)
        mexErrMsgIdAndTxt("lbmpc:args",
                          "usage: hist = lbmpc_mex('closed_loop', h, steps, q, use_oracle, warm_shift, x_eq, u_eq, x_init, wbar, seed)");
    e = lookup(get_handle(prhs[1]));
    steps = (int)mxGetScalar(prhs[2]);
    q = (int)mxGetScalar(prhs[3]);
    batch = (int64_t)mxGetN(prhs[8]);
    d3[0] = (mwSize)e->nx; d3[1] = (mwSize)(steps + 1); d3[2] = (mwSize)batch;
    x = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
    u = mxCreateDoubleMatrix((mwSize)steps, (mwSize)batch, mxREAL);
    th = mxCreateDoubleMatrix((mwSize)steps, (mwSize)batch, mxREAL);
    it = mxCreateNumericMatrix((mwSize)steps, (mwSize)batch, mxINT32_CLASS, mxREAL);
    st = mxCreateNumericMatrix((mwSize)steps, (mwSize)batch, mxINT32_CLASS, mxREAL);
    rc = lbmpc_closed_loop(e->h, batch, steps, q, (int)mxGetScalar(prhs[4]), (int)mxGetScalar(prhs[5]), dptr(prhs[6]),
                           mxGetScalar(prhs[7]), dptr(prhs[8]), nrhs > 9 ? dptr(prhs[9]) : NULL,
                           nrhs > 10 ? (uint64_t)mxGetScalar(prhs[10]) : 0, 0, mxGetPr(x), mxGetPr(u), mxGetPr(th),
                           (int32_t *)mxGetData(it), (int32_t *)mxGetData(st), NULL);
    if (rc != LBMPC_OK) fail_rc("lbmpc_closed_loop", rc);
    plhs[0] = mxCreateStructMatrix(1, 1, 5, names);
    mxSetField(plhs[0], 0, "x", x); mxSetField(plhs[0], 0, "u", u); mxSetField(plhs[0], 0, "theta", th);
    mxSetField(plhs[0], 0, "iters", it); mxSetField(plhs[0], 0, "status", st);
    (void)nlhs;
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    char cmd[MAX_CMD];
    if (nrhs < 1 || mxGetString(prhs[0], cmd, sizeof cmd))
        mexErrMsgIdAndTxt("lbmpc:args", "first argument must be a command string");
    if (strcmp(cmd, "create") == 0) cmd_create(nlhs, plhs, nrhs, prhs);
    else if (strcmp(cmd, "solve") == 0) cmd_solve(nlhs, plhs, nrhs, prhs);
    else if (strcmp(cmd, "oracle") == 0) cmd_oracle(nlhs, plhs, nrhs, prhs);
    else if (strcmp(cmd, "solve_sqp") == 0) cmd_solve_sqp(nlhs, plhs, nrhs, prhs);
    else if (strcmp(cmd, "closed_loop") == 0) cmd_closed_loop(nlhs, plhs, nrhs, prhs);
    else if (strcmp(cmd, "destroy") == 0) {
        entry_t *e;
        if (nrhs < 2) mexErrMsgIdAndTxt("lbmpc:args", "usage: lbmpc_mex('destroy', h)");
        e = lookup(get_handle(prhs[1]));
        lbmpc_destroy(e->h);
        e->h = NULL;
    } else if (strcmp(cmd, "version") == 0) {
        plhs[0] = mxCreateString(lbmpc_version());
    } else {
        mexErrMsgIdAndTxt("lbmpc:args", "unknown command '%s'", cmd);
    }
}
