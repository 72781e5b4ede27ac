/*
 * lbmpc.h — C ABI of the B200-native batched interior-point QP engine for (LB)MPC.
 *
 * Drop-in seam: the reference (bevanda/Learning-Based-MPC, 100 % MATLAB) has no FFI of its own;
 * the seam is the third-party solver call inside its closed-loop drivers
 *     opt_var = fmincon(COSTFUN,opt_var,[],[],[],[],[],[],CONSFUN,options);   functions/ocpLBMPC.m:27-31
 *                                                                           functions/ocpLMPC.m:20-24
 *     res = solver('x0',y_init,'lbx',lb,'ubx',ub,'lbg',con_lb,'ubg',con_ub);   examples/DMS_tracking_LMPC_casadi.m:163-167
 *                                                                           examples/LBMPC_casadi.m:170-174
 * Every entry point below replaces one piece of that path; the MATLAB-side binding (MEX gateway
 * + ocpLBMPC_gpu.m) is in learning-based-mpc_b200/matlab/ and described in INTEGRATION.md.
 *
 * Conventions
 *  - all matrices are COLUMN-major doubles (MATLAB layout), dimensions passed explicitly;
 *  - batch arrays are "one column per QP": dx0 is nx x batch, u_or_c is (nu*N) x batch, ...;
 *  - the caller owns every array; the handle owns device scratch sized at create;
 *  - a handle is NOT re-entrant: at most ONE call may be in flight per handle (work-queue counter, timing events and
 *    staging buffers are per handle).  Device-pointer calls are asynchronous on `stream`; issue the next call of the same
 *    handle on the SAME stream (stream order serialises them) or synchronise first.  Use one handle per host thread /
 *    stream for concurrent solves;
 *  - return value 0 = ok, negative = error (text via lbmpc_last_error()); per-QP outcome only
 *    through status[];
 *  - there is NO CPU fallback: every call fails with LBMPC_ECUDA if no sm_100 device is usable.
 */
#ifndef LBMPC_H
#define LBMPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes */
#define LBMPC_OK 0
#define LBMPC_EINVAL (-1)       /* bad argument / NULL pointer                         */
#define LBMPC_ESHAPE (-2)       /* unsupported dimensions or constraint structure      */
#define LBMPC_ECUDA (-3)        /* CUDA runtime error or no usable device              */
#define LBMPC_ENOMEM (-4)

/* problem form: which reference script family the QP reproduces */
#define LBMPC_FORM_F 0          /* fmincon scripts: opt_var=[c;theta], u = K x + c (transitionNominal.m:12),
                                   cost costLMPC.m:25-45, rows constraintsLMPC.m:20-41                      */
#define LBMPC_FORM_C 1          /* CasADi scripts: y=[x;u;theta], delta-scaled stage cost
                                   (DMS_tracking_LMPC_casadi.m:223-251), rows :254-287                       */
#define LBMPC_VARIANT_LMPC 0    /* terminal invariant set F_w_N [x_last;theta] <= h_w_N (getCONS.m:57-58)    */
#define LBMPC_VARIANT_LBMPC 1   /* robust rows on x_1: F_x_d, F_w_N (constraintsLBMPC.m:26-31,
                                   LBMPC_casadi.m:286-290)                                                   */

/* per-QP status */
#define LBMPC_ST_OPTIMAL 0
#define LBMPC_ST_MAXITER 1
#define LBMPC_ST_INFEASIBLE 2
#define LBMPC_ST_NUMERICAL 3

/* thread mapping of the solve kernel (lbmpc_set_kernel).  AUTO picks by batch size and problem shape. */
#define LBMPC_KERNEL_AUTO 0
#define LBMPC_KERNEL_WARP 1        /* one warp per QP, iterate in shared memory (throughput at moderate batches)      */
#define LBMPC_KERNEL_CTA 2         /* one 4-warp CTA per QP (latency: few QPs per SM, 616-row sets, long horizons)    */
#define LBMPC_KERNEL_STREAM 3      /* one THREAD per QP, iterate streamed from HBM (large batches, any horizon)       */
#define LBMPC_KERNEL_STREAM_MIXED 4 /* STREAM with the direction / factor records stored in FP32 ("f32+f64": iterate,
                                      residuals and all arithmetic stay FP64; reported separately, never the default) */

typedef struct lbmpc_handle lbmpc_handle;

/* The reference's model/constraint matrices, exactly the arguments of ocpLBMPC.m:1-6
 * (A,B,Kstabil,Q,R,P,T,LAMBDA,PSI,F_x,h_x,F_u,h_u,F_w_N,h_w_N,F_x_d,h_x_d). */
typedef struct lbmpc_model {
    int32_t nx, nu, nt;        /* states, inputs, dim(theta) (= m in matOCP.m:14-16)            */
    const double *A;           /* nx x nx   mgcmDLTI.m:38 / nominalModel.m:14-17                 */
    const double *B;           /* nx x nu   mgcmDLTI.m:39 / nominalModel.m:18-21                 */
    const double *K;           /* nu x nx   Kstabil, matOCP.m:7-9                                */
    const double *Q;           /* nx x nx   matOCP.m:27                                          */
    const double *R;           /* nu x nu                                                        */
    const double *P;           /* nx x nx   matOCP.m:30                                          */
    const double *T;           /* nx x nx, or 1 x 1 when T_is_scalar (matOCP.m:31: T = 1000)     */
    int32_t T_is_scalar;
    const double *LAMBDA;      /* nx x nt   matOCP.m:15                                          */
    const double *PSI;         /* nu x nt   matOCP.m:16                                          */
    const double *F_x, *h_x;   /* n_Fx x nx, n_Fx     getCONS.m:16 — one non-zero per row (box)  */
    int32_t n_Fx;
    const double *F_u, *h_u;   /* n_Fu x nu, n_Fu     getCONS.m:15 — one non-zero per row (box)  */
    int32_t n_Fu;
    const double *F_w_N, *h_w_N; /* n_Fw x (nx+nt), n_Fw   getCONS.m:57-58 / getCONSPOLY.m:69    */
    int32_t n_Fw;
    const double *F_x_d, *h_x_d; /* n_Fxd x nx, n_Fxd (LBMPC only)  getCONSPOLY.m:28-30          */
    int32_t n_Fxd;
} lbmpc_model;

typedef struct lbmpc_config {
    int32_t form;              /* LBMPC_FORM_*                                                    */
    int32_t variant;           /* LBMPC_VARIANT_*                                                 */
    int32_t N;                 /* horizon (LBMPC_RunExample.m:22; N_t/delta in the CasADi files)  */
    double delta;              /* C-form stage weight (DMS_tracking_LMPC_casadi.m:84); ignored in F-form */
    double tol_res;            /* 0 -> 1e-9   |r_p|inf and |r_d|inf / max(1, 100 |lambda|inf)      */
    double tol_mu;             /* 0 -> 1e-10  complementarity gap                                  */
    double inf_radius;         /* 0 -> auto   Farkas test: status 2 when h'lambda < 0 and
                                  1.01 sum_j |(G'lambda)_j| ybar_j <= -h'lambda in the reduced space y = [u;theta], ybar_j
                                  = input bound (10 per unbounded variable); a value R > 0 replaces the factor 1.01
                                  by R / sum_j ybar_j (i.e. R plays the role of a radius of the feasible set)  */
    int32_t max_iter;          /* 0 -> 60                                                          */
    int64_t max_batch;         /* largest batch of one solve call (device I/O staging is sized on it) */
    int32_t pointers_on_device;/* 0: batch arrays are host pointers (calls are synchronous);
                                  1: device pointers (calls are asynchronous on `stream`)          */
} lbmpc_config;

/* replaces the setup the reference does once per script before its loop (LBMPC_RunExample.m:10-85):
 * copies the model to the device and builds the canonical stage form. */
int lbmpc_create(const lbmpc_model *model, const lbmpc_config *cfg, int device, lbmpc_handle **out);

/* replaces the solver call fmincon(...) ocpLBMPC.m:31 / solver(...) DMS_tracking_LMPC_casadi.m:163-167
 * for `batch` independent QPs.
 *   dx0      nx x batch          measured state minus working point        (ocpLBMPC.m:21-23)
 *   dx_ref   nx x batch or NULL  tracking reference xs                     (costLMPC.m:38)
 *   d_off    nx x N x batch or NULL  per-stage dynamics offset d_k = oracle correction frozen on
 *                                a linearisation trajectory                (learnedModel.m:25)
 *   warm     (nu*N+nt) x batch or NULL  initial [c;theta] / [du;theta]      (opt_var, ocpLBMPC.m:31)
 *   u_or_c   (nu*N) x batch  OUT c (F-form) or du = u - u_wp (C-form)
 *   theta    nt x batch      OUT
 *   x_traj   nx x (N+1) x batch or NULL  OUT predicted states (delta coordinates)
 *   obj      batch OUT objective J (costLMPC.m / `res.f`)
 *   iters, status  batch OUT
 *   stream   cudaStream_t (may be NULL = default stream)
 * Host-pointer handles (pointers_on_device = 0): the call returns when the results are in the caller's arrays.  Page-locked
 * arrays (cudaHostAlloc / cudaHostRegister) are read and written in place by the kernel; pageable arrays of a small call
 * (<= 512 KB in total, e.g. one closed-loop step) go through a pinned mapped block of the handle, those of a large call
 * through device staging buffers sized by max_batch.  The three routes give bit-identical results. */
int lbmpc_solve_batch(lbmpc_handle *h, int64_t batch, const double *dx0, const double *dx_ref,
                      const double *d_off, const double *warm, double *u_or_c, double *theta,
                      double *x_traj, double *obj, int32_t *iters, int32_t *status, void *stream);

/* lbmpc_solve_batch with a per-stage COST SHIFT: the objective is evaluated at x_k + cost_shift_k while the dynamics and
 * every constraint row act on x_k (cost_shift: nx x (N+1) x batch, NULL = lbmpc_solve_batch).  This is the twin-sequence
 * problem of DMS_LBMPC_casadi.m:252-319 — learned states xl in costfunction (:252-268), nominal states x in
 * nonlinearconstraints (:283-319) — with the oracle frozen: e_k = xl_k - x_k obeys e_{k+1} = A e_k + g_k, e_0 = 0 and does
 * not depend on the optimisation variables (lbmpc_solve_sqp with twin = 1 builds it). */
int lbmpc_solve_batch_shifted(lbmpc_handle *h, int64_t batch, const double *dx0, const double *dx_ref,
                              const double *d_off, const double *cost_shift, const double *warm,
                              double *u_or_c, double *theta, double *x_traj, double *obj, int32_t *iters,
                              int32_t *status, void *stream);

/* The same with a state AND input shift: cost_shift_xu is (nx+nu) x (N+1) x batch, records [ex_k ; eu_k] (eu of stage N
 * ignored); the objective is evaluated at [x_k + ex_k ; theta ; u_k + eu_k].  This is what the F-form LBMPC problem needs:
 * costLBMPC.m:27 rolls the LEARNED model (transitionLearned.m:13-14, u = K x + c on the learned state) while
 * constraintsLBMPC.m:23 rolls the NOMINAL one, so with the oracle frozen the two sequences differ by e_{k+1} = (A + B K) e_k
 * + d_k in the states and by K e_k in the inputs, independently of the decision variables [c; theta]. */
int lbmpc_solve_batch_shifted_xu(lbmpc_handle *h, int64_t batch, const double *dx0, const double *dx_ref,
                                 const double *d_off, const double *cost_shift_xu, const double *warm,
                                 double *u_or_c, double *theta, double *x_traj, double *obj, int32_t *iters,
                                 int32_t *status, void *stream);

/* replaces learnedModel.m:25 + oracleL2NW.m:2-36 (mask variant casadiL2NW.m:14-28) applied along
 * the horizon: rolls the learned model x+ = A x + B u + g([x1;x2;u]) for the given input sequence
 * and returns the per-stage corrections d_k = g(.) (feed them to lbmpc_solve_batch as d_off).
 *   dx0 nx x batch; du (nu*N) x batch; X 3 x q x batch; Y nx x q x batch; valid q x batch or NULL;
 *   d_off nx x N x batch OUT.  Pointers follow cfg.pointers_on_device.
 * F-form handles: `du` holds the decision variables c and the rollout applies u = K x + c on the learned state
 * (transitionLearned.m:13-14); C-form handles: `du` is the input sequence itself (LBMPC_casadi.m:307-347). */
int lbmpc_oracle_apply(lbmpc_handle *h, int64_t batch, int32_t q, double bandwidth, double lambda,
                       const double *dx0, const double *du, const double *X, const double *Y,
                       const double *valid, double *d_off, void *stream);

/* replaces the closed-loop drivers ocpLBMPC.m:10-47 / LBMPC_casadi.m:160-223 for `batch`
 * independent scenarios: solve, apply the first move to the Moore-Greitzer plant (RK4,
 * DMS_tracking_LMPC_casadi.m:297-304), add a bounded uniform disturbance (RunExample_robust.m:250-252),
 * update the q-sample data window (get_data.m:3-9), shift the warm start (…casadi.m:187-189).
 * Warm-start shift, stated exactly: the next initial guess is [u_1 .. u_{N-1}, u_{N-1}; theta] (the last input is
 * repeated).  The reference appends K_loc * x_OL(end-n-m+1:end-m) instead (DMS_tracking_LMPC_casadi.m:187), a slice of the
 * ABSOLUTE state vector that straddles x_{N-1}(4) and x_N(1:3); that is only IPOPT's starting point and is deliberately not
 * reproduced: the initial guess changes iteration counts (by at most one in our sweeps), never the minimiser, and oracle
 * and kernels use the same rule so iteration parity stays exact.
 *   x_init nx x batch (absolute); x_eq nx; wbar nx or NULL;
 *   x_hist nx x (steps+1) x batch OUT; u_hist steps x batch OUT; theta_hist steps x batch OUT;
 *   iters_hist/status_hist steps x batch OUT (any OUT may be NULL).
 * F-form handles run the loops of ocpLBMPC.m:10-47 / ocpLMPC.m:11-40 instead: the decision variables are c and the plant input is
 * u = K (x - x_wp) + c + u_wp (transitionTrue.m:11-12); every solve starts from the previous, UNSHIFTED opt_var (ocpLBMPC.m:31;
 * warm_shift is ignored); the window follows update_data.m:3-10 and starts from one all-zero sample like the scripts; an LBMPC
 * handle with use_oracle = 1 solves each step by two first-order SQP iterations with the learned term in the cost only
 * (costLBMPC.m:27 vs constraintsLBMPC.m:23; lbmpc_solve_sqp_ex twin = 1, order = 1), otherwise by one exact QP.  u_hist holds
 * the absolute plant input, theta_hist the artificial-reference parameter.  Per-step launches only (the fused kernel is C-form). */
int lbmpc_closed_loop(lbmpc_handle *h, int64_t batch, int32_t steps, int32_t q, int32_t use_oracle,
                      int32_t warm_shift, const double *x_eq, double u_eq, const double *x_init,
                      const double *wbar, uint64_t seed, uint64_t scenario0, double *x_hist,
                      double *u_hist, double *theta_hist, int32_t *iters_hist,
                      int32_t *status_hist, void *stream);

/* next row of the scope table: the learned-oracle problem (costLBMPC.m:27, DMS_LBMPC_casadi.m:252-319 roll the
 * L2NW oracle inside the optimisation, a non-convex NLP) as a SEQUENCE of QPs.  Outer iteration j = 0..sqp_iters-1:
 *     d_k = g([x1;x2;u]_k ; X, Y) along the learned-model rollout of the previous inputs (u^-1 = warm or 0)   [lbmpc_oracle_apply]
 *     one QP with the frozen offsets d_k, started from the previous solution                                   [lbmpc_solve_batch]
 * twin = 0: ONE state sequence, the frozen d_k enter the dynamics and therefore the cost AND the rows (the C-form script
 * LBMPC_casadi.m with its learned-dynamics line :292-293 switched on).  twin = 1: the cost sees the LEARNED sequence, the
 * constraint rows and the dynamics the NOMINAL one — the twin sequences of DMS_LBMPC_casadi.m:252-319 (C-form: e_{k+1} =
 * A e_k + d_k, e_0 = 0) and the F-form problem of costLBMPC.m:25-45 / constraintsLBMPC.m:18-45 (e_{k+1} = (A + B K) e_k + d_k,
 * input gap K e_k; lbmpc_solve_batch_shifted_xu); x_traj returns the nominal sequence.  F-form handles accept twin = 1 only.
 * Outputs are those of the last QP; du_step (batch x sqp_iters, may be NULL) = |u^j - u^{j-1}|_inf per outer iteration.
 * 4-state model; array conventions as in lbmpc_solve_batch / lbmpc_oracle_apply. */
int lbmpc_solve_sqp(lbmpc_handle *h, int64_t batch, int32_t sqp_iters, int32_t twin, int32_t q, double bandwidth, double lambda,
                    const double *dx0, const double *dx_ref, const double *X, const double *Y, const double *valid,
                    const double *warm, double *u, double *theta, double *x_traj, double *obj, int32_t *iters,
                    int32_t *status, double *du_step, void *stream);

/* lbmpc_solve_sqp with the ORDER of the oracle model chosen per call.
 *   order = 0  the oracle VALUE is frozen along the previous solution (lbmpc_solve_sqp): cheap, but the fixed point of the
 *              outer iteration misses the dg/d(x,u) terms of the stationarity condition (error O(|dg/dxi|));
 *   order = 1  value AND Jacobian (the derivative CasADi's AD of casadiL2NW.m:14-28 gives IPOPT, and fmincon's finite
 *              differences of costLBMPC.m): every outer iteration solves an LTV QP on the LEARNED state sequence,
 *                  x+ = (A + [J_k(:,1:2) 0 0]) x + (B + J_k(:,3)) u + d_k,  d_k = g(xibar_k) - J_k xibar_k,
 *              i.e. a genuine SQP step whose fixed point is a stationary point of the reference's NLP.  twin = 1: the rows
 *              follow the NOMINAL sequence (constraintsLBMPC.m:23, DMS_LBMPC_casadi.m:283-319) — they act on x_k - e_k,
 *              u_k - K e_k with the gap e_k between the learned and the nominal rollout of the previous solution (exact
 *              row values at the fixed point, row Jacobians to O(|dg/dxi|)); x_traj then returns the LEARNED sequence.
 *              Reproduces the reference's saved LBMPC_N50_sys_full.mat history (40 steps, learned term active from step
 *              2) to 5e-6 on the inputs (tests/test_oracle_golden.py, tests/test_gpu_parity.py).
 * order = 1 runs on the stream mapping (the only one with per-stage dynamics). */
int lbmpc_solve_sqp_ex(lbmpc_handle *h, int64_t batch, int32_t sqp_iters, int32_t twin, int32_t order, int32_t q,
                       double bandwidth, double lambda, const double *dx0, const double *dx_ref, const double *X,
                       const double *Y, const double *valid, const double *warm, double *u, double *theta,
                       double *x_traj, double *obj, int32_t *iters, int32_t *status, double *du_step, void *stream);

/* Force a thread mapping (LBMPC_KERNEL_*) and the lock-step tick of the warp kernel (lockstep: -1 auto, 0 off, 1 on).
 * Every mapping runs the same algorithm: results agree to round-off (tests/test_gpu_parity.py forces each one).
 * The environment variables LBMPC_KERNEL=warp|cta|stream|mixed and LBMPC_LOCKSTEP=0|1 set the same two values ONCE, when
 * the handle is created. */
int lbmpc_set_kernel(lbmpc_handle *h, int32_t kernel, int32_t lockstep);
/* mapping the last solve call used (LBMPC_KERNEL_WARP / _CTA / _STREAM / _STREAM_MIXED; 0 before the first call) */
int lbmpc_last_kernel(const lbmpc_handle *h);

/* introspection used by the host mirrors, tests and bench */
int lbmpc_num_rows(const lbmpc_handle *h);            /* inequality rows m of one QP                     */
int lbmpc_slots_per_cta(const lbmpc_handle *h);       /* QPs resident per CTA (shared-memory bound)      */
int64_t lbmpc_kernel_launches(const lbmpc_handle *h); /* kernels launched by this handle so far          */
/* last solve call: device time of the IPM kernel in ms (CUDA events on the launch stream) */
float lbmpc_last_kernel_ms(lbmpc_handle *h);

/* diagnostic: SM-cycle counters (clock64) of the first warp / CTA of the solve kernel, accumulated per phase over the solve
 * calls made while enabled.  out16[0..8] = {load + rollout, E+A update + assembly, B factorisation (+ dual residual, affine
 * backward substitution, Farkas), B2 affine forward substitution, C affine rows + sigma, D corrector substitution,
 * E final rows + step length, number of iterations, store}; out16[9..15] = sub-phases (corrector sweep P1/P2/P3, polytope
 * rows and reductions of the assembly; warp-per-QP kernel only).  Reads and clears the counters, then sets `enable`. */
int lbmpc_debug_phase_cycles(lbmpc_handle *h, int enable, uint64_t *out16);

/* roofline denominator: measured FP64-FMA throughput of the device in TFLOP/s (register-resident DFMA
 * chains, best of 5 after one warm-up; MEASURED_PEAKS.json carries no FP64 entry) */
int lbmpc_measure_fp64_peak(int device, double *tflops);

void lbmpc_destroy(lbmpc_handle *h);
const char *lbmpc_last_error(void);                  /* thread-local message of the last failure         */
const char *lbmpc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LBMPC_H */
