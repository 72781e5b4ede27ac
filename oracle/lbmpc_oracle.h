/*
 * lbmpc_oracle.h — CPU FP64 restatement of the reference's per-step (LB)MPC optimisation.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (learning-based-mpc_b200/, include/)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker / reported CPU baseline.
 *
 * What it restates (reference = bevanda/Learning-Based-MPC, paths relative to matlab/LBMPC/):
 *   F-form problem  functions/costLMPC.m:20-45, costLBMPC.m:20-45, constraintsLMPC.m:18-41,
 *                   constraintsLBMPC.m:18-45, transitionNominal.m:12, models/nominalModel.m:14-28
 *   C-form problem  examples/DMS_tracking_LMPC_casadi.m:109-119,223-291, LBMPC_casadi.m:111-120,240-305
 *   oracle          functions/oracleL2NW.m:9-36, functions/casadiL2NW.m:14-28
 *   data window     utilities/update_data.m:3-10, utilities/get_data.m:3-9
 *   plant           examples/DMS_tracking_LMPC_casadi.m:215-221,297-304 (RK4, delta = 0.01)
 *   closed loop     functions/ocpLBMPC.m:10-47, functions/ocpLMPC.m:11-40,
 *                   examples/DMS_tracking_LMPC_casadi.m:153-205, LBMPC_casadi.m:160-223
 *
 * The optimiser itself is NOT in the reference tree: the reference calls MATLAB Optimization
 * Toolbox `fmincon(...,'Algorithm','sqp')` (ocpLBMPC.m:31; version unpinned, dump header says
 * R2019a) or CasADi 3.4.5 `nlpsol('solver','ipopt',nlp)` (DMS_tracking_LMPC_casadi.m:127).
 * Neither is vendored nor installable here.  The exact-QP cases are strictly convex, so the
 * minimiser is solver independent; this file solves them with a Mehrotra predictor-corrector
 * primal-dual interior-point method whose KKT systems are solved by a stage-wise Riccati
 * recursion — the algorithm the GPU kernels implement — and is pinned against the reference's
 * saved first-step results (tests/golden/reference_fixtures.npz, tests/test_oracle_golden.py).
 * Iteration counts and feasibility verdicts are DEFINED by this file (the reference discards
 * solver status).
 *
 * All matrices in this API are ROW-major (C order).
 */
#ifndef LBMPC_ORACLE_H
#define LBMPC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define LBO_FORM_F 0      /* fmincon scripts: decision vars [c;theta], u = K x + c        */
#define LBO_FORM_C 1      /* CasADi scripts : decision vars [x;u;theta], delta-scaled cost */
#define LBO_VAR_LMPC 0    /* terminal invariant set on the last constrained state          */
#define LBO_VAR_LBMPC 1   /* tightened set X(-)D and robust set on x_1                     */

#define LBO_ST_OPTIMAL 0
#define LBO_ST_MAXITER 1
#define LBO_ST_INFEASIBLE 2
#define LBO_ST_NUMERICAL 3

typedef struct lbo_problem lbo_problem;

/* Build a problem.  Dimensions: A nx*nx, B nx*nu, K nu*nx, Q nx*nx, R nu*nu, P nx*nx, T nx*nx,
 * Lam nx*nt, Psi nu*nt, Fx nFx*nx, hx nFx, Fu nFu*nu, hu nFu, Fw nFw*(nx+nt), hw nFw,
 * Fxd nFxd*nx, hxd nFxd (LBMPC only; NULL/0 otherwise).  Fx/Fu rows must have exactly one
 * non-zero (the reference builds them as [I;-I], getCONS.m:15-16).  Returns NULL on error
 * (message via lbo_last_error()). */
lbo_problem *lbo_create(int form, int variant, int nx, int nu, int nt, int N, double delta,
                        const double *A, const double *B, const double *K, const double *Q,
                        const double *R, const double *P, const double *T, const double *Lam,
                        const double *Psi, const double *Fx, const double *hx, int nFx,
                        const double *Fu, const double *hu, int nFu, const double *Fw,
                        const double *hw, int nFw, const double *Fxd, const double *hxd, int nFxd);
void lbo_destroy(lbo_problem *p);
const char *lbo_last_error(void);
void lbo_set_options(lbo_problem *p, double tol_res, double tol_mu, int max_iter, double inf_radius);
/* General custom-shape entry used for the trackingMPC double integrator
 * (trackingMPC/costFunction.m:20-39, constraintsFunction.m:20-38): stage weights and ranges given
 * explicitly.  wq/wr: N entries, kP: stage carrying the P and T terms, x rows on k in [kx0,kx1],
 * u rows on [ku0,ku1], general rows on stage kg. */
lbo_problem *lbo_create_custom(int nx, int nu, int nt, int N, const double *A, const double *B,
                               const double *Kinit, const double *Q, const double *R,
                               const double *P, const double *T, const double *Lam,
                               const double *Psi, const double *wq, const double *wr, int kP,
                               const double *lo_x, const double *hi_x, int kx0, int kx1,
                               const double *lo_u, const double *hi_u, int ku0, int ku1,
                               const double *G, const double *hg, int ng, int kg);

int lbo_num_rows(const lbo_problem *p);

/* Solve one QP.
 *  dx0   nx      measured state minus working point
 *  dx_ref nx     tracking reference (NULL = 0)      (costLMPC.m:38 `xs`)
 *  d_off nx*N    per-stage additive dynamics offset d_k (row-major [k][i]); NULL = 0
 *  warm  N*nu+nt initial [c or du ; theta] (NULL = zeros, LBMPC_RunExample.m:69-71)
 *  uc    N*nu    OUT  c (F-form) or du = u - u_wp (C-form)
 *  theta nt      OUT
 *  xtraj (N+1)*nx OUT predicted delta states (NULL to skip)
 *  obj, iters, status  OUT scalars;  stats[6] OUT {|r_d|inf, |r_p|inf, mu, flops, |lambda|inf, lambda'slack} (NULL ok)
 */
int lbo_solve(const lbo_problem *p, const double *dx0, const double *dx_ref, const double *d_off,
              const double *warm, double *uc, double *theta, double *xtraj, double *obj,
              int *iters, int *status, double *stats);

/* Batch of independent solves (POSIX threads over QPs, dynamic chunks).  Arrays are
 * batch-major: dx0[b*nx+i], d_off[b*nx*N+...], uc[b*N*nu+...]. */
int lbo_solve_batch(const lbo_problem *p, long batch, const double *dx0, const double *dx_ref,
                    const double *d_off, const double *warm, double *uc, double *theta,
                    double *xtraj, double *obj, int *iters, int *status, int nthreads);

/* Same with a per-stage COST SHIFT: the objective is evaluated at x_k + cost_shift_k (nx*(N+1), [k][i]; NULL = 0) while
 * the dynamics and every constraint row act on x_k.  This is the twin-sequence problem of DMS_LBMPC_casadi.m:252-319
 * (learned states xl in costfunction :252-268, nominal states x in nonlinearconstraints :283-319) with the oracle frozen:
 * xl_k - x_k = e_k obeys e_{k+1} = A e_k + g_k, e_0 = 0, independent of the optimisation variables. */
int lbo_solve_shifted(const lbo_problem *p, const double *dx0, const double *dx_ref, const double *d_off,
                      const double *cost_shift, const double *warm, double *uc, double *theta, double *xtraj,
                      double *obj, int *iters, int *status, double *stats);
int lbo_solve_batch_shifted(const lbo_problem *p, long batch, const double *dx0, const double *dx_ref,
                            const double *d_off, const double *cost_shift, const double *warm, double *uc,
                            double *theta, double *xtraj, double *obj, int *iters, int *status, int nthreads);

/* Same with a state AND input shift: cost_shift holds (N+1) records of cs_stride doubles, [ex_k (nx) | eu_k (nu)] when
 * cs_stride = nx+nu; the objective is evaluated at [x_k + ex_k; theta; u_k + eu_k].  F-form LBMPC (costLBMPC.m:27 learned
 * rollout with u = K x + c, constraintsLBMPC.m:23 nominal rollout): e_{k+1} = (A + B K) e_k + d_k, eu_k = K e_k. */
int lbo_solve_batch_shifted_xu(const lbo_problem *p, long batch, const double *dx0, const double *dx_ref,
                               const double *d_off, const double *cost_shift, int cs_stride, const double *warm, double *uc,
                               double *theta, double *xtraj, double *obj, int *iters, int *status, int nthreads);

/* General form: `shift` ((N+1) records of cs_stride doubles, or NULL) moves the COST (row_shift = 0, as above) or the ROWS
 * (row_shift = 1: every box / polytope row acts on x_k - ex_k, u_k - eu_k while cost and dynamics act on x_k, u_k);
 * jac (N*nx*3 per QP, or NULL): per-stage Jacobians J_k of the learned term w.r.t. xi = [x1;x2;u] -> LTV dynamics
 * A_k = A + [J_k(:,1:2) 0 0], B_k = B + J_k(:,3).  Together they state one first-order SQP step of the learned-oracle problem
 * on the LEARNED state sequence (costLBMPC.m:27 / DMS_LBMPC_casadi.m:252-268), the rows following the NOMINAL sequence
 * (constraintsLBMPC.m:23 / DMS_LBMPC_casadi.m:283-319) at the distance e_k frozen from the previous outer iteration. */
int lbo_solve_batch_ex(const lbo_problem *p, long batch, const double *dx0, const double *dx_ref,
                       const double *d_off, const double *shift, int cs_stride, int row_shift, const double *jac,
                       const double *warm, double *uc, double *theta, double *xtraj, double *obj, int *iters, int *status,
                       int nthreads);

/* L2NW oracle value and Jacobian dg/dxi (nout x nin row-major); X nin*q, Y nout*q row-major as in lbo_oracle_l2nw */
void lbo_oracle_l2nw_jac(const double *X, const double *Y, const double *valid, int q, int nin, int nout,
                         const double *xi, double bandwidth, double lambda, double *g, double *J);
/* offsets d_k = g(xibar_k) - J_k xibar_k, Jacobians J_k (N*nx*3) and optionally g(xibar_k) along the learned rollout */
void lbo_oracle_offsets_jac(const lbo_problem *p, const double *dx0, const double *du,
                            const double *X, const double *Y, const double *valid, int q,
                            double bandwidth, double lambda, double *d_off, double *jac, double *g_out);

/* Nadaraya-Watson oracle g(xi), xi=[dx1;dx2;du]  (oracleL2NW.m:26-36).  X 3*q, Y 4*q row-major
 * ([i][j] = component i of sample j), valid q or NULL (mask variant casadiL2NW.m:18-21). */
void lbo_oracle_l2nw(const double *X, const double *Y, const double *valid, int q, int nin,
                     int nout, const double *xi, double bandwidth, double lambda, double *g);

/* Moore-Greitzer plant, one RK4 step of length delta (DMS_tracking_LMPC_casadi.m:297-304). */
void lbo_plant_rk4(const double *x, double u, double delta, double *xnext);

/* Frozen-affine oracle offsets along the learned-model rollout of the sequence du (N) from dx0:
 * d_k = g([dx_k(1:2); u_k]) with dx_{k+1} = A dx_k + B u_k + d_k, u_k = du_k (C-form) or K dx_k + du_k (F-form: du holds c,
 * transitionLearned.m:13-14). */
void lbo_oracle_offsets(const lbo_problem *p, const double *dx0, const double *du,
                        const double *X, const double *Y, const double *valid, int q,
                        double bandwidth, double lambda, double *d_off);

/* Closed loop, C-form conventions (LBMPC_casadi.m:160-223 / DMS_tracking_LMPC_casadi.m:153-205):
 * RK4 plant, optional uniform disturbance |w_i| <= wbar_i (counter-based RNG, seed+scenario),
 * sliding data window of q samples, frozen-affine oracle offsets (use_oracle), warm start shift.
 * Outputs xhist (steps+1)*nx absolute states, uhist steps absolute inputs, thist steps,
 * itershist/statushist steps. */
int lbo_closed_loop(const lbo_problem *p, const double *x_eq, double u_eq, const double *x_init,
                    int steps, int q, int use_oracle, int warm_shift, const double *wbar,
                    unsigned long long seed, unsigned long long scenario, double *xhist,
                    double *uhist, double *thist, int *itershist, int *statushist);

#ifdef __cplusplus
}
#endif
#endif
