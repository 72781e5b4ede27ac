/*
 * lbmpc_oracle.c — CPU FP64 restatement of the reference's per-step (LB)MPC optimisation.
 * TEST INFRASTRUCTURE ONLY (see lbmpc_oracle.h for the reference file:line map and for why the
 * optimiser is a restated algorithm rather than a compiled reference).
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against every saved
 * first-step result the reference ships (LMPC_N{20,40,50}, LBMPC_N{40,50,60}, DMS_N50_tLMPC,
 * DMS_N50_tLBMPC_q100, DMS_tLBMPC_q100; tolerances 2e-7 abs = fmincon/IPOPT tolerance of the
 * fixtures) and against the independent dense formulation oracle/dense_mehrotra.py (1e-9).
 *
 * Algorithm (the definition of iteration counts / verdicts for the GPU path):
 *   variables  u_0..u_{N-1}, theta ; states x_k follow x_{k+1} = A x_k + B u_k + d_k exactly
 *   rows       a_i v + s_i = h_i, s_i>0, lambda_i>0 (two-sided boxes on x_k,u_k; one dense
 *              polytope block G [x_kg;theta] <= hg)
 *   start      u_k = Kinit x_k + c_k (c = warm start or 0), theta = 0, s = max(h - a v, 1), lambda = 1
 *   iterate    r_p = a v + s - h, mu = s'lambda/m, r_d = reduced gradient of the Lagrangian
 *              (adjoint recursion); stop when |r_d|_inf < tol_res max(1, 100 |lambda|_inf) (the
 *              round-off floor of r_d scales with the multipliers),
 *              |r_p|_inf < tol_res and mu < tol_mu
 *              predictor: Newton step with r_c = s.lambda ; alpha_aff ; sigma = (mu_aff/mu)^2, sigma mu >= 0.1 tol_mu
 *              corrector: r_c = s.lambda + ds_a.dl_a - sigma mu ; alpha = min(1, 0.99 alpha_max)
 *              both Newton systems solved by ONE Riccati factorisation over the augmented state
 *              z = [x;theta] (backward sweep) + two backward/forward substitution sweeps
 *   infeasible when lambda is a Farkas certificate on the box |y_j| <= ybar_j that contains the feasible set
 *              (y = [u;theta]):  h_red'lambda < 0 and 1.01 sum_j |(G_red'lambda)_j| ybar_j <= -h_red'lambda
 *              (checked once |lambda|_inf >= 1e2; ybar = input bounds, 10 per unbounded variable)
 */
#include "lbmpc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define MAXNX 6
#define MAXNT 2
#define MAXNU 2
#define MAXNZ (MAXNX + MAXNT)
#define MAXNV (MAXNZ + MAXNU)
#define MAXTYPES 8

static __thread char g_err[256];
const char *lbo_last_error(void) { return g_err; }

struct lbo_problem {
    int nx, nt, nu, nz, nv, nvb, N;
    double A[MAXNX * MAXNX], B[MAXNX * MAXNU];
    double Abar[MAXNZ * MAXNZ], Bbar[MAXNZ * MAXNU]; /* diag(A,I), [B;0] */
    double Kinit[MAXNU * MAXNX], Kout[MAXNU * MAXNX];
    int ntypes;
    double W[MAXTYPES][MAXNV * MAXNV];
    double wtuple[MAXTYPES][4];
    int *wtype; /* N+1 */
    double Lref[MAXNZ * MAXNX], Tm[MAXNX * MAXNX];
    int kT;
    double lo[MAXNX + MAXNU], hi[MAXNX + MAXNU]; /* boxes on [x;u] */
    int kx0, kx1, ku0, ku1;
    int ng, kg;
    double *G, *hg; /* ng*nz, ng */
    int m_rows;
    double tol_res, tol_mu, inf_trigger, inf_scale, inf_bound_sum;
    double fk_u[MAXNU], fk_free; /* Farkas test: upper bounds of |u_i| (inside the input-row stage range) / of a free variable */
    double fk_th[MAXNT];         /* ... of |theta_t| on the feasible set, derived from the polytope block and the state box */
    int max_iter;
};

/* ------------------------------------------------------------------------------------------ */
/* small dense helpers                                                                          */
/* ------------------------------------------------------------------------------------------ */
/* inverse of a small symmetric positive definite matrix via Cholesky; returns 0 on success */
static int spd_inv(int n, const double *M, double *Minv) {
    double L[MAXNZ * MAXNZ], Li[MAXNZ * MAXNZ];
    memset(L, 0, sizeof L);
    memset(Li, 0, sizeof Li);
    for (int j = 0; j < n; ++j) {
        double d = M[j * n + j];
        for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
        if (!(d > 0.0)) return 1;
        d = sqrt(d);
        L[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double v = M[i * n + j];
            for (int k = 0; k < j; ++k) v -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = v / d;
        }
    }
    for (int j = 0; j < n; ++j) { /* Li = L^-1 (lower) */
        Li[j * n + j] = 1.0 / L[j * n + j];
        for (int i = j + 1; i < n; ++i) {
            double v = 0.0;
            for (int k = j; k < i; ++k) v -= L[i * n + k] * Li[k * n + j];
            Li[i * n + j] = v / L[i * n + i];
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double v = 0.0;
            for (int k = (i > j ? i : j); k < n; ++k) v += Li[k * n + i] * Li[k * n + j];
            Minv[i * n + j] = v;
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* problem construction                                                                         */
/* ------------------------------------------------------------------------------------------ */
static void build_W(const lbo_problem *p, const double *Q, const double *R, const double *P,
                    const double *T, const double *Lam, const double *Psi, double wq, double wr,
                    double wp, double wt, double *W) {
    const int nx = p->nx, nt = p->nt, nu = p->nu, nv = p->nv;
    double Mx[MAXNX * MAXNV] = {0}, Mu[MAXNU * MAXNV] = {0}, Mt[MAXNX * MAXNV] = {0};
    for (int i = 0; i < nx; ++i) {
        Mx[i * nv + i] = 1.0;
        for (int j = 0; j < nt; ++j) {
            Mx[i * nv + nx + j] = -Lam[i * nt + j];
            Mt[i * nv + nx + j] = Lam[i * nt + j];
        }
    }
    for (int i = 0; i < nu; ++i) {
        Mu[i * nv + nx + nt + i] = 1.0;
        for (int j = 0; j < nt; ++j) Mu[i * nv + nx + j] = -Psi[i * nt + j];
    }
    for (int a = 0; a < nv; ++a)
        for (int b = 0; b < nv; ++b) {
            double v = 0.0;
            for (int i = 0; i < nx; ++i)
                for (int j = 0; j < nx; ++j) {
                    v += wq * Mx[i * nv + a] * Q[i * nx + j] * Mx[j * nv + b];
                    v += wp * Mx[i * nv + a] * P[i * nx + j] * Mx[j * nv + b];
                    v += wt * Mt[i * nv + a] * T[i * nx + j] * Mt[j * nv + b];
                }
            for (int i = 0; i < nu; ++i)
                for (int j = 0; j < nu; ++j)
                    v += wr * Mu[i * nv + a] * R[i * nu + j] * Mu[j * nv + b];
            W[a * nv + b] = 2.0 * v;
        }
}

static lbo_problem *alloc_problem(int nx, int nu, int nt, int N) {
    if (nx < 1 || nx > MAXNX || nu < 1 || nu > MAXNU || nt < 1 || nt > MAXNT || N < 3) {
        snprintf(g_err, sizeof g_err, "unsupported dims nx=%d nu=%d nt=%d N=%d", nx, nu, nt, N);
        return NULL;
    }
    lbo_problem *p = (lbo_problem *)calloc(1, sizeof *p);
    p->nx = nx; p->nu = nu; p->nt = nt; p->nz = nx + nt; p->nv = nx + nt + nu; p->nvb = nx + nu;
    p->N = N;
    p->wtype = (int *)calloc((size_t)N + 1, sizeof(int));
    p->tol_res = 1e-9; p->tol_mu = 1e-10; p->inf_trigger = 1e2;
    p->max_iter = 60;
    return p;
}

static void set_dynamics(lbo_problem *p, const double *A, const double *B) {
    const int nx = p->nx, nu = p->nu, nz = p->nz;
    memcpy(p->A, A, sizeof(double) * nx * nx);
    memcpy(p->B, B, sizeof(double) * nx * nu);
    memset(p->Abar, 0, sizeof p->Abar);
    memset(p->Bbar, 0, sizeof p->Bbar);
    for (int i = 0; i < nx; ++i) {
        for (int j = 0; j < nx; ++j) p->Abar[i * nz + j] = A[i * nx + j];
        for (int j = 0; j < nu; ++j) p->Bbar[i * nu + j] = B[i * nu + j];
    }
    for (int i = nx; i < nz; ++i) p->Abar[i * nz + i] = 1.0;
}

static int stage_type(lbo_problem *p, const double *Q, const double *R, const double *P,
                      const double *T, const double *Lam, const double *Psi, double wq, double wr,
                      double wp, double wt) {
    for (int t = 0; t < p->ntypes; ++t)
        if (p->wtuple[t][0] == wq && p->wtuple[t][1] == wr && p->wtuple[t][2] == wp &&
            p->wtuple[t][3] == wt)
            return t;
    if (p->ntypes == MAXTYPES) return -1;
    int t = p->ntypes++;
    p->wtuple[t][0] = wq; p->wtuple[t][1] = wr; p->wtuple[t][2] = wp; p->wtuple[t][3] = wt;
    build_W(p, Q, R, P, T, Lam, Psi, wq, wr, wp, wt, p->W[t]);
    return t;
}

static void count_rows(lbo_problem *p) {
    int m = p->ng;
    for (int k = 0; k <= p->N; ++k)
        for (int j = 0; j < p->nvb; ++j) {
            int in = (j < p->nx) ? (k >= p->kx0 && k <= p->kx1) : (k >= p->ku0 && k <= p->ku1);
            if (!in) continue;
            m += isfinite(p->hi[j]) ? 1 : 0;
            m += isfinite(p->lo[j]) ? 1 : 0;
        }
    p->m_rows = m;
    /* Farkas test: per-variable upper bounds ybar_j of |y_j| over the feasible set (input box bounds where they
     * exist, 10 per unbounded variable), 1 % margin for round-off:
     *     h_red'lambda < 0  and  1.01 sum_j |(G_red'lambda)_j| ybar_j <= -h_red'lambda
     * (raising the box multipliers of u_j by |(G_red'lambda)_j| cancels that component exactly at the price ybar_j each,
     *  so the test says: the corrected lambda is an exact Farkas certificate)                                 */
    /* theta: every polytope row with ONE non-zero theta coefficient g_t bounds that component once x_kg is boxed:
     *     g_t theta_t <= hg_i + sum_j |G_ij| max(|lo_j|, |hi_j|)
     * the tightest upper and lower bound over the rows give ybar_theta (10 if a side stays unbounded).  For the reference's
     * sets this is the exact range of theta on the feasible set (0.98 for the Moore-Greitzer model), so models with a larger
     * steady-state range are judged by their own range, not by a constant. */
    double R = 0.0;
    for (int t = 0; t < p->nt; ++t) {
        double ub = INFINITY, lb = -INFINITY;
        const int boxed = p->kg >= p->kx0 && p->kg <= p->kx1;
        for (int i = 0; i < p->ng && boxed; ++i) {
            const double *g = p->G + (size_t)i * p->nz;
            int others = 0;
            for (int a = 0; a < p->nt; ++a) others += (a != t && g[p->nx + a] != 0.0);
            const double gt = g[p->nx + t];
            if (others || gt == 0.0) continue;
            double rhs = p->hg[i];
            for (int j = 0; j < p->nx; ++j)
                if (g[j] != 0.0) rhs += fabs(g[j]) * fmax(fabs(p->lo[j]), fabs(p->hi[j]));
            if (!isfinite(rhs)) continue;
            if (gt > 0.0) { if (rhs / gt < ub) ub = rhs / gt; }
            else          { if (rhs / gt > lb) lb = rhs / gt; }
        }
        p->fk_th[t] = (isfinite(ub) && isfinite(lb)) ? fmax(fabs(ub), fabs(lb)) : 10.0;
        R += p->fk_th[t];
    }
    p->fk_free = 10.0;
    for (int j = p->nx; j < p->nvb; ++j) {
        const double b = fmax(fabs(p->lo[j]), fabs(p->hi[j]));
        p->fk_u[j - p->nx] = isfinite(b) ? b : 10.0;
    }
    for (int k = 0; k < p->N; ++k)
        for (int j = p->nx; j < p->nvb; ++j) R += (k >= p->ku0 && k <= p->ku1) ? p->fk_u[j - p->nx] : 10.0;
    p->inf_bound_sum = R;
    p->inf_scale = 1.01;
}

static void set_ref_terms(lbo_problem *p, const double *T, const double *Lam, int kT) {
    const int nx = p->nx, nt = p->nt;
    memcpy(p->Tm, T, sizeof(double) * nx * nx);
    memset(p->Lref, 0, sizeof p->Lref);
    for (int a = 0; a < nt; ++a) /* lin_theta = -2 Lam' T x_ref */
        for (int j = 0; j < nx; ++j) {
            double v = 0.0;
            for (int i = 0; i < nx; ++i) v += Lam[i * nt + a] * T[i * nx + j];
            p->Lref[(nx + a) * nx + j] = -2.0 * v;
        }
    p->kT = kT;
}

/* boxes from H-rep rows with a single non-zero each (getCONS.m:15-16) */
static int boxes_from_hrep(const double *F, const double *h, int nrow, int nvar, double *lo,
                           double *hi) {
    for (int r = 0; r < nrow; ++r) {
        int col = -1;
        for (int j = 0; j < nvar; ++j)
            if (F[r * nvar + j] != 0.0) {
                if (col >= 0) return 1;
                col = j;
            }
        if (col < 0) return 1;
        double c = F[r * nvar + col], b = h[r] / c;
        if (c > 0) { if (b < hi[col]) hi[col] = b; }
        else       { if (b > lo[col]) lo[col] = b; }
    }
    return 0;
}

lbo_problem *lbo_create(int form, int variant, int nx, int nu, int nt, int N, double delta,
                        const double *A, const double *B, const double *K, const double *Q,
                        const double *R, const double *P, const double *T, const double *Lam,
                        const double *Psi, const double *Fx, const double *hx, int nFx,
                        const double *Fu, const double *hu, int nFu, const double *Fw,
                        const double *hw, int nFw, const double *Fxd, const double *hxd, int nFxd) {
    lbo_problem *p = alloc_problem(nx, nu, nt, N);
    if (!p) return NULL;
    set_dynamics(p, A, B);
    if (form == LBO_FORM_F) { /* u = K x + c  (transitionNominal.m:12) */
        memcpy(p->Kinit, K, sizeof(double) * nu * nx);
        memcpy(p->Kout, K, sizeof(double) * nu * nx);
    }
    for (int j = 0; j < nx + nu; ++j) { p->lo[j] = -INFINITY; p->hi[j] = INFINITY; }
    if (boxes_from_hrep(Fx, hx, nFx, nx, p->lo, p->hi) ||
        boxes_from_hrep(Fu, hu, nFu, nu, p->lo + nx, p->hi + nx)) {
        snprintf(g_err, sizeof g_err, "F_x / F_u rows must have exactly one non-zero (box rows)");
        lbo_destroy(p);
        return NULL;
    }
    for (int k = 0; k <= N; ++k) {
        double wq = 0, wr = 0, wp = 0, wt = 0;
        if (k == N) { wp = 1; wt = 1; }
        else if (form == LBO_FORM_C) { wq = delta; wr = delta; }   /* …casadi.m:233-237 */
        else if (k <= N - 3) { wq = 1; wr = 1; }                   /* costLMPC.m:30 "k < N-1" */
        int t = stage_type(p, Q, R, P, T, Lam, Psi, wq, wr, wp, wt);
        p->wtype[k] = t;
    }
    set_ref_terms(p, T, Lam, N);
    if (form == LBO_FORM_C) { p->kx0 = 1; p->kx1 = N; p->ku0 = 0; p->ku1 = N - 1; }
    else { p->kx0 = 1; p->kx1 = N - 1; p->ku0 = 0; p->ku1 = N - 2; } /* constraintsLMPC.m:21 */
    const int nz = p->nz;
    if (variant == LBO_VAR_LMPC) {
        p->kg = (form == LBO_FORM_C) ? N : N - 1; /* constraintsLMPC.m:36-38 reuses x_{N-1} */
        p->ng = nFw;
        p->G = (double *)malloc(sizeof(double) * (size_t)nFw * nz);
        p->hg = (double *)malloc(sizeof(double) * (size_t)nFw);
        memcpy(p->G, Fw, sizeof(double) * (size_t)nFw * nz);
        memcpy(p->hg, hw, sizeof(double) * (size_t)nFw);
    } else { /* constraintsLBMPC.m:26-31, LBMPC_casadi.m:286-290: [F_x_d 0; F_w_N] on x_1 */
        p->kg = 1;
        p->ng = nFxd + nFw;
        p->G = (double *)calloc((size_t)p->ng * nz, sizeof(double));
        p->hg = (double *)malloc(sizeof(double) * (size_t)p->ng);
        for (int r = 0; r < nFxd; ++r) {
            for (int j = 0; j < nx; ++j) p->G[r * nz + j] = Fxd[r * nx + j];
            p->hg[r] = hxd[r];
        }
        memcpy(p->G + (size_t)nFxd * nz, Fw, sizeof(double) * (size_t)nFw * nz);
        memcpy(p->hg + nFxd, hw, sizeof(double) * (size_t)nFw);
    }
    count_rows(p);
    return p;
}

lbo_problem *lbo_create_custom(int nx, int nu, int nt, int N, const double *A, const double *B,
                               const double *Kinit, const double *Q, const double *R,
                               const double *P, const double *T, const double *Lam,
                               const double *Psi, const double *wq, const double *wr, int kP,
                               const double *lo_x, const double *hi_x, int kx0, int kx1,
                               const double *lo_u, const double *hi_u, int ku0, int ku1,
                               const double *G, const double *hg, int ng, int kg) {
    lbo_problem *p = alloc_problem(nx, nu, nt, N);
    if (!p) return NULL;
    set_dynamics(p, A, B);
    if (Kinit) memcpy(p->Kinit, Kinit, sizeof(double) * nu * nx);
    for (int j = 0; j < nx; ++j) { p->lo[j] = lo_x[j]; p->hi[j] = hi_x[j]; }
    for (int j = 0; j < nu; ++j) { p->lo[nx + j] = lo_u[j]; p->hi[nx + j] = hi_u[j]; }
    for (int k = 0; k <= N; ++k) {
        double q = (k < N) ? wq[k] : 0.0, r = (k < N) ? wr[k] : 0.0, pp = (k == kP) ? 1.0 : 0.0;
        int t = stage_type(p, Q, R, P, T, Lam, Psi, q, r, pp, pp);
        if (t < 0) { snprintf(g_err, sizeof g_err, "too many stage cost types"); lbo_destroy(p); return NULL; }
        p->wtype[k] = t;
    }
    set_ref_terms(p, T, Lam, kP);
    p->kx0 = kx0; p->kx1 = kx1; p->ku0 = ku0; p->ku1 = ku1;
    p->ng = ng; p->kg = kg;
    p->G = (double *)malloc(sizeof(double) * (size_t)(ng > 0 ? ng : 1) * p->nz);
    p->hg = (double *)malloc(sizeof(double) * (size_t)(ng > 0 ? ng : 1));
    if (ng > 0) {
        memcpy(p->G, G, sizeof(double) * (size_t)ng * p->nz);
        memcpy(p->hg, hg, sizeof(double) * (size_t)ng);
    }
    count_rows(p);
    return p;
}

void lbo_destroy(lbo_problem *p) {
    if (!p) return;
    free(p->wtype); free(p->G); free(p->hg); free(p);
}

void lbo_set_options(lbo_problem *p, double tol_res, double tol_mu, int max_iter, double inf_radius) {
    if (tol_res > 0) p->tol_res = tol_res;
    if (tol_mu > 0) p->tol_mu = tol_mu;
    if (max_iter > 0) p->max_iter = max_iter;
    if (inf_radius > 0) p->inf_scale = inf_radius / p->inf_bound_sum; /* a caller-supplied radius rescales the bounds */
}

int lbo_num_rows(const lbo_problem *p) { return p->m_rows; }

/* ------------------------------------------------------------------------------------------ */
/* solver workspace                                                                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double *x, *u, th[MAXNT];                /* iterate */
    double *sb, *lb;                         /* box rows [(k*nvb+j)*2+side] */
    double *sg, *lg;                         /* general rows */
    double *rpb, *rpg;                       /* primal residuals */
    double *ccb, *ccg;                       /* corrector terms ds_a*dl_a */
    double *Qd;                              /* (N+1)*nvb barrier diagonal */
    double *gres;                            /* (N+1)*nv : cost gradient + G'lambda */
    double *gcon;                            /* (N+1)*nv : G'lambda only (certificate) */
    double *dgp;                             /* (N+1)*nvb: (newton rhs) - (G'lambda) on box vars */
    double HG[MAXNZ * MAXNZ], gGl[MAXNZ], dG[MAXNZ];
    double *L, *Ri, *kap;                    /* factors: N*nu*nz, N*nu*nu, N*nu */
    double P0tt_inv[MAXNT * MAXNT];
    double *dx, *du, dth[MAXNT];
    double lin[MAXNZ], cconst;
    const double *csh;                       /* (N+1)*cs_stride or NULL: the cost is evaluated at [x_k + ex_k; u_k + eu_k] (twin sequences) */
    int cs_stride;                           /* nx: state shift only; nx+nu: [ex_k | eu_k] */
    int row_shift;                           /* 0: the shift moves the COST; 1: it moves the ROWS (rows act on x_k - ex_k, u_k - eu_k:
                                                first-order SQP on the learned sequence, the rows follow the nominal one) */
    const double *jac;                       /* N*nx*3 or NULL: per-stage Jacobian J_k of the learned term w.r.t. [x1;x2;u]:
                                                A_k = A + [J(:,1:2) 0 0], B_k = B + J(:,3) (LTV dynamics) */
    double flops;
} lbo_ws;

static lbo_ws *ws_alloc(const lbo_problem *p) {
    lbo_ws *w = (lbo_ws *)calloc(1, sizeof *w);
    const size_t N1 = (size_t)p->N + 1;
    w->x = calloc(N1 * p->nx, 8); w->u = calloc(N1 * p->nu, 8);
    w->sb = calloc(N1 * p->nvb * 2, 8); w->lb = calloc(N1 * p->nvb * 2, 8);
    w->rpb = calloc(N1 * p->nvb * 2, 8); w->ccb = calloc(N1 * p->nvb * 2, 8);
    size_t ng = p->ng > 0 ? p->ng : 1;
    w->sg = calloc(ng, 8); w->lg = calloc(ng, 8); w->rpg = calloc(ng, 8); w->ccg = calloc(ng, 8);
    w->Qd = calloc(N1 * p->nvb, 8); w->gres = calloc(N1 * p->nv, 8); w->gcon = calloc(N1 * p->nv, 8);
    w->dgp = calloc(N1 * p->nvb, 8);
    w->L = calloc(N1 * p->nu * p->nz, 8); w->Ri = calloc(N1 * p->nu * p->nu, 8);
    w->kap = calloc(N1 * p->nu, 8);
    w->dx = calloc(N1 * p->nx, 8); w->du = calloc(N1 * p->nu, 8);
    return w;
}
static void ws_free(lbo_ws *w) {
    free(w->x); free(w->u); free(w->sb); free(w->lb); free(w->rpb); free(w->ccb); free(w->sg);
    free(w->lg); free(w->rpg); free(w->ccg); free(w->Qd); free(w->gres); free(w->gcon);
    free(w->dgp); free(w->L); free(w->Ri); free(w->kap); free(w->dx); free(w->du); free(w);
}

static inline int row_on(const lbo_problem *p, int k, int j, int side) {
    int in = (j < p->nx) ? (k >= p->kx0 && k <= p->kx1) : (k >= p->ku0 && k <= p->ku1 && k < p->N);
    if (!in) return 0;
    return side == 0 ? isfinite(p->hi[j]) : isfinite(p->lo[j]);
}
static inline double var_at(const lbo_problem *p, const lbo_ws *w, int k, int j) {
    return j < p->nx ? w->x[k * p->nx + j] : w->u[k * p->nu + (j - p->nx)];
}
static inline double dvar_at(const lbo_problem *p, const lbo_ws *w, int k, int j) {
    return j < p->nx ? w->dx[k * p->nx + j] : w->du[k * p->nu + (j - p->nx)];
}
/* slack of a box row at the current iterate: upper hi - v, lower v - lo */
static inline double row_shift_at(const lbo_problem *p, const lbo_ws *w, int k, int j);
static inline double box_slack(const lbo_problem *p, const lbo_ws *w, int k, int j, int side) {
    double v = var_at(p, w, k, j) - row_shift_at(p, w, k, j);
    return side == 0 ? p->hi[j] - v : v - p->lo[j];
}
static double gen_dot(const lbo_problem *p, int i, const double *xk, const double *th);
/* slack of polytope row i at the current iterate (rows follow x_kg - ex_kg when the shift moves the rows) */
static double gen_slack_ws(const lbo_problem *p, const lbo_ws *w, int i) {
    double xs[MAXNX];
    for (int j = 0; j < p->nx; ++j) xs[j] = w->x[(size_t)p->kg * p->nx + j] - row_shift_at(p, w, p->kg, j);
    return p->hg[i] - gen_dot(p, i, xs, w->th);
}
static double gen_dot(const lbo_problem *p, int i, const double *xk, const double *th) {
    const double *g = p->G + (size_t)i * p->nz;
    double v = 0.0;
    for (int j = 0; j < p->nx; ++j) v += g[j] * xk[j];
    for (int j = 0; j < p->nt; ++j) v += g[p->nx + j] * th[j];
    return v;
}

/* dynamics of stage k: LTI (p->A, p->B) or LTV with the oracle Jacobian (nu = 1, nx >= 2) */
static void stage_ab(const lbo_problem *p, const lbo_ws *w, int k, double *Ak, double *Bk, double *Abar, double *Bbar) {
    const int nx = p->nx, nu = p->nu, nz = p->nz;
    for (int a = 0; a < nx; ++a) {
        for (int b = 0; b < nx; ++b) Ak[a * nx + b] = p->A[a * nx + b] + ((w->jac && b < 2) ? w->jac[((size_t)k * nx + a) * 3 + b] : 0.0);
        for (int i = 0; i < nu; ++i) Bk[a * nu + i] = p->B[a * nu + i] + ((w->jac && i == 0) ? w->jac[((size_t)k * nx + a) * 3 + 2] : 0.0);
    }
    if (Abar) {
        memset(Abar, 0, sizeof(double) * nz * nz);
        memset(Bbar, 0, sizeof(double) * nz * nu);
        for (int a = 0; a < nx; ++a) {
            for (int b = 0; b < nx; ++b) Abar[a * nz + b] = Ak[a * nx + b];
            for (int i = 0; i < nu; ++i) Bbar[a * nu + i] = Bk[a * nu + i];
        }
        for (int a = nx; a < nz; ++a) Abar[a * nz + a] = 1.0;
    }
}
/* shift of bounded variable j at stage k when the shift moves the rows */
static inline double row_shift_at(const lbo_problem *p, const lbo_ws *w, int k, int j) {
    if (!w->csh || !w->row_shift) return 0.0;
    if (j < p->nx) return w->csh[k * w->cs_stride + j];
    return (w->cs_stride > p->nx && k < p->N) ? w->csh[k * w->cs_stride + j] : 0.0;
}

/* ------------------------------------------------------------------------------------------ */
/* assembly of barrier terms and gradients                                                      */
/* corr = 0: predictor rhs  t_i = w_i r_p,i ; corr = 1: t_i = w_i r_p,i - (cc_i - sigma mu)/s_i */
/* ------------------------------------------------------------------------------------------ */
static void assemble(const lbo_problem *p, lbo_ws *w, int corr, double sigmu, double *rp_inf,
                     double *mu, double *lam_inf, double *hlam) {
    const int nx = p->nx, nu = p->nu, nz = p->nz, nv = p->nv, nvb = p->nvb, N = p->N;
    double rpi = 0.0, sl = 0.0, li = 0.0, hl = 0.0;
    if (!corr) {
        /* cost gradient  W_k v_k (+ lin at kT) */
        for (int k = 0; k <= N; ++k) {
            const double *W = p->W[p->wtype[k]];
            double v[MAXNV];
            const int cshift = w->csh && !w->row_shift;
            for (int j = 0; j < nx; ++j) v[j] = w->x[k * nx + j] + (cshift ? w->csh[k * w->cs_stride + j] : 0.0);
            for (int j = 0; j < p->nt; ++j) v[nx + j] = w->th[j];
            for (int j = 0; j < nu; ++j)
                v[nz + j] = (k < N) ? w->u[k * nu + j] + ((cshift && w->cs_stride > nx) ? w->csh[k * w->cs_stride + nx + j] : 0.0) : 0.0;
            int lim = (k < N) ? nv : nz;
            for (int a = 0; a < nv; ++a) {
                double g = 0.0;
                if (a < lim)
                    for (int b = 0; b < lim; ++b) g += W[a * nv + b] * v[b];
                if (k == p->kT && a < nz) g += w->lin[a];
                w->gres[k * nv + a] = g;
                w->gcon[k * nv + a] = 0.0;
            }
            w->flops += 2.0 * lim * lim;
        }
        memset(w->HG, 0, sizeof w->HG);
        memset(w->gGl, 0, sizeof w->gGl);
    }
    memset(w->dG, 0, sizeof w->dG);
    for (int k = 0; k <= N; ++k)
        for (int j = 0; j < nvb; ++j) {
            double qd = 0.0, gl = 0.0, gp = 0.0;
            for (int side = 0; side < 2; ++side) {
                if (!row_on(p, k, j, side)) continue;
                const int r = (k * nvb + j) * 2 + side;
                const double sgn = side == 0 ? 1.0 : -1.0;
                const double s = w->sb[r], l = w->lb[r];
                double t;
                if (!corr) {
                    const double slack = box_slack(p, w, k, j, side);
                    const double rp = s - slack;
                    w->rpb[r] = rp;
                    if (fabs(rp) > rpi) rpi = fabs(rp);
                    sl += s * l; if (l > li) li = l; hl += l * slack;
                    t = (l / s) * rp;
                } else {
                    t = (l / s) * w->rpb[r] - (w->ccb[r] - sigmu) / s;
                }
                qd += l / s; gl += sgn * l; gp += sgn * t;
                w->flops += 8.0;
            }
            const int a = (j < nx) ? j : nz + (j - nx);
            if (!corr) {
                w->Qd[k * nvb + j] = qd;
                w->gres[k * nv + a] += gl;
                w->gcon[k * nv + a] += gl;
            }
            w->dgp[k * nvb + j] = gp - gl;
        }
    for (int i = 0; i < p->ng; ++i) {
        const double *g = p->G + (size_t)i * nz;
        const double s = w->sg[i], l = w->lg[i];
        double t;
        if (!corr) {
            const double slack = gen_slack_ws(p, w, i);
            const double rp = s - slack;
            w->rpg[i] = rp;
            if (fabs(rp) > rpi) rpi = fabs(rp);
            sl += s * l; if (l > li) li = l; hl += l * slack;
            const double wi = l / s;
            t = wi * rp;
            for (int a = 0; a < nz; ++a) {
                for (int b = 0; b < nz; ++b) w->HG[a * nz + b] += wi * g[a] * g[b];
                w->gGl[a] += g[a] * l;
            }
            w->flops += 2.0 * nz + 4.0 + nz * (nz + 1.0) + 2.0 * nz;
        } else {
            t = (l / s) * w->rpg[i] - (w->ccg[i] - sigmu) / s;
        }
        for (int a = 0; a < nz; ++a) w->dG[a] += g[a] * (t - l);
        w->flops += 2.0 * nz + 4.0;
    }
    if (!corr) {
        *rp_inf = rpi; *mu = sl / p->m_rows; *lam_inf = li; *hlam = hl;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* backward sweep.  factor=1: Riccati factorisation (stores L, Ri) + adjoint recursion for r_d   */
/* (+ certificate adjoint when cert!=NULL).  Always: gradient recursion -> kap, dth.             */
/* returns non-zero if a pivot is not positive                                                   */
/* ------------------------------------------------------------------------------------------ */
static int backward(const lbo_problem *p, lbo_ws *w, int factor, double *rd_inf, double *cert) {
    const int nx = p->nx, nt = p->nt, nu = p->nu, nz = p->nz, nv = p->nv, nvb = p->nvb, N = p->N;
    double P[MAXNZ * MAXNZ], pv[MAXNZ], pi[MAXNZ], pc[MAXNZ];
    double rdi = 0.0, ci = 0.0, ydot = 0.0;
    /* terminal stage */
    {
        const double *W = p->W[p->wtype[N]];
        for (int a = 0; a < nz; ++a) {
            for (int b = 0; b < nz; ++b) P[a * nz + b] = W[a * nv + b];
            if (a < nx) P[a * nz + a] += w->Qd[N * nvb + a];
            pv[a] = w->gres[N * nv + a] + (a < nx ? w->dgp[N * nvb + a] : 0.0);
            pi[a] = w->gres[N * nv + a];
            pc[a] = w->gcon[N * nv + a];
        }
        if (p->kg == N)
            for (int a = 0; a < nz; ++a) {
                for (int b = 0; b < nz; ++b) P[a * nz + b] += w->HG[a * nz + b];
                pv[a] += w->gGl[a] + w->dG[a];
                pi[a] += w->gGl[a];
                pc[a] += w->gGl[a];
            }
    }
    for (int k = N - 1; k >= 0; --k) {
        const double *W = p->W[p->wtype[k]];
        double *L = w->L + (size_t)k * nu * nz, *Ri = w->Ri + (size_t)k * nu * nu;
        double Ak_[MAXNX * MAXNX], Bk_[MAXNX * MAXNU], Abar_k[MAXNZ * MAXNZ], Bbar_k[MAXNZ * MAXNU];
        stage_ab(p, w, k, Ak_, Bk_, Abar_k, Bbar_k);
        if (factor) {
            double M[MAXNZ * MAXNZ], F[MAXNZ * MAXNZ], Rt[MAXNU * MAXNU], PB[MAXNZ * MAXNU];
            for (int a = 0; a < nz; ++a)
                for (int b = 0; b < nz; ++b) {
                    double v = 0.0;
                    for (int c = 0; c < nz; ++c) v += P[a * nz + c] * Abar_k[c * nz + b];
                    M[a * nz + b] = v;
                }
            for (int a = 0; a < nz; ++a)
                for (int b = 0; b < nz; ++b) {
                    double v = 0.0;
                    for (int c = 0; c < nz; ++c) v += Abar_k[c * nz + a] * M[c * nz + b];
                    F[a * nz + b] = v;
                }
            for (int i = 0; i < nu; ++i)
                for (int b = 0; b < nz; ++b) {
                    double v = W[(nz + i) * nv + b];
                    for (int c = 0; c < nz; ++c) v += Bbar_k[c * nu + i] * M[c * nz + b];
                    L[i * nz + b] = v;
                }
            for (int a = 0; a < nz; ++a)
                for (int i = 0; i < nu; ++i) {
                    double v = 0.0;
                    for (int c = 0; c < nz; ++c) v += P[a * nz + c] * Bbar_k[c * nu + i];
                    PB[a * nu + i] = v;
                }
            for (int i = 0; i < nu; ++i)
                for (int j = 0; j < nu; ++j) {
                    double v = W[(nz + i) * nv + nz + j];
                    if (i == j) v += w->Qd[k * nvb + nx + i];
                    for (int c = 0; c < nz; ++c) v += Bbar_k[c * nu + i] * PB[c * nu + j];
                    Rt[i * nu + j] = v;
                }
            if (spd_inv(nu, Rt, Ri)) return 1;
            for (int a = 0; a < nz; ++a)
                for (int b = 0; b < nz; ++b) {
                    double v = W[a * nv + b] + F[a * nz + b];
                    for (int i = 0; i < nu; ++i)
                        for (int j = 0; j < nu; ++j) v -= L[i * nz + a] * Ri[i * nu + j] * L[j * nz + b];
                    P[a * nz + b] = v;
                }
            for (int a = 0; a < nx; ++a) P[a * nz + a] += w->Qd[k * nvb + a];
            if (p->kg == k)
                for (int a = 0; a < nz * nz; ++a) P[a] += w->HG[a];
            /* symmetrise (kills round-off drift) */
            for (int a = 0; a < nz; ++a)
                for (int b = a + 1; b < nz; ++b) {
                    double v = 0.5 * (P[a * nz + b] + P[b * nz + a]);
                    P[a * nz + b] = v; P[b * nz + a] = v;
                }
            w->flops += 4.0 * nz * nz * nz + 4.0 * nz * nz * nu + 2.0 * nz * nu * nu + 2.0 * nz * nz * nu * nu;
        }
        /* gradient recursion */
        double rt[MAXNU], kap[MAXNU], Atp[MAXNZ];
        for (int i = 0; i < nu; ++i) {
            double v = w->gres[k * nv + nz + i] + w->dgp[k * nvb + nx + i];
            for (int c = 0; c < nz; ++c) v += Bbar_k[c * nu + i] * pv[c];
            rt[i] = v;
        }
        for (int i = 0; i < nu; ++i) {
            double v = 0.0;
            for (int j = 0; j < nu; ++j) v -= Ri[i * nu + j] * rt[j];
            kap[i] = v;
            w->kap[k * nu + i] = v;
        }
        for (int a = 0; a < nz; ++a) {
            double v = 0.0;
            for (int c = 0; c < nz; ++c) v += Abar_k[c * nz + a] * pv[c];
            Atp[a] = v;
        }
        for (int a = 0; a < nz; ++a) {
            double v = w->gres[k * nv + a] + (a < nx ? w->dgp[k * nvb + a] : 0.0) + Atp[a];
            for (int i = 0; i < nu; ++i) v += L[i * nz + a] * kap[i];
            pv[a] = v;
        }
        if (p->kg == k)
            for (int a = 0; a < nz; ++a) pv[a] += w->gGl[a] + w->dG[a];
        w->flops += 2.0 * nz * nz + 4.0 * nz * nu + 2.0 * nu * nu;
        if (factor) { /* adjoint recursion: reduced gradient of the Lagrangian */
            double npi[MAXNZ];
            for (int i = 0; i < nu; ++i) {
                double v = w->gres[k * nv + nz + i];
                for (int c = 0; c < nz; ++c) v += Bbar_k[c * nu + i] * pi[c];
                if (fabs(v) > rdi) rdi = fabs(v);
                if (v != v) rdi = NAN;
            }
            for (int a = 0; a < nz; ++a) {
                double v = w->gres[k * nv + a];
                for (int c = 0; c < nz; ++c) v += Abar_k[c * nz + a] * pi[c];
                npi[a] = v;
            }
            memcpy(pi, npi, sizeof(double) * nz);
            if (p->kg == k)
                for (int a = 0; a < nz; ++a) pi[a] += w->gGl[a];
            w->flops += 2.0 * nz * nz + 2.0 * nz * nu;
            if (cert) {
                for (int i = 0; i < nu; ++i) {
                    double v = w->gcon[k * nv + nz + i];
                    for (int c = 0; c < nz; ++c) v += Bbar_k[c * nu + i] * pc[c];
                    ci += fabs(v) * ((k >= p->ku0 && k <= p->ku1) ? p->fk_u[i] : p->fk_free);
                    ydot += v * w->u[k * nu + i];
                }
                for (int a = 0; a < nz; ++a) {
                    double v = w->gcon[k * nv + a];
                    for (int c = 0; c < nz; ++c) v += Abar_k[c * nz + a] * pc[c];
                    npi[a] = v;
                }
                memcpy(pc, npi, sizeof(double) * nz);
                if (p->kg == k)
                    for (int a = 0; a < nz; ++a) pc[a] += w->gGl[a];
            }
        }
    }
    if (factor) {
        double Ptt[MAXNT * MAXNT];
        for (int a = 0; a < nt; ++a)
            for (int b = 0; b < nt; ++b) Ptt[a * nt + b] = P[(nx + a) * nz + nx + b];
        if (spd_inv(nt, Ptt, w->P0tt_inv)) return 1;
        for (int a = 0; a < nt; ++a) {
            if (fabs(pi[nx + a]) > rdi) rdi = fabs(pi[nx + a]);
            if (pi[nx + a] != pi[nx + a]) rdi = NAN;
            if (cert) {
                ci += fabs(pc[nx + a]) * p->fk_th[a];
                ydot += pc[nx + a] * w->th[a];
            }
        }
        *rd_inf = rdi;
        if (cert) { cert[0] = ci; cert[1] = ydot; }
    }
    for (int a = 0; a < nt; ++a) {
        double v = 0.0;
        for (int b = 0; b < nt; ++b) v -= w->P0tt_inv[a * nt + b] * pv[nx + b];
        w->dth[a] = v;
    }
    return 0;
}

static void forward(const lbo_problem *p, lbo_ws *w) {
    const int nx = p->nx, nt = p->nt, nu = p->nu, nz = p->nz, N = p->N;
    for (int j = 0; j < nx; ++j) w->dx[j] = 0.0;
    for (int k = 0; k < N; ++k) {
        const double *L = w->L + (size_t)k * nu * nz, *Ri = w->Ri + (size_t)k * nu * nu;
        double Lz[MAXNU];
        for (int i = 0; i < nu; ++i) {
            double v = 0.0;
            for (int c = 0; c < nx; ++c) v += L[i * nz + c] * w->dx[k * nx + c];
            for (int c = 0; c < nt; ++c) v += L[i * nz + nx + c] * w->dth[c];
            Lz[i] = v;
        }
        for (int i = 0; i < nu; ++i) {
            double v = w->kap[k * nu + i];
            for (int j = 0; j < nu; ++j) v -= Ri[i * nu + j] * Lz[j];
            w->du[k * nu + i] = v;
        }
        double Ak_[MAXNX * MAXNX], Bk_[MAXNX * MAXNU];
        stage_ab(p, w, k, Ak_, Bk_, NULL, NULL);
        for (int a = 0; a < nx; ++a) {
            double v = 0.0;
            for (int c = 0; c < nx; ++c) v += Ak_[a * nx + c] * w->dx[k * nx + c];
            for (int i = 0; i < nu; ++i) v += Bk_[a * nu + i] * w->du[k * nu + i];
            w->dx[(k + 1) * nx + a] = v;
        }
        w->flops += 2.0 * nu * nz + 2.0 * nu * nu + 2.0 * nx * (nx + nu);
    }
}

/* row directions; mode 0: predictor (stores ds*dl in cc, returns alpha_max and the three sums
 * for mu_aff); mode 1: final direction (returns alpha_max).  mode 2: apply step alpha. */
static double rows_step(const lbo_problem *p, lbo_ws *w, int mode, double sigmu, double alpha,
                        double *sums) {
    const int nvb = p->nvb, N = p->N, nz = p->nz;
    double amax = INFINITY, s_sl = 0, s_cross = 0, s_dd = 0;
    for (int k = 0; k <= N; ++k)
        for (int j = 0; j < nvb; ++j)
            for (int side = 0; side < 2; ++side) {
                if (!row_on(p, k, j, side)) continue;
                const int r = (k * nvb + j) * 2 + side;
                const double sgn = side == 0 ? 1.0 : -1.0;
                const double s = w->sb[r], l = w->lb[r];
                const double rc = (mode == 0) ? s * l : s * l + w->ccb[r] - sigmu;
                const double ds = -w->rpb[r] - sgn * dvar_at(p, w, k, j);
                const double dl = (-rc - l * ds) / s;
                if (mode == 2) { w->sb[r] = s + alpha * ds; w->lb[r] = l + alpha * dl; continue; }
                if (ds < 0 && -s / ds < amax) amax = -s / ds;
                if (dl < 0 && -l / dl < amax) amax = -l / dl;
                if (mode == 0) { w->ccb[r] = ds * dl; s_sl += s * l; s_cross += s * dl + l * ds; s_dd += ds * dl; }
                w->flops += 10.0;
            }
    double dz[MAXNZ];
    for (int j = 0; j < p->nx; ++j) dz[j] = w->dx[(size_t)p->kg * p->nx + j];
    for (int j = 0; j < p->nt; ++j) dz[p->nx + j] = w->dth[j];
    for (int i = 0; i < p->ng; ++i) {
        const double *g = p->G + (size_t)i * nz;
        double adv = 0.0;
        for (int a = 0; a < nz; ++a) adv += g[a] * dz[a];
        const double s = w->sg[i], l = w->lg[i];
        const double rc = (mode == 0) ? s * l : s * l + w->ccg[i] - sigmu;
        const double ds = -w->rpg[i] - adv;
        const double dl = (-rc - l * ds) / s;
        if (mode == 2) { w->sg[i] = s + alpha * ds; w->lg[i] = l + alpha * dl; continue; }
        if (ds < 0 && -s / ds < amax) amax = -s / ds;
        if (dl < 0 && -l / dl < amax) amax = -l / dl;
        if (mode == 0) { w->ccg[i] = ds * dl; s_sl += s * l; s_cross += s * dl + l * ds; s_dd += ds * dl; }
        w->flops += 2.0 * nz + 10.0;
    }
    if (sums) { sums[0] = s_sl; sums[1] = s_cross; sums[2] = s_dd; }
    return amax;
}

static int solve_ws(const lbo_problem *p, lbo_ws *w, const double *dx0, const double *dx_ref,
                    const double *d_off, const double *warm, double *uc, double *theta,
                    double *xtraj, double *obj, int *iters, int *status, double *stats) {
    const int nx = p->nx, nt = p->nt, nu = p->nu, nz = p->nz, nv = p->nv, nvb = p->nvb, N = p->N;
    w->flops = 0.0;
    /* ---- initial point: rollout of u = Kinit x + c ---- */
    for (int j = 0; j < nt; ++j) w->th[j] = warm ? warm[N * nu + j] : 0.0;
    for (int j = 0; j < nx; ++j) w->x[j] = dx0[j];
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < nu; ++i) {
            double v = warm ? warm[k * nu + i] : 0.0;
            for (int j = 0; j < nx; ++j) v += p->Kinit[i * nx + j] * w->x[k * nx + j];
            w->u[k * nu + i] = v;
        }
        double Ak_[MAXNX * MAXNX], Bk_[MAXNX * MAXNU];
        stage_ab(p, w, k, Ak_, Bk_, NULL, NULL);
        for (int a = 0; a < nx; ++a) {
            double v = d_off ? d_off[k * nx + a] : 0.0;
            for (int j = 0; j < nx; ++j) v += Ak_[a * nx + j] * w->x[k * nx + j];
            for (int i = 0; i < nu; ++i) v += Bk_[a * nu + i] * w->u[k * nu + i];
            w->x[(k + 1) * nx + a] = v;
        }
    }
    memset(w->lin, 0, sizeof w->lin);
    w->cconst = 0.0;
    if (dx_ref) {
        for (int a = 0; a < nz; ++a)
            for (int j = 0; j < nx; ++j) w->lin[a] += p->Lref[a * nx + j] * dx_ref[j];
        for (int i = 0; i < nx; ++i)
            for (int j = 0; j < nx; ++j) w->cconst += dx_ref[i] * p->Tm[i * nx + j] * dx_ref[j];
    }
    for (int k = 0; k <= N; ++k)
        for (int j = 0; j < nvb; ++j)
            for (int side = 0; side < 2; ++side) {
                if (!row_on(p, k, j, side)) continue;
                const int r = (k * nvb + j) * 2 + side;
                const double slack = box_slack(p, w, k, j, side);
                w->sb[r] = slack > 1.0 ? slack : 1.0;
                w->lb[r] = 1.0;
            }
    for (int i = 0; i < p->ng; ++i) {
        const double slack = gen_slack_ws(p, w, i);
        w->sg[i] = slack > 1.0 ? slack : 1.0;
        w->lg[i] = 1.0;
    }
    int st = LBO_ST_MAXITER, it = 0;
    double rp_inf = 0, rd_inf = 0, mu = 0, lam_inf = 0, hlam = 0;
    for (it = 0;; ++it) {
        assemble(p, w, 0, 0.0, &rp_inf, &mu, &lam_inf, &hlam);
        double cert[2] = {0, 0};
        const int want_cert = lam_inf >= p->inf_trigger;
        if (backward(p, w, 1, &rd_inf, want_cert ? cert : NULL)) { st = LBO_ST_NUMERICAL; break; }
        if (!(rd_inf == rd_inf) || !(rp_inf == rp_inf) || !(mu == mu) || isinf(rd_inf) || isinf(mu)) {
            st = LBO_ST_NUMERICAL; break;
        }
        {
            const double rd_tol = p->tol_res * (100.0 * lam_inf > 1.0 ? 100.0 * lam_inf : 1.0);
            if (rd_inf < rd_tol && rp_inf < p->tol_res && mu < p->tol_mu) { st = LBO_ST_OPTIMAL; break; }
        }
        if (want_cert && hlam + cert[1] < 0.0 && cert[0] * p->inf_scale <= -(hlam + cert[1])) {
            st = LBO_ST_INFEASIBLE; break;
        }
        if (it >= p->max_iter) break; /* the iterate of the last allowed iteration has been tested: LBO_ST_MAXITER */
        forward(p, w);
        double sums[3];
        double aaff = rows_step(p, w, 0, 0.0, 0.0, sums);
        if (aaff > 1.0) aaff = 1.0;
        const double mu_aff = (sums[0] + aaff * sums[1] + aaff * aaff * sums[2]) / p->m_rows;
        const double sr = mu_aff / mu;
        double sigmu = sr * sr * mu;
        if (sigmu < 0.1 * p->tol_mu) sigmu = 0.1 * p->tol_mu; /* do not centre below the target gap: weights s/lambda beyond ~1/tol_mu
                                                                   only add round-off to the Newton steps (degenerate vertices) */
        double dummy;
        assemble(p, w, 1, sigmu, &dummy, &dummy, &dummy, &dummy);
        backward(p, w, 0, NULL, NULL);
        forward(p, w);
        double amax = rows_step(p, w, 1, sigmu, 0.0, NULL);
        double alpha = 0.99 * amax;
        if (alpha > 1.0) alpha = 1.0;
        for (int k = 0; k <= N; ++k) {
            for (int j = 0; j < nx; ++j) w->x[k * nx + j] += alpha * w->dx[k * nx + j];
            if (k < N)
                for (int j = 0; j < nu; ++j) w->u[k * nu + j] += alpha * w->du[k * nu + j];
        }
        for (int j = 0; j < nt; ++j) w->th[j] += alpha * w->dth[j];
        rows_step(p, w, 2, sigmu, alpha, NULL);
        w->flops += 2.0 * (N + 1) * (nx + nu);
    }
    /* ---- outputs ---- */
    double J = w->cconst;
    for (int k = 0; k <= N; ++k) {
        const double *W = p->W[p->wtype[k]];
        double v[MAXNV];
        const int cshift = w->csh && !w->row_shift;
        for (int j = 0; j < nx; ++j) v[j] = w->x[k * nx + j] + (cshift ? w->csh[k * w->cs_stride + j] : 0.0);
        for (int j = 0; j < nt; ++j) v[nx + j] = w->th[j];
        for (int j = 0; j < nu; ++j)
            v[nz + j] = (k < N) ? w->u[k * nu + j] + ((cshift && w->cs_stride > nx) ? w->csh[k * w->cs_stride + nx + j] : 0.0) : 0.0;
        int lim = (k < N) ? nv : nz;
        for (int a = 0; a < lim; ++a)
            for (int b = 0; b < lim; ++b) J += 0.5 * v[a] * W[a * nv + b] * v[b];
        if (k == p->kT)
            for (int a = 0; a < nz; ++a) J += w->lin[a] * v[a];
    }
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < nu; ++i) {
            double v = w->u[k * nu + i];
            for (int j = 0; j < nx; ++j) v -= p->Kout[i * nx + j] * w->x[k * nx + j];
            uc[k * nu + i] = v;
        }
    for (int j = 0; j < nt; ++j) theta[j] = w->th[j];
    if (xtraj) memcpy(xtraj, w->x, sizeof(double) * (size_t)(N + 1) * nx);
    *obj = J; *iters = it; *status = st;
    if (stats) { stats[0] = rd_inf; stats[1] = rp_inf; stats[2] = mu; stats[3] = w->flops; stats[4] = lam_inf; stats[5] = hlam; }
    return 0;
}

int lbo_solve(const lbo_problem *p, const double *dx0, const double *dx_ref, const double *d_off,
              const double *warm, double *uc, double *theta, double *xtraj, double *obj,
              int *iters, int *status, double *stats) {
    lbo_ws *w = ws_alloc(p);
    int rc = solve_ws(p, w, dx0, dx_ref, d_off, warm, uc, theta, xtraj, obj, iters, status, stats);
    ws_free(w);
    return rc;
}

int lbo_solve_shifted(const lbo_problem *p, const double *dx0, const double *dx_ref, const double *d_off,
                      const double *cost_shift, const double *warm, double *uc, double *theta, double *xtraj,
                      double *obj, int *iters, int *status, double *stats) {
    lbo_ws *w = ws_alloc(p);
    w->csh = cost_shift;
    w->cs_stride = p->nx;
    int rc = solve_ws(p, w, dx0, dx_ref, d_off, warm, uc, theta, xtraj, obj, iters, status, stats);
    ws_free(w);
    return rc;
}

typedef struct {
    const lbo_problem *p; long batch; const double *dx0, *dx_ref, *d_off, *warm, *csh;
    double *uc, *theta, *xtraj, *obj; int *iters, *status; long *next; int cs_stride;
    int row_shift; const double *jac;
} lbo_batch_job;

static void *batch_worker(void *arg) {
    lbo_batch_job *j = (lbo_batch_job *)arg;
    const lbo_problem *p = j->p;
    const int nx = p->nx, nu = p->nu, nt = p->nt, N = p->N;
    lbo_ws *w = ws_alloc(p); /* one workspace per thread, reused across QPs */
    for (;;) {
        long b0 = __atomic_fetch_add(j->next, 4, __ATOMIC_RELAXED);
        if (b0 >= j->batch) break;
        long b1 = b0 + 4 < j->batch ? b0 + 4 : j->batch;
        for (long b = b0; b < b1; ++b) {
            w->csh = j->csh ? j->csh + b * (long)(N + 1) * j->cs_stride : NULL;
            w->cs_stride = j->cs_stride;
            w->row_shift = j->row_shift;
            w->jac = j->jac ? j->jac + b * (long)N * nx * 3 : NULL;
            solve_ws(p, w, j->dx0 + b * nx, j->dx_ref ? j->dx_ref + b * nx : NULL,
                      j->d_off ? j->d_off + b * (long)nx * N : NULL,
                      j->warm ? j->warm + b * (long)(N * nu + nt) : NULL, j->uc + b * (long)N * nu,
                      j->theta + b * nt, j->xtraj ? j->xtraj + b * (long)(N + 1) * nx : NULL,
                      j->obj + b, j->iters + b, j->status + b, NULL);
        }
    }
    ws_free(w);
    return NULL;
}

/* independent QPs over `nthreads` POSIX threads (dynamic chunks of 4) */
int lbo_solve_batch(const lbo_problem *p, long batch, const double *dx0, const double *dx_ref,
                    const double *d_off, const double *warm, double *uc, double *theta,
                    double *xtraj, double *obj, int *iters, int *status, int nthreads) {
    return lbo_solve_batch_shifted(p, batch, dx0, dx_ref, d_off, NULL, warm, uc, theta, xtraj, obj, iters, status, nthreads);
}

int lbo_solve_batch_shifted(const lbo_problem *p, long batch, const double *dx0, const double *dx_ref,
                            const double *d_off, const double *cost_shift, const double *warm, double *uc,
                            double *theta, double *xtraj, double *obj, int *iters, int *status, int nthreads) {
    return lbo_solve_batch_shifted_xu(p, batch, dx0, dx_ref, d_off, cost_shift, p->nx, warm, uc, theta, xtraj, obj, iters, status, nthreads);
}

int lbo_solve_batch_shifted_xu(const lbo_problem *p, long batch, const double *dx0, const double *dx_ref,
                               const double *d_off, const double *cost_shift, int cs_stride, const double *warm, double *uc,
                               double *theta, double *xtraj, double *obj, int *iters, int *status, int nthreads) {
    return lbo_solve_batch_ex(p, batch, dx0, dx_ref, d_off, cost_shift, cs_stride, 0, NULL, warm, uc, theta, xtraj, obj, iters, status, nthreads);
}

int lbo_solve_batch_ex(const lbo_problem *p, long batch, const double *dx0, const double *dx_ref,
                       const double *d_off, const double *shift, int cs_stride, int row_shift, const double *jac,
                       const double *warm, double *uc, double *theta, double *xtraj, double *obj, int *iters, int *status,
                       int nthreads) {
    long next = 0;
    lbo_batch_job job = {p, batch, dx0, dx_ref, d_off, warm, shift, uc, theta, xtraj, obj, iters, status, &next, cs_stride,
                         row_shift, jac};
    if (jac && (p->nu != 1 || p->nx < 2)) { snprintf(g_err, sizeof g_err, "LTV Jacobians: nu = 1 and nx >= 2 (xi = [x1;x2;u])"); return -1; }
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, batch_worker, &job);
    batch_worker(&job);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* learned-oracle, plant, data window, closed loop                                              */
/* ------------------------------------------------------------------------------------------ */
void lbo_oracle_l2nw(const double *X, const double *Y, const double *valid, int q, int nin,
                     int nout, const double *xi, double bandwidth, double lambda, double *g) {
    /* oracleL2NW.m:26-36 (valid==NULL) / casadiL2NW.m:14-28 (mask in the denominator only) */
    double sk = 0.0;
    for (int o = 0; o < nout; ++o) g[o] = 0.0;
    double *kv = (double *)malloc(sizeof(double) * (size_t)(q > 0 ? q : 1));
    for (int i = 0; i < q; ++i) {
        double d2 = 0.0;
        for (int a = 0; a < nin; ++a) { double d = X[a * q + i] - xi[a]; d2 += d * d; }
        kv[i] = exp(-d2 / (bandwidth * bandwidth));
        sk += valid ? kv[i] * valid[i] : kv[i];
    }
    for (int i = 0; i < q; ++i) {
        const double wgt = kv[i] / (lambda + sk);
        for (int o = 0; o < nout; ++o) g[o] += Y[o * q + i] * wgt;
    }
    free(kv);
}

static void mg_rhs(const double *x, double u, double *f) { /* …casadi.m:215-221 */
    f[0] = -x[1] + 1.0 + 3.0 * (x[0] / 2.0) - (x[0] * x[0] * x[0] / 2.0);
    f[1] = x[0] + 1.0 - x[2] * sqrt(x[1]);
    f[2] = x[3];
    f[3] = -1000.0 * x[2] - 2.0 * sqrt(500.0) * x[3] + 1000.0 * u;
}
void lbo_plant_rk4(const double *x, double u, double delta, double *xn) { /* …casadi.m:297-304 */
    double k1[4], k2[4], k3[4], k4[4], t[4];
    mg_rhs(x, u, k1);
    for (int i = 0; i < 4; ++i) t[i] = x[i] + delta / 2 * k1[i];
    mg_rhs(t, u, k2);
    for (int i = 0; i < 4; ++i) t[i] = x[i] + delta / 2 * k2[i];
    mg_rhs(t, u, k3);
    for (int i = 0; i < 4; ++i) t[i] = x[i] + delta * k3[i];
    mg_rhs(t, u, k4);
    for (int i = 0; i < 4; ++i) xn[i] = x[i] + delta / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
}

/* g(xi) and its Jacobian dg/dxi (nout x nin, row-major) of the L2NW oracle:
 *   k_i = exp(-|X_i - xi|^2 / h^2), D = lambda + sum_i v_i k_i, g = sum_i Y_i k_i / D
 *   dk_i/dxi = k_i 2 (X_i - xi) / h^2 ;  dg/dxi = (sum_i Y_i dk_i - g sum_i v_i dk_i) / D
 * (what CasADi's algorithmic differentiation of casadiL2NW.m:14-28 evaluates inside IPOPT) */
void lbo_oracle_l2nw_jac(const double *X, const double *Y, const double *valid, int q, int nin, int nout,
                         const double *xi, double bandwidth, double lambda, double *g, double *J) {
    const double ih2 = 1.0 / (bandwidth * bandwidth);
    double D = lambda, num[8] = {0}, dnum[8][4] = {{0}}, dD[4] = {0};
    for (int i = 0; i < q; ++i) {
        double d2 = 0.0, dxv[4];
        for (int c = 0; c < nin; ++c) { dxv[c] = X[c * q + i] - xi[c]; d2 += dxv[c] * dxv[c]; }
        const double k = exp(-d2 * ih2), vi = valid ? valid[i] : 1.0;
        D += vi * k;
        for (int c = 0; c < nin; ++c) dD[c] += vi * k * 2.0 * dxv[c] * ih2;
        for (int a = 0; a < nout; ++a) {
            num[a] += Y[a * q + i] * k;
            for (int c = 0; c < nin; ++c) dnum[a][c] += Y[a * q + i] * k * 2.0 * dxv[c] * ih2;
        }
    }
    for (int a = 0; a < nout; ++a) {
        g[a] = num[a] / D;
        for (int c = 0; c < nin; ++c) J[a * nin + c] = (dnum[a][c] - g[a] * dD[c]) / D;
    }
}

/* Oracle offsets AND Jacobians along the learned-model rollout of the sequence du from dx0 (see lbo_oracle_offsets):
 * first-order model of the learned dynamics around the rollout (xbar_k, ubar_k):
 *   x+ = (A + [J_k(:,1:2) 0 0]) x + (B + J_k(:,3)) u + d_k ,   d_k = g(xibar_k) - J_k xibar_k */
void lbo_oracle_offsets_jac(const lbo_problem *p, const double *dx0, const double *du,
                            const double *X, const double *Y, const double *valid, int q,
                            double bandwidth, double lambda, double *d_off, double *jac, double *g_out) {
    const int nx = p->nx, nu = p->nu, N = p->N;
    double x[MAXNX], xn[MAXNX];
    memcpy(x, dx0, sizeof(double) * nx);
    for (int k = 0; k < N; ++k) {
        double u0 = du[k * nu];
        for (int j = 0; j < nx; ++j) u0 += p->Kinit[j] * x[j];
        double xi[3] = {x[0], x[1], u0}, g[MAXNX], *J = jac + (size_t)k * nx * 3;
        lbo_oracle_l2nw_jac(X, Y, valid, q, 3, nx, xi, bandwidth, lambda, g, J);
        for (int a = 0; a < nx; ++a) {
            if (g_out) g_out[k * nx + a] = g[a];
            d_off[k * nx + a] = g[a] - (J[a * 3] * xi[0] + J[a * 3 + 1] * xi[1] + J[a * 3 + 2] * xi[2]);
            double v = g[a];
            for (int j = 0; j < nx; ++j) v += p->A[a * nx + j] * x[j];
            v += p->B[a * nu] * u0;
            xn[a] = v;
        }
        memcpy(x, xn, sizeof(double) * nx);
    }
}

void lbo_oracle_offsets(const lbo_problem *p, const double *dx0, const double *du,
                        const double *X, const double *Y, const double *valid, int q,
                        double bandwidth, double lambda, double *d_off) {
    const int nx = p->nx, nu = p->nu, N = p->N;
    double x[MAXNX], xn[MAXNX];
    memcpy(x, dx0, sizeof(double) * nx);
    for (int k = 0; k < N; ++k) {
        /* F-form: the sequence holds c and u = K x + c on the learned state (transitionLearned.m:13); Kinit = 0 in the C-form */
        double u0 = du[k * nu];
        for (int j = 0; j < nx; ++j) u0 += p->Kinit[j] * x[j];
        double xi[3] = {x[0], x[1], u0};
        lbo_oracle_l2nw(X, Y, valid, q, 3, nx, xi, bandwidth, lambda, d_off + (size_t)k * nx);
        for (int a = 0; a < nx; ++a) {
            double v = d_off[k * nx + a];
            for (int j = 0; j < nx; ++j) v += p->A[a * nx + j] * x[j];
            v += p->B[a * nu] * u0;
            xn[a] = v;
        }
        memcpy(x, xn, sizeof(double) * nx);
    }
}

/* counter-based uniform in [0,1): splitmix64 finaliser over (seed, scenario, step, component) */
static double lbo_uniform(unsigned long long seed, unsigned long long scen, unsigned long long step,
                          unsigned long long comp) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (scen * 0x100000001B3ULL + step * 8ULL + comp + 1ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

int lbo_closed_loop(const lbo_problem *p, const double *x_eq, double u_eq, const double *x_init,
                    int steps, int q, int use_oracle, int warm_shift, const double *wbar,
                    unsigned long long seed, unsigned long long scenario, double *xhist,
                    double *uhist, double *thist, int *itershist, int *statushist) {
    const int nx = p->nx, nu = p->nu, nt = p->nt, N = p->N;
    if (nx != 4 || nu != 1) { snprintf(g_err, sizeof g_err, "closed loop needs the 4-state plant"); return -1; }
    double *X = calloc((size_t)3 * q, 8), *Y = calloc((size_t)4 * q, 8), *V = calloc((size_t)q, 8);
    double *warm = calloc((size_t)N * nu + nt, 8), *uc = calloc((size_t)N * nu, 8);
    double *xtraj = calloc((size_t)(N + 1) * nx, 8), *doff = calloc((size_t)N * nx, 8);
    double x[4], th[MAXNT];
    int nd = 0, have_warm = 0;
    memcpy(x, x_init, sizeof x);
    memcpy(xhist, x, sizeof x);
    for (int it = 0; it < steps; ++it) {
        double dx0[4];
        for (int j = 0; j < 4; ++j) dx0[j] = x[j] - x_eq[j];
        const double *dptr = NULL;
        if (use_oracle && nd > 0) {
            lbo_oracle_offsets(p, dx0, warm, X, Y, V, q, 0.5, 0.001, doff);
            dptr = doff;
        }
        double J; int iters, st;
        lbo_solve(p, dx0, NULL, dptr, (warm_shift && have_warm) ? warm : NULL, uc, th, xtraj, &J, &iters, &st, NULL);
        /* A QP without an optimal solution (infeasible: the disturbance pushed the state out of the robust set; or the
         * iteration cap) has no first move to apply: the loop then keeps executing the LAST OPTIMAL plan, i.e. applies the
         * first element of the shifted previous plan (zero before any plan exists) and keeps shifting it.  The reference
         * discards the solver status and would apply IPOPT's last iterate. */
        const int ok = (st == LBO_ST_OPTIMAL);
        if (!ok) { memcpy(uc, warm, sizeof(double) * (size_t)N * nu); th[0] = warm[N * nu]; }
        const double u0 = uc[0] + u_eq; /* C-form: uc = du */
        double xn[4];
        lbo_plant_rk4(x, u0, 0.01, xn);
        if (wbar)
            for (int j = 0; j < 4; ++j)
                xn[j] += wbar[j] * (2.0 * lbo_uniform(seed, scenario, (unsigned long long)it, (unsigned long long)j) - 1.0);
        /* data acquisition (LBMPC_casadi.m:194-198): X=[dx1;dx2;du], Y = dx+ - (A dx + B du) */
        double xs[3] = {dx0[0], dx0[1], uc[0]}, ys[4];
        for (int a = 0; a < 4; ++a) {
            double v = xn[a] - x_eq[a];
            for (int j = 0; j < 4; ++j) v -= p->A[a * 4 + j] * dx0[j];
            v -= p->B[a] * uc[0];
            ys[a] = v;
        }
        if (nd < q) { /* get_data.m:3-6 */
            for (int a = 0; a < 3; ++a) X[a * q + nd] = xs[a];
            for (int a = 0; a < 4; ++a) Y[a * q + nd] = ys[a];
            V[nd] = 1.0; nd++;
        } else {      /* get_data.m:8 : shift left, append */
            for (int a = 0; a < 3; ++a) { memmove(X + a * q, X + a * q + 1, 8 * (size_t)(q - 1)); X[a * q + q - 1] = xs[a]; }
            for (int a = 0; a < 4; ++a) { memmove(Y + a * q, Y + a * q + 1, 8 * (size_t)(q - 1)); Y[a * q + q - 1] = ys[a]; }
        }
        /* warm start shift (DMS_tracking_LMPC_casadi.m:187-189): u0=[u_OL(2:end); K x_N] */
        for (int k = 0; k + 1 < N; ++k) warm[k] = uc[k + 1];
        warm[N - 1] = uc[N - 1];
        for (int j = 0; j < nt; ++j) warm[N * nu + j] = th[j];
        have_warm = 1;
        memcpy(x, xn, sizeof x);
        memcpy(xhist + (size_t)(it + 1) * 4, x, sizeof x);
        uhist[it] = u0; thist[it] = th[0]; itershist[it] = iters; statushist[it] = st;
    }
    free(X); free(Y); free(V); free(warm); free(uc); free(xtraj); free(doff);
    return 0;
}
