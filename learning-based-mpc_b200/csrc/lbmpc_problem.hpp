// lbmpc_problem.hpp — host side: turn the reference's model/constraint matrices (the argument
// list of functions/ocpLBMPC.m:1-6) into the canonical stage form the kernels consume.
//
// Canonical form (one QP; z = [x;theta], v = [x;theta;u]):
//   min  sum_k 0.5 v_k' W_{type(k)} v_k + lin' z_{kT} + const      (W already carries the factor 2
//                                                                   of the reference's cost, which
//                                                                   has no 1/2: costLMPC.m:32-38)
//   s.t. x_{k+1} = A x_k + B u_k + d_k,  x_0 = dx0
//        lo <= [x_k;u_k] <= hi on stage ranges [kx0,kx1], [ku0,ku1]
//        G [x_kg;theta] <= hg
// Reference statements this encodes:
//   F-form  costLMPC.m:25-45 (running cost only for k < N-1 -> stages 0..N-3, terminal on x_N),
//           constraintsLMPC.m:20-41 (rows only for k < N -> x_1..x_{N-1}, u_0..u_{N-2};
//           terminal set on the LAST COMPUTED state x_{N-1}), constraintsLBMPC.m:26-31 (X(-)D and
//           robust set on x_1), transitionNominal.m:12 (u = K x + c: a change of variables,
//           undone on output: c = u - K x)
//   C-form  DMS_tracking_LMPC_casadi.m:233-238 (delta-scaled running cost on stages 0..N-1,
//           terminal cost on x_N), :264-286 (rows on x_1..x_N, u_0..u_{N-1}, terminal set on x_N),
//           LBMPC_casadi.m:286-290 (robust rows on x_1)
#pragma once

#include <cmath>
#include <cstdint>
#include <limits>
#include <string>
#include <vector>

#include "../../include/lbmpc.h"
#include "lbmpc_core.cuh"

namespace lbmpc {

struct HostProblem {
    int nx = 0, nt = 0, nu = 0, N = 0, form = 0, variant = 0;
    int ng = 0, ngp = 0, kg = 0, kT = 0, kx0 = 0, kx1 = 0, ku0 = 0, ku1 = 0, ntypes = 0, m_rows = 0;
    int max_iter = 60;
    int tseg[kMaxTypes + 1] = {0, 0, 0, 0, 0};
    unsigned rowmask = 0;
    std::vector<double> A, B, Kinit, Kout, W, lo, hi, Lref, Tm;  // row-major
    std::vector<double> G, hg;                                   // G component-major [NZ][ngp]
    double tol_res = 1e-9, tol_mu = 1e-10, inf_trigger = 1e2, inf_scale = 1.01;
    std::vector<double> fk_u, fk_th;
};

namespace detail {
// column-major (MATLAB) -> row-major
inline std::vector<double> to_rm(const double* M, int rows, int cols) {
    std::vector<double> out((size_t)rows * cols);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) out[(size_t)r * cols + c] = M[(size_t)c * rows + r];
    return out;
}
inline bool boxes_from_hrep(const double* F, const double* h, int nrow, int nvar, double* lo,
                            double* hi) {
    for (int r = 0; r < nrow; ++r) {  // F column-major nrow x nvar
        int col = -1;
        for (int j = 0; j < nvar; ++j)
            if (F[(size_t)j * nrow + r] != 0.0) {
                if (col >= 0) return false;
                col = j;
            }
        if (col < 0) return false;
        const double c = F[(size_t)col * nrow + r], b = h[r] / c;
        if (c > 0) { if (b < hi[col]) hi[col] = b; }
        else       { if (b > lo[col]) lo[col] = b; }
    }
    return true;
}
}  // namespace detail

// stage Hessian for weights (wq, wr, wp, wt):
//   2 [ wq Mx'Q Mx + wp Mx'P Mx + wr Mu'R Mu + wt Mt'T Mt ],  Mx = [I -Lam 0], Mu = [0 -Psi I], Mt = [0 Lam 0]
inline void stage_hessian(const HostProblem& hp, const std::vector<double>& Q, const std::vector<double>& R,
                          const std::vector<double>& Pm, const std::vector<double>& T,
                          const std::vector<double>& Lam, const std::vector<double>& Psi, double wq,
                          double wr, double wp, double wt, double* W) {
    const int nx = hp.nx, nt = hp.nt, nu = hp.nu, nz = nx + nt, nv = nz + nu;
    std::vector<double> Mx((size_t)nx * nv, 0.0), Mu((size_t)nu * nv, 0.0), Mt((size_t)nx * nv, 0.0);
    for (int i = 0; i < nx; ++i) {
        Mx[(size_t)i * nv + i] = 1.0;
        for (int j = 0; j < nt; ++j) {
            Mx[(size_t)i * nv + nx + j] = -Lam[(size_t)i * nt + j];
            Mt[(size_t)i * nv + nx + j] = Lam[(size_t)i * nt + j];
        }
    }
    for (int i = 0; i < nu; ++i) {
        Mu[(size_t)i * nv + nz + i] = 1.0;
        for (int j = 0; j < nt; ++j) Mu[(size_t)i * nv + nx + j] = -Psi[(size_t)i * nt + j];
    }
    for (int a = 0; a < nv; ++a)
        for (int b = 0; b < nv; ++b) {
            double v = 0.0;
            for (int i = 0; i < nx; ++i)
                for (int j = 0; j < nx; ++j) {
                    const double mxa = Mx[(size_t)i * nv + a], mxb = Mx[(size_t)j * nv + b];
                    v += (wq * Q[(size_t)i * nx + j] + wp * Pm[(size_t)i * nx + j]) * mxa * mxb;
                    v += wt * T[(size_t)i * nx + j] * Mt[(size_t)i * nv + a] * Mt[(size_t)j * nv + b];
                }
            for (int i = 0; i < nu; ++i)
                for (int j = 0; j < nu; ++j)
                    v += wr * R[(size_t)i * nu + j] * Mu[(size_t)i * nv + a] * Mu[(size_t)j * nv + b];
            W[(size_t)a * nv + b] = 2.0 * v;
        }
}

inline int build_problem(const lbmpc_model* m, const lbmpc_config* c, HostProblem& hp, std::string& err) {
    if (!m || !c) { err = "null model/config"; return LBMPC_EINVAL; }
    if (!m->A || !m->B || !m->K || !m->Q || !m->R || !m->P || !m->T || !m->LAMBDA || !m->PSI || !m->F_x ||
        !m->h_x || !m->F_u || !m->h_u || !m->F_w_N || !m->h_w_N) {
        err = "null matrix pointer in lbmpc_model";
        return LBMPC_EINVAL;
    }
    if (c->form != LBMPC_FORM_F && c->form != LBMPC_FORM_C) { err = "config.form must be LBMPC_FORM_F or _C"; return LBMPC_EINVAL; }
    if (c->variant != LBMPC_VARIANT_LMPC && c->variant != LBMPC_VARIANT_LBMPC) { err = "config.variant must be LMPC or LBMPC"; return LBMPC_EINVAL; }
    if (c->variant == LBMPC_VARIANT_LBMPC && (!m->F_x_d || !m->h_x_d || m->n_Fxd <= 0)) {
        err = "LBMPC variant needs F_x_d/h_x_d (getCONSPOLY.m:28-30)";
        return LBMPC_EINVAL;
    }
    if (c->N < 3 || c->N > 4096) { err = "horizon N must be in [3, 4096]"; return LBMPC_ESHAPE; }
    if (m->n_Fw < 0 || m->n_Fx <= 0 || m->n_Fu <= 0) { err = "bad constraint row counts"; return LBMPC_EINVAL; }
    hp.nx = m->nx; hp.nu = m->nu; hp.nt = m->nt; hp.N = c->N; hp.form = c->form; hp.variant = c->variant;
    const int nx = hp.nx, nu = hp.nu, nt = hp.nt, nz = nx + nt, nv = nz + nu, N = hp.N;
    if (nx < 1 || nu < 1 || nt < 1 || nx > 8 || nu > 4 || nt > 4) { err = "unsupported dimensions"; return LBMPC_ESHAPE; }
    hp.A = detail::to_rm(m->A, nx, nx);
    hp.B = detail::to_rm(m->B, nx, nu);
    std::vector<double> K = detail::to_rm(m->K, nu, nx), Q = detail::to_rm(m->Q, nx, nx),
                        R = detail::to_rm(m->R, nu, nu), Pm = detail::to_rm(m->P, nx, nx),
                        Lam = detail::to_rm(m->LAMBDA, nx, nt), Psi = detail::to_rm(m->PSI, nu, nt);
    std::vector<double> T((size_t)nx * nx, 0.0);
    if (m->T_is_scalar) for (int i = 0; i < nx; ++i) T[(size_t)i * nx + i] = m->T[0];
    else T = detail::to_rm(m->T, nx, nx);
    hp.Kinit.assign((size_t)nu * nx, 0.0);
    hp.Kout.assign((size_t)nu * nx, 0.0);
    if (c->form == LBMPC_FORM_F) { hp.Kinit = K; hp.Kout = K; }
    // ---- boxes ----
    const double inf = std::numeric_limits<double>::infinity();
    hp.lo.assign(nx + nu, -inf);
    hp.hi.assign(nx + nu, inf);
    if (!detail::boxes_from_hrep(m->F_x, m->h_x, m->n_Fx, nx, hp.lo.data(), hp.hi.data()) ||
        !detail::boxes_from_hrep(m->F_u, m->h_u, m->n_Fu, nu, hp.lo.data() + nx, hp.hi.data() + nx)) {
        err = "F_x / F_u rows must have exactly one non-zero each (boxes, getCONS.m:15-16)";
        return LBMPC_ESHAPE;
    }
    hp.rowmask = 0;
    for (int j = 0; j < nx + nu; ++j) {
        if (std::isfinite(hp.hi[j])) hp.rowmask |= 1u << (2 * j);
        if (std::isfinite(hp.lo[j])) hp.rowmask |= 1u << (2 * j + 1);
    }
    // ---- stage ranges and cost segments ----
    const double delta = c->delta > 0 ? c->delta : 0.01;
    struct Wt { double q, r, p, t; };
    std::vector<Wt> w(N + 1);
    for (int k = 0; k <= N; ++k) {
        if (k == N) w[k] = {0, 0, 1, 1};
        else if (c->form == LBMPC_FORM_C) w[k] = {delta, delta, 0, 0};
        else if (k <= N - 3) w[k] = {1, 1, 0, 0};
        else w[k] = {0, 0, 0, 0};
    }
    hp.ntypes = 0;
    hp.W.clear();
    for (int k = 0; k <= N; ++k) {
        const bool same = k > 0 && w[k].q == w[k - 1].q && w[k].r == w[k - 1].r && w[k].p == w[k - 1].p &&
                          w[k].t == w[k - 1].t;
        if (!same) {
            if (hp.ntypes == kMaxTypes) { err = "too many stage-cost segments"; return LBMPC_ESHAPE; }
            hp.tseg[hp.ntypes] = k;
            hp.W.resize((size_t)(hp.ntypes + 1) * nv * nv);
            stage_hessian(hp, Q, R, Pm, T, Lam, Psi, w[k].q, w[k].r, w[k].p, w[k].t,
                          hp.W.data() + (size_t)hp.ntypes * nv * nv);
            hp.ntypes++;
        }
    }
    for (int t = hp.ntypes; t <= kMaxTypes; ++t) hp.tseg[t] = N + 1;
    hp.kT = N;
    hp.Tm = T;
    hp.Lref.assign((size_t)nz * nx, 0.0);  // lin_theta = -2 Lam' T x_ref   (costLMPC.m:38)
    for (int a = 0; a < nt; ++a)
        for (int j = 0; j < nx; ++j) {
            double v = 0.0;
            for (int i = 0; i < nx; ++i) v += Lam[(size_t)i * nt + a] * T[(size_t)i * nx + j];
            hp.Lref[(size_t)(nx + a) * nx + j] = -2.0 * v;
        }
    if (c->form == LBMPC_FORM_C) { hp.kx0 = 1; hp.kx1 = N; hp.ku0 = 0; hp.ku1 = N - 1; }
    else { hp.kx0 = 1; hp.kx1 = N - 1; hp.ku0 = 0; hp.ku1 = N - 2; }
    // ---- polytope block ----
    const int nFxd = c->variant == LBMPC_VARIANT_LBMPC ? m->n_Fxd : 0;
    hp.ng = nFxd + m->n_Fw;
    hp.ngp = (hp.ng + 3) & ~3;
    if (hp.ngp == 0) hp.ngp = 4;
    hp.kg = c->variant == LBMPC_VARIANT_LBMPC ? 1 : (c->form == LBMPC_FORM_C ? N : N - 1);
    hp.G.assign((size_t)nz * hp.ngp, 0.0);
    hp.hg.assign(hp.ngp, 1.0);
    for (int r = 0; r < nFxd; ++r) {
        for (int j = 0; j < nx; ++j) hp.G[(size_t)j * hp.ngp + r] = m->F_x_d[(size_t)j * nFxd + r];
        hp.hg[r] = m->h_x_d[r];
    }
    for (int r = 0; r < m->n_Fw; ++r) {
        for (int j = 0; j < nz; ++j) hp.G[(size_t)j * hp.ngp + nFxd + r] = m->F_w_N[(size_t)j * m->n_Fw + r];
        hp.hg[nFxd + r] = m->h_w_N[r];
    }
    // ---- row count (mu = s'lambda / m) ----
    int rows = hp.ng;
    for (int k = 0; k <= N; ++k)
        for (int j = 0; j < nx + nu; ++j) {
            const bool in = j < nx ? (k >= hp.kx0 && k <= hp.kx1) : (k >= hp.ku0 && k <= hp.ku1);
            if (!in) continue;
            rows += ((hp.rowmask >> (2 * j)) & 1u) + ((hp.rowmask >> (2 * j + 1)) & 1u);
        }
    hp.m_rows = rows;
    // Farkas infeasibility test: per-variable upper bounds ybar_j of |y_j| on the feasible set (input box bounds
    // where they exist, 10 per unbounded variable), margin 1 % for round-off:  1.01 sum_j |(G_red'lambda)_j| ybar_j <= -h_red'lambda
    // (raising the box multipliers of u_j by |(G_red'lambda)_j| cancels that component exactly and costs ybar_j each: the
    // test says the corrected lambda is an exact Farkas certificate)
    // theta: every polytope row with ONE non-zero theta coefficient g_t bounds that component once x_kg is boxed,
    //     g_t theta_t <= hg_i + sum_j |G_ij| max(|lo_j|, |hi_j|);
    // tightest upper / lower bound over the rows -> ybar_theta (10 if a side stays unbounded).  Exact for the reference's sets
    // (0.98 for Moore-Greitzer); a model with a wider steady-state range is judged by its own range.
    double Rb = 0.0;
    hp.fk_th.assign(nt, 10.0);
    for (int t = 0; t < nt; ++t) {
        double ub = inf, lb = -inf;
        const bool boxed = hp.kg >= hp.kx0 && hp.kg <= hp.kx1;
        for (int i = 0; i < hp.ng && boxed; ++i) {
            int others = 0;
            for (int a = 0; a < nt; ++a) others += (a != t && hp.G[(size_t)(nx + a) * hp.ngp + i] != 0.0);
            const double gt = hp.G[(size_t)(nx + t) * hp.ngp + i];
            if (others || gt == 0.0) continue;
            double rhs = hp.hg[i];
            for (int j = 0; j < nx; ++j) {
                const double gj = hp.G[(size_t)j * hp.ngp + i];
                if (gj != 0.0) rhs += std::fabs(gj) * std::fmax(std::fabs(hp.lo[j]), std::fabs(hp.hi[j]));
            }
            if (!std::isfinite(rhs)) continue;
            if (gt > 0.0) { if (rhs / gt < ub) ub = rhs / gt; }
            else          { if (rhs / gt > lb) lb = rhs / gt; }
        }
        if (std::isfinite(ub) && std::isfinite(lb)) hp.fk_th[t] = std::fmax(std::fabs(ub), std::fabs(lb));
        Rb += hp.fk_th[t];
    }
    hp.fk_u.assign(nu, 10.0);
    for (int j = nx; j < nx + nu; ++j) {
        const double bnd = std::fmax(std::fabs(hp.lo[j]), std::fabs(hp.hi[j]));
        if (std::isfinite(bnd)) hp.fk_u[j - nx] = bnd;
    }
    for (int k = 0; k < N; ++k)
        for (int j = nx; j < nx + nu; ++j) Rb += (k >= hp.ku0 && k <= hp.ku1) ? hp.fk_u[j - nx] : 10.0;
    hp.inf_scale = 1.01;
    if (c->tol_res > 0) hp.tol_res = c->tol_res;
    if (c->tol_mu > 0) hp.tol_mu = c->tol_mu;
    if (c->inf_radius > 0) hp.inf_scale = c->inf_radius / Rb;  // a caller-supplied radius rescales the bounds
    if (c->max_iter > 0) hp.max_iter = c->max_iter;
    return LBMPC_OK;
}

template <int NX, int NT, int NU>
inline Params<NX, NT, NU> to_params(const HostProblem& hp) {
    Params<NX, NT, NU> p{};
    constexpr int NV = NX + NT + NU, NZ = NX + NT, NVB = NX + NU;
    p.N = hp.N; p.ng = hp.ng; p.ngp = hp.ngp; p.kg = hp.kg; p.kT = hp.kT;
    p.kx0 = hp.kx0; p.kx1 = hp.kx1; p.ku0 = hp.ku0; p.ku1 = hp.ku1;
    p.ntypes = hp.ntypes; p.max_iter = hp.max_iter; p.m_rows = hp.m_rows;
    for (int t = 0; t <= kMaxTypes; ++t) p.tseg[t] = hp.tseg[t];
    p.rowmask = hp.rowmask;
    for (int i = 0; i < NX * NX; ++i) p.A[i] = hp.A[i];
    {  // A^bm for the blocked adjoint recursion
        const int bm = Layout<NX, NT, NU>::block_len(hp.N);
        double acc[NX * NX], tmp[NX * NX];
        for (int i = 0; i < NX * NX; ++i) acc[i] = (i / NX == i % NX) ? 1.0 : 0.0;
        for (int r = 0; r < bm; ++r) {
            for (int i = 0; i < NX; ++i)
                for (int j = 0; j < NX; ++j) {
                    double v = 0.0;
                    for (int k = 0; k < NX; ++k) v += acc[i * NX + k] * hp.A[k * NX + j];
                    tmp[i * NX + j] = v;
                }
            for (int i = 0; i < NX * NX; ++i) acc[i] = tmp[i];
        }
        for (int i = 0; i < NX * NX; ++i) p.Apow[i] = acc[i];
    }
    for (int i = 0; i < NX * NU; ++i) p.B[i] = hp.B[i];
    for (int i = 0; i < NU * NX; ++i) { p.Kinit[i] = hp.Kinit[i]; p.Kout[i] = hp.Kout[i]; }
    for (int t = 0; t < hp.ntypes; ++t)
        for (int i = 0; i < NV * NV; ++i) p.W[t][i] = hp.W[(size_t)t * NV * NV + i];
    for (int j = 0; j < NVB; ++j) {  // keep the constant bank finite: masked rows never read these
        p.lo[j] = std::isfinite(hp.lo[j]) ? hp.lo[j] : 0.0;
        p.hi[j] = std::isfinite(hp.hi[j]) ? hp.hi[j] : 0.0;
    }
    for (int i = 0; i < NZ * NX; ++i) p.Lref[i] = hp.Lref[i];
    for (int i = 0; i < NX * NX; ++i) p.Tm[i] = hp.Tm[i];
    p.tol_res = hp.tol_res; p.tol_mu = hp.tol_mu; p.inf_trigger = hp.inf_trigger; p.inf_scale = hp.inf_scale; p.fk_free = 10.0;
    for (int i = 0; i < NU; ++i) p.fk_u[i] = hp.fk_u[i];
    for (int i = 0; i < NT; ++i) p.fk_th[i] = hp.fk_th[i];
    p.inv_m = 1.0 / (double)hp.m_rows;
    return p;
}

}  // namespace lbmpc
