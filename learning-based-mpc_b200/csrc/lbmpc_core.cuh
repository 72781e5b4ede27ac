// lbmpc_core.cuh — per-QP math of the batched Mehrotra/Riccati interior-point engine.
//
// Everything here is a small __host__ __device__ function that works on ONE QP whose whole
// iterate lives in a flat double buffer (a shared-memory "slot" on the GPU).  Three kinds, matching
// how the kernel (lbmpc_kernels.cuh) maps them onto threads:
//   * stage/row functions  — independent per horizon stage k or per polytope row i; the kernel
//     runs them with one warp per QP, lanes striding over stages/rows, and combines the small
//     reduction structs with warp shuffles;
//   * vector sweeps        — the sequential backward / forward substitution recursions over the
//     N stages (and the adjoint recursions that give the dual residual and the Farkas test);
//     thread-local, ONE LANE PER QP, so one warp advances all QPs of a CTA in lock step;
//   * the Riccati factorisation — the one recursion whose per-stage work (two 5x5 congruences)
//     is large enough to spread over lanes: Coop<NX> runs it with 16 lanes per QP (one lane per
//     unique entry of the symmetric cost-to-go matrix), the stage-to-stage dependency chain
//     shortened by carrying the matrix scaled by the previous pivot so that the reciprocal of the
//     pivot is off the critical path.  Core::factor_serial is the thread-local form of the same
//     recursion for the other compiled shape (nu = 2).
//
// The optimisation problem is the reference's per-step (LB)MPC problem in canonical stage form
// (reference files cited in include/lbmpc.h and lbmpc_problem.hpp where the
// canonical form is built).  Host compilation (no __CUDACC__) exists only so that
// tests/emul/ can unit-test this arithmetic without a GPU; the product never runs it on the CPU.
#pragma once

#include <math.h>

#ifdef __CUDACC__
#define LB_HD __host__ __device__ __forceinline__
#else
#define LB_HD inline
#endif

namespace lbmpc {

constexpr int kMaxTypes = 4;  // distinct stage-cost segments along the horizon

enum SlotState : int { SLOT_EMPTY = 0, SLOT_FRESH = 1, SLOT_RUN = 2, SLOT_DONE = 3 };

// ------------------------------------------------------------------------------------------------
// problem constants: passed by value as a __grid_constant__ kernel parameter (constant bank)
// ------------------------------------------------------------------------------------------------
template <int NX, int NT, int NU>
struct Params {
    static constexpr int NZ = NX + NT, NV = NX + NT + NU, NVB = NX + NU;
    int N, ng, ngp, kg, kT, kx0, kx1, ku0, ku1, ntypes, max_iter, m_rows;
    int tseg[kMaxTypes + 1];  // cost type t covers stages [tseg[t], tseg[t+1])
    unsigned rowmask;         // bit (2*j+side) set: bound (variable j, side) is finite
    double A[NX * NX], B[NX * NU], Kinit[NU * NX], Kout[NU * NX];  // row-major
    double W[kMaxTypes][NV * NV];                                  // stage Hessians, v=[x;theta;u]
    double lo[NVB], hi[NVB];
    double Lref[NZ * NX], Tm[NX * NX];
    double tol_res, tol_mu, inf_trigger, inf_radius, inv_m;
};

// ------------------------------------------------------------------------------------------------
// slot layout (offsets in doubles).  Arrays are [component][stage] so that lanes striding over
// stages hit consecutive banks; the slot stride is odd so that lanes striding over slots do too.
// ------------------------------------------------------------------------------------------------
template <int NX, int NT, int NU>
struct Layout {
    static constexpr int NZ = NX + NT, NV = NX + NT + NU, NVB = NX + NU;
    static constexpr int NH = NZ * (NZ + 1) / 2;
    // misc block
    static constexpr int M_TH = 0, M_DTH = M_TH + NT, M_DTHA = M_DTH + NT, M_LIN = M_DTHA + NT,
                         M_HG = M_LIN + NZ, M_GGL = M_HG + NH, M_DG = M_GGL + NZ, M_PTT = M_DG + NZ,
                         M_RP = M_PTT + NT * NT, M_MU = M_RP + 1, M_LAM = M_MU + 1, M_HLAM = M_LAM + 1,
                         M_RD = M_HLAM + 1, M_ALPHA = M_RD + 1, M_SIGMU = M_ALPHA + 1,
                         M_CCONST = M_SIGMU + 1, M_CERT = M_CCONST + 1, M_OBJ = M_CERT + 1,
                         M_PIV = M_OBJ + 1, M_SIZE = M_PIV + 1;
    // [qd | q | g | RL] are dead once the corrector sweeps are done: the final step-length pass
    // parks the row directions ds, dl (2 x 2 NVB per stage) there
    static_assert(NV + NU * NZ >= 2 * NVB, "scratch for the row directions does not fit");
    int Np, ngp;
    int o_x, o_u, o_sb, o_lb, o_qd, o_q, o_g, o_L, o_Ri, o_kap, o_dx, o_du, o_dxa, o_dua, o_sg,
        o_lg, o_misc, stride;
    LB_HD Layout(int N, int ngp_) {
        Np = (N + 1) | 1;  // odd: lanes striding over components hit different banks too
        ngp = ngp_;
        int o = 0;
        o_x = o;   o += NX * Np;
        o_u = o;   o += NU * Np;
        o_sb = o;  o += NVB * 2 * Np;
        o_lb = o;  o += NVB * 2 * Np;
        o_qd = o;  o += NVB * Np;      // barrier diagonal
        o_q = o;   o += NVB * Np;      // Newton rhs on the bounded variables (cost gradient + barrier terms)
        o_g = o;   o += NV * Np;       // cost gradient + G'lambda (dual residual recursion)
        o_L = o;   o += NU * NZ * Np;  // RL_k = Ri_k L_k
        o_Ri = o;  o += NU * NU * Np;
        o_kap = o; o += NU * Np;
        o_dx = o;  o += NX * Np;       // (o_dx, o_du) contiguous: scratch of the corrector assembly
        o_du = o;  o += NU * Np;
        o_dxa = o; o += NX * Np;
        o_dua = o; o += NU * Np;
        o_sg = o;  o += ngp;
        o_lg = o;  o += ngp;
        o_misc = o; o += M_SIZE;
        stride = o | 1;  // odd
    }
};

struct RedAsm {  // reductions of the assembly pass
    double rp, sl, lam, hl;  // max |r_p|, sum s*lambda, max lambda, sum lambda*slack
};
struct RedStep {  // reductions of a step-length pass
    double ratio, s0, s1, s2;  // max(-ds/s, -dl/l), sum s l, sum (s dl + l ds), sum ds dl
};

LB_HD double lb_max(double a, double b) { return a > b ? a : b; }
LB_HD double lb_abs(double a) { return a < 0 ? -a : a; }
// max that keeps a NaN once seen (residual norms: status 3 must be visible)
LB_HD double lb_nanmax(double acc, double v) { return (v == v) ? (acc == acc ? lb_max(acc, v) : acc) : v; }

// reciprocal of a positive normal double: hardware seed + two Newton steps (4 DFMA) on the device
LB_HD double lb_rcp(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
#else
    return 1.0 / x;
#endif
}

template <int NX, int NT, int NU>
struct Core {
    using P = Params<NX, NT, NU>;
    using L = Layout<NX, NT, NU>;
    static constexpr int NZ = NX + NT, NV = NX + NT + NU, NVB = NX + NU, NH = L::NH;

    static LB_HD int stage_type(const P& p, int k) {
        int t = 0;
#pragma unroll
        for (int i = 1; i < kMaxTypes; ++i) t += (i < p.ntypes && k >= p.tseg[i]) ? 1 : 0;
        return t;
    }
    // which rows exist at stage k: bit (2*j+side)
    static LB_HD unsigned stage_rows(const P& p, int k) {
        const bool inx = k >= p.kx0 && k <= p.kx1, inu = k >= p.ku0 && k <= p.ku1 && k < p.N;
        constexpr unsigned xbits = (1u << (2 * NX)) - 1u;
        return p.rowmask & ((inx ? xbits : 0u) | (inu ? ~xbits : 0u));
    }
    static LB_HD bool row_on(const P& p, int k, int j, int side) { return (stage_rows(p, k) >> (2 * j + side)) & 1u; }
    static LB_HD int sym(int a, int b) {  // packed index of symmetric NZ x NZ, a<=b
        return a * NZ - a * (a - 1) / 2 + (b - a);
    }
    static LB_HD int zidx(int j) { return j < NX ? j : NZ + (j - NX); }  // bounded var -> index in v=[x;theta;u]
    // value of bounded variable j (x then u) at stage k from arrays ax (NX x Np), au (NU x Np)
    static LB_HD double bvar(const L& l, const double* s, int ox, int ou, int k, int j) {
        return j < NX ? s[ox + j * l.Np + k] : s[ou + (j - NX) * l.Np + k];
    }

    // ============================================================================================
    // sweep: initial rollout, in place.  On entry u(:,k) holds the warm-start c_k (or 0) and
    // x(:,k+1) holds the dynamics offset d_k (or 0); x(:,0) = dx0.
    //   u_k = Kinit x_k + c_k (transitionNominal.m:12) ; x_{k+1} = A x_k + B u_k + d_k (nominalModel.m:28)
    // ============================================================================================
    static LB_HD void rollout(const P& p, const L& l, double* s) {
        double x[NX], u[NU], xn[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) x[j] = s[l.o_x + j * l.Np];
        for (int k = 0; k < p.N; ++k) {
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                double v = s[l.o_u + i * l.Np + k];
#pragma unroll
                for (int j = 0; j < NX; ++j) v += p.Kinit[i * NX + j] * x[j];
                u[i] = v;
                s[l.o_u + i * l.Np + k] = v;
            }
#pragma unroll
            for (int a = 0; a < NX; ++a) {
                double v = s[l.o_x + a * l.Np + k + 1];
#pragma unroll
                for (int j = 0; j < NX; ++j) v += p.A[a * NX + j] * x[j];
#pragma unroll
                for (int i = 0; i < NU; ++i) v += p.B[a * NU + i] * u[i];
                xn[a] = v;
                s[l.o_x + a * l.Np + k + 1] = v;
            }
#pragma unroll
            for (int j = 0; j < NX; ++j) x[j] = xn[j];
        }
    }

    // ============================================================================================
    // stage/row: initial slacks and multipliers  s = max(h - a v, 1), lambda = 1
    // ============================================================================================
    static LB_HD void init_rows_stage(const P& p, const L& l, double* s, int k) {
        const unsigned rows = stage_rows(p, k);
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
            if (!((rows >> (2 * j)) & 3u)) continue;
            const double v = bvar(l, s, l.o_x, l.o_u, k, j);
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                if (!((rows >> (2 * j + side)) & 1u)) continue;
                const double slack = side == 0 ? p.hi[j] - v : v - p.lo[j];
                const int r = (2 * j + side) * l.Np + k;
                s[l.o_sb + r] = slack > 1.0 ? slack : 1.0;
                s[l.o_lb + r] = 1.0;
            }
        }
    }
    static LB_HD double gen_slack(const P& p, const L& l, const double* s, const double* G,
                                  const double* hg, int i) {
        double v = hg[i];
#pragma unroll
        for (int a = 0; a < NX; ++a) v -= G[a * p.ngp + i] * s[l.o_x + a * l.Np + p.kg];
#pragma unroll
        for (int a = 0; a < NT; ++a) v -= G[(NX + a) * p.ngp + i] * s[l.o_misc + L::M_TH + a];
        return v;
    }
    static LB_HD void init_rows_gen(const P& p, const L& l, double* s, const double* G,
                                    const double* hg, int i) {
        const double slack = gen_slack(p, l, s, G, hg, i);
        s[l.o_sg + i] = slack > 1.0 ? slack : 1.0;
        s[l.o_lg + i] = 1.0;
    }

    // ============================================================================================
    // stage: predictor assembly.  Writes Qd (barrier diagonal), g (cost gradient + G'lambda) and
    // q (Newton rhs on the bounded variables) of stage k; accumulates reductions.
    // ============================================================================================
    static LB_HD void assemble_stage(const P& p, const L& l, double* s, int k, RedAsm& red) {
        double v[NV], g[NV];
#pragma unroll
        for (int j = 0; j < NX; ++j) v[j] = s[l.o_x + j * l.Np + k];
#pragma unroll
        for (int j = 0; j < NT; ++j) v[NX + j] = s[l.o_misc + L::M_TH + j];
        const bool last = k >= p.N;
#pragma unroll
        for (int j = 0; j < NU; ++j) v[NZ + j] = last ? 0.0 : s[l.o_u + j * l.Np + k];
        const double* W = p.W[stage_type(p, k)];
#pragma unroll
        for (int a = 0; a < NV; ++a) {
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < NV; ++b) acc += W[a * NV + b] * v[b];
            g[a] = (last && a >= NZ) ? 0.0 : acc;
        }
        if (k == p.kT) {
#pragma unroll
            for (int a = 0; a < NZ; ++a) g[a] += s[l.o_misc + L::M_LIN + a];
        }
        const unsigned rows = stage_rows(p, k);
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
            double qd = 0.0, gl = 0.0, gp = 0.0;
            const int a = zidx(j);
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                if (!((rows >> (2 * j + side)) & 1u)) continue;
                const int r = (2 * j + side) * l.Np + k;
                const double S = s[l.o_sb + r], Lm = s[l.o_lb + r];
                const double sgn = side == 0 ? 1.0 : -1.0;
                const double slack = side == 0 ? p.hi[j] - v[a] : v[a] - p.lo[j];
                const double rp = S - slack;
                const double w = Lm * lb_rcp(S);
                qd += w;
                gl += sgn * Lm;
                gp += sgn * (w * rp);
                red.rp = lb_nanmax(red.rp, lb_abs(rp));
                red.sl += S * Lm;
                red.lam = lb_max(red.lam, Lm);
                red.hl += Lm * slack;
            }
            s[l.o_qd + j * l.Np + k] = qd;
            s[l.o_q + j * l.Np + k] = g[a] + gp;
            g[a] += gl;
        }
#pragma unroll
        for (int a = 0; a < NV; ++a) s[l.o_g + a * l.Np + k] = g[a];
    }

    // row: predictor assembly of polytope row i.  acc = {HG (NH packed), gGl (NZ), dG (NZ)}
    static LB_HD void assemble_gen_row(const P& p, const L& l, const double* s, const double* G,
                                       const double* hg, int i, double* acc, RedAsm& red) {
        const double slack = gen_slack(p, l, s, G, hg, i);
        const double S = s[l.o_sg + i], Lm = s[l.o_lg + i];
        const double rp = S - slack, w = Lm * lb_rcp(S), t = w * rp;
        red.rp = lb_nanmax(red.rp, lb_abs(rp));
        red.sl += S * Lm;
        red.lam = lb_max(red.lam, Lm);
        red.hl += Lm * slack;
        double g[NZ];
#pragma unroll
        for (int a = 0; a < NZ; ++a) g[a] = G[a * p.ngp + i];
        int idx = 0;
#pragma unroll
        for (int a = 0; a < NZ; ++a) {
            const double wa = w * g[a];
#pragma unroll
            for (int b = a; b < NZ; ++b) acc[idx++] += wa * g[b];
        }
#pragma unroll
        for (int a = 0; a < NZ; ++a) {
            acc[NH + a] += g[a] * Lm;
            acc[NH + NZ + a] += g[a] * (t - Lm);
        }
    }

    // ============================================================================================
    // sweep: Riccati factorisation over z=[x;theta] with Abar=diag(A,I), Bbar=[B;0], thread-local
    // form.  Stores RL_k = Ri_k L_k, Ri_k and the inverse theta block of P_0.  Returns false when a
    // pivot is not positive / not finite.
    // ============================================================================================
    static LB_HD bool factor_serial(const P& p, const L& l, double* s) {
        double Pxx[NX][NX], Pxt[NX][NT], Ptt[NT][NT];
        double* m = s + l.o_misc;
        const int N = p.N;
        bool ok = true;
        {
            const double* W = p.W[stage_type(p, N)];
            const bool kg = (p.kg == N);
#pragma unroll
            for (int a = 0; a < NX; ++a) {
#pragma unroll
                for (int b = 0; b < NX; ++b)
                    Pxx[a][b] = W[a * NV + b] + (kg ? m[L::M_HG + (a <= b ? sym(a, b) : sym(b, a))] : 0.0);
                Pxx[a][a] += s[l.o_qd + a * l.Np + N];
#pragma unroll
                for (int t = 0; t < NT; ++t)
                    Pxt[a][t] = W[a * NV + NX + t] + (kg ? m[L::M_HG + sym(a, NX + t)] : 0.0);
            }
#pragma unroll
            for (int a = 0; a < NT; ++a)
#pragma unroll
                for (int b = 0; b < NT; ++b)
                    Ptt[a][b] = W[(NX + a) * NV + NX + b] +
                                (kg ? m[L::M_HG + (a <= b ? sym(NX + a, NX + b) : sym(NX + b, NX + a))] : 0.0);
        }
        for (int k = N - 1; k >= 0; --k) {
            const double* W = p.W[stage_type(p, k)];
            const bool kg = (p.kg == k);
            double Lk[NU][NZ], Ri[NU][NU];
            double M[NX][NX], PB[NX][NU], Rt[NU][NU];
#pragma unroll
            for (int a = 0; a < NX; ++a)
#pragma unroll
                for (int b = 0; b < NX; ++b) {
                    double v = 0.0;
#pragma unroll
                    for (int c = 0; c < NX; ++c) v += Pxx[a][c] * p.A[c * NX + b];
                    M[a][b] = v;
                }
#pragma unroll
            for (int a = 0; a < NX; ++a)
#pragma unroll
                for (int i = 0; i < NU; ++i) {
                    double v = 0.0;
#pragma unroll
                    for (int c = 0; c < NX; ++c) v += Pxx[a][c] * p.B[c * NU + i];
                    PB[a][i] = v;
                }
            // L = Wuz + Bbar' P Abar
#pragma unroll
            for (int i = 0; i < NU; ++i) {
#pragma unroll
                for (int b = 0; b < NX; ++b) {
                    double v = W[(NZ + i) * NV + b];
#pragma unroll
                    for (int c = 0; c < NX; ++c) v += p.B[c * NU + i] * M[c][b];
                    Lk[i][b] = v;
                }
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    double v = W[(NZ + i) * NV + NX + t];
#pragma unroll
                    for (int c = 0; c < NX; ++c) v += p.B[c * NU + i] * Pxt[c][t];
                    Lk[i][NX + t] = v;
                }
            }
            // Rt = Wuu + Qd_u + B' Pxx B
#pragma unroll
            for (int i = 0; i < NU; ++i)
#pragma unroll
                for (int j = 0; j < NU; ++j) {
                    double v = W[(NZ + i) * NV + NZ + j];
#pragma unroll
                    for (int c = 0; c < NX; ++c) v += p.B[c * NU + i] * PB[c][j];
                    Rt[i][j] = v;
                }
#pragma unroll
            for (int i = 0; i < NU; ++i) Rt[i][i] += s[l.o_qd + (NX + i) * l.Np + k];
            if (NU == 1) {
                ok = ok && (Rt[0][0] > 0.0);
                Ri[0][0] = 1.0 / Rt[0][0];
            } else {  // NU == 2 closed form
                const double det = Rt[0][0] * Rt[NU - 1][NU - 1] - Rt[0][NU - 1] * Rt[NU - 1][0];
                ok = ok && (Rt[0][0] > 0.0) && (det > 0.0);
                const double id = 1.0 / det;
                Ri[0][0] = Rt[NU - 1][NU - 1] * id;
                Ri[NU - 1][NU - 1] = Rt[0][0] * id;
                Ri[0][NU - 1] = -Rt[0][NU - 1] * id;
                Ri[NU - 1][0] = -Rt[NU - 1][0] * id;
            }
            // RL = Ri L  (NU x NZ)
            double RL[NU][NZ];
#pragma unroll
            for (int i = 0; i < NU; ++i)
#pragma unroll
                for (int b = 0; b < NZ; ++b) {
                    double v = 0.0;
#pragma unroll
                    for (int j = 0; j < NU; ++j) v += Ri[i][j] * Lk[j][b];
                    RL[i][b] = v;
                }
            // new P = Wzz + diag(Qd) (+HG) + Abar' P Abar - L' Ri L   (upper triangle, mirrored)
            double Nxx[NX][NX], Nxt[NX][NT];
#pragma unroll
            for (int a = 0; a < NX; ++a) {
#pragma unroll
                for (int b = a; b < NX; ++b) {
                    double v = W[a * NV + b] + (kg ? m[L::M_HG + sym(a, b)] : 0.0);
#pragma unroll
                    for (int c = 0; c < NX; ++c) v += p.A[c * NX + a] * M[c][b];
#pragma unroll
                    for (int i = 0; i < NU; ++i) v -= Lk[i][a] * RL[i][b];
                    Nxx[a][b] = v;
                }
                Nxx[a][a] += s[l.o_qd + a * l.Np + k];
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    double v = W[a * NV + NX + t] + (kg ? m[L::M_HG + sym(a, NX + t)] : 0.0);
#pragma unroll
                    for (int c = 0; c < NX; ++c) v += p.A[c * NX + a] * Pxt[c][t];
#pragma unroll
                    for (int i = 0; i < NU; ++i) v -= Lk[i][a] * RL[i][NX + t];
                    Nxt[a][t] = v;
                }
            }
#pragma unroll
            for (int a = 0; a < NT; ++a)
#pragma unroll
                for (int b = a; b < NT; ++b) {
                    double v = Ptt[a][b] + W[(NX + a) * NV + NX + b] +
                               (kg ? m[L::M_HG + sym(NX + a, NX + b)] : 0.0);
#pragma unroll
                    for (int i = 0; i < NU; ++i) v -= Lk[i][NX + a] * RL[i][NX + b];
                    Ptt[a][b] = v;
                    Ptt[b][a] = v;
                }
#pragma unroll
            for (int a = 0; a < NX; ++a) {
#pragma unroll
                for (int b = a; b < NX; ++b) {
                    Pxx[a][b] = Nxx[a][b];
                    Pxx[b][a] = Nxx[a][b];
                }
#pragma unroll
                for (int t = 0; t < NT; ++t) Pxt[a][t] = Nxt[a][t];
            }
#pragma unroll
            for (int i = 0; i < NU; ++i) {
#pragma unroll
                for (int b = 0; b < NZ; ++b) s[l.o_L + (i * NZ + b) * l.Np + k] = RL[i][b];
#pragma unroll
                for (int j = 0; j < NU; ++j) s[l.o_Ri + (i * NU + j) * l.Np + k] = Ri[i][j];
            }
        }
        // theta block of P_0 -> inverse (NT = 1 or 2)
        if (NT == 1) {
            ok = ok && (Ptt[0][0] > 0.0);
            m[L::M_PTT] = 1.0 / Ptt[0][0];
        } else {
            const double det = Ptt[0][0] * Ptt[NT - 1][NT - 1] - Ptt[0][NT - 1] * Ptt[NT - 1][0];
            ok = ok && (Ptt[0][0] > 0.0) && (det > 0.0);
            const double id = 1.0 / det;
            m[L::M_PTT + 0] = Ptt[NT - 1][NT - 1] * id;
            m[L::M_PTT + NT * NT - 1] = Ptt[0][0] * id;
            m[L::M_PTT + (NT - 1)] = -Ptt[0][NT - 1] * id;
            m[L::M_PTT + (NT - 1) * NT] = -Ptt[NT - 1][0] * id;
        }
        m[L::M_PIV] = ok ? 1.0 : 0.0;
        return ok;
    }

    // ============================================================================================
    // sweep: adjoint recursion -> |r_d|inf, the reduced gradient of the Lagrangian w.r.t. (u, theta)
    // ============================================================================================
    static LB_HD void adjoint_sweep(const P& p, const L& l, double* s) {
        double pi[NZ];
        double* m = s + l.o_misc;
        const int N = p.N;
        double rdi = 0.0;
#pragma unroll
        for (int a = 0; a < NZ; ++a) pi[a] = s[l.o_g + a * l.Np + N] + (p.kg == N ? m[L::M_GGL + a] : 0.0);
        for (int k = N - 1; k >= 0; --k) {
            const bool kg = (p.kg == k);
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                double v = s[l.o_g + (NZ + i) * l.Np + k];
#pragma unroll
                for (int c = 0; c < NX; ++c) v += p.B[c * NU + i] * pi[c];
                rdi = lb_nanmax(rdi, lb_abs(v));
            }
            double np_[NX];
#pragma unroll
            for (int a = 0; a < NX; ++a) {
                double v = s[l.o_g + a * l.Np + k];
#pragma unroll
                for (int c = 0; c < NX; ++c) v += p.A[c * NX + a] * pi[c];
                np_[a] = v;
            }
#pragma unroll
            for (int a = 0; a < NX; ++a) pi[a] = np_[a] + (kg ? m[L::M_GGL + a] : 0.0);
#pragma unroll
            for (int t = 0; t < NT; ++t)
                pi[NX + t] += s[l.o_g + (NX + t) * l.Np + k] + (kg ? m[L::M_GGL + NX + t] : 0.0);
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) rdi = lb_nanmax(rdi, lb_abs(pi[NX + t]));
        m[L::M_RD] = rdi;
    }

    // ============================================================================================
    // sweep: Farkas adjoint (same recursion with G'lambda only) -> |G_red'lambda|inf at M_CERT and
    // h_red'lambda = lambda'slack + y'(G'lambda)_red accumulated into M_HLAM
    // ============================================================================================
    static LB_HD void farkas_sweep(const P& p, const L& l, double* s) {
        double pc[NZ];
        double* m = s + l.o_misc;
        const int N = p.N;
        double ci = 0.0, ydot = 0.0;
        {
            const unsigned rows = stage_rows(p, N);
#pragma unroll
            for (int a = 0; a < NZ; ++a) pc[a] = (p.kg == N) ? m[L::M_GGL + a] : 0.0;
#pragma unroll
            for (int j = 0; j < NX; ++j)
#pragma unroll
                for (int side = 0; side < 2; ++side)
                    if ((rows >> (2 * j + side)) & 1u)
                        pc[j] += (side == 0 ? 1.0 : -1.0) * s[l.o_lb + (2 * j + side) * l.Np + N];
        }
        for (int k = N - 1; k >= 0; --k) {
            const bool kg = (p.kg == k);
            const unsigned rows = stage_rows(p, k);
            double gc[NVB];
#pragma unroll
            for (int j = 0; j < NVB; ++j) {
                double v = 0.0;
#pragma unroll
                for (int side = 0; side < 2; ++side)
                    if ((rows >> (2 * j + side)) & 1u)
                        v += (side == 0 ? 1.0 : -1.0) * s[l.o_lb + (2 * j + side) * l.Np + k];
                gc[j] = v;
            }
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                double v = gc[NX + i];
#pragma unroll
                for (int c = 0; c < NX; ++c) v += p.B[c * NU + i] * pc[c];
                ci = lb_max(ci, lb_abs(v));
                ydot += v * s[l.o_u + i * l.Np + k];
            }
            double np_[NX];
#pragma unroll
            for (int a = 0; a < NX; ++a) {
                double v = gc[a];
#pragma unroll
                for (int c = 0; c < NX; ++c) v += p.A[c * NX + a] * pc[c];
                np_[a] = v;
            }
#pragma unroll
            for (int a = 0; a < NX; ++a) pc[a] = np_[a] + (kg ? m[L::M_GGL + a] : 0.0);
#pragma unroll
            for (int t = 0; t < NT; ++t) pc[NX + t] += (kg ? m[L::M_GGL + NX + t] : 0.0);
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            ci = lb_max(ci, lb_abs(pc[NX + t]));
            ydot += pc[NX + t] * m[L::M_TH + t];
        }
        m[L::M_CERT] = ci;
        m[L::M_HLAM] += ydot;
    }

    // ============================================================================================
    // sweep: backward substitution (gradient recursion) with the stored factors:
    //   rt = q_u + B'pv ; kap_k = -Ri rt ; pv <- q_z + Abar'pv - RL' rt ; d(theta) = -Ptt^-1 pv_theta
    // d(theta) goes to M_DTHA (aff) or M_DTH.
    // ============================================================================================
    static LB_HD void backward_vec(const P& p, const L& l, double* s, bool aff) {
        double pv[NZ];
        double* m = s + l.o_misc;
        const int N = p.N;
#pragma unroll
        for (int a = 0; a < NZ; ++a) {
            const double base = a < NX ? s[l.o_q + a * l.Np + N] : s[l.o_g + a * l.Np + N];
            pv[a] = base + (p.kg == N ? m[L::M_GGL + a] + m[L::M_DG + a] : 0.0);
        }
        for (int k = N - 1; k >= 0; --k) {
            const bool kg = (p.kg == k);
            double rt[NU];
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                double v0 = s[l.o_q + (NX + i) * l.Np + k], v1 = 0.0;
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                    if (c & 1) v1 += p.B[c * NU + i] * pv[c];
                    else v0 += p.B[c * NU + i] * pv[c];
                }
                rt[i] = v0 + v1;
            }
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < NU; ++j) v -= s[l.o_Ri + (i * NU + j) * l.Np + k] * rt[j];
                s[l.o_kap + i * l.Np + k] = v;
            }
            double np_[NZ];
#pragma unroll
            for (int a = 0; a < NX; ++a) {
                double v0 = s[l.o_q + a * l.Np + k], v1 = 0.0;
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                    if (c & 1) v1 += p.A[c * NX + a] * pv[c];
                    else v0 += p.A[c * NX + a] * pv[c];
                }
                double v = v0 + v1;
#pragma unroll
                for (int i = 0; i < NU; ++i) v -= s[l.o_L + (i * NZ + a) * l.Np + k] * rt[i];
                np_[a] = v;
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                double v = s[l.o_g + (NX + t) * l.Np + k] + pv[NX + t];
#pragma unroll
                for (int i = 0; i < NU; ++i) v -= s[l.o_L + (i * NZ + NX + t) * l.Np + k] * rt[i];
                np_[NX + t] = v;
            }
#pragma unroll
            for (int a = 0; a < NZ; ++a) pv[a] = np_[a] + (kg ? m[L::M_GGL + a] + m[L::M_DG + a] : 0.0);
        }
#pragma unroll
        for (int a = 0; a < NT; ++a) {
            double v = 0.0;
#pragma unroll
            for (int b = 0; b < NT; ++b) v -= m[L::M_PTT + a * NT + b] * pv[NX + b];
            m[(aff ? L::M_DTHA : L::M_DTH) + a] = v;
        }
    }

    // ============================================================================================
    // sweep: forward substitution  du_k = kap_k - RL_k dz_k ; dx_{k+1} = A dx_k + B du_k
    // ============================================================================================
    static LB_HD void forward_vec(const P& p, const L& l, double* s, bool aff) {
        const int ox = aff ? l.o_dxa : l.o_dx, ou = aff ? l.o_dua : l.o_du;
        const double* m = s + l.o_misc;
        double dx[NX], dth[NT], du[NU];
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            dx[j] = 0.0;
            s[ox + j * l.Np] = 0.0;
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) dth[t] = m[(aff ? L::M_DTHA : L::M_DTH) + t];
        for (int k = 0; k < p.N; ++k) {
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                double v0 = s[l.o_kap + i * l.Np + k], v1 = 0.0;
#pragma unroll
                for (int t = 0; t < NT; ++t) v0 -= s[l.o_L + (i * NZ + NX + t) * l.Np + k] * dth[t];
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                    if (c & 1) v1 -= s[l.o_L + (i * NZ + c) * l.Np + k] * dx[c];
                    else v0 -= s[l.o_L + (i * NZ + c) * l.Np + k] * dx[c];
                }
                du[i] = v0 + v1;
                s[ou + i * l.Np + k] = du[i];
            }
            double xn[NX];
#pragma unroll
            for (int a = 0; a < NX; ++a) {
                double v0 = 0.0, v1 = 0.0;
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                    if (c & 1) v1 += p.A[a * NX + c] * dx[c];
                    else v0 += p.A[a * NX + c] * dx[c];
                }
                double v = v0 + v1;
#pragma unroll
                for (int i = 0; i < NU; ++i) v += p.B[a * NU + i] * du[i];
                xn[a] = v;
                s[ox + a * l.Np + k + 1] = v;
            }
#pragma unroll
            for (int a = 0; a < NX; ++a) dx[a] = xn[a];
        }
    }

    // ============================================================================================
    // stage: affine (predictor) row directions of stage k.  Accumulates the step-length ratio and
    // the three sums of mu_aff; writes the sigma-independent part of the corrector rhs to q and the
    // coefficient of sigma*mu to the (dx,du) scratch:   q_cor = q + sigmu * scratch   (corr_stage)
    //   ds_a = -r_p - a dv_a ; dl_a = -lambda - w ds_a ;  -ds_a/s = -r, -dl_a/lambda = 1 + r, r = ds_a/s
    //   t = w r_p - (ds_a dl_a - sigma mu)/s
    // ============================================================================================
    static LB_HD void affine_stage(const P& p, const L& l, double* s, int k, RedStep& red) {
        const unsigned rows = stage_rows(p, k);
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
            if (j >= NX && k >= p.N) continue;
            double aj = 0.0, bj = 0.0;
            if ((rows >> (2 * j)) & 3u) {
                const double v = bvar(l, s, l.o_x, l.o_u, k, j), dva = bvar(l, s, l.o_dxa, l.o_dua, k, j);
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    if (!((rows >> (2 * j + side)) & 1u)) continue;
                    const int r = (2 * j + side) * l.Np + k;
                    const double S = s[l.o_sb + r], Lm = s[l.o_lb + r];
                    const double sgn = side == 0 ? 1.0 : -1.0;
                    const double slack = side == 0 ? p.hi[j] - v : v - p.lo[j];
                    const double rp = S - slack, is = lb_rcp(S), w = Lm * is;
                    const double dsa = -rp - sgn * dva, dla = -Lm - w * dsa;
                    const double rr = dsa * is;
                    red.ratio = lb_max(red.ratio, lb_max(-rr, 1.0 + rr));
                    red.s0 += S * Lm;
                    red.s1 += S * dla + Lm * dsa;
                    red.s2 += dsa * dla;
                    aj += sgn * (w * rp - dsa * dla * is - Lm);
                    bj += sgn * is;
                }
            }
            s[l.o_q + j * l.Np + k] = s[l.o_g + zidx(j) * l.Np + k] + aj;
            s[l.o_dx + j * l.Np + k] = bj;
        }
    }
    static LB_HD void corr_stage(const P& p, const L& l, double* s, int k, double sigmu) {
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
            if (j >= NX && k >= p.N) continue;
            s[l.o_q + j * l.Np + k] += sigmu * s[l.o_dx + j * l.Np + k];
        }
    }
    // polytope row i: acc[0..NZ) += G (t1 - lambda), acc[NZ..2NZ) += G / s   (dG = acc1 + sigmu acc2)
    static LB_HD void affine_gen_row(const P& p, const L& l, const double* s, const double* G,
                                     const double* hg, int i, double* acc, RedStep& red) {
        const double* m = s + l.o_misc;
        const double slack = gen_slack(p, l, s, G, hg, i);
        const double S = s[l.o_sg + i], Lm = s[l.o_lg + i];
        const double rp = S - slack, is = lb_rcp(S), w = Lm * is;
        double g[NZ], adva = 0.0;
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            g[a] = G[a * p.ngp + i];
            adva += g[a] * s[l.o_dxa + a * l.Np + p.kg];
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            g[NX + t] = G[(NX + t) * p.ngp + i];
            adva += g[NX + t] * m[L::M_DTHA + t];
        }
        const double dsa = -rp - adva, dla = -Lm - w * dsa, rr = dsa * is;
        red.ratio = lb_max(red.ratio, lb_max(-rr, 1.0 + rr));
        red.s0 += S * Lm;
        red.s1 += S * dla + Lm * dsa;
        red.s2 += dsa * dla;
        const double t1 = w * rp - dsa * dla * is - Lm;
#pragma unroll
        for (int a = 0; a < NZ; ++a) {
            acc[a] += g[a] * t1;
            acc[NZ + a] += g[a] * is;
        }
    }

    // ============================================================================================
    // stage: final (corrector) row directions of stage k -> parks ds, dl in the scratch block that
    // starts at o_qd, returns the step-length ratio.  update_stage applies the step.
    // ============================================================================================
    static LB_HD int scr_ds(const L& l, int j, int side, int k) { return l.o_qd + (2 * j + side) * l.Np + k; }
    static LB_HD int scr_dl(const L& l, int j, int side, int k) { return l.o_qd + (2 * NVB + 2 * j + side) * l.Np + k; }
    static LB_HD double final_stage(const P& p, const L& l, double* s, int k, double sigmu) {
        const unsigned rows = stage_rows(p, k);
        double ratio = 0.0;
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
            if (!((rows >> (2 * j)) & 3u)) continue;
            const double v = bvar(l, s, l.o_x, l.o_u, k, j), dva = bvar(l, s, l.o_dxa, l.o_dua, k, j),
                         dv = bvar(l, s, l.o_dx, l.o_du, k, j);
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                if (!((rows >> (2 * j + side)) & 1u)) continue;
                const int r = (2 * j + side) * l.Np + k;
                const double S = s[l.o_sb + r], Lm = s[l.o_lb + r];
                const double sgn = side == 0 ? 1.0 : -1.0;
                const double slack = side == 0 ? p.hi[j] - v : v - p.lo[j];
                const double rp = S - slack, is = lb_rcp(S), w = Lm * is;
                const double dsa = -rp - sgn * dva, dla = -Lm - w * dsa;
                const double ds = -rp - sgn * dv;
                const double rc = S * Lm + dsa * dla - sigmu;
                const double dl = (-rc - Lm * ds) * is;
                ratio = lb_max(ratio, lb_max(-ds * is, -dl * lb_rcp(Lm)));
                s[scr_ds(l, j, side, k)] = ds;
                s[scr_dl(l, j, side, k)] = dl;
            }
        }
        return ratio;
    }
    static LB_HD void update_stage(const P& p, const L& l, double* s, int k, double alpha) {
        const unsigned rows = stage_rows(p, k);
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                if (!((rows >> (2 * j + side)) & 1u)) continue;
                const int r = (2 * j + side) * l.Np + k;
                s[l.o_sb + r] += alpha * s[scr_ds(l, j, side, k)];
                s[l.o_lb + r] += alpha * s[scr_dl(l, j, side, k)];
            }
        }
#pragma unroll
        for (int j = 0; j < NX; ++j) s[l.o_x + j * l.Np + k] += alpha * s[l.o_dx + j * l.Np + k];
        if (k < p.N) {
#pragma unroll
            for (int j = 0; j < NU; ++j) s[l.o_u + j * l.Np + k] += alpha * s[l.o_du + j * l.Np + k];
        }
    }
    // polytope row i, final direction (recomputed by final_gen_row and update_gen_row: no scratch)
    static LB_HD void gen_final_dir(const P& p, const L& l, const double* s, const double* G, const double* hg,
                                    int i, double sigmu, double& S, double& Lm, double& ds, double& dl, double& is) {
        const double* m = s + l.o_misc;
        const double slack = gen_slack(p, l, s, G, hg, i);
        S = s[l.o_sg + i];
        Lm = s[l.o_lg + i];
        const double rp = S - slack;
        is = lb_rcp(S);
        const double w = Lm * is;
        double adva = 0.0, adv = 0.0;
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            const double g = G[a * p.ngp + i];
            adva += g * s[l.o_dxa + a * l.Np + p.kg];
            adv += g * s[l.o_dx + a * l.Np + p.kg];
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const double g = G[(NX + t) * p.ngp + i];
            adva += g * m[L::M_DTHA + t];
            adv += g * m[L::M_DTH + t];
        }
        const double dsa = -rp - adva, dla = -Lm - w * dsa;
        ds = -rp - adv;
        const double rc = S * Lm + dsa * dla - sigmu;
        dl = (-rc - Lm * ds) * is;
    }
    static LB_HD double final_gen_row(const P& p, const L& l, const double* s, const double* G, const double* hg,
                                      int i, double sigmu) {
        double S, Lm, ds, dl, is;
        gen_final_dir(p, l, s, G, hg, i, sigmu, S, Lm, ds, dl, is);
        return lb_max(-ds * is, -dl * lb_rcp(Lm));
    }
    static LB_HD void update_gen_row(const P& p, const L& l, double* s, const double* G, const double* hg, int i,
                                     double sigmu, double alpha) {
        double S, Lm, ds, dl, is;
        gen_final_dir(p, l, s, G, hg, i, sigmu, S, Lm, ds, dl, is);
        s[l.o_sg + i] = S + alpha * ds;
        s[l.o_lg + i] = Lm + alpha * dl;
    }

    // stage: objective contribution 0.5 v'W v (+ lin'z at kT)
    static LB_HD double objective_stage(const P& p, const L& l, const double* s, int k) {
        double v[NV];
#pragma unroll
        for (int j = 0; j < NX; ++j) v[j] = s[l.o_x + j * l.Np + k];
#pragma unroll
        for (int j = 0; j < NT; ++j) v[NX + j] = s[l.o_misc + L::M_TH + j];
        const bool last = k >= p.N;
#pragma unroll
        for (int j = 0; j < NU; ++j) v[NZ + j] = last ? 0.0 : s[l.o_u + j * l.Np + k];
        const double* W = p.W[stage_type(p, k)];
        double J = 0.0;
#pragma unroll
        for (int a = 0; a < NV; ++a) {
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < NV; ++b) acc += W[a * NV + b] * v[b];
            J += 0.5 * v[a] * acc;
        }
        if (k == p.kT) {
#pragma unroll
            for (int a = 0; a < NZ; ++a) J += s[l.o_misc + L::M_LIN + a] * v[a];
        }
        return J;
    }

    // termination / verdict after the factor sweep (same rule as oracle/lbmpc_oracle.c)
    //   returns -1 to continue, else the LBMPC_ST_* status
    static LB_HD int verdict(const P& p, const double* m, bool cert) {
        const double rd = m[L::M_RD], rp = m[L::M_RP], mu = m[L::M_MU], lam = m[L::M_LAM];
        const bool pivots_ok = m[L::M_PIV] > 0.5;
        if (!pivots_ok || !(rd == rd) || !(rp == rp) || !(mu == mu) || isinf(rd) || isinf(mu)) return 3;
        const double rd_tol = p.tol_res * (100.0 * lam > 1.0 ? 100.0 * lam : 1.0);
        if (rd < rd_tol && rp < p.tol_res && mu < p.tol_mu) return 0;
        if (cert && m[L::M_HLAM] < 0.0 && m[L::M_CERT] * p.inf_radius <= -m[L::M_HLAM]) return 2;
        return -1;
    }
};

// ------------------------------------------------------------------------------------------------
// Cooperative Riccati factorisation for NT = NU = 1 (the Moore-Greitzer shape): 16 lanes per QP.
//
// Lane h < NH owns one unique entry (a,b), a <= b, of the symmetric NZ x NZ cost-to-go matrix, kept
// SCALED by the pivot of the stage it came from:  Pt = rho P, ir = 1/rho.  One stage is
//     F    = Abar_e' Pt Abar_e,  Abar_e = [Abar Bbar]        (every entry a fixed linear form in the
//                                                             NH unique entries of Pt: dot products
//                                                             with per-lane coefficient vectors)
//     Rt   = Wuu + Qd_u + F_uu ir ,  L = Wuz + F_uz ir
//     Pt'  = Rt (Wzz + Qd + [HG] + F_zz ir) - L'L ,  rho' = Rt
// which is P' = Wzz + Qd + Abar'P Abar - L'L/Rt multiplied through by Rt: the reciprocal of the
// pivot is only needed one stage later (and for the stored factors RL = L/Rt, Ri = 1/Rt), so it
// overlaps the next stage's exchange and dot products instead of sitting on the dependency chain.
// Lanes 0..NX-1 also produce F_uz[a] (a < NX), lane NX produces F_uu, lane 15 produces F_uz[theta].
// Two exchanges per stage through a small shared buffer: xch (the NH entries), xf (F_uz, F_uu).
// The three lane-phases st1/st2/st3 are separated by __syncwarp() in the kernel; tests/emul runs
// them as loops over the 16 lanes.
// ------------------------------------------------------------------------------------------------
template <int NX>
struct Coop {
    static constexpr int NT = 1, NU = 1, NZ = NX + 1, NV = NZ + 1, NH = NZ * (NZ + 1) / 2, NXX = NX * (NX + 1) / 2;
    static constexpr int kLanes = 16, kXch = 16, kXf = 8;
    static_assert(NH <= 15 && NZ + 1 <= kXf, "shape does not fit the 16-lane mapping");
    using P = Params<NX, 1, 1>;
    using L = Layout<NX, 1, 1>;
    using C = Core<NX, 1, 1>;

    // unique entries ordered: xx block row-wise, then (a,theta), then (theta,theta)
    static LB_HD constexpr int ent(int a, int b) {  // a <= b
        return b < NX ? a * NX - a * (a - 1) / 2 + (b - a) : (a < NX ? NXX + a : NH - 1);
    }
    struct Lane {
        int a, b;           // entry owned (h < NH)
        bool isP, diag;
        double c1[NH];      // F_zz[a][b] (h < NH) or F_uz[theta] (h == 15) as a linear form in the entries
        double c2[NXX];     // F_uz[h] (h < NX) or F_uu (h == NX) as a linear form in the xx entries
        double wzz, wa, wb, wuu;
        double pt, ir, d1;
        bool ok;
    };
    static LB_HD double abar(const P& p, int c, int a) {  // Abar[c][a]
        return c < NX ? (a < NX ? p.A[c * NX + a] : 0.0) : (a == NX ? 1.0 : 0.0);
    }
    static LB_HD void lane_init(const P& p, int h, Lane& ln) {
        ln.isP = h < NH;
        ln.a = 0;
        ln.b = 0;
        // decode (a,b) of entry h
#pragma unroll
        for (int a = 0; a < NZ; ++a)
#pragma unroll
            for (int b = a; b < NZ; ++b)
                if (ent(a, b) == h) {
                    ln.a = a;
                    ln.b = b;
                }
        ln.diag = ln.isP && ln.a == ln.b;
#pragma unroll
        for (int c = 0; c < NZ; ++c)
#pragma unroll
            for (int d = c; d < NZ; ++d) {
                double v = 0.0;
                if (ln.isP) {
                    v = abar(p, c, ln.a) * abar(p, d, ln.b);
                    if (c != d) v += abar(p, d, ln.a) * abar(p, c, ln.b);
                } else if (h == 15) {  // F_uz[theta] = sum_c B[c] Pt[c][theta]
                    v = (c < NX && d == NX) ? p.B[c] : 0.0;
                }
                ln.c1[ent(c, d)] = v;
            }
#pragma unroll
        for (int c = 0; c < NX; ++c)
#pragma unroll
            for (int d = c; d < NX; ++d) {
                double v = 0.0;
                if (h < NX) {
                    v = p.A[c * NX + h] * p.B[d];
                    if (c != d) v += p.A[d * NX + h] * p.B[c];
                } else if (h == NX) {
                    v = p.B[c] * p.B[d] * (c != d ? 2.0 : 1.0);
                }
                ln.c2[ent(c, d)] = v;
            }
        ln.ok = true;
        ln.pt = 0.0;
        ln.ir = 1.0;
        ln.d1 = 0.0;
        ln.wzz = ln.wa = ln.wb = ln.wuu = 0.0;
    }
    static LB_HD void load_type(const P& p, int t, Lane& ln) {
        const double* W = p.W[t];
        ln.wzz = W[ln.a * NV + ln.b];
        ln.wa = W[NZ * NV + ln.a];
        ln.wb = W[NZ * NV + ln.b];
        ln.wuu = W[NZ * NV + NZ];
    }
    // terminal stage: Pt = Wzz + Qd (+HG), rho = 1
    static LB_HD void terminal(const P& p, const L& l, const double* s, Lane& ln) {
        const double* m = s + l.o_misc;
        const int N = p.N;
        load_type(p, C::stage_type(p, N), ln);
        double v = ln.wzz;
        if (ln.diag && ln.a < NX) v += s[l.o_qd + ln.a * l.Np + N];
        if (p.kg == N) v += m[L::M_HG + C::sym(ln.a, ln.b)];
        ln.pt = ln.isP ? v : 0.0;
        ln.ir = 1.0;
        ln.ok = true;
    }
    static LB_HD void st1(const Lane& ln, int h, double* xch) {
        if (h < NH) xch[h] = ln.pt;
    }
    static LB_HD void st2(Lane& ln, int h, const double* xch, double* xf) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
        for (int j = 0; j < NH; ++j) {
            const double v = xch[j];
            if (j % 3 == 0) a0 += ln.c1[j] * v;
            else if (j % 3 == 1) a1 += ln.c1[j] * v;
            else a2 += ln.c1[j] * v;
            if (j < NXX) {
                if (j & 1) b1 += ln.c2[j] * v;
                else b0 += ln.c2[j] * v;
            }
        }
        ln.d1 = (a0 + a1) + a2;
        const double d2 = b0 + b1;
        if (h < NX) xf[h] = d2;            // F_uz[h]
        else if (h == NX) xf[NZ] = d2;     // F_uu
        else if (h == 15) xf[NX] = ln.d1;  // F_uz[theta]
    }
    // act = false: the half-warp has no running QP (nothing is stored)
    static LB_HD void st3(const P& p, const L& l, double* s, int k, int h, Lane& ln, const double* xf, bool act) {
        if (!ln.isP) return;
        const double fa = xf[ln.a], fb = xf[ln.b], fuu = xf[NZ];
        const double ir = ln.ir;
        const double Rt = (ln.wuu + s[l.o_qd + NX * l.Np + k]) + fuu * ir;
        const double La = ln.wa + fa * ir, Lb = ln.wb + fb * ir;
        double hz = ln.wzz + ln.d1 * ir;
        if (ln.diag && ln.a < NX) hz += s[l.o_qd + ln.a * l.Np + k];
        if (k == p.kg) hz += s[l.o_misc + L::M_HG + C::sym(ln.a, ln.b)];
        ln.pt = Rt * hz - La * Lb;
        ln.ok = ln.ok && (Rt > 0.0);
        const double irn = lb_rcp(Rt);
        ln.ir = irn;
        if (act) {
            if (ln.diag) s[l.o_L + ln.a * l.Np + k] = La * irn;  // RL_k[a]
            if (h == 0) s[l.o_Ri + k] = irn;
        }
    }
    // after stage 0: inverse of the theta block of P_0 (lane of the (theta,theta) entry)
    static LB_HD void finish(const L& l, double* s, int h, Lane& ln, bool act) {
        if (h == NH - 1) {
            const double ptt = ln.pt * ln.ir;
            ln.ok = ln.ok && (ptt > 0.0);
            if (act) s[l.o_misc + L::M_PTT] = 1.0 / ptt;
        }
    }
};

}  // namespace lbmpc
