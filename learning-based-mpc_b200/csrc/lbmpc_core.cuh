// lbmpc_core.cuh — per-QP math of the batched Mehrotra/Riccati interior-point engine.
//
// Everything here is a small __host__ __device__ function that works on ONE QP whose whole
// iterate lives in a flat double buffer (a shared-memory "slot" on the GPU).  The functions come
// in two kinds, matching how the kernel (lbmpc_kernels.cu) maps them onto threads:
//   * stage/row functions  — independent per horizon stage k or per polytope row i; the kernel
//     runs them with one warp per QP, lanes striding over stages/rows, and combines the small
//     reduction structs with warp shuffles;
//   * sweep functions      — the sequential backward Riccati / forward substitution recursions
//     over the N stages; the kernel runs them thread-local, ONE LANE PER QP, so that the 32
//     lanes of a sweep warp advance 32 different QPs in lock step with no communication.
//
// The optimisation problem is the reference's per-step (LB)MPC problem in canonical stage form
// (reference files cited in include/lbmpc.h and lbmpc_capi.cu where the
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
                         M_SIZE = M_OBJ + 1;
    int Np, ngp;
    int o_x, o_u, o_sb, o_lb, o_qd, o_g, o_dgp, o_L, o_Ri, o_kap, o_dx, o_du, o_dxa, o_dua, o_sg,
        o_lg, o_misc, stride;
    LB_HD Layout(int N, int ngp_) {
        Np = N + 1;
        ngp = ngp_;
        int o = 0;
        o_x = o;   o += NX * Np;
        o_u = o;   o += NU * Np;
        o_sb = o;  o += NVB * 2 * Np;
        o_lb = o;  o += NVB * 2 * Np;
        o_qd = o;  o += NVB * Np;
        o_g = o;   o += NV * Np;
        o_dgp = o; o += NVB * Np;
        o_L = o;   o += NU * NZ * Np;
        o_Ri = o;  o += NU * NU * Np;
        o_kap = o; o += NU * Np;
        o_dx = o;  o += NX * Np;
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
    static LB_HD bool row_on(const P& p, int k, int j, int side) {
        const bool in = (j < NX) ? (k >= p.kx0 && k <= p.kx1) : (k >= p.ku0 && k <= p.ku1);
        return in && ((p.rowmask >> (2 * j + side)) & 1u);
    }
    static LB_HD int sym(int a, int b) {  // packed index of symmetric NZ x NZ, a<=b
        return a * NZ - a * (a - 1) / 2 + (b - a);
    }
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
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
            if (j >= NX && k >= p.N) continue;
            const double v = bvar(l, s, l.o_x, l.o_u, k, j);
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                if (!row_on(p, k, j, side)) continue;
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
    // dgp (Newton rhs minus G'lambda on the bounded variables) of stage k; accumulates reductions.
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
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
            double qd = 0.0, gl = 0.0, gp = 0.0;
            const int a = j < NX ? j : NZ + (j - NX);
            if (!(j >= NX && last)) {
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    if (!row_on(p, k, j, side)) continue;
                    const int r = (2 * j + side) * l.Np + k;
                    const double S = s[l.o_sb + r], Lm = s[l.o_lb + r];
                    const double sgn = side == 0 ? 1.0 : -1.0;
                    const double slack = side == 0 ? p.hi[j] - v[a] : v[a] - p.lo[j];
                    const double rp = S - slack;
                    const double w = Lm / S;
                    qd += w;
                    gl += sgn * Lm;
                    gp += sgn * (w * rp);
                    red.rp = lb_max(red.rp, lb_abs(rp));
                    red.sl += S * Lm;
                    red.lam = lb_max(red.lam, Lm);
                    red.hl += Lm * slack;
                }
            }
            s[l.o_qd + j * l.Np + k] = qd;
            s[l.o_dgp + j * l.Np + k] = gp - gl;
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
        const double rp = S - slack, w = Lm / S, t = w * rp;
        red.rp = lb_max(red.rp, lb_abs(rp));
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
    // sweep: backward pass.  FACTOR: Riccati factorisation over z=[x;theta] with Abar=diag(A,I),
    // Bbar=[B;0] (stores L_k, Ri_k) + adjoint recursion -> |r_d|inf (+ Farkas adjoint when cert).
    // Always: gradient recursion -> kap_k and d(theta) (stored at M_DTHA if aff else M_DTH).
    // Returns false when a pivot is not positive / not finite.
    // ============================================================================================
    template <bool FACTOR>
    static LB_HD bool backward(const P& p, const L& l, double* s, bool aff, bool cert) {
        double Pxx[NX][NX], Pxt[NX][NT], Ptt[NT][NT];
        double pv[NZ], pi[NZ], pc[NZ];
        double* m = s + l.o_misc;
        const int N = p.N;
        double rdi = 0.0, ci = 0.0, ydot = 0.0;
        bool ok = true;
        // ---- terminal stage ----
        {
            const double* W = p.W[stage_type(p, N)];
            const bool kg = (p.kg == N);
            if (FACTOR) {
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
#pragma unroll
            for (int a = 0; a < NZ; ++a) {
                const double g = s[l.o_g + a * l.Np + N];
                const double ggl = kg ? m[L::M_GGL + a] : 0.0;
                pi[a] = g + ggl;
                pv[a] = g + ggl + (a < NX ? s[l.o_dgp + a * l.Np + N] : 0.0) + (kg ? m[L::M_DG + a] : 0.0);
                pc[a] = ggl;
            }
            if (FACTOR && cert) {
#pragma unroll
                for (int j = 0; j < NX; ++j)
#pragma unroll
                    for (int side = 0; side < 2; ++side)
                        if (row_on(p, N, j, side))
                            pc[j] += (side == 0 ? 1.0 : -1.0) * s[l.o_lb + (2 * j + side) * l.Np + N];
            }
        }
        for (int k = N - 1; k >= 0; --k) {
            const double* W = p.W[stage_type(p, k)];
            const bool kg = (p.kg == k);
            double Lk[NU][NZ], Ri[NU][NU];
            if (FACTOR) {
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
                    for (int b = 0; b < NZ; ++b) s[l.o_L + (i * NZ + b) * l.Np + k] = Lk[i][b];
#pragma unroll
                    for (int j = 0; j < NU; ++j) s[l.o_Ri + (i * NU + j) * l.Np + k] = Ri[i][j];
                }
            } else {
#pragma unroll
                for (int i = 0; i < NU; ++i) {
#pragma unroll
                    for (int b = 0; b < NZ; ++b) Lk[i][b] = s[l.o_L + (i * NZ + b) * l.Np + k];
#pragma unroll
                    for (int j = 0; j < NU; ++j) Ri[i][j] = s[l.o_Ri + (i * NU + j) * l.Np + k];
                }
            }
            // ---- gradient recursion (both solves) ----
            double gz[NZ], gu[NU], rt[NU], kap[NU];
#pragma unroll
            for (int a = 0; a < NZ; ++a) gz[a] = s[l.o_g + a * l.Np + k];
#pragma unroll
            for (int i = 0; i < NU; ++i) gu[i] = s[l.o_g + (NZ + i) * l.Np + k];
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                double v = gu[i] + s[l.o_dgp + (NX + i) * l.Np + k];
#pragma unroll
                for (int c = 0; c < NX; ++c) v += p.B[c * NU + i] * pv[c];
                rt[i] = v;
            }
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < NU; ++j) v -= Ri[i][j] * rt[j];
                kap[i] = v;
                s[l.o_kap + i * l.Np + k] = v;
            }
            double np_[NZ];
#pragma unroll
            for (int a = 0; a < NX; ++a) {
                double v = gz[a] + s[l.o_dgp + a * l.Np + k];
#pragma unroll
                for (int c = 0; c < NX; ++c) v += p.A[c * NX + a] * pv[c];
#pragma unroll
                for (int i = 0; i < NU; ++i) v += Lk[i][a] * kap[i];
                np_[a] = v;
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                double v = gz[NX + t] + pv[NX + t];
#pragma unroll
                for (int i = 0; i < NU; ++i) v += Lk[i][NX + t] * kap[i];
                np_[NX + t] = v;
            }
#pragma unroll
            for (int a = 0; a < NZ; ++a) pv[a] = np_[a] + (kg ? m[L::M_GGL + a] + m[L::M_DG + a] : 0.0);
            if (FACTOR) {
                // ---- adjoint recursion: reduced gradient of the Lagrangian ----
#pragma unroll
                for (int i = 0; i < NU; ++i) {
                    double v = gu[i];
#pragma unroll
                    for (int c = 0; c < NX; ++c) v += p.B[c * NU + i] * pi[c];
                    rdi = (v == v) ? lb_max(rdi, lb_abs(v)) : v;
                }
#pragma unroll
                for (int a = 0; a < NX; ++a) {
                    double v = gz[a];
#pragma unroll
                    for (int c = 0; c < NX; ++c) v += p.A[c * NX + a] * pi[c];
                    np_[a] = v;
                }
#pragma unroll
                for (int a = 0; a < NX; ++a) pi[a] = np_[a] + (kg ? m[L::M_GGL + a] : 0.0);
#pragma unroll
                for (int t = 0; t < NT; ++t) pi[NX + t] += gz[NX + t] + (kg ? m[L::M_GGL + NX + t] : 0.0);
                if (cert) {
                    // Farkas adjoint: same recursion with G'lambda only
                    double gc[NVB];
#pragma unroll
                    for (int j = 0; j < NVB; ++j) {
                        double v = 0.0;
#pragma unroll
                        for (int side = 0; side < 2; ++side)
                            if (row_on(p, k, j, side))
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
            }
        }
        if (FACTOR) {
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
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const double v = pi[NX + t];
                rdi = (v == v) ? lb_max(rdi, lb_abs(v)) : v;
                if (cert) {
                    ci = lb_max(ci, lb_abs(pc[NX + t]));
                    ydot += pc[NX + t] * m[L::M_TH + t];
                }
            }
            m[L::M_RD] = rdi;
            if (cert) {
                m[L::M_CERT] = ci;
                m[L::M_HLAM] += ydot;  // h_red' lambda = lambda' slack + y' (G' lambda)_red
            }
        }
#pragma unroll
        for (int a = 0; a < NT; ++a) {
            double v = 0.0;
#pragma unroll
            for (int b = 0; b < NT; ++b) v -= m[L::M_PTT + a * NT + b] * pv[NX + b];
            m[(aff ? L::M_DTHA : L::M_DTH) + a] = v;
        }
        return ok;
    }

    // ============================================================================================
    // sweep: forward substitution  du_k = kap_k - Ri_k L_k dz_k ; dx_{k+1} = A dx_k + B du_k
    // ============================================================================================
    static LB_HD void forward(const P& p, const L& l, double* s, bool aff) {
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
            double Lz[NU];
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                double v = 0.0;
#pragma unroll
                for (int c = 0; c < NX; ++c) v += s[l.o_L + (i * NZ + c) * l.Np + k] * dx[c];
#pragma unroll
                for (int t = 0; t < NT; ++t) v += s[l.o_L + (i * NZ + NX + t) * l.Np + k] * dth[t];
                Lz[i] = v;
            }
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                double v = s[l.o_kap + i * l.Np + k];
#pragma unroll
                for (int j = 0; j < NU; ++j) v -= s[l.o_Ri + (i * NU + j) * l.Np + k] * Lz[j];
                du[i] = v;
                s[ou + i * l.Np + k] = v;
            }
            double xn[NX];
#pragma unroll
            for (int a = 0; a < NX; ++a) {
                double v = 0.0;
#pragma unroll
                for (int c = 0; c < NX; ++c) v += p.A[a * NX + c] * dx[c];
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
    // per-row Newton directions.  MODE 0: affine (predictor) ; MODE 1: final (corrector).
    // ============================================================================================
    struct RowDir {
        double ds, dl, inv_s, rp, w;
    };
    // box row (k, j, side); S, Lm current slack/multiplier; sigmu only used in MODE 1
    template <int MODE>
    static LB_HD RowDir box_dir(const P& p, const L& l, const double* s, int k, int j, int side,
                                double S, double Lm, double sigmu) {
        RowDir r;
        const double v = bvar(l, s, l.o_x, l.o_u, k, j);
        const double sgn = side == 0 ? 1.0 : -1.0;
        const double slack = side == 0 ? p.hi[j] - v : v - p.lo[j];
        r.rp = S - slack;
        r.inv_s = 1.0 / S;
        r.w = Lm * r.inv_s;
        const double dsa = -r.rp - sgn * bvar(l, s, l.o_dxa, l.o_dua, k, j);
        const double dla = -Lm - r.w * dsa;
        if (MODE == 0) {
            r.ds = dsa;
            r.dl = dla;
        } else {
            const double ds = -r.rp - sgn * bvar(l, s, l.o_dx, l.o_du, k, j);
            const double rc = S * Lm + dsa * dla - sigmu;
            r.ds = ds;
            r.dl = (-rc - Lm * ds) * r.inv_s;
        }
        return r;
    }
    template <int MODE>
    static LB_HD RowDir gen_dir(const P& p, const L& l, const double* s, const double* G,
                                const double* hg, int i, double S, double Lm, double sigmu) {
        RowDir r;
        const double* m = s + l.o_misc;
        const double slack = gen_slack(p, l, s, G, hg, i);
        r.rp = S - slack;
        r.inv_s = 1.0 / S;
        r.w = Lm * r.inv_s;
        double adva = 0.0, adv = 0.0;
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            const double g = G[a * p.ngp + i];
            adva += g * s[l.o_dxa + a * l.Np + p.kg];
            if (MODE == 1) adv += g * s[l.o_dx + a * l.Np + p.kg];
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const double g = G[(NX + t) * p.ngp + i];
            adva += g * m[L::M_DTHA + t];
            if (MODE == 1) adv += g * m[L::M_DTH + t];
        }
        const double dsa = -r.rp - adva;
        const double dla = -Lm - r.w * dsa;
        if (MODE == 0) {
            r.ds = dsa;
            r.dl = dla;
        } else {
            const double ds = -r.rp - adv;
            const double rc = S * Lm + dsa * dla - sigmu;
            r.ds = ds;
            r.dl = (-rc - Lm * ds) * r.inv_s;
        }
        return r;
    }
    static LB_HD void step_acc(RedStep& red, const RowDir& r, double S, double Lm) {
        if (r.ds < 0.0) red.ratio = lb_max(red.ratio, -r.ds * r.inv_s);
        if (r.dl < 0.0) red.ratio = lb_max(red.ratio, -r.dl / Lm);
        red.s0 += S * Lm;
        red.s1 += S * r.dl + Lm * r.ds;
        red.s2 += r.ds * r.dl;
    }

    // stage: step-length pass over the box rows of stage k
    template <int MODE>
    static LB_HD void step_stage(const P& p, const L& l, const double* s, int k, double sigmu,
                                 RedStep& red) {
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
            if (j >= NX && k >= p.N) continue;
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                if (!row_on(p, k, j, side)) continue;
                const int r = (2 * j + side) * l.Np + k;
                const double S = s[l.o_sb + r], Lm = s[l.o_lb + r];
                step_acc(red, box_dir<MODE>(p, l, s, k, j, side, S, Lm, sigmu), S, Lm);
            }
        }
    }
    template <int MODE>
    static LB_HD void step_gen_row(const P& p, const L& l, const double* s, const double* G,
                                   const double* hg, int i, double sigmu, RedStep& red) {
        const double S = s[l.o_sg + i], Lm = s[l.o_lg + i];
        step_acc(red, gen_dir<MODE>(p, l, s, G, hg, i, S, Lm, sigmu), S, Lm);
    }

    // stage: corrector rhs  t = w r_p - (ds_a dl_a - sigma mu)/s  -> dgp
    static LB_HD void corrector_stage(const P& p, const L& l, double* s, int k, double sigmu) {
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
            if (j >= NX && k >= p.N) continue;
            double gp = 0.0, gl = 0.0;
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                if (!row_on(p, k, j, side)) continue;
                const int r = (2 * j + side) * l.Np + k;
                const double S = s[l.o_sb + r], Lm = s[l.o_lb + r];
                const double sgn = side == 0 ? 1.0 : -1.0;
                const RowDir d = box_dir<0>(p, l, s, k, j, side, S, Lm, 0.0);
                const double t = d.w * d.rp - (d.ds * d.dl - sigmu) * d.inv_s;
                gp += sgn * t;
                gl += sgn * Lm;
            }
            s[l.o_dgp + j * l.Np + k] = gp - gl;
        }
    }
    // row: corrector rhs of polytope row i -> acc[NZ] (dG)
    static LB_HD void corrector_gen_row(const P& p, const L& l, const double* s, const double* G,
                                        const double* hg, int i, double sigmu, double* acc) {
        const double S = s[l.o_sg + i], Lm = s[l.o_lg + i];
        const RowDir d = gen_dir<0>(p, l, s, G, hg, i, S, Lm, 0.0);
        const double t = d.w * d.rp - (d.ds * d.dl - sigmu) * d.inv_s;
#pragma unroll
        for (int a = 0; a < NZ; ++a) acc[a] += G[a * p.ngp + i] * (t - Lm);
    }

    // stage: apply the step to the rows of stage k, then to x_k, u_k
    static LB_HD void update_stage(const P& p, const L& l, double* s, int k, double sigmu, double alpha) {
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
            if (j >= NX && k >= p.N) continue;
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                if (!row_on(p, k, j, side)) continue;
                const int r = (2 * j + side) * l.Np + k;
                const double S = s[l.o_sb + r], Lm = s[l.o_lb + r];
                const RowDir d = box_dir<1>(p, l, s, k, j, side, S, Lm, sigmu);
                s[l.o_sb + r] = S + alpha * d.ds;
                s[l.o_lb + r] = Lm + alpha * d.dl;
            }
        }
#pragma unroll
        for (int j = 0; j < NX; ++j) s[l.o_x + j * l.Np + k] += alpha * s[l.o_dx + j * l.Np + k];
        if (k < p.N) {
#pragma unroll
            for (int j = 0; j < NU; ++j) s[l.o_u + j * l.Np + k] += alpha * s[l.o_du + j * l.Np + k];
        }
    }
    static LB_HD void update_gen_row(const P& p, const L& l, double* s, const double* G,
                                     const double* hg, int i, double sigmu, double alpha) {
        const double S = s[l.o_sg + i], Lm = s[l.o_lg + i];
        const RowDir d = gen_dir<1>(p, l, s, G, hg, i, S, Lm, sigmu);
        s[l.o_sg + i] = S + alpha * d.ds;
        s[l.o_lg + i] = Lm + alpha * d.dl;
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
    static LB_HD int verdict(const P& p, const double* m, bool pivots_ok, bool cert) {
        const double rd = m[L::M_RD], rp = m[L::M_RP], mu = m[L::M_MU], lam = m[L::M_LAM];
        if (!pivots_ok || !(rd == rd) || !(rp == rp) || !(mu == mu) || isinf(rd) || isinf(mu)) return 3;
        const double rd_tol = p.tol_res * (100.0 * lam > 1.0 ? 100.0 * lam : 1.0);
        if (rd < rd_tol && rp < p.tol_res && mu < p.tol_mu) return 0;
        if (cert && m[L::M_HLAM] < 0.0 && m[L::M_CERT] * p.inf_radius <= -m[L::M_HLAM]) return 2;
        return -1;
    }
};

}  // namespace lbmpc
