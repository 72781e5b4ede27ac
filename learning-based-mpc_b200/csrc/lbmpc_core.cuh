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
    double Apow[NX * NX];                                          // A^bm, bm = Layout::block_len(N) (blocked adjoint recursion)
    double W[kMaxTypes][NV * NV];                                  // stage Hessians, v=[x;theta;u]
    double lo[NVB], hi[NVB];
    double Lref[NZ * NX], Tm[NX * NX];
    double tol_res, tol_mu, inf_trigger, inv_m;
    // Farkas test: status 2 when h_red'lambda < 0 and inf_scale * sum_j |(G_red'lambda)_j| ybar_j <= -h_red'lambda, ybar_j an
    // upper bound of |y_j| on the feasible set: fk_u (input bound) inside the stage range of the input rows, fk_free else
    double fk_u[NU], fk_th[NT], fk_free, inf_scale;   // fk_th: |theta_t| on the feasible set, derived from the polytope block (lbmpc_problem.hpp)
};

// ------------------------------------------------------------------------------------------------
// slot layout (offsets in doubles).  Three arrays of per-stage RECORDS ([stage][field]) so that
//  - a lane that walks the horizon sequentially (the sweeps) addresses every operand of a stage as
//    base + compile-time offset (one pointer bump per stage, no index arithmetic), and
//  - lanes that stride over stages (the row phases) hit different banks: record sizes are odd.
// The slot stride is odd too, so lanes striding over slots are conflict-free as well.
//   R1 (iterate)    x[NX] u[NU] s[2 NVB] lambda[2 NVB]
//   R2 (Newton)     qd[NVB] q[NVB] g[NV] RL[NU NZ] Ri[NU NU] kap[NU]
//   R3 (directions) dx[NX] du[NU] dxa[NX] dua[NU]
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
                         M_PIV = M_OBJ + 1, M_GTH = M_PIV + 1, M_SIZE = M_GTH + 1;
    // field offsets inside the records
    static constexpr int F_X = 0, F_U = NX, F_S = NVB, F_LB = 3 * NVB, RS1 = (5 * NVB) | 1;
    static constexpr int F_QD = 0, F_Q = NVB, F_G = 2 * NVB, F_RL = F_G + NV, F_RI = F_RL + NU * NZ,
                         F_KAP = F_RI + NU * NU, RS2 = (F_KAP + NU) | 1;
    static constexpr int F_DX = 0, F_DU = NX, F_DXA = NVB, F_DUA = NVB + NX, F_RDK = 2 * NVB, RS3 = (2 * NVB) | 1;
    static_assert(NV <= RS3, "Farkas scratch does not fit record R3");
    // [qd | q | g | RL] are dead once the corrector sweeps are done: the final step-length pass
    // parks the row directions ds, dl (2 x 2 NVB per stage) there
    static_assert(F_RI >= 4 * NVB, "scratch for the row directions does not fit");
    // blocked substitution sweeps: per block of bm stages the NX columns of the block transfer matrix
    // ([Tx; t_theta], NZ values each) and one NZ-vector (local result / incoming state)
    static constexpr int TS = NX * NZ, BS = (TS + NZ) | 1;
    int Np, ngp, bm, nb;
    int o_r1, o_r2, o_r3, o_blk, o_sg, o_lg, o_misc, stride;
    // large polytope blocks (616 rows = 9.9 KB of slacks / multipliers per QP): optionally kept in a global (L2-resident) array
    // instead of the slot, which is what bounds the resident QPs per SM for that set; lanes stride over rows, so the accesses
    // are coalesced.  nullptr: in the slot.
    double* gsg = nullptr;
    LB_HD double* sgp(double* s) const { return gsg ? gsg : s + o_sg; }
    LB_HD double* lgp(double* s) const { return gsg ? gsg + ngp : s + o_lg; }
    LB_HD const double* sgp(const double* s) const { return gsg ? gsg : s + o_sg; }
    LB_HD const double* lgp(const double* s) const { return gsg ? gsg + ngp : s + o_lg; }
    static LB_HD int block_len(int N) {  // ~0.6 sqrt(N), at least N/32 (one lane per block)
        int m = 1;
        while (25 * m * m < 9 * N) ++m;
        const int mmin = (N + 31) / 32;
        return m > mmin ? m : mmin;
    }
    LB_HD Layout(int N, int ngp_, bool poly_global = false) {
        Np = N + 1;
        ngp = ngp_;
        bm = block_len(N);
        nb = (N + bm - 1) / bm;
        int o = 0;
        o_r1 = o;  o += RS1 * Np;
        o_r2 = o;  o += RS2 * Np;
        o_r3 = o;  o += RS3 * Np;
        o_blk = o; o += BS * nb;
        o_sg = o;  o += poly_global ? 0 : ngp;
        o_lg = o;  o += poly_global ? 0 : ngp;
        o_misc = o; o += M_SIZE;
        stride = o | 1;  // odd
    }
    LB_HD int r1(int k) const { return o_r1 + k * RS1; }
    LB_HD int r2(int k) const { return o_r2 + k * RS2; }
    LB_HD int r3(int k) const { return o_r3 + k * RS3; }
    LB_HD int blk(int b) const { return o_blk + b * BS; }
    LB_HD int i_x(int j, int k) const { return r1(k) + F_X + j; }
    LB_HD int i_u(int i, int k) const { return r1(k) + F_U + i; }
    LB_HD int i_v(int j, int k) const { return r1(k) + j; }            // bounded variable j of [x;u]
    LB_HD int i_sb(int r, int k) const { return r1(k) + F_S + r; }     // r = 2*j+side
    LB_HD int i_lb(int r, int k) const { return r1(k) + F_LB + r; }
    LB_HD int i_qd(int j, int k) const { return r2(k) + F_QD + j; }
    LB_HD int i_q(int j, int k) const { return r2(k) + F_Q + j; }
    LB_HD int i_g(int a, int k) const { return r2(k) + F_G + a; }
    LB_HD int i_L(int j, int k) const { return r2(k) + F_RL + j; }     // RL[i*NZ+b]
    LB_HD int i_Ri(int j, int k) const { return r2(k) + F_RI + j; }
    LB_HD int i_kap(int i, int k) const { return r2(k) + F_KAP + i; }
    LB_HD int i_scr_ds(int r, int k) const { return r2(k) + r; }
    LB_HD int i_scr_dl(int r, int k) const { return r2(k) + 2 * NVB + r; }
    LB_HD int i_dv(int j, int k, bool aff) const { return r3(k) + (aff ? NVB : 0) + j; }  // [dx;du] / [dxa;dua]
};

// Cost shift of one QP: the objective is evaluated at [x_k + ex_k ; theta ; u_k + eu_k] while the dynamics and every row act
// on (x_k, u_k).  p points at (N+1) records of `stride` doubles: ex (NX) and, when stride > NX, eu (NU) behind it.
//   twin state sequences with the oracle frozen: C-form e_{k+1} = A e_k + d_k (DMS_LBMPC_casadi.m:252-319); F-form, where the
//   input is u = K x + c (transitionLearned.m:13 vs transitionNominal.m:12): e_{k+1} = (A + B K) e_k + d_k and eu_k = K e_k
//   (costLBMPC.m:27 rolls the learned model, constraintsLBMPC.m:23 the nominal one).
struct CShift {
    const double* p;
    int stride;
};

struct RedAsm {  // reductions of the assembly pass
    double rp, sl, lam, hl;  // max |r_p|, sum s*lambda, max lambda, sum lambda*slack
    double gth;              // sum over the stages of g_theta (NT = 1): theta component of the dual residual
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

// ------------------------------------------------------------------------------------------------
// Shared-memory addresses of the latency-critical loops.  On the device they are 32-bit shared-space
// addresses used through ld.shared / st.shared (no generic-address arithmetic inside the loops); on
// the host (tests/emul) they are plain pointers.
// ------------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
typedef unsigned SA;
__device__ __forceinline__ SA sa_of(const double* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ SA sa_add(SA a, int n) { return a + 8u * (unsigned)n; }
__device__ __forceinline__ double sa_ld(SA a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sa_ld2(SA a, double& x, double& y) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ void sa_st(SA a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
#else
typedef double* SA;
inline SA sa_of(const double* p) { return const_cast<double*>(p); }
inline SA sa_add(SA a, int n) { return a + n; }
inline double sa_ld(SA a) { return *a; }
inline void sa_ld2(SA a, double& x, double& y) { x = a[0]; y = a[1]; }
inline void sa_st(SA a, double v) { *a = v; }
#endif

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
    // ============================================================================================
    // sweep: initial rollout, in place.  On entry u(:,k) holds the warm-start c_k (or 0) and
    // x(:,k+1) holds the dynamics offset d_k (or 0); x(:,0) = dx0.
    //   u_k = Kinit x_k + c_k (transitionNominal.m:12) ; x_{k+1} = A x_k + B u_k + d_k (nominalModel.m:28)
    // ============================================================================================
    static LB_HD void rollout(const P& p, const L& l, double* s) {
        double x[NX], u[NU], xn[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) x[j] = s[l.i_x(j, 0)];
        for (int k = 0; k < p.N; ++k) {
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                double v = s[l.i_u(i, k)];
#pragma unroll
                for (int j = 0; j < NX; ++j) v += p.Kinit[i * NX + j] * x[j];
                u[i] = v;
                s[l.i_u(i, k)] = v;
            }
#pragma unroll
            for (int a = 0; a < NX; ++a) {
                double v = s[l.i_x(a, k + 1)];
#pragma unroll
                for (int j = 0; j < NX; ++j) v += p.A[a * NX + j] * x[j];
#pragma unroll
                for (int i = 0; i < NU; ++i) v += p.B[a * NU + i] * u[i];
                xn[a] = v;
                s[l.i_x(a, k + 1)] = v;
            }
#pragma unroll
            for (int j = 0; j < NX; ++j) x[j] = xn[j];
        }
    }

    // ============================================================================================
    // Row phases.  The box rows of a stage are processed WITHOUT branches: rows that do not exist at
    // a stage (stage_rows mask) are replaced by the neutral pair s = 1, lambda = 0 through selects,
    // so the ten rows of a stage form one basic block and their dependency chains (reciprocal,
    // directions, ratios) interleave instead of running one after the other.
    // ============================================================================================
    struct StageRows {  // iterate of one stage in registers
        double v[NVB], S[2 * NVB], Lm[2 * NVB];
    };
    static LB_HD void load_rows(const L& l, const double* s, int k, StageRows& r) {
        const double* r1 = s + l.r1(k);
#pragma unroll
        for (int j = 0; j < NVB; ++j) r.v[j] = r1[j];
#pragma unroll
        for (int q = 0; q < 2 * NVB; ++q) {
            r.S[q] = r1[L::F_S + q];
            r.Lm[q] = r1[L::F_LB + q];
        }
    }
    static LB_HD double gen_slack(const P& p, const L& l, const double* s, const double* G,
                                  const double* hg, int i) {
        double v = hg[i];
#pragma unroll
        for (int a = 0; a < NX; ++a) v -= G[a * p.ngp + i] * s[l.i_x(a, p.kg)];
#pragma unroll
        for (int a = 0; a < NT; ++a) v -= G[(NX + a) * p.ngp + i] * s[l.o_misc + L::M_TH + a];
        return v;
    }
    static LB_HD void init_rows_gen(const P& p, const L& l, double* s, const double* G,
                                    const double* hg, int i) {
        const double slack = gen_slack(p, l, s, G, hg, i);
        l.sgp(s)[i] = slack > 1.0 ? slack : 1.0;
        l.lgp(s)[i] = 1.0;
    }

    // ============================================================================================
    // stage: predictor assembly from the register copy r of stage k.  Writes Qd (barrier diagonal),
    // g (cost gradient + G'lambda), q (Newton rhs on the bounded variables) and the Farkas input
    // (G'lambda of the box rows, g-layout, first NV fields of record R3); accumulates reductions.
    // ============================================================================================
    static LB_HD void asm_core(const P& p, const L& l, double* s, int k, unsigned rows, const StageRows& r, RedAsm& red,
                              CShift csh = CShift{nullptr, 0}) {
        double v[NV], g[NV], e[NV];
        const bool last = k >= p.N;
#pragma unroll
        for (int j = 0; j < NV; ++j) e[j] = 0.0;
        if (csh.p) {  // cost evaluated at [x + ex_k; theta; u + eu_k] (twin state sequences)
#pragma unroll
            for (int j = 0; j < NX; ++j) e[j] = csh.p[k * csh.stride + j];
            if (csh.stride > NX && !last) {
#pragma unroll
                for (int j = 0; j < NU; ++j) e[NZ + j] = csh.p[k * csh.stride + NX + j];
            }
        }
#pragma unroll
        for (int j = 0; j < NX; ++j) v[j] = r.v[j];
#pragma unroll
        for (int j = 0; j < NT; ++j) v[NX + j] = s[l.o_misc + L::M_TH + j];
#pragma unroll
        for (int j = 0; j < NU; ++j) v[NZ + j] = last ? 0.0 : r.v[NX + j];
        const double* W = p.W[stage_type(p, k)];
#pragma unroll
        for (int a = 0; a < NV; ++a) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int b = 0; b < NV; ++b) {
                const double vb = v[b] + e[b];
                if (b & 1) a1 += W[a * NV + b] * vb;
                else a0 += W[a * NV + b] * vb;
            }
            g[a] = (last && a >= NZ) ? 0.0 : a0 + a1;
        }
        if (k == p.kT) {
#pragma unroll
            for (int a = 0; a < NZ; ++a) g[a] += s[l.o_misc + L::M_LIN + a];
        }
        double* r2 = s + l.r2(k);
        double* r3 = s + l.r3(k);
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
            double qd = 0.0, gl = 0.0, gp = 0.0;
            const int a = zidx(j);
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const int q = 2 * j + side;
                const bool act = (rows >> q) & 1u;
                const double S = act ? r.S[q] : 1.0, Lm = act ? r.Lm[q] : 0.0;
                const double sgn = side == 0 ? 1.0 : -1.0;
                const double slack = side == 0 ? p.hi[j] - v[a] : v[a] - p.lo[j];
                const double rp = act ? S - slack : 0.0;
                const double w = Lm * lb_rcp(S);
                qd += w;
                gl += sgn * Lm;
                gp += sgn * (w * rp);
                red.rp = lb_max(red.rp, lb_abs(rp));
                red.sl += S * Lm;
                red.lam = lb_max(red.lam, Lm);
                red.hl += Lm * slack;
            }
            r2[L::F_QD + j] = qd;
            r2[L::F_Q + j] = g[a] + gp;
            g[a] += gl;
            r3[a] = gl;
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) r3[NX + t] = 0.0;
#pragma unroll
        for (int a = 0; a < NV; ++a) r2[L::F_G + a] = g[a];
        red.gth += g[NX];
    }
    // fresh QP: initial slacks and multipliers s = max(h - a v, 1), lambda = 1, then the assembly
    static LB_HD void init_assemble_stage(const P& p, const L& l, double* s, int k, RedAsm& red, CShift csh = CShift{nullptr, 0}) {
        const unsigned rows = stage_rows(p, k);
        StageRows r;
        double* r1 = s + l.r1(k);
#pragma unroll
        for (int j = 0; j < NVB; ++j) r.v[j] = r1[j];
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const int q = 2 * j + side;
                const double slack = side == 0 ? p.hi[j] - r.v[j] : r.v[j] - p.lo[j];
                r.S[q] = slack > 1.0 ? slack : 1.0;
                r.Lm[q] = 1.0;
                if ((rows >> q) & 1u) {
                    r1[L::F_S + q] = r.S[q];
                    r1[L::F_LB + q] = 1.0;
                }
            }
        }
        asm_core(p, l, s, k, rows, r, red, csh);
    }
    // running QP: apply the step parked by final_stage (ds, dl in the scratch block; dx, du), then the assembly
    static LB_HD void update_assemble_stage(const P& p, const L& l, double* s, int k, double alpha, RedAsm& red,
                                           CShift csh = CShift{nullptr, 0}) {
        const unsigned rows = stage_rows(p, k);
        StageRows r;
        load_rows(l, s, k, r);
        double* r1 = s + l.r1(k);
        const double* r2 = s + l.r2(k);
        const double* r3 = s + l.r3(k);
#pragma unroll
        for (int q = 0; q < 2 * NVB; ++q) {
            r.S[q] += alpha * r2[q];
            r.Lm[q] += alpha * r2[2 * NVB + q];
            if ((rows >> q) & 1u) {
                r1[L::F_S + q] = r.S[q];
                r1[L::F_LB + q] = r.Lm[q];
            }
        }
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
            if (j < NX || k < p.N) {
                r.v[j] += alpha * r3[j];
                r1[j] = r.v[j];
            }
        }
        asm_core(p, l, s, k, rows, r, red, csh);
    }

    // ============================================================================================
    // Row phases over ITEMS (CTA-per-QP kernel: one thread per item).  Item (k, j), j < NVB: the
    // bounded variable j of [x;u] at stage k with its two box rows; theta_item: the theta rows of
    // the cost gradient of stage k.  Rows that do not exist at a stage (stage_rows mask) are replaced
    // by the neutral pair s = 1, lambda = 0 through selects.  Thread-indexed constants come from a
    // RowTab in shared memory (the constant bank serialises divergent indices).
    // ============================================================================================
    struct RowTab {
        double W[kMaxTypes][NV * NV];
        double lo[NVB], hi[NVB];
    };
    static LB_HD void fill_rowtab(const P& p, RowTab& T, int tid, int nthreads) {
        for (int i = tid; i < kMaxTypes * NV * NV; i += nthreads) T.W[i / (NV * NV)][i % (NV * NV)] = p.W[i / (NV * NV)][i % (NV * NV)];
        for (int i = tid; i < NVB; i += nthreads) {
            T.lo[i] = p.lo[i];
            T.hi[i] = p.hi[i];
        }
    }
    // fresh QP: s = max(h - a v, 1), lambda = 1
    static LB_HD void init_item(const P& p, const RowTab& T, const L& l, double* s, int k, int j) {
        const unsigned rows = stage_rows(p, k);
        double* r1 = s + l.r1(k);
        const double v = r1[j];
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const int q = 2 * j + side;
            const double slack = side == 0 ? T.hi[j] - v : v - T.lo[j];
            if ((rows >> q) & 1u) {
                r1[L::F_S + q] = slack > 1.0 ? slack : 1.0;
                r1[L::F_LB + q] = 1.0;
            }
        }
    }
    // running QP: apply the step parked by fin_item (ds, dl in the scratch block of R2; dx, du in R3)
    static LB_HD void upd_item(const P& p, const L& l, double* s, int k, int j, double alpha) {
        const unsigned rows = stage_rows(p, k);
        double* r1 = s + l.r1(k);
        const double* r2 = s + l.r2(k);
        const double* r3 = s + l.r3(k);
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const int q = 2 * j + side;
            if ((rows >> q) & 1u) {
                r1[L::F_S + q] += alpha * r2[q];
                r1[L::F_LB + q] += alpha * r2[2 * NVB + q];
            }
        }
        if (j < NX || k < p.N) r1[j] += alpha * r3[j];
    }
    // cost gradient row a of stage k: W_type(k)[a,:] v_k (+ lin at kT)
    static LB_HD double grad_row(const P& p, const RowTab& T, const L& l, const double* s, int k, int a, CShift csh = CShift{nullptr, 0}) {
        const bool last = k >= p.N;
        const double* r1 = s + l.r1(k);
        const double* W = T.W[stage_type(p, k)] + a * NV;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int b = 0; b < NV; ++b) {
            const double vb = b < NX ? r1[b] + (csh.p ? csh.p[k * csh.stride + (b < NX ? b : 0)] : 0.0)
                                     : (b < NZ ? s[l.o_misc + L::M_TH + (b - NX)]
                                               : (last ? 0.0 : r1[NX + (b - NZ)] + ((csh.p && csh.stride > NX) ? csh.p[k * csh.stride + NX + (b >= NZ ? b - NZ : 0)] : 0.0)));
            if (b & 1) a1 += W[b] * vb;
            else a0 += W[b] * vb;
        }
        double g = (last && a >= NZ) ? 0.0 : a0 + a1;
        if (k == p.kT && a < NZ) g += s[l.o_misc + L::M_LIN + a];
        return g;
    }
    // theta rows of stage k: gradient only
    static LB_HD void asm_theta_item(const P& p, const RowTab& T, const L& l, double* s, int k, RedAsm& red, CShift csh = CShift{nullptr, 0}) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const double g = grad_row(p, T, l, s, k, NX + t, csh);
            s[l.r2(k) + L::F_G + NX + t] = g;
            s[l.r3(k) + NX + t] = 0.0;
            if (t == 0) red.gth += g;
        }
    }
    // predictor assembly of item (k, j): cost gradient row, barrier diagonal, Newton rhs, Farkas input, reductions
    static LB_HD void asm_item(const P& p, const RowTab& T, const L& l, double* s, int k, int j, RedAsm& red, CShift csh = CShift{nullptr, 0}) {
        const bool last = k >= p.N;
        const double* r1 = s + l.r1(k);
        double* r2 = s + l.r2(k);
        double* r3 = s + l.r3(k);
        const int a = zidx(j);
        const double g = grad_row(p, T, l, s, k, a, csh);
        const unsigned rows = stage_rows(p, k);
        const double vj = (last && j >= NX) ? 0.0 : r1[j];
        double qd = 0.0, gl = 0.0, gp = 0.0;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const int q = 2 * j + side;
            const bool act = (rows >> q) & 1u;
            const double S = act ? r1[L::F_S + q] : 1.0, Lm = act ? r1[L::F_LB + q] : 0.0;
            const double slack = side == 0 ? T.hi[j] - vj : vj - T.lo[j];
            const double rp = act ? S - slack : 0.0;
            const double w = Lm * lb_rcp(S);
            qd += w;
            if (side == 0) {
                gl += Lm;
                gp += w * rp;
            } else {
                gl -= Lm;
                gp -= w * rp;
            }
            red.rp = lb_max(red.rp, lb_abs(rp));
            red.sl += S * Lm;
            red.lam = lb_max(red.lam, Lm);
            red.hl += Lm * slack;
        }
        r2[L::F_QD + j] = qd;
        r2[L::F_Q + j] = g + gp;
        r2[L::F_G + a] = g + gl;
        r3[a] = gl;
    }
    // affine (predictor) row directions of item (k, j): step-length ratio, the three sums of mu_aff, the
    // sigma-independent part of the corrector rhs (to q) and the coefficient of sigma*mu (to the dx/du scratch)
    static LB_HD void aff_item(const P& p, const RowTab& T, const L& l, double* s, int k, int j, RedStep& red) {
        const unsigned rows = stage_rows(p, k);
        const double* r1 = s + l.r1(k);
        double* r2 = s + l.r2(k);
        double* r3 = s + l.r3(k);
        double aj = 0.0, bj = 0.0;
        const double v = r1[j], dva = r3[NVB + j];
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const int q = 2 * j + side;
            const bool act = (rows >> q) & 1u;
            const double S = act ? r1[L::F_S + q] : 1.0, Lm = act ? r1[L::F_LB + q] : 0.0;
            const double slack = side == 0 ? T.hi[j] - v : v - T.lo[j];
            const double rp = act ? S - slack : 0.0, is = lb_rcp(S), w = Lm * is;
            const double sdva = side == 0 ? dva : -dva;
            const double dsa = act ? -rp - sdva : 0.0, dla = -Lm - w * dsa;
            const double rr = dsa * is;
            red.ratio = lb_max(red.ratio, act ? lb_max(-rr, 1.0 + rr) : 0.0);
            red.s0 += S * Lm;
            red.s1 += S * dla + Lm * dsa;
            red.s2 += dsa * dla;
            const double t = w * rp - dsa * dla * is - Lm;
            if (side == 0) {
                aj += t;
                bj += act ? is : 0.0;
            } else {
                aj -= t;
                bj -= act ? is : 0.0;
            }
        }
        if (j < NX || k < p.N) {
            r2[L::F_Q + j] = r2[L::F_G + zidx(j)] + aj;
            r3[j] = bj;
        }
    }
    static LB_HD void corr_item(const P& p, const L& l, double* s, int k, int j, double sigmu) {
        if (j >= NX && k >= p.N) return;
        s[l.i_q(j, k)] += sigmu * s[l.i_dv(j, k, false)];
    }
    // final (corrector) row directions of item (k, j): parks ds, dl in the scratch block at the start of record
    // R2, returns the step-length ratio
    static LB_HD double fin_item(const P& p, const RowTab& T, const L& l, double* s, int k, int j, double sigmu) {
        const unsigned rows = stage_rows(p, k);
        const double* r1 = s + l.r1(k);
        double* r2 = s + l.r2(k);
        const double* r3 = s + l.r3(k);
        double ratio = 0.0;
        const double v = r1[j], dva = r3[NVB + j], dv = r3[j];
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const int q = 2 * j + side;
            const bool act = (rows >> q) & 1u;
            const double S = act ? r1[L::F_S + q] : 1.0, Lm = act ? r1[L::F_LB + q] : 1.0;
            const double slack = side == 0 ? T.hi[j] - v : v - T.lo[j];
            const double rp = S - slack, is = lb_rcp(S), w = Lm * is;
            const double sdva = side == 0 ? dva : -dva, sdv = side == 0 ? dv : -dv;
            const double dsa = -rp - sdva, dla = -Lm - w * dsa;
            const double ds = -rp - sdv;
            const double rc = S * Lm + dsa * dla - sigmu;
            const double dl = (-rc - Lm * ds) * is;
            ratio = lb_max(ratio, act ? lb_max(-ds * is, -dl * lb_rcp(Lm)) : 0.0);
            r2[q] = act ? ds : 0.0;
            r2[2 * NVB + q] = act ? dl : 0.0;
        }
        return ratio;
    }

    // row: predictor assembly of polytope row i.  acc = {HG (NH packed), gGl (NZ), dG (NZ)}
    static LB_HD void assemble_gen_row(const P& p, const L& l, const double* s, const double* G,
                                       const double* hg, int i, double* acc, RedAsm& red) {
        const double slack = gen_slack(p, l, s, G, hg, i);
        const double S = l.sgp(s)[i], Lm = l.lgp(s)[i];
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
                Pxx[a][a] += s[l.i_qd(a, N)];
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
            for (int i = 0; i < NU; ++i) Rt[i][i] += s[l.i_qd(NX + i, k)];
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
                Nxx[a][a] += s[l.i_qd(a, k)];
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
                for (int b = 0; b < NZ; ++b) s[l.i_L(i * NZ + b, k)] = RL[i][b];
#pragma unroll
                for (int j = 0; j < NU; ++j) s[l.i_Ri(i * NU + j, k)] = Ri[i][j];
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
    // The four thread-local recursions below are latency-critical (one lane per QP, in-order issue,
    // FP64 issue-bound).  Each walks the per-stage records with ONE pointer, fetches every stage's
    // operands one stage ahead (two register sets, loop unrolled by two: no copies) so the 29-cycle
    // LDS latency never sits on the recursion's dependency chain, and keeps A, B in registers.
    // ============================================================================================
    struct AB {
        double A[NX * NX], B[NX * NU];
    };
    static LB_HD void load_ab(const P& p, AB& c) {
#pragma unroll
        for (int i = 0; i < NX * NX; ++i) c.A[i] = p.A[i];
#pragma unroll
        for (int i = 0; i < NX * NU; ++i) c.B[i] = p.B[i];
    }

    // ---- adjoint recursions.  One instruction stream serves two uses, selected per lane:
    //   farkas = false: input g (cost gradient + G'lambda, record R2) -> |r_d|inf at M_RD, the reduced
    //                   gradient of the Lagrangian w.r.t. (u, theta);
    //   farkas = true : input G'lambda only (parked by assemble_stage in the first NV fields of record
    //                   R3, free until the affine forward sweep) -> |G_red'lambda|inf at M_CERT and
    //                   h_red'lambda = lambda'slack + y'(G'lambda)_red accumulated into M_HLAM. ----
    struct AdjOps {
        double g[NV], u[NU];
    };
    static LB_HD void adj_load(const double* src, const double* r1, AdjOps& o) {
#pragma unroll
        for (int a = 0; a < NV; ++a) o.g[a] = src[a];
#pragma unroll
        for (int i = 0; i < NU; ++i) o.u[i] = r1[L::F_U + i];
    }
    // wu: weights of the certificate norm (Farkas use: nrm accumulates sum |v| wu) or nullptr (dual residual: max |v|)
    static LB_HD void adj_step(const AB& c, const AdjOps& o, double* pi, double& nrm, double& ydot, const double* wu) {
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            double v = o.g[NZ + i];
#pragma unroll
            for (int cc = 0; cc < NX; ++cc) v += c.B[cc * NU + i] * pi[cc];
            nrm = wu ? nrm + lb_abs(v) * wu[i] : lb_nanmax(nrm, lb_abs(v));
            ydot += v * o.u[i];
        }
        double np_[NX];
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            double v = o.g[a];
#pragma unroll
            for (int cc = 0; cc < NX; ++cc) v += c.A[cc * NX + a] * pi[cc];
            np_[a] = v;
        }
#pragma unroll
        for (int a = 0; a < NX; ++a) pi[a] = np_[a];
#pragma unroll
        for (int t = 0; t < NT; ++t) pi[NX + t] += o.g[NX + t];
    }
    static LB_HD void farkas_weights(const P& p, int k, double* wu) {
        const bool in = k >= p.ku0 && k <= p.ku1;
#pragma unroll
        for (int i = 0; i < NU; ++i) wu[i] = in ? p.fk_u[i] : p.fk_free;
    }
    static LB_HD void adjoint_sweep(const P& p, const L& l, double* s, bool farkas) {
        double pi[NZ];
        double* m = s + l.o_misc;
        const int N = p.N;
        AB c;
        load_ab(p, c);
        double nrm = 0.0, ydot = 0.0, wu[NU];
        const int step = farkas ? L::RS3 : L::RS2;
        const double* src = s + (farkas ? l.r3(N) : l.r2(N) + L::F_G);
        const double* r1 = s + l.r1(N - 1);
#pragma unroll
        for (int a = 0; a < NZ; ++a) pi[a] = src[a] + (p.kg == N ? m[L::M_GGL + a] : 0.0);
        src -= step;
        AdjOps o0, o1;
        adj_load(src, r1, o0);
        int k = N - 1;
        for (; k >= 1; k -= 2) {
            adj_load(src - step, r1 - L::RS1, o1);
            farkas_weights(p, k, wu);
            adj_step(c, o0, pi, nrm, ydot, farkas ? wu : nullptr);
            if (p.kg == k) {
#pragma unroll
                for (int a = 0; a < NZ; ++a) pi[a] += m[L::M_GGL + a];
            }
            src -= 2 * step;
            r1 -= 2 * L::RS1;
            adj_load(k >= 2 ? src : src + step, k >= 2 ? r1 : r1 + L::RS1, o0);
            farkas_weights(p, k - 1, wu);
            adj_step(c, o1, pi, nrm, ydot, farkas ? wu : nullptr);
            if (p.kg == k - 1) {
#pragma unroll
                for (int a = 0; a < NZ; ++a) pi[a] += m[L::M_GGL + a];
            }
        }
        if (k == 0) {
            farkas_weights(p, 0, wu);
            adj_step(c, o0, pi, nrm, ydot, farkas ? wu : nullptr);
            if (p.kg == 0) {
#pragma unroll
                for (int a = 0; a < NZ; ++a) pi[a] += m[L::M_GGL + a];
            }
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            nrm = farkas ? nrm + lb_abs(pi[NX + t]) * p.fk_th[t] : lb_nanmax(nrm, lb_abs(pi[NX + t]));
            ydot += pi[NX + t] * m[L::M_TH + t];
        }
        if (farkas) {
            m[L::M_CERT] = nrm;
            m[L::M_HLAM] += ydot;
        } else {
            m[L::M_RD] = nrm;
        }
    }

    // ---- the same adjoint recursion BLOCKED over the horizon (Farkas use; the transition matrix is the constant
    //      Abar, so the block transfer matrix is A^bm from Params): P1 lanes = blocks, zero incoming costate ->
    //      local result; P2 one lane chains the blocks; P3 lanes = blocks, true incoming costate -> per-lane
    //      pieces of sum_j |(G_red'lambda)_j| ybar_j and y'(G'lambda)_red, combined by the caller with warp reductions.
    //      The vector slots of the block scratch hold the local results / incoming costates. ----
    static LB_HD void adj_block(const P& p, const L& l, const double* s, int b, double* pi, double& nrm, double& ydot) {
        const double* m = s + l.o_misc;
        AB c;
        load_ab(p, c);
        const int lo = b * l.bm, hi = (lo + l.bm < p.N) ? lo + l.bm : p.N;
        const double* src = s + l.r3(hi - 1);
        const double* r1 = s + l.r1(hi - 1);
        for (int k = hi - 1; k >= lo; --k) {
            AdjOps o;
            double wu[NU];
            adj_load(src, r1, o);
            farkas_weights(p, k, wu);
            adj_step(c, o, pi, nrm, ydot, wu);
            if (p.kg == k) {
#pragma unroll
                for (int a = 0; a < NZ; ++a) pi[a] += m[L::M_GGL + a];
            }
            src -= L::RS3;
            r1 -= L::RS1;
        }
    }
    static LB_HD void farkas_p1(const P& p, const L& l, double* s, int b) {
        const double* m = s + l.o_misc;
        double pi[NZ], nrm = 0.0, ydot = 0.0;
#pragma unroll
        for (int a = 0; a < NZ; ++a) pi[a] = 0.0;
        if (b == l.nb - 1) {
            const double* src = s + l.r3(p.N);
#pragma unroll
            for (int a = 0; a < NZ; ++a) pi[a] = src[a] + (p.kg == p.N ? m[L::M_GGL + a] : 0.0);
        }
        adj_block(p, l, s, b, pi, nrm, ydot);
        double* v = s + l.blk(b) + L::TS;
#pragma unroll
        for (int a = 0; a < NZ; ++a) v[a] = pi[a];
    }
    // one lane: costate entering block b from above = Abar'^bm (costate entering b+1) + local(b+1); left in the
    // vector slot of block b+1 (block nb-1 starts from the terminal costate: its slot is not read)
    static LB_HD void farkas_p2(const P& p, const L& l, double* s) {
        double out[NZ];
        {
            const double* v = s + l.blk(l.nb - 1) + L::TS;
#pragma unroll
            for (int a = 0; a < NZ; ++a) out[a] = v[a];
        }
        for (int b = l.nb - 2; b >= 0; --b) {
            double* v = s + l.blk(b) + L::TS;      // local result of block b (zero incoming costate)
            double* vin = s + l.blk(b + 1) + L::TS;  // becomes the costate entering block b
            double nv[NZ];
#pragma unroll
            for (int a = 0; a < NX; ++a) {
                double acc = v[a];
#pragma unroll
                for (int cc = 0; cc < NX; ++cc) acc += p.Apow[cc * NX + a] * out[cc];
                nv[a] = acc;
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) nv[NX + t] = v[NX + t] + out[NX + t];
#pragma unroll
            for (int a = 0; a < NZ; ++a) {
                vin[a] = out[a];
                out[a] = nv[a];
            }
        }
        double* v0 = s + l.blk(0) + L::TS;  // costate at stage 0 (theta part = (G_red'lambda)_theta)
#pragma unroll
        for (int a = 0; a < NZ; ++a) v0[a] = out[a];
    }
    // incoming costate of block b (read by every block lane BEFORE lane 0's slot is reused)
    static LB_HD void farkas_p3(const P& p, const L& l, const double* s, int b, double& nrm, double& ydot) {
        const double* m = s + l.o_misc;
        double pi[NZ];
        if (b == l.nb - 1) {
            const double* src = s + l.r3(p.N);
#pragma unroll
            for (int a = 0; a < NZ; ++a) pi[a] = src[a] + (p.kg == p.N ? m[L::M_GGL + a] : 0.0);
        } else {
            const double* v = s + l.blk(b + 1) + L::TS;
#pragma unroll
            for (int a = 0; a < NZ; ++a) pi[a] = v[a];
        }
        adj_block(p, l, s, b, pi, nrm, ydot);
        if (b == 0) {
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                nrm += lb_abs(pi[NX + t]) * p.fk_th[t];
                ydot += pi[NX + t] * m[L::M_TH + t];
            }
        }
    }

    // ---- backward / forward substitution with the stored factors, BLOCKED over the horizon ----
    //   backward (gradient recursion): rt = q_u + B'pv ; kap_k = Ri rt ; pv <- q_z + Abar'pv - RL' rt
    //   forward                      : du_k = -kap_k - RL_k dz_k ; dx_{k+1} = A dx_k + B du_k
    // Both recursions are affine in their state, so the horizon is cut into nb blocks of bm stages and
    // one LANE per block runs them concurrently (one instruction stream for the whole warp instead of
    // one per stage chain):
    //   P1  every block from a zero state -> local result r_b; once per factorisation also the NX unit
    //       states through the homogeneous recursion -> block transfer matrix T_b (lanes = block x vector;
    //       the forward recursion uses the transposed x-block of the same matrices)
    //   P2  one lane chains the blocks: state_in(b) = T_b state_in(b+1) + r_b   (nb small mat-vecs)
    //   P3  every block again from its true incoming state, storing kap (backward) / du, dx (forward).
    // The homogeneous runs read their affine inputs from an all-zero record (pointer step 0).
    struct BwdOps {  // shared-memory operands of one stage
        double q[NVB], gt[NT], RL[NU * NZ], Ri[NU * NU];
    };
    static LB_HD void bwd_load(const double* r2, const double* qrec, BwdOps& o) {
#pragma unroll
        for (int j = 0; j < NVB; ++j) o.q[j] = qrec[L::F_Q + j];
#pragma unroll
        for (int t = 0; t < NT; ++t) o.gt[t] = qrec[L::F_G + NX + t];
#pragma unroll
        for (int j = 0; j < NU * NZ; ++j) o.RL[j] = r2[L::F_RL + j];
#pragma unroll
        for (int j = 0; j < NU * NU; ++j) o.Ri[j] = r2[L::F_RI + j];
    }
    static LB_HD void bwd_step(const AB& c, const BwdOps& o, double* pv, double* r2, bool store) {
        double rt[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            double v0 = o.q[NX + i], v1 = 0.0;
#pragma unroll
            for (int cc = 0; cc < NX; ++cc) {
                if (cc & 1) v1 += c.B[cc * NU + i] * pv[cc];
                else v0 += c.B[cc * NU + i] * pv[cc];
            }
            rt[i] = v0 + v1;
        }
        double np_[NZ];
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            double v = o.q[a];
#pragma unroll
            for (int cc = 0; cc < NX; ++cc) v += c.A[cc * NX + a] * pv[cc];
#pragma unroll
            for (int i = 0; i < NU; ++i) v -= o.RL[i * NZ + a] * rt[i];
            np_[a] = v;
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            double v = o.gt[t] + pv[NX + t];
#pragma unroll
            for (int i = 0; i < NU; ++i) v -= o.RL[i * NZ + NX + t] * rt[i];
            np_[NX + t] = v;
        }
        if (store) {
#pragma unroll
            for (int i = 0; i < NU; ++i) {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < NU; ++j) v += o.Ri[i * NU + j] * rt[j];
                r2[L::F_KAP + i] = v;  // kappa_n = Ri rt (the forward sweep subtracts it)
            }
        }
#pragma unroll
        for (int a = 0; a < NZ; ++a) pv[a] = np_[a];
    }
    static LB_HD void bwd_kg(const double* m, double* pv) {
#pragma unroll
        for (int a = 0; a < NZ; ++a) pv[a] += m[L::M_GGL + a] + m[L::M_DG + a];
    }
    // costate at stage N (start of the recursion)
    static LB_HD void bwd_terminal(const P& p, const L& l, const double* s, double* pv) {
        const double* m = s + l.o_misc;
        const int N = p.N;
#pragma unroll
        for (int a = 0; a < NZ; ++a) {
            const double base = a < NX ? s[l.i_q(a, N)] : s[l.i_g(a, N)];
            pv[a] = base + (p.kg == N ? m[L::M_GGL + a] + m[L::M_DG + a] : 0.0);
        }
    }
    // stages hi-1 .. lo from the costate pv (in/out).  qrec: record with q / g_theta of stage hi-1, stepping by
    // qstep per stage (the slot's R2, or an all-zero record with qstep = 0 for the homogeneous recursion);
    // kgl: stage after which the polytope terms are added (-1: never); store: write kap_k
    static LB_HD void bwd_range(const AB& c, const L& l, double* s, int lo, int hi, double* pv, const double* qrec,
                                int qstep, int kgl, bool store) {
        const double* m = s + l.o_misc;
        double* r2 = s + l.r2(hi - 1);
        BwdOps o0, o1;
        bwd_load(r2, qrec, o0);
        int k = hi - 1;
        for (; k >= lo + 1; k -= 2) {
            bwd_load(r2 - L::RS2, qrec - qstep, o1);
            bwd_step(c, o0, pv, r2, store);
            if (kgl == k) bwd_kg(m, pv);
            r2 -= 2 * L::RS2;
            qrec -= 2 * qstep;
            const bool more = k >= lo + 2;
            bwd_load(more ? r2 : r2 + L::RS2, more ? qrec : qrec + qstep, o0);
            bwd_step(c, o1, pv, r2 + L::RS2, store);
            if (kgl == k - 1) bwd_kg(m, pv);
        }
        if (k == lo) {
            bwd_step(c, o0, pv, r2, store);
            if (kgl == lo) bwd_kg(m, pv);
        }
    }
    // P1 task t of the backward solve (with_T: tasks = block x (1 + NX) vectors, else one task per block)
    static LB_HD void bwd_p1(const P& p, const L& l, double* s, const double* zero_rec, int t, bool with_T) {
        bwd_p1_task(p, l, s, zero_rec, with_T ? t / (NX + 1) : t, with_T ? t % (NX + 1) : 0);
    }
    // homogeneous tasks only (block transfer matrices): t = block x unit vector
    static LB_HD void bwd_p1_T(const P& p, const L& l, double* s, const double* zero_rec, int t) {
        bwd_p1_task(p, l, s, zero_rec, t / NX, 1 + t % NX);
    }
    // vec = 0: particular solution of block b (local result); vec = j+1: unit costate e_j through the homogeneous recursion
    static LB_HD void bwd_p1_task(const P& p, const L& l, double* s, const double* zero_rec, int b, int vec) {
        AB c;
        load_ab(p, c);
        const int lo = b * l.bm, hi = (lo + l.bm < p.N) ? lo + l.bm : p.N;
        double pv[NZ];
#pragma unroll
        for (int a = 0; a < NZ; ++a) pv[a] = (vec > 0 && a == vec - 1) ? 1.0 : 0.0;
        if (vec == 0 && b == l.nb - 1) bwd_terminal(p, l, s, pv);
        const bool part = vec == 0;
        bwd_range(c, l, s, lo, hi, pv, part ? s + l.r2(hi - 1) : zero_rec, part ? L::RS2 : 0, part ? p.kg : -1, false);
        double* dst = s + l.blk(b) + (part ? L::TS : (vec - 1) * NZ);
#pragma unroll
        for (int a = 0; a < NZ; ++a) dst[a] = pv[a];
    }
    // P2 of the backward solve (one lane): chain the blocks top-down, leave the incoming costate of block b in the
    // vector slot of block b+1, finish with d(theta) = -Ptt^-1 pv_theta(0) -> M_DTHA (aff) or M_DTH
    static LB_HD void bwd_p2(const L& l, double* s, bool aff) {
        double* m = s + l.o_misc;
        double out[NZ];
        {
            const double* v = s + l.blk(l.nb - 1) + L::TS;
#pragma unroll
            for (int a = 0; a < NZ; ++a) out[a] = v[a];
        }
        for (int b = l.nb - 2; b >= 0; --b) {
            double* vin = s + l.blk(b + 1) + L::TS;  // consumed: now holds the incoming costate of block b
            const double* T = s + l.blk(b);
            double nv[NZ];
#pragma unroll
            for (int a = 0; a < NZ; ++a) nv[a] = T[L::TS + a] + (a >= NX ? out[a] : 0.0);
#pragma unroll
            for (int j = 0; j < NX; ++j)
#pragma unroll
                for (int a = 0; a < NZ; ++a) nv[a] += T[j * NZ + a] * out[j];
#pragma unroll
            for (int a = 0; a < NZ; ++a) {
                vin[a] = out[a];
                out[a] = nv[a];
            }
        }
#pragma unroll
        for (int a = 0; a < NT; ++a) {
            double v = 0.0;
#pragma unroll
            for (int bb = 0; bb < NT; ++bb) v -= m[L::M_PTT + a * NT + bb] * out[NX + bb];
            m[(aff ? L::M_DTHA : L::M_DTH) + a] = v;
        }
    }
    // incoming costate of block b for P3 (read by every block lane BEFORE any lane overwrites the slots)
    static LB_HD void bwd_p3_in(const P& p, const L& l, const double* s, int b, double* pv) {
        if (b == l.nb - 1) {
            bwd_terminal(p, l, s, pv);
        } else {
            const double* v = s + l.blk(b + 1) + L::TS;
#pragma unroll
            for (int a = 0; a < NZ; ++a) pv[a] = v[a];
        }
    }

    struct FwdOps {
        double RL[NU * NZ], kap[NU];
    };
    static LB_HD void fwd_load(const double* r2, FwdOps& o) {
#pragma unroll
        for (int j = 0; j < NU * NZ; ++j) o.RL[j] = r2[L::F_RL + j];
#pragma unroll
        for (int i = 0; i < NU; ++i) o.kap[i] = r2[L::F_KAP + i];
    }
    // r3: record of stage k (du_k is written there, dx_{k+1} to the next record); off: 0 or NVB (affine)
    static LB_HD void fwd_step(const AB& c, const FwdOps& o, const double* dth, double* dx, double* r3, int off, bool store) {
        double du[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            double v0 = -o.kap[i], v1 = 0.0;
#pragma unroll
            for (int t = 0; t < NT; ++t) v0 -= o.RL[i * NZ + NX + t] * dth[t];
#pragma unroll
            for (int cc = 0; cc < NX; ++cc) {
                if (cc & 1) v1 -= o.RL[i * NZ + cc] * dx[cc];
                else v0 -= o.RL[i * NZ + cc] * dx[cc];
            }
            du[i] = v0 + v1;
        }
        double xn[NX];
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            double v = 0.0;
#pragma unroll
            for (int cc = 0; cc < NX; ++cc) v += c.A[a * NX + cc] * dx[cc];
#pragma unroll
            for (int i = 0; i < NU; ++i) v += c.B[a * NU + i] * du[i];
            xn[a] = v;
        }
        if (store) {
#pragma unroll
            for (int i = 0; i < NU; ++i) r3[off + L::F_DU + i] = du[i];
        }
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            dx[a] = xn[a];
            if (store) r3[L::RS3 + off + L::F_DX + a] = xn[a];
        }
    }
    // stages lo .. hi-1 from the state dx (in/out)
    static LB_HD void fwd_range(const AB& c, const L& l, double* s, int lo, int hi, double* dx, const double* dth,
                                int off, bool store) {
        double* r3 = s + l.r3(lo);
        const double* r2 = s + l.r2(lo);
        FwdOps o0, o1;
        fwd_load(r2, o0);
        int k = lo;
        for (; k + 1 < hi; k += 2) {
            fwd_load(r2 + L::RS2, o1);
            fwd_step(c, o0, dth, dx, r3, off, store);
            r2 += 2 * L::RS2;
            fwd_load(k + 2 < hi ? r2 : r2 - L::RS2, o0);
            fwd_step(c, o1, dth, dx, r3 + L::RS3, off, store);
            r3 += 2 * L::RS3;
        }
        if (k < hi) fwd_step(c, o0, dth, dx, r3, off, store);
    }
    // backward P3 (true incoming costate pv, stores kap) fused with forward P1 (zero state -> local result in
    // the vector slot of block b) for block b
    static LB_HD void bwd_p3_fwd_p1(const P& p, const L& l, double* s, int b, double* pv, bool aff) {
        AB c;
        load_ab(p, c);
        const int lo = b * l.bm, hi = (lo + l.bm < p.N) ? lo + l.bm : p.N;
        bwd_range(c, l, s, lo, hi, pv, s + l.r2(hi - 1), L::RS2, p.kg, true);
        fwd_p1_ab(c, p, l, s, b, aff);
    }
    // forward P1 for block b: zero state -> local result in the vector slot of block b
    static LB_HD void fwd_p1(const P& p, const L& l, double* s, int b, bool aff) {
        AB c;
        load_ab(p, c);
        fwd_p1_ab(c, p, l, s, b, aff);
    }
    static LB_HD void fwd_p1_ab(const AB& c, const P& p, const L& l, double* s, int b, bool aff) {
        const int lo = b * l.bm, hi = (lo + l.bm < p.N) ? lo + l.bm : p.N;
        const double* m = s + l.o_misc;
        double dx[NX], dth[NT];
#pragma unroll
        for (int j = 0; j < NX; ++j) dx[j] = 0.0;
#pragma unroll
        for (int t = 0; t < NT; ++t) dth[t] = m[(aff ? L::M_DTHA : L::M_DTH) + t];
        fwd_range(c, l, s, lo, hi, dx, dth, aff ? NVB : 0, false);
        double* v = s + l.blk(b) + L::TS;
#pragma unroll
        for (int j = 0; j < NX; ++j) v[j] = dx[j];
    }
    // forward P2 (one lane): dx_in(b+1) = Tx_b' dx_in(b) + local(b), left in the vector slot of block b
    static LB_HD void fwd_p2(const L& l, double* s) {
        double din[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) din[j] = 0.0;
        for (int b = 0; b + 1 < l.nb; ++b) {
            double* v = s + l.blk(b) + L::TS;
            const double* T = s + l.blk(b);
            double nv[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                double a0 = v[i];
#pragma unroll
                for (int j = 0; j < NX; ++j) a0 += T[i * NZ + j] * din[j];
                nv[i] = a0;
            }
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                din[i] = nv[i];
                v[i] = nv[i];
            }
        }
    }
    // forward P3 for block b: from the true incoming state, storing du_k, dx_{k+1}
    static LB_HD void fwd_p3(const P& p, const L& l, double* s, int b, bool aff) {
        AB c;
        load_ab(p, c);
        const int lo = b * l.bm, hi = (lo + l.bm < p.N) ? lo + l.bm : p.N;
        const double* m = s + l.o_misc;
        const int off = aff ? NVB : 0;
        double dx[NX], dth[NT];
#pragma unroll
        for (int j = 0; j < NX; ++j) dx[j] = b > 0 ? s[l.blk(b - 1) + L::TS + j] : 0.0;
#pragma unroll
        for (int t = 0; t < NT; ++t) dth[t] = m[(aff ? L::M_DTHA : L::M_DTH) + t];
        if (b == 0) {
#pragma unroll
            for (int j = 0; j < NX; ++j) s[l.r3(0) + off + L::F_DX + j] = 0.0;
        }
        fwd_range(c, l, s, lo, hi, dx, dth, off, true);
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
        const double* r1 = s + l.r1(k);
        double* r2 = s + l.r2(k);
        double* r3 = s + l.r3(k);
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
            double aj = 0.0, bj = 0.0;
            const double v = r1[j], dva = r3[NVB + j];
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const int q = 2 * j + side;
                const bool act = (rows >> q) & 1u;
                const double S = act ? r1[L::F_S + q] : 1.0, Lm = act ? r1[L::F_LB + q] : 0.0;
                const double sgn = side == 0 ? 1.0 : -1.0;
                const double slack = side == 0 ? p.hi[j] - v : v - p.lo[j];
                const double rp = act ? S - slack : 0.0, is = lb_rcp(S), w = Lm * is;
                const double dsa = act ? -rp - sgn * dva : 0.0, dla = -Lm - w * dsa;
                const double rr = dsa * is;
                red.ratio = lb_max(red.ratio, act ? lb_max(-rr, 1.0 + rr) : 0.0);
                red.s0 += S * Lm;
                red.s1 += S * dla + Lm * dsa;
                red.s2 += dsa * dla;
                aj += sgn * (w * rp - dsa * dla * is - Lm);
                bj += act ? sgn * is : 0.0;
            }
            if (j < NX || k < p.N) {
                r2[L::F_Q + j] = r2[L::F_G + zidx(j)] + aj;
                r3[j] = bj;
            }
        }
    }
    static LB_HD void corr_stage(const P& p, const L& l, double* s, int k, double sigmu) {
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
            if (j >= NX && k >= p.N) continue;
            s[l.i_q(j, k)] += sigmu * s[l.i_dv(j, k, false)];
        }
    }
    // ---- polytope rows, CTA-per-QP kernel: a thread per row writes the row's scalars to a row buffer
    //      (rb: 3 x ngp doubles), then a thread per OUTPUT entry sums over the rows (no cross-thread reduction) ----
    static LB_HD void gen_row_asm_scalars(const P& p, const L& l, const double* s, const double* G, const double* hg, int i,
                                          double* rb, RedAsm& red) {
        const double slack = gen_slack(p, l, s, G, hg, i);
        const double S = l.sgp(s)[i], Lm = l.lgp(s)[i];
        const double rp = S - slack, w = Lm * lb_rcp(S), t = w * rp;
        red.rp = lb_nanmax(red.rp, lb_abs(rp));
        red.sl += S * Lm;
        red.lam = lb_max(red.lam, Lm);
        red.hl += Lm * slack;
        rb[i] = w;
        rb[p.ngp + i] = Lm;
        rb[2 * p.ngp + i] = t - Lm;
    }
    // output e of the predictor assembly: e < NH packed Hessian entry, then G'lambda (NZ), then the rhs part dG (NZ)
    static LB_HD double gen_output_asm(const P& p, const double* G, const double* rb, int e) {
        int a = 0, b = 0, c = 0;
        if (e < NH) {
            int idx = 0;
#pragma unroll
            for (int aa = 0; aa < NZ; ++aa)
#pragma unroll
                for (int bb = aa; bb < NZ; ++bb) {
                    if (idx == e) {
                        a = aa;
                        b = bb;
                    }
                    ++idx;
                }
        } else {
            c = e < NH + NZ ? 1 : 2;
            a = e - NH - (c - 1) * NZ;
            b = -1;
        }
        const double* ga = G + a * p.ngp;
        const double* gb = G + (b < 0 ? 0 : b) * p.ngp;
        const double* co = rb + c * p.ngp;
        double a0 = 0.0, a1 = 0.0;
        for (int i = 0; i + 1 < p.ng; i += 2) {
            a0 += co[i] * ga[i] * (b < 0 ? 1.0 : gb[i]);
            a1 += co[i + 1] * ga[i + 1] * (b < 0 ? 1.0 : gb[i + 1]);
        }
        if (p.ng & 1) a0 += co[p.ng - 1] * ga[p.ng - 1] * (b < 0 ? 1.0 : gb[p.ng - 1]);
        return a0 + a1;
    }
    static LB_HD void gen_row_aff_scalars(const P& p, const L& l, const double* s, const double* G, const double* hg, int i,
                                          double* rb, RedStep& red) {
        const double* m = s + l.o_misc;
        const double slack = gen_slack(p, l, s, G, hg, i);
        const double S = l.sgp(s)[i], Lm = l.lgp(s)[i];
        const double rp = S - slack, is = lb_rcp(S), w = Lm * is;
        double adva = 0.0;
#pragma unroll
        for (int a = 0; a < NX; ++a) adva += G[a * p.ngp + i] * s[l.i_dv(a, p.kg, true)];
#pragma unroll
        for (int t = 0; t < NT; ++t) adva += G[(NX + t) * p.ngp + i] * m[L::M_DTHA + t];
        const double dsa = -rp - adva, dla = -Lm - w * dsa, rr = dsa * is;
        red.ratio = lb_max(red.ratio, lb_max(-rr, 1.0 + rr));
        red.s0 += S * Lm;
        red.s1 += S * dla + Lm * dsa;
        red.s2 += dsa * dla;
        rb[i] = w * rp - dsa * dla * is - Lm;
        rb[p.ngp + i] = is;
    }
    // corrector rhs part of component a: sum_i G[a][i] (t1_i + sigmu / s_i)
    static LB_HD double gen_output_aff(const P& p, const double* G, const double* rb, int a, double sigmu) {
        const double* ga = G + a * p.ngp;
        double a0 = 0.0, a1 = 0.0;
        for (int i = 0; i < p.ng; ++i) {
            a0 += ga[i] * rb[i];
            a1 += ga[i] * rb[p.ngp + i];
        }
        return a0 + sigmu * a1;
    }

    // polytope row i: acc[0..NZ) += G (t1 - lambda), acc[NZ..2NZ) += G / s   (dG = acc1 + sigmu acc2)
    static LB_HD void affine_gen_row(const P& p, const L& l, const double* s, const double* G,
                                     const double* hg, int i, double* acc, RedStep& red) {
        const double* m = s + l.o_misc;
        const double slack = gen_slack(p, l, s, G, hg, i);
        const double S = l.sgp(s)[i], Lm = l.lgp(s)[i];
        const double rp = S - slack, is = lb_rcp(S), w = Lm * is;
        double g[NZ], adva = 0.0;
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            g[a] = G[a * p.ngp + i];
            adva += g[a] * s[l.i_dv(a, p.kg, true)];
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
    // stage: final (corrector) row directions of stage k -> parks ds, dl in the scratch block at the
    // start of record R2, returns the step-length ratio.  update_assemble_stage applies the step.
    // ============================================================================================
    static LB_HD double final_stage(const P& p, const L& l, double* s, int k, double sigmu) {
        const unsigned rows = stage_rows(p, k);
        const double* r1 = s + l.r1(k);
        double* r2 = s + l.r2(k);
        const double* r3 = s + l.r3(k);
        double ratio = 0.0;
#pragma unroll
        for (int j = 0; j < NVB; ++j) {
            const double v = r1[j], dva = r3[NVB + j], dv = r3[j];
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const int q = 2 * j + side;
                const bool act = (rows >> q) & 1u;
                const double S = act ? r1[L::F_S + q] : 1.0, Lm = act ? r1[L::F_LB + q] : 1.0;
                const double sgn = side == 0 ? 1.0 : -1.0;
                const double slack = side == 0 ? p.hi[j] - v : v - p.lo[j];
                const double rp = S - slack, is = lb_rcp(S), w = Lm * is;
                const double dsa = -rp - sgn * dva, dla = -Lm - w * dsa;
                const double ds = -rp - sgn * dv;
                const double rc = S * Lm + dsa * dla - sigmu;
                const double dl = (-rc - Lm * ds) * is;
                ratio = lb_max(ratio, act ? lb_max(-ds * is, -dl * lb_rcp(Lm)) : 0.0);
                r2[q] = act ? ds : 0.0;
                r2[2 * NVB + q] = act ? dl : 0.0;
            }
        }
        return ratio;
    }
    // polytope row i, final direction (recomputed by final_gen_row and update_gen_row: no scratch)
    static LB_HD void gen_final_dir(const P& p, const L& l, const double* s, const double* G, const double* hg,
                                    int i, double sigmu, double& S, double& Lm, double& ds, double& dl, double& is) {
        const double* m = s + l.o_misc;
        const double slack = gen_slack(p, l, s, G, hg, i);
        S = l.sgp(s)[i];
        Lm = l.lgp(s)[i];
        const double rp = S - slack;
        is = lb_rcp(S);
        const double w = Lm * is;
        double adva = 0.0, adv = 0.0;
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            const double g = G[a * p.ngp + i];
            adva += g * s[l.i_dv(a, p.kg, true)];
            adv += g * s[l.i_dv(a, p.kg, false)];
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
        l.sgp(s)[i] = S + alpha * ds;
        l.lgp(s)[i] = Lm + alpha * dl;
    }

    // stage: objective contribution 0.5 v'W v (+ lin'z at kT)
    static LB_HD double objective_stage(const P& p, const L& l, const double* s, int k, CShift csh = CShift{nullptr, 0}) {
        double v[NV];
#pragma unroll
        for (int j = 0; j < NX; ++j) v[j] = s[l.i_x(j, k)] + (csh.p ? csh.p[k * csh.stride + j] : 0.0);
#pragma unroll
        for (int j = 0; j < NT; ++j) v[NX + j] = s[l.o_misc + L::M_TH + j];
        const bool last = k >= p.N;
#pragma unroll
        for (int j = 0; j < NU; ++j)
            v[NZ + j] = last ? 0.0 : s[l.i_u(j, k)] + ((csh.p && csh.stride > NX) ? csh.p[k * csh.stride + NX + j] : 0.0);
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
        if (cert && m[L::M_HLAM] < 0.0 && m[L::M_CERT] * p.inf_scale <= -m[L::M_HLAM]) return 2;
        return -1;
    }
};

// ------------------------------------------------------------------------------------------------
// Cooperative Riccati factorisation for NT = NU = 1 (the Moore-Greitzer shape): ONE WARP per QP,
// one dot product per lane per stage.  Three recursions ride on the same instruction stream:
//   * the factorisation itself (cost-to-go matrix P over z = [x;theta]),
//   * the adjoint recursion that gives the dual residual |r_d| (input g = cost gradient + G'lambda),
//   * the backward substitution of the AFFINE (predictor) Newton system (input q),
// so that after the sweep only forward substitutions are left for the predictor.
//
// The stage time is set by the dependency chain publish -> read -> dot -> shuffle -> update
// (~130 cycles with the latencies measured in tools/microbench/lat.cu); everything else — the
// reciprocal of the pivot, the stored factors, the stage inputs — is kept off that chain.
//
// Lane roles (NX = 4: NH = 15 unique entries of the symmetric NZ x NZ matrix):
//   0..NH-1      entry (a,b), a <= b, kept SCALED by the pivot of the stage it came from: Pt = rho P
//   kFz..+NZ-1   F_uz[c] = (Abar' Pt Bbar)[c]          kFu   F_uu = Bbar' Pt Bbar
//   kPi..+NX-1   adjoint pi[a] (unscaled)              kRd   running max |g_u + B'pi|
//   kPv..+NZ-1   affine costate pv[a], scaled: pvt = rho pv        kRt   B'pvt
// One stage:  F = Abar_e' Pt Abar_e with Abar_e = [Abar Bbar]: every entry is a fixed linear form in
// the unique entries of Pt, i.e. a dot product with a per-lane coefficient vector over NLD values
// read from a lane-type dependent base of the exchange buffer.  With ir = 1/rho:
//     Rt  = Wuu + Qd_u + F_uu ir ,  L = Wuz + F_uz ir
//     Pt' = Rt (Wzz + Qd + [HG] + F_zz ir) - L'L ,  rho' = Rt            (P' = Wzz+Qd+Abar'P Abar - L'L/Rt)
//     pvt'= Rt (q + [GGL+DG] + (Abar'pvt) ir) - L' rtn ,  rtn = q_u + (B'pvt) ir ,  kappa = -rtn/Rt
//     pi' = g + [GGL] + Abar'pi
// The reciprocal of the pivot is needed only one stage later (and for the stored factors RL = L/Rt,
// Ri = 1/Rt, kappa): its hardware seed is issued as soon as Rt exists and its Newton refinement is
// interleaved with the next stage's dot products, so no division sits on the stage-to-stage chain.
// The theta component of the adjoint needs no recursion (Abar is the identity there): it is the
// plain sum of g_theta over the stages, reduced in the assembly phase (M_GTH).
//
// Exchange 1 (shared memory, double buffered: one __syncwarp per stage): publishing lanes store
// their value, every lane reads NLD consecutive doubles from its base.  Layout of one buffer:
// [xx block | (a,theta) | (theta,theta) | pad | pi | pad | pvt | pad], pads are zero and only ever
// multiplied by zero coefficients.  Exchange 2 (warp shuffles): F_uz[a], F_uz[b] / B'pvt, F_uu.
// The lane phases st1/st2/st3 are separated by __syncwarp()/shuffles in the kernel; tests/emul runs
// them as loops over the 32 lanes.
// ------------------------------------------------------------------------------------------------
template <int NX>
struct Coop {
    static constexpr int NT = 1, NU = 1, NZ = NX + 1, NV = NZ + 1, NH = NZ * (NZ + 1) / 2, NXX = NX * (NX + 1) / 2;
    static constexpr int NLD = (NXX + 1) & ~1;            // values each lane reads (even: 16-byte loads)
    static constexpr int kFz = NH, kFu = NH + NZ, kPi = kFu + 1, kRd = kPi + NX, kPv = kRd + 1, kRt = kPv + NZ;
    // exchange buffer (doubles); every read base is even
    static constexpr int oXT = (NXX + 1) & ~1, oTT = (oXT + NX + 1) & ~1, oPi = (oTT + NLD + 1) & ~1,
                         oPv = (oPi + NLD + 1) & ~1, oDummy = (oPv + NLD + 1) & ~1, kBuf = oDummy + 8;
    static constexpr int kXch = 2 * kBuf;                 // two buffers, each with 8 never-read slots for the lanes that publish nothing
    static_assert(kRt < 32 && NX >= 2 && NZ <= NLD, "shape does not fit the one-warp mapping");
    using P = Params<NX, 1, 1>;
    using L = Layout<NX, 1, 1>;
    using C = Core<NX, 1, 1>;

    // unique entries ordered: xx block row-wise, then (a,theta), then (theta,theta)
    static LB_HD constexpr int ent(int a, int b) {  // a <= b
        return b < NX ? a * NX - a * (a - 1) / 2 + (b - a) : (a < NX ? NXX + a : NH - 1);
    }
    struct Lane {
        // constants of the lane
        double c[NLD];                 // linear form over xch[base .. base+NLD)
        int a, b, in_off, in_step, out_off, out_step, srcA, srcB, pubi, basei;
        bool isP, isPi, isPv, isRd, isRt, plike, stDef, stRi;
        SA pub[2], base[2];            // publish / read addresses in the two exchange buffers
        // constants of the cost segment / of this factorisation
        double wzz, wa, wb, wuu, hg;
        // recursion state
        double val;                    // Pt entry | pi | pvt
        double rt, y0;                 // pivot of the matrix being published and its reciprocal seed
        double dmul;                   // deferred factor store: L_a (diagonal lanes), rtn (kRt)
        // stage temporaries handed from phase to phase
        double qdu, wbk, czz, d1, ir;
        SA in_ptr;                     // stage input of this lane (walks the R2 records; a zero double otherwise)
        SA out_ptr;                    // store target: deferred factors (R2, one stage behind) / |r_d| of the stage (kRd, R3)
    };
    static LB_HD double rcp_seed(double x) {
#ifdef __CUDA_ARCH__
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        return y;
#else
        return 1.0 / x;
#endif
    }
    static LB_HD double rcp_refine(double x, double y) {
        double e = fma(-x, y, 1.0);
        y = fma(y, e, y);
        e = fma(-x, y, 1.0);
        return fma(y, e, y);
    }
    // keep a lane constant in a register: without this the compiler re-derives it from threadIdx inside the
    // stage loop (dozens of integer instructions per stage)
    template <typename T>
    static LB_HD void pin(T& v) {
#ifdef __CUDA_ARCH__
        if constexpr (sizeof(T) == 8) asm volatile("" : "+d"(*reinterpret_cast<double*>(&v)));
        else asm volatile("" : "+r"(*reinterpret_cast<int*>(&v)));
#else
        (void)v;
#endif
    }
    static LB_HD void lane_init(const P& p, int h, double* xch, Lane& ln) {
        ln.isP = h < NH;
        ln.isPi = h >= kPi && h < kPi + NX;
        ln.isRd = h == kRd;
        ln.isPv = h >= kPv && h < kPv + NZ;
        ln.isRt = h == kRt;
        ln.plike = ln.isP || ln.isPv;
        ln.a = 0;
        ln.b = 0;
#pragma unroll
        for (int a = 0; a < NZ; ++a)
#pragma unroll
            for (int b = a; b < NZ; ++b)
                if (ent(a, b) == h) {
                    ln.a = a;
                    ln.b = b;
                }
        if (ln.isPi) ln.a = ln.b = h - kPi;
        if (ln.isPv) ln.a = ln.b = h - kPv;
        const bool xx = h < NXX, xt = (h >= NXX && h < NH - 1), tt = (h == NH - 1);
        const int fz = h - kFz;  // F_uz component for lanes kFz..kFz+NZ-1
        const bool fzx = (fz >= 0 && fz < NX), fzt = (fz == NX), fu = (h == kFu);
        const int base = (xt || fzt) ? oXT : (tt ? oTT : ((ln.isPi || ln.isRd) ? oPi : ((ln.isPv || ln.isRt) ? oPv : 0)));
        // where the lane publishes its value (a never-read slot if it has nothing to publish)
        const int pubi = xx ? h : (xt ? oXT + ln.a : (tt ? oTT : (ln.isPi ? oPi + ln.a : (ln.isPv ? oPv + ln.a
                              : oDummy + (fz >= 0 && fz < NZ ? fz : (fu ? 5 : (ln.isRd ? 6 : 7)))))));
        ln.pubi = pubi;
        ln.basei = base;
        const SA x0 = sa_of(xch);
        ln.pub[0] = sa_add(x0, pubi);
        ln.pub[1] = sa_add(x0, kBuf + pubi);
        ln.base[0] = sa_add(x0, base);
        ln.base[1] = sa_add(x0, kBuf + base);
        // stage input: Qd_a (diagonal x entries), g_a (adjoint), g_u (r_d lane), q_a / g_theta (affine costate), q_u (kRt)
        const bool dgx = ln.isP && ln.a == ln.b && ln.a < NX;
        const bool has_in = dgx || ln.isPi || ln.isRd || ln.isPv || ln.isRt;
        ln.in_off = dgx ? L::F_QD + ln.a
                        : (ln.isPi ? L::F_G + ln.a
                                   : (ln.isRd ? L::F_G + NZ
                                              : (ln.isPv ? (ln.a < NX ? L::F_Q + ln.a : L::F_G + NX) : (ln.isRt ? L::F_Q + NX : 0))));
        ln.in_step = has_in ? L::RS2 : 0;
        // stores: RL[a] from the diagonal lanes, kappa_n = rtn/Rt from kRt, Ri from kFu (all one stage behind, record R2);
        // |g_u + B'pi| of the stage from kRd (spare field of record R3)
        ln.stDef = (ln.isP && ln.a == ln.b) || ln.isRt;
        ln.stRi = fu;
        ln.out_off = ln.isRt ? L::F_KAP : (fu ? L::F_RI : L::F_RL + ln.a);
        ln.out_step = ln.isRd ? L::RS3 : L::RS2;
        // shuffle sources: F_uz[a] and F_uz[b]; the affine costate lanes take B'pvt through the b channel
        ln.srcA = kFz + ln.a;
        ln.srcB = (ln.isPv || ln.isRt) ? kRt : kFz + ln.b;
#pragma unroll
        for (int j = 0; j < NLD; ++j) ln.c[j] = 0.0;
#pragma unroll
        for (int c = 0; c < NX; ++c)
#pragma unroll
            for (int d = c; d < NX; ++d) {
                double v = 0.0;
                if (xx) {  // F_zz[a][b], a,b < NX
                    v = p.A[c * NX + ln.a] * p.A[d * NX + ln.b];
                    if (c != d) v += p.A[d * NX + ln.a] * p.A[c * NX + ln.b];
                } else if (fzx) {  // F_uz[fz]
                    v = p.A[c * NX + fz] * p.B[d];
                    if (c != d) v += p.A[d * NX + fz] * p.B[c];
                } else if (fu) {  // F_uu
                    v = p.B[c] * p.B[d] * (c != d ? 2.0 : 1.0);
                }
                if (xx || fzx || fu) ln.c[ent(c, d)] = v;
            }
#pragma unroll
        for (int c = 0; c < NX; ++c) {
            if (xt) ln.c[c] = p.A[c * NX + ln.a];                                   // F_zz[a][theta]
            if (fzt) ln.c[c] = p.B[c];                                              // F_uz[theta]
            if ((ln.isPi || ln.isPv) && ln.a < NX) ln.c[c] = p.A[c * NX + ln.a];    // (Abar'pi)[a]
            if (ln.isRd || ln.isRt) ln.c[c] = p.B[c];                               // B'pi
        }
        if (tt) ln.c[0] = 1.0;                                                      // F_zz[theta][theta]
        if (ln.isPv && ln.a == NX) ln.c[NX] = 1.0;                                  // (Abar'pv)[theta]
#pragma unroll
        for (int j = 0; j < NLD; ++j) pin(ln.c[j]);
        pin(ln.in_off); pin(ln.in_step); pin(ln.out_off); pin(ln.out_step); pin(ln.srcA); pin(ln.srcB);
        pin(ln.pub[0]); pin(ln.pub[1]); pin(ln.base[0]); pin(ln.base[1]);
        ln.val = 0.0;
        ln.rt = ln.y0 = 1.0;
        ln.dmul = 0.0;
        ln.wzz = ln.wa = ln.wb = ln.wuu = ln.hg = 0.0;
        ln.qdu = ln.wbk = ln.czz = ln.d1 = 0.0;
        ln.ir = 1.0;
        ln.in_ptr = ln.out_ptr = x0;
    }
    // The lane constants are the same for every warp and every factorisation.  The CTA-per-QP kernel computes them
    // once per CTA into a shared-memory table (field-major: conflict-free) and reloads them at the start of each
    // factorisation, so that they occupy no registers during the other phases of the iteration.
    static constexpr int kTabInts = 11;
    struct LaneTab {
        double c[NLD][32];
        int iv[kTabInts][32];
    };
    static LB_HD void lane_store(const Lane& ln, int h, LaneTab& t) {
#pragma unroll
        for (int j = 0; j < NLD; ++j) t.c[j][h] = ln.c[j];
        const int flags = (ln.isP ? 1 : 0) | (ln.isPi ? 2 : 0) | (ln.isPv ? 4 : 0) | (ln.isRd ? 8 : 0) | (ln.isRt ? 16 : 0) |
                          (ln.stDef ? 32 : 0) | (ln.stRi ? 64 : 0);
        const int iv[kTabInts] = {ln.a, ln.b, ln.in_off, ln.in_step, ln.out_off, ln.out_step, ln.srcA, ln.srcB, ln.pubi, ln.basei, flags};
#pragma unroll
        for (int j = 0; j < kTabInts; ++j) t.iv[j][h] = iv[j];
    }
    static LB_HD void lane_load(const LaneTab& t, int h, double* xch, Lane& ln) {
#pragma unroll
        for (int j = 0; j < NLD; ++j) ln.c[j] = t.c[j][h];
        ln.a = t.iv[0][h]; ln.b = t.iv[1][h]; ln.in_off = t.iv[2][h]; ln.in_step = t.iv[3][h];
        ln.out_off = t.iv[4][h]; ln.out_step = t.iv[5][h]; ln.srcA = t.iv[6][h]; ln.srcB = t.iv[7][h];
        ln.pubi = t.iv[8][h]; ln.basei = t.iv[9][h];
        const int flags = t.iv[10][h];
        ln.isP = flags & 1; ln.isPi = flags & 2; ln.isPv = flags & 4; ln.isRd = flags & 8; ln.isRt = flags & 16;
        ln.stDef = flags & 32; ln.stRi = flags & 64;
        ln.plike = ln.isP || ln.isPv;
        const SA x0 = sa_of(xch);
        ln.pub[0] = sa_add(x0, ln.pubi);
        ln.pub[1] = sa_add(x0, kBuf + ln.pubi);
        ln.base[0] = sa_add(x0, ln.basei);
        ln.base[1] = sa_add(x0, kBuf + ln.basei);
        ln.val = 0.0;
        ln.rt = ln.y0 = 1.0;
        ln.dmul = 0.0;
        ln.wzz = ln.wa = ln.wb = ln.wuu = ln.hg = 0.0;
        ln.qdu = ln.wbk = ln.czz = ln.d1 = 0.0;
        ln.ir = 1.0;
        ln.in_ptr = ln.out_ptr = x0;
    }
    static LB_HD void load_type(const P& p, int t, Lane& ln) {
        const double* W = p.W[t];
        ln.wzz = ln.isP ? W[ln.a * NV + ln.b] : 0.0;
        ln.wa = ln.plike ? W[NZ * NV + ln.a] : 0.0;
        ln.wb = ln.isP ? W[NZ * NV + ln.b] : 0.0;
        ln.wuu = W[NZ * NV + NZ];
    }
    // zero the exchange buffers once (pads are read with zero coefficients)
    static LB_HD void xch_init(int h, double* xch) {
        for (int j = h; j < kXch; j += 32) xch[j] = 0.0;
    }
    // start of a factorisation: polytope terms of stage kg, pointers, terminal stage
    //   Pt = Wzz + Qd (+HG), rho = 1 ; pi = g (+GGL) ; pvt = q | g_theta (+GGL+DG)
    static LB_HD void begin(const P& p, const L& l, double* s, const double* zero, Lane& ln) {
        const int N = p.N;
        const double* m = s + l.o_misc;
        double hv = 0.0;
        if (ln.isP) hv = m[L::M_HG + C::sym(ln.a, ln.b)];
        else if (ln.isPi) hv = m[L::M_GGL + ln.a];
        else if (ln.isPv) hv = m[L::M_GGL + ln.a] + m[L::M_DG + ln.a];
        ln.hg = hv;
        load_type(p, C::stage_type(p, N), ln);
        ln.in_ptr = ln.in_step ? sa_of(s + l.r2(N) + ln.in_off) : sa_of(zero);
        // stage N has no factors: the first deferred store lands in its (unused) factor fields.  The r_d lane
        // stores in phase 3, after the pointer has moved to the stage being processed: start one record up.
        ln.out_ptr = ln.isRd ? sa_of(s + l.r3(N) + L::F_RDK) : sa_of(s + l.r2(N) + ln.out_off);
        ln.val = ln.isRt ? 0.0 : (ln.wzz + sa_ld(ln.in_ptr)) + (p.kg == N ? ln.hg : 0.0);
        ln.in_ptr = sa_add(ln.in_ptr, -ln.in_step);
        ln.rt = ln.y0 = 1.0;
        ln.dmul = 0.0;
    }
    // phase 1 of stage k: publish, fetch the stage inputs (rec = record R2 of stage k, B = exchange buffer k & 1)
    template <int B>
    static LB_HD void st1(Lane& ln, SA rec, bool atkg) {
        sa_st(ln.pub[B], ln.val);
        ln.qdu = sa_ld(sa_add(rec, L::F_QD + NX));
        const double qu = sa_ld(sa_add(rec, L::F_Q + NX));
        const double in = sa_ld(ln.in_ptr);
        ln.in_ptr = sa_add(ln.in_ptr, -ln.in_step);
        double czz = ln.wzz + in;
        if (atkg) czz += ln.hg;
        ln.czz = czz;
        double wbk = ln.wb;
        if (ln.isPv) wbk = qu;
        ln.wbk = wbk;
    }
    // phase 2: the dot product; refinement of the previous pivot's reciprocal and the factor stores that wait for it
    template <int B>
    static LB_HD void st2(Lane& ln) {
        double v[NLD];
#pragma unroll
        for (int j = 0; j < NLD / 2; ++j) sa_ld2(sa_add(ln.base[B], 2 * j), v[2 * j], v[2 * j + 1]);
        const double irn = rcp_refine(ln.rt, ln.y0);  // 1/rho of the matrix being read
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int j = 0; j < NLD; ++j) {
            if (j & 1) a1 += ln.c[j] * v[j];
            else a0 += ln.c[j] * v[j];
        }
        ln.d1 = a0 + a1;
        ln.ir = irn;
        if (ln.stDef) sa_st(ln.out_ptr, ln.dmul * irn);
        if (ln.stRi) sa_st(ln.out_ptr, irn);
        ln.out_ptr = sa_add(ln.out_ptr, -ln.out_step);
    }
    // phase 3.  fa, fb, fuu: d1 of lanes srcA, srcB, kFu
    static LB_HD void st3(Lane& ln, double fa, double fb, double fuu) {
        const double ir = ln.ir;
        const double Rt = (ln.wuu + ln.qdu) + fuu * ir;
        const double La = ln.wa + fa * ir;
        const double Lb = ln.wbk + fb * ir;
        const double hz = ln.czz + ln.d1 * ir;            // kRt: rtn = q_u + (B'pvt) ir
        const double anew = ln.czz + ln.d1;               // adjoint lanes: pi' = g + Abar'pi (+GGL at kg)
        double nv = anew;
        if (ln.plike) nv = Rt * hz - La * Lb;             // entry lanes, affine costate lanes
        ln.val = nv;
        if (ln.isRd) sa_st(ln.out_ptr, lb_abs(anew));     // |g_u + B'pi| of this stage
        ln.rt = Rt;
        ln.y0 = rcp_seed(Rt);
        double dm = La;
        if (ln.isRt) dm = hz;
        ln.dmul = dm;
    }
    // after stage 0: store its factors; returns val / rho_0 (lane NH-1: P_0[theta][theta]; lane kPv+NX: pv_theta(0))
    static LB_HD double finish(Lane& ln) {
        const double irn = rcp_refine(ln.rt, ln.y0);
        if (ln.stDef) sa_st(ln.out_ptr, ln.dmul * irn);
        if (ln.stRi) sa_st(ln.out_ptr, irn);
        return ln.val * irn;
    }
    // pivots and dual residual of the sweep, stage k (lanes stride over the stages afterwards)
    static LB_HD void check_stage(const L& l, const double* s, int k, bool& ok, double& rd) {
        const double ri = s[l.r2(k) + L::F_RI];
        ok = ok && (ri > 0.0) && (ri < 1e300);
        rd = lb_nanmax(rd, s[l.r3(k) + L::F_RDK]);
    }
};

}  // namespace lbmpc
