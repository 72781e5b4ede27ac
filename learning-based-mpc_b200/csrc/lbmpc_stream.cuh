// lbmpc_stream.cuh — the "stream" mapping of the batched Mehrotra/Riccati interior-point engine:
// ONE THREAD PER QP, the whole iterate of a QP in HBM, laid out structure-of-arrays over the 32 QPs
// of a warp (element e of lane l at base[e * 32 + l]) so that every load / store of a warp is one
// fully used 256-byte line.
//
// Why a second mapping next to the warp-per-QP kernel (lbmpc_kernels.cuh): with one warp per QP the
// iterate lives in shared memory, which caps an SM at 8 resident QPs, and most issued instructions
// are shuffles, shared-memory exchanges and partially filled FP64 operations (14.8 k warp
// instructions per QP-iteration, profiles/r1_*).  With one thread per QP every instruction is 32
// useful lanes, there is no inter-thread communication at all (dot products, duality-gap sums and
// step-length minima are thread-local), the horizon length is not bounded by shared memory, and the
// per-stage dynamics may differ per QP (LTV: oracle Jacobians, lbmpc_solve_sqp).  The price is that
// the iterate streams through HBM four times per iteration, so this mapping is HBM-bound by
// construction; everything below is organised to move as few bytes as possible:
//
//   pass BU (backward)  apply the previous step (recomputed from the stored directions, not stored
//                       row by row) + predictor assembly + Riccati factorisation + dual residual
//                       adjoint + Farkas adjoint + affine backward substitution, all in ONE sweep
//                       (reads iterate + both directions, writes iterate + factors)
//   pass F1 (forward)   affine forward substitution + affine row directions + sigma sums, and the
//                       sigma-independent part of the corrector right-hand side (so that pass B2
//                       never touches the iterate)
//   pass B2 (backward)  corrector backward substitution from the stored right-hand side pieces
//   pass F2 (forward)   corrector forward substitution + final row directions + step length
//
// FP64 record of a stage: x[NX] u s[2 NVB] lambda[2 NVB] (25 doubles at NX = 4); direction / factor
// record (storage type FT: double, or float for the mixed mode): dva[NVB] dv[NVB] RL[NZ] Ri kap
// qc[NV] t2[NVB] (28).  Traffic per stage and iteration (FT = double): BU 35 R + 32 W, F1 31 R +
// 21 W, B2 17 R + 1 W, F2 36 R + 5 W = 178 doubles = 1.4 KB; N = 50: 73 KB per QP-iteration.
//
// The lanes of a warp are PERSISTENT and independent: a lane whose QP reaches its verdict writes the
// result and takes the next QP from the global queue while its neighbours continue with theirs (a
// fresh QP's first pass BU initialises the rows instead of applying a step, selected per lane), so
// QPs with 6 and with 16 iterations share a warp without waiting for each other.
//
// The arithmetic is the algorithm of oracle/lbmpc_oracle.c (the definition of iteration counts and
// verdicts) and of lbmpc_core.cuh; reference statements: costLBMPC.m:25-45, constraintsLBMPC.m:18-45,
// DMS_tracking_LMPC_casadi.m:223-291, LBMPC_casadi.m:240-305 (via lbmpc_problem.hpp).  NT = NU = 1.
// Host compilation exists only for tests/emul (LS = 1).
#pragma once

#include "lbmpc_core.cuh"

namespace lbmpc {

template <int NX>
struct StreamLayout {
    static constexpr int NZ = NX + 1, NV = NX + 2, NVB = NX + 1, NH = NZ * (NZ + 1) / 2;
    static constexpr int F_X = 0, F_U = NX, F_S = NVB, F_LB = 3 * NVB, RS_IT = 5 * NVB;
    static constexpr int D_DA = 0, D_DV = NVB, D_RL = 2 * NVB, D_RI = D_RL + NZ, D_KAP = D_RI + 1, D_QC = D_KAP + 1,
                         D_T2 = D_QC + NV, RS_D = D_T2 + NVB;
    static constexpr int NJ = NX * 3;  // LTV: Jacobian of the learned term w.r.t. xi = [x1; x2; u] per stage
    int N, ng, o_it, o_cs, o_j, o_sg, o_lg, n64, nft;  // offsets / sizes in elements (per QP)
    LB_HD StreamLayout(int N_, int ng_, bool cs, bool ltv) {
        N = N_;
        ng = ng_;
        int o = 0;
        o_it = o; o += RS_IT * (N + 1);
        o_cs = o; o += cs ? NX * (N + 1) : 0;
        o_j = o;  o += ltv ? NJ * N : 0;
        o_sg = o; o += ng;
        o_lg = o; o += ng;
        n64 = o;
        nft = RS_D * (N + 1);
    }
};

// per-lane state that lives across the passes of one QP (registers)
template <int NX>
struct StreamLane {
    static constexpr int NZ = NX + 1;
    long long q;          // QP index, -1: idle
    int iters;
    bool fresh;           // no step to apply yet: the next pass BU initialises the rows
    double th, dtha, dth, alpha, sigmu, iptt, mu;
    double gGl[NZ], dG1[NZ], dG2[NZ], lin[NZ], cconst;
    // results of pass BU
    double rp, lam, hlam, rd, cert, obj;
    bool piv_ok;
};

template <typename FT>
struct StreamIO {  // batch-major caller arrays (include/lbmpc.h)
    long long batch;
    const double *dx0, *dx_ref, *d_off, *warm, *cshift, *jac;
    double *uc, *theta, *xtraj, *obj;
    int *iters, *status;
    unsigned long long* queue;
    double* ws64;  // workspace: per warp n64 * 32 doubles ...
    FT* wsft;      // ... and nft * 32 FT
};

template <int NX, bool LTV, typename FT, int LS>
struct Stream {
    using P = Params<NX, 1, 1>;
    using SL = StreamLayout<NX>;
    using Lane = StreamLane<NX>;
    using C = Core<NX, 1, 1>;
    static constexpr int NZ = NX + 1, NV = NX + 2, NVB = NX + 1, NH = SL::NH;

    static LB_HD double ld(const double* p, int e) { return p[e * LS]; }
    static LB_HD void st(double* p, int e, double v) { p[e * LS] = v; }
    static LB_HD double ldf(const FT* p, int e) { return (double)p[e * LS]; }
    static LB_HD void stf(FT* p, int e, double v) { p[e * LS] = (FT)v; }
    static LB_HD int sym(int a, int b) { return a <= b ? C::sym(a, b) : C::sym(b, a); }

    struct AB {  // dynamics of the stage being processed
        double A[NX * NX], B[NX];
    };
    // LTI: the constant bank; LTV: A_k = A + [J(:,1:2) 0 0], B_k = B + J(:,3), J read from the workspace
    static LB_HD void load_ab(const P& p, const SL& l, const double* w64, int k, AB& ab) {
#pragma unroll
        for (int c = 0; c < NX; ++c) {
#pragma unroll
            for (int b = 0; b < NX; ++b) {
                double v = p.A[c * NX + b];
                if (LTV && b < 2) v += ld(w64, l.o_j + k * SL::NJ + c * 3 + b);
                ab.A[c * NX + b] = v;
            }
            double v = p.B[c];
            if (LTV) v += ld(w64, l.o_j + k * SL::NJ + c * 3 + 2);
            ab.B[c] = v;
        }
    }

    // cost gradient g = W_type(k) [x + e; theta; u] (u part dropped at the last stage), objective piece 0.5 v'g (+ lin'z at kT)
    static LB_HD void cost_grad(const P& p, const SL& l, const Lane& ln, const double* w64, int k, const double* v,
                                bool has_cs, double* g, double* Jacc) {
        const bool last = k >= p.N;
        double vv[NV];
#pragma unroll
        for (int j = 0; j < NX; ++j) vv[j] = v[j] + (has_cs ? ld(w64, l.o_cs + k * NX + j) : 0.0);
        vv[NX] = ln.th;
        vv[NZ] = last ? 0.0 : v[NX];
        const double* W = p.W[C::stage_type(p, k)];
        double J = 0.0;
#pragma unroll
        for (int a = 0; a < NV; ++a) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int b = 0; b < NV; ++b) {
                if (b & 1) a1 += W[a * NV + b] * vv[b];
                else a0 += W[a * NV + b] * vv[b];
            }
            const double ga = (last && a >= NZ) ? 0.0 : a0 + a1;
            J += 0.5 * vv[a] * ga;
            g[a] = ga;
        }
        if (k == p.kT) {
#pragma unroll
            for (int a = 0; a < NZ; ++a) {
                g[a] += ln.lin[a];
                J += ln.lin[a] * vv[a];
            }
        }
        if (Jacc) *Jacc += J;
    }

    // ============================================================================================
    // a new QP: inputs -> workspace, initial rollout u_k = Kinit x_k + c_k (transitionNominal.m:12),
    // x_{k+1} = A_k x_k + B_k u_k + d_k (nominalModel.m:28)
    // ============================================================================================
    static LB_HD void init_qp(const P& p, const SL& l, Lane& ln, const StreamIO<FT>& io, long long q, double* w64, bool has_cs) {
        const int N = p.N;
        ln.q = q;
        ln.iters = 0;
        ln.fresh = true;
        ln.alpha = 0.0;
        ln.sigmu = 0.0;
        ln.dth = ln.dtha = 0.0;
        ln.th = io.warm ? io.warm[q * (N + 1) + N] : 0.0;
        double cconst = 0.0;
#pragma unroll
        for (int a = 0; a < NZ; ++a) {
            double v = 0.0;
            if (io.dx_ref) {
#pragma unroll
                for (int j = 0; j < NX; ++j) v += p.Lref[a * NX + j] * io.dx_ref[q * NX + j];
            }
            ln.lin[a] = v;
            ln.gGl[a] = ln.dG1[a] = ln.dG2[a] = 0.0;
        }
        if (io.dx_ref) {
#pragma unroll
            for (int i = 0; i < NX; ++i)
#pragma unroll
                for (int j = 0; j < NX; ++j) cconst += io.dx_ref[q * NX + i] * p.Tm[i * NX + j] * io.dx_ref[q * NX + j];
        }
        ln.cconst = cconst;
        if (LTV) {
            const double* jq = io.jac + q * (long long)(N * SL::NJ);
            for (int i = 0; i < N * SL::NJ; ++i) st(w64, l.o_j + i, jq[i]);
        }
        if (has_cs) {
            const double* cq = io.cshift + q * (long long)((N + 1) * NX);
            for (int i = 0; i < (N + 1) * NX; ++i) st(w64, l.o_cs + i, cq[i]);
        }
        double x[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) x[j] = io.dx0[q * NX + j];
        for (int k = 0; k <= N; ++k) {
            double* it = w64 + (l.o_it + k * SL::RS_IT) * LS;
#pragma unroll
            for (int j = 0; j < NX; ++j) st(it, SL::F_X + j, x[j]);
            if (k == N) break;
            AB ab;
            load_ab(p, l, w64, k, ab);
            double u = io.warm ? io.warm[q * (N + 1) + k] : 0.0;
#pragma unroll
            for (int j = 0; j < NX; ++j) u += p.Kinit[j] * x[j];
            st(it, SL::F_U, u);
            double xn[NX];
#pragma unroll
            for (int a = 0; a < NX; ++a) {
                double v = io.d_off ? io.d_off[(q * N + k) * NX + a] : 0.0;
#pragma unroll
                for (int j = 0; j < NX; ++j) v += ab.A[a * NX + j] * x[j];
                v += ab.B[a] * u;
                xn[a] = v;
            }
#pragma unroll
            for (int a = 0; a < NX; ++a) x[a] = xn[a];
        }
    }

    // ============================================================================================
    // pass BU.  Returns the verdict (LBMPC_ST_*) or -1 to continue.
    // ============================================================================================
    static LB_HD int pass_bu(const P& p, const SL& l, Lane& ln, const double* __restrict__ G, const double* __restrict__ hg,
                             double* w64, FT* wft, bool has_cs) {
        const int N = p.N;
        const bool fresh = ln.fresh;
        const double alpha = fresh ? 0.0 : ln.alpha, sigmu = ln.sigmu;
        const double th_old = ln.th;
        ln.th = th_old + alpha * ln.dth;
        const double th = ln.th;
        double Pm[NH], pv[NZ], pi[NZ], pc[NZ];
        double HG[NH], gGl[NZ], dG[NZ];
        double rpm = 0.0, sl = 0.0, lam = 0.0, hl = 0.0, rd = 0.0, ci = 0.0, yd = 0.0, J = 0.0;
        bool ok = true;
#pragma unroll
        for (int a = 0; a < NH; ++a) HG[a] = 0.0;
#pragma unroll
        for (int a = 0; a < NZ; ++a) gGl[a] = dG[a] = 0.0;
        for (int k = N; k >= 0; --k) {
            double* it = w64 + (l.o_it + k * SL::RS_IT) * LS;
            FT* d = wft + k * SL::RS_D * LS;
            const unsigned rows = C::stage_rows(p, k);
            const bool last = k >= N;
            // ---- apply the step / initialise the rows, then the predictor assembly at the new iterate ----
            double vo[NVB], vn[NVB], dva[NVB], dv[NVB];
#pragma unroll
            for (int j = 0; j < NVB; ++j) {
                const bool ex = j < NX || !last;
                vo[j] = ex ? ld(it, j) : 0.0;
                dva[j] = (ex && !fresh) ? ldf(d, SL::D_DA + j) : 0.0;
                dv[j] = (ex && !fresh) ? ldf(d, SL::D_DV + j) : 0.0;
                vn[j] = vo[j] + alpha * dv[j];
                if (ex) st(it, j, vn[j]);
            }
            double g[NV];
            cost_grad(p, l, ln, w64, k, vn, has_cs, g, &J);
            double qd[NVB], q[NVB], gl[NVB];
#pragma unroll
            for (int j = 0; j < NVB; ++j) {
                double qdj = 0.0, glj = 0.0, gpj = 0.0;
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    const int r = 2 * j + side;
                    const bool act = (rows >> r) & 1u;
                    const double sgn = side == 0 ? 1.0 : -1.0;
                    double S = 1.0, Lm = 0.0;
                    if (act) {
                        S = ld(it, SL::F_S + r);
                        Lm = ld(it, SL::F_LB + r);
                    }
                    const double slack_o = side == 0 ? p.hi[j] - vo[j] : vo[j] - p.lo[j];
                    const double rp_o = S - slack_o, is = lb_rcp(S), w = Lm * is;
                    const double dsa = -rp_o - sgn * dva[j], dla = -Lm - w * dsa;
                    const double ds = -rp_o - sgn * dv[j];
                    const double rc = S * Lm + dsa * dla - sigmu;
                    const double dl = (-rc - Lm * ds) * is;
                    double Sn = fresh ? (slack_o > 1.0 ? slack_o : 1.0) : S + alpha * ds;
                    double Ln = fresh ? 1.0 : Lm + alpha * dl;
                    Sn = act ? Sn : 1.0;
                    Ln = act ? Ln : 0.0;
                    if (act) {
                        st(it, SL::F_S + r, Sn);
                        st(it, SL::F_LB + r, Ln);
                    }
                    const double slack = side == 0 ? p.hi[j] - vn[j] : vn[j] - p.lo[j];
                    const double rp = act ? Sn - slack : 0.0;
                    const double wn = Ln * lb_rcp(Sn);
                    qdj += wn;
                    glj += sgn * Ln;
                    gpj += sgn * (wn * rp);
                    rpm = lb_max(rpm, lb_abs(rp));
                    sl += Sn * Ln;
                    lam = lb_max(lam, Ln);
                    hl += Ln * slack;
                }
                const int a = C::zidx(j);
                qd[j] = qdj;
                gl[j] = glj;
                q[j] = g[a] + gpj;  // Newton right-hand side on the bounded variables
                g[a] += glj;        // cost gradient + G'lambda (dual residual input)
            }
            // ---- polytope block ----
            if (k == p.kg) {
                double zo[NZ], zn[NZ], dza[NZ], dz[NZ];
#pragma unroll
                for (int a = 0; a < NX; ++a) {
                    zo[a] = vo[a]; zn[a] = vn[a]; dza[a] = dva[a]; dz[a] = dv[a];
                }
                zo[NX] = th_old; zn[NX] = th; dza[NX] = ln.dtha; dz[NX] = ln.dth;
                for (int i = 0; i < p.ng; ++i) {
                    double gi[NZ], so = hg[i], sn = hg[i], adva = 0.0, adv = 0.0;
#pragma unroll
                    for (int a = 0; a < NZ; ++a) {
                        gi[a] = G[a * p.ngp + i];
                        so -= gi[a] * zo[a];
                        sn -= gi[a] * zn[a];
                        adva += gi[a] * dza[a];
                        adv += gi[a] * dz[a];
                    }
                    const double S = ld(w64, l.o_sg + i), Lm = ld(w64, l.o_lg + i);
                    const double rp_o = S - so, is = lb_rcp(S), w = Lm * is;
                    const double dsa = -rp_o - adva, dla = -Lm - w * dsa;
                    const double ds = -rp_o - adv;
                    const double rc = S * Lm + dsa * dla - sigmu;
                    const double dl = (-rc - Lm * ds) * is;
                    const double Sn = fresh ? (so > 1.0 ? so : 1.0) : S + alpha * ds;
                    const double Ln = fresh ? 1.0 : Lm + alpha * dl;
                    st(w64, l.o_sg + i, Sn);
                    st(w64, l.o_lg + i, Ln);
                    const double rp = Sn - sn, wn = Ln * lb_rcp(Sn), t = wn * rp;
                    rpm = lb_nanmax(rpm, lb_abs(rp));
                    sl += Sn * Ln;
                    lam = lb_max(lam, Ln);
                    hl += Ln * sn;
                    int idx = 0;
#pragma unroll
                    for (int a = 0; a < NZ; ++a) {
                        const double wa = wn * gi[a];
#pragma unroll
                        for (int b = a; b < NZ; ++b) HG[idx++] += wa * gi[b];
                        gGl[a] += gi[a] * Ln;
                        dG[a] += gi[a] * (t - Ln);
                    }
                }
            }
            const bool atkg = k == p.kg;
            if (last) {
                // ---- terminal stage: P = Wzz + Qd (+HG); pv = q | g_theta; pi = g; pc = G'lambda ----
                const double* W = p.W[C::stage_type(p, N)];
#pragma unroll
                for (int a = 0; a < NZ; ++a) {
#pragma unroll
                    for (int b = a; b < NZ; ++b) {
                        double v = W[a * NV + b] + (atkg ? HG[C::sym(a, b)] : 0.0);
                        if (a == b && a < NX) v += qd[a];
                        Pm[C::sym(a, b)] = v;
                    }
                    const double kgp = atkg ? gGl[a] : 0.0;
                    pv[a] = (a < NX ? q[a] : g[a]) + kgp + (atkg ? dG[a] : 0.0);
                    pi[a] = g[a] + kgp;
                    pc[a] = (a < NX ? gl[a] : 0.0) + kgp;
                }
                continue;
            }
            // ---- Riccati step: P (stage k+1) -> P (stage k), factors RL = L / Rt, Ri = 1 / Rt ----
            AB ab;
            load_ab(p, l, w64, k, ab);
            const double* W = p.W[C::stage_type(p, k)];
            double M[NZ][NX], L[NZ], Rt;
#pragma unroll
            for (int a = 0; a < NZ; ++a)
#pragma unroll
                for (int b = 0; b < NX; ++b) {
                    double v = 0.0;
#pragma unroll
                    for (int c = 0; c < NX; ++c) v += Pm[sym(a, c)] * ab.A[c * NX + b];
                    M[a][b] = v;
                }
            {
                double acc = W[NZ * NV + NZ] + qd[NX];
#pragma unroll
                for (int c = 0; c < NX; ++c) {
                    double pb = 0.0;
#pragma unroll
                    for (int e = 0; e < NX; ++e) pb += Pm[sym(c, e)] * ab.B[e];
                    acc += ab.B[c] * pb;
                }
                Rt = acc;
            }
#pragma unroll
            for (int b = 0; b < NZ; ++b) {
                double v = W[NZ * NV + b];
#pragma unroll
                for (int c = 0; c < NX; ++c) v += ab.B[c] * (b < NX ? M[c][b] : Pm[sym(c, NX)]);
                L[b] = v;
            }
            ok = ok && (Rt > 0.0) && (Rt < 1e300);
            const double Ri = 1.0 / Rt;
            double RL[NZ];
#pragma unroll
            for (int b = 0; b < NZ; ++b) {
                RL[b] = Ri * L[b];
                stf(d, SL::D_RL + b, RL[b]);
            }
            stf(d, SL::D_RI, Ri);
            double Pn[NH];
#pragma unroll
            for (int a = 0; a < NZ; ++a)
#pragma unroll
                for (int b = a; b < NZ; ++b) {
                    double v = W[a * NV + b] + (atkg ? HG[C::sym(a, b)] : 0.0);
                    if (a < NX) {
#pragma unroll
                        for (int c = 0; c < NX; ++c) v += ab.A[c * NX + a] * (b < NX ? M[c][b] : Pm[sym(c, NX)]);
                    } else {
                        v += Pm[C::sym(NX, NX)];
                    }
                    v -= L[a] * RL[b];
                    if (a == b && a < NX) v += qd[a];
                    Pn[C::sym(a, b)] = v;
                }
#pragma unroll
            for (int a = 0; a < NH; ++a) Pm[a] = Pn[a];
            // ---- affine backward substitution, dual-residual adjoint, Farkas adjoint ----
            double rt = q[NX], vu = g[NZ], vc = gl[NX];
#pragma unroll
            for (int c = 0; c < NX; ++c) {
                rt += ab.B[c] * pv[c];
                vu += ab.B[c] * pi[c];
                vc += ab.B[c] * pc[c];
            }
            const double kap = -Ri * rt;
            stf(d, SL::D_KAP, kap);
            rd = lb_nanmax(rd, lb_abs(vu));
            ci += lb_abs(vc) * ((k >= p.ku0 && k <= p.ku1) ? p.fk_u[0] : p.fk_free);
            yd += vc * vn[NX];
            double pvn[NZ], pin[NZ], pcn[NZ];
#pragma unroll
            for (int a = 0; a < NZ; ++a) {
                double s0 = a < NX ? q[a] : g[a], s1 = g[a], s2 = a < NX ? gl[a] : 0.0;
                if (a < NX) {
#pragma unroll
                    for (int c = 0; c < NX; ++c) {
                        s0 += ab.A[c * NX + a] * pv[c];
                        s1 += ab.A[c * NX + a] * pi[c];
                        s2 += ab.A[c * NX + a] * pc[c];
                    }
                } else {
                    s0 += pv[NX];
                    s1 += pi[NX];
                    s2 += pc[NX];
                }
                s0 += L[a] * kap;
                if (atkg) {
                    s0 += gGl[a] + dG[a];
                    s1 += gGl[a];
                    s2 += gGl[a];
                }
                pvn[a] = s0; pin[a] = s1; pcn[a] = s2;
            }
#pragma unroll
            for (int a = 0; a < NZ; ++a) {
                pv[a] = pvn[a]; pi[a] = pin[a]; pc[a] = pcn[a];
            }
        }
        const double ptt = Pm[C::sym(NX, NX)];
        ok = ok && (ptt > 0.0);
        ln.iptt = 1.0 / ptt;
        ln.dtha = -ln.iptt * pv[NX];
        rd = lb_nanmax(rd, lb_abs(pi[NX]));
        ci += lb_abs(pc[NX]) * p.fk_free;
        yd += pc[NX] * th;
#pragma unroll
        for (int a = 0; a < NZ; ++a) ln.gGl[a] = gGl[a];
        ln.rp = rpm;
        ln.mu = sl * p.inv_m;
        ln.lam = lam;
        ln.hlam = hl + yd;
        ln.rd = rd;
        ln.cert = ci;
        ln.obj = J + ln.cconst;
        ln.piv_ok = ok;
        // verdict (oracle/lbmpc_oracle.c solve_ws)
        const double mu = ln.mu;
        if (!ok || !(rd == rd) || !(rpm == rpm) || !(mu == mu) || isinf(rd) || isinf(mu)) return 3;
        const double rd_tol = p.tol_res * (100.0 * lam > 1.0 ? 100.0 * lam : 1.0);
        if (rd < rd_tol && rpm < p.tol_res && mu < p.tol_mu) return 0;
        if (lam >= p.inf_trigger && ln.hlam < 0.0 && ci * p.inf_scale <= -ln.hlam) return 2;
        return -1;
    }

    // ============================================================================================
    // pass F1: affine forward substitution, affine row directions, sigma
    // ============================================================================================
    static LB_HD void pass_f1(const P& p, const SL& l, Lane& ln, const double* __restrict__ G, const double* __restrict__ hg,
                              const double* w64, FT* wft, bool has_cs) {
        const int N = p.N;
        double dxa[NX], ratio = 0.0, s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int j = 0; j < NX; ++j) dxa[j] = 0.0;
#pragma unroll
        for (int a = 0; a < NZ; ++a) ln.dG1[a] = ln.dG2[a] = 0.0;
        for (int k = 0; k <= N; ++k) {
            const double* it = w64 + (l.o_it + k * SL::RS_IT) * LS;
            FT* d = wft + k * SL::RS_D * LS;
            const unsigned rows = C::stage_rows(p, k);
            const bool last = k >= N;
            double v[NVB], dva[NVB];
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                v[j] = ld(it, j);
                dva[j] = dxa[j];
            }
            v[NX] = last ? 0.0 : ld(it, NX);
            dva[NX] = 0.0;
            if (!last) {
                double acc = ldf(d, SL::D_RL + NX) * ln.dtha;
#pragma unroll
                for (int c = 0; c < NX; ++c) acc += ldf(d, SL::D_RL + c) * dxa[c];
                dva[NX] = ldf(d, SL::D_KAP) - acc;
            }
#pragma unroll
            for (int j = 0; j < NVB; ++j)
                if (j < NX || !last) stf(d, SL::D_DA + j, dva[j]);
            double g[NV];
            cost_grad(p, l, ln, w64, k, v, has_cs, g, nullptr);
            stf(d, SL::D_QC + NX, g[NX]);
#pragma unroll
            for (int j = 0; j < NVB; ++j) {
                double t1 = 0.0, t2 = 0.0;
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    const int r = 2 * j + side;
                    const bool act = (rows >> r) & 1u;
                    const double sgn = side == 0 ? 1.0 : -1.0;
                    double S = 1.0, Lm = 0.0;
                    if (act) {
                        S = ld(it, SL::F_S + r);
                        Lm = ld(it, SL::F_LB + r);
                    }
                    const double slack = side == 0 ? p.hi[j] - v[j] : v[j] - p.lo[j];
                    const double rp = act ? S - slack : 0.0, is = lb_rcp(S), w = Lm * is;
                    const double dsa = act ? -rp - sgn * dva[j] : 0.0, dla = -Lm - w * dsa;
                    const double rr = dsa * is;
                    ratio = lb_max(ratio, act ? lb_max(-rr, 1.0 + rr) : 0.0);
                    s0 += S * Lm;
                    s1 += S * dla + Lm * dsa;
                    s2 += dsa * dla;
                    t1 += sgn * (w * rp - dsa * dla * is);
                    t2 += act ? sgn * is : 0.0;
                }
                if (j < NX || !last) {
                    stf(d, SL::D_QC + C::zidx(j), g[C::zidx(j)] + t1);
                    stf(d, SL::D_T2 + j, t2);
                }
            }
            if (k == p.kg) {
                double dza[NZ], z[NZ];
#pragma unroll
                for (int a = 0; a < NX; ++a) {
                    dza[a] = dxa[a];
                    z[a] = v[a];
                }
                dza[NX] = ln.dtha;
                z[NX] = ln.th;
                for (int i = 0; i < p.ng; ++i) {
                    double gi[NZ], slack = hg[i], adva = 0.0;
#pragma unroll
                    for (int a = 0; a < NZ; ++a) {
                        gi[a] = G[a * p.ngp + i];
                        slack -= gi[a] * z[a];
                        adva += gi[a] * dza[a];
                    }
                    const double S = ld(w64, l.o_sg + i), Lm = ld(w64, l.o_lg + i);
                    const double rp = S - slack, is = lb_rcp(S), w = Lm * is;
                    const double dsa = -rp - adva, dla = -Lm - w * dsa, rr = dsa * is;
                    ratio = lb_max(ratio, lb_max(-rr, 1.0 + rr));
                    s0 += S * Lm;
                    s1 += S * dla + Lm * dsa;
                    s2 += dsa * dla;
                    const double t1 = w * rp - dsa * dla * is - Lm;
#pragma unroll
                    for (int a = 0; a < NZ; ++a) {
                        ln.dG1[a] += gi[a] * t1;
                        ln.dG2[a] += gi[a] * is;
                    }
                }
            }
            if (!last) {
                AB ab;
                load_ab(p, l, w64, k, ab);
                double xn[NX];
#pragma unroll
                for (int a = 0; a < NX; ++a) {
                    double acc = ab.B[a] * dva[NX];
#pragma unroll
                    for (int c = 0; c < NX; ++c) acc += ab.A[a * NX + c] * dxa[c];
                    xn[a] = acc;
                }
#pragma unroll
                for (int a = 0; a < NX; ++a) dxa[a] = xn[a];
            }
        }
        const double aaff = ratio > 1.0 ? 1.0 / ratio : 1.0;
        const double mu_aff = (s0 + aaff * s1 + aaff * aaff * s2) * p.inv_m;
        const double sr = mu_aff / ln.mu;
        const double sg = sr * sr * ln.mu;
        ln.sigmu = sg > 0.1 * p.tol_mu ? sg : 0.1 * p.tol_mu;  // no centring below the target gap
    }

    // ============================================================================================
    // pass B2: corrector backward substitution from the stored right-hand side pieces
    // ============================================================================================
    static LB_HD void pass_b2(const P& p, const SL& l, Lane& ln, const double* w64, FT* wft) {
        const int N = p.N;
        const double sigmu = ln.sigmu;
        double pv[NZ], kgt[NZ];
#pragma unroll
        for (int a = 0; a < NZ; ++a) kgt[a] = ln.gGl[a] + ln.dG1[a] + sigmu * ln.dG2[a];
        {
            const FT* d = wft + N * SL::RS_D * LS;
#pragma unroll
            for (int a = 0; a < NZ; ++a)
                pv[a] = ldf(d, SL::D_QC + a) + (a < NX ? sigmu * ldf(d, SL::D_T2 + a) : 0.0) + (p.kg == N ? kgt[a] : 0.0);
        }
        for (int k = N - 1; k >= 0; --k) {
            FT* d = wft + k * SL::RS_D * LS;
            AB ab;
            load_ab(p, l, w64, k, ab);
            double rt = ldf(d, SL::D_QC + NZ) + sigmu * ldf(d, SL::D_T2 + NX);
#pragma unroll
            for (int c = 0; c < NX; ++c) rt += ab.B[c] * pv[c];
            stf(d, SL::D_KAP, -ldf(d, SL::D_RI) * rt);
            double pvn[NZ];
#pragma unroll
            for (int a = 0; a < NZ; ++a) {
                double v = ldf(d, SL::D_QC + a) + (a < NX ? sigmu * ldf(d, SL::D_T2 + a) : 0.0);
                if (a < NX) {
#pragma unroll
                    for (int c = 0; c < NX; ++c) v += ab.A[c * NX + a] * pv[c];
                } else {
                    v += pv[NX];
                }
                v -= ldf(d, SL::D_RL + a) * rt;
                if (k == p.kg) v += kgt[a];
                pvn[a] = v;
            }
#pragma unroll
            for (int a = 0; a < NZ; ++a) pv[a] = pvn[a];
        }
        ln.dth = -ln.iptt * pv[NX];
    }

    // ============================================================================================
    // pass F2: corrector forward substitution, final row directions, step length
    // ============================================================================================
    static LB_HD void pass_f2(const P& p, const SL& l, Lane& ln, const double* __restrict__ G, const double* __restrict__ hg,
                              const double* w64, FT* wft) {
        const int N = p.N;
        const double sigmu = ln.sigmu;
        double dx[NX], ratio = 0.0;
#pragma unroll
        for (int j = 0; j < NX; ++j) dx[j] = 0.0;
        for (int k = 0; k <= N; ++k) {
            const double* it = w64 + (l.o_it + k * SL::RS_IT) * LS;
            FT* d = wft + k * SL::RS_D * LS;
            const unsigned rows = C::stage_rows(p, k);
            const bool last = k >= N;
            double v[NVB], dva[NVB], dv[NVB];
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                v[j] = ld(it, j);
                dva[j] = ldf(d, SL::D_DA + j);
                dv[j] = dx[j];
            }
            v[NX] = last ? 0.0 : ld(it, NX);
            dva[NX] = last ? 0.0 : ldf(d, SL::D_DA + NX);
            dv[NX] = 0.0;
            if (!last) {
                double acc = ldf(d, SL::D_RL + NX) * ln.dth;
#pragma unroll
                for (int c = 0; c < NX; ++c) acc += ldf(d, SL::D_RL + c) * dx[c];
                dv[NX] = ldf(d, SL::D_KAP) - acc;
            }
#pragma unroll
            for (int j = 0; j < NVB; ++j)
                if (j < NX || !last) stf(d, SL::D_DV + j, dv[j]);
#pragma unroll
            for (int j = 0; j < NVB; ++j) {
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    const int r = 2 * j + side;
                    const bool act = (rows >> r) & 1u;
                    const double sgn = side == 0 ? 1.0 : -1.0;
                    double S = 1.0, Lm = 1.0;
                    if (act) {
                        S = ld(it, SL::F_S + r);
                        Lm = ld(it, SL::F_LB + r);
                    }
                    const double slack = side == 0 ? p.hi[j] - v[j] : v[j] - p.lo[j];
                    const double rp = S - slack, is = lb_rcp(S), w = Lm * is;
                    const double dsa = -rp - sgn * dva[j], dla = -Lm - w * dsa;
                    const double ds = -rp - sgn * dv[j];
                    const double rc = S * Lm + dsa * dla - sigmu;
                    const double dl = (-rc - Lm * ds) * is;
                    ratio = lb_max(ratio, act ? lb_max(-ds * is, -dl * lb_rcp(Lm)) : 0.0);
                }
            }
            if (k == p.kg) {
                double dza[NZ], dz[NZ], z[NZ];
#pragma unroll
                for (int a = 0; a < NX; ++a) {
                    dza[a] = dva[a];
                    dz[a] = dx[a];
                    z[a] = v[a];
                }
                dza[NX] = ln.dtha;
                dz[NX] = ln.dth;
                z[NX] = ln.th;
                for (int i = 0; i < p.ng; ++i) {
                    double slack = hg[i], adva = 0.0, adv = 0.0;
#pragma unroll
                    for (int a = 0; a < NZ; ++a) {
                        const double gia = G[a * p.ngp + i];
                        slack -= gia * z[a];
                        adva += gia * dza[a];
                        adv += gia * dz[a];
                    }
                    const double S = ld(w64, l.o_sg + i), Lm = ld(w64, l.o_lg + i);
                    const double rp = S - slack, is = lb_rcp(S), w = Lm * is;
                    const double dsa = -rp - adva, dla = -Lm - w * dsa;
                    const double ds = -rp - adv;
                    const double rc = S * Lm + dsa * dla - sigmu;
                    const double dl = (-rc - Lm * ds) * is;
                    ratio = lb_max(ratio, lb_max(-ds * is, -dl * lb_rcp(Lm)));
                }
            }
            if (!last) {
                AB ab;
                load_ab(p, l, w64, k, ab);
                double xn[NX];
#pragma unroll
                for (int a = 0; a < NX; ++a) {
                    double acc = ab.B[a] * dv[NX];
#pragma unroll
                    for (int c = 0; c < NX; ++c) acc += ab.A[a * NX + c] * dx[c];
                    xn[a] = acc;
                }
#pragma unroll
                for (int a = 0; a < NX; ++a) dx[a] = xn[a];
            }
        }
        double alpha = ratio > 0.0 ? 0.99 / ratio : 1.0;
        ln.alpha = alpha > 1.0 ? 1.0 : alpha;
    }

    // ============================================================================================
    // results of a finished QP: c = u - Kout x (transitionNominal.m:12 undone), theta, objective
    // ============================================================================================
    static LB_HD void finish(const P& p, const SL& l, const Lane& ln, const StreamIO<FT>& io, const double* w64, int status) {
        const int N = p.N;
        const long long q = ln.q;
        for (int k = 0; k <= N; ++k) {
            const double* it = w64 + (l.o_it + k * SL::RS_IT) * LS;
            double x[NX];
#pragma unroll
            for (int j = 0; j < NX; ++j) x[j] = ld(it, j);
            if (io.xtraj) {
#pragma unroll
                for (int j = 0; j < NX; ++j) io.xtraj[(q * (N + 1) + k) * NX + j] = x[j];
            }
            if (k < N) {
                double v = ld(it, NX);
#pragma unroll
                for (int j = 0; j < NX; ++j) v -= p.Kout[j] * x[j];
                io.uc[q * N + k] = v;
            }
        }
        io.theta[q] = ln.th;
        io.obj[q] = ln.obj;
        io.iters[q] = ln.iters;
        io.status[q] = status;
    }

    // one whole solve by one lane (tests/emul; the kernel interleaves the same calls with the queue logic)
    static LB_HD void solve_one(const P& p, const SL& l, const StreamIO<FT>& io, long long q, const double* G, const double* hg,
                                double* w64, FT* wft) {
        Lane ln;
        const bool has_cs = io.cshift != nullptr;
        init_qp(p, l, ln, io, q, w64, has_cs);
        for (;;) {
            int st = pass_bu(p, l, ln, G, hg, w64, wft, has_cs);
            if (st < 0 && ln.iters >= p.max_iter) st = 1;
            if (st >= 0) {
                finish(p, l, ln, io, w64, st);
                return;
            }
            pass_f1(p, l, ln, G, hg, w64, wft, has_cs);
            pass_b2(p, l, ln, w64, wft);
            pass_f2(p, l, ln, G, hg, w64, wft);
            ln.iters += 1;
            ln.fresh = false;
        }
    }
};

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// the kernel: persistent lanes, QPs from a global queue (warp-aggregated atomics)
// ---------------------------------------------------------------------------------------------
template <int NX, bool LTV, typename FT>
__global__ void __launch_bounds__(128, 2)
ipm_stream_kernel(const __grid_constant__ Params<NX, 1, 1> p, const StreamIO<FT> io, const double* __restrict__ G,
                  const double* __restrict__ hg) {
    using S = Stream<NX, LTV, FT, 32>;
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const bool has_cs = io.cshift != nullptr;
    const StreamLayout<NX> l(p.N, p.ng, has_cs, LTV);
    double* const w64 = io.ws64 + warp * (long long)l.n64 * 32 + lane;
    FT* const wft = io.wsft + warp * (long long)l.nft * 32 + lane;
    typename S::Lane ln;
    ln.q = -1;
    bool drained = false;  // warp-uniform: the queue has run dry
    for (;;) {
        const bool need = ln.q < 0;
        const unsigned want = drained ? 0u : __ballot_sync(0xffffffffu, need);
        if (want) {
            unsigned long long base = 0;
            const int leader = __ffs(want) - 1;
            if (lane == leader) base = atomicAdd(io.queue, (unsigned long long)__popc(want));
            base = __shfl_sync(0xffffffffu, base, leader);
            const long long mine = (long long)base + __popc(want & ((1u << lane) - 1u));
            if (need && mine < io.batch) S::init_qp(p, l, ln, io, mine, w64, has_cs);
            drained = (long long)base + __popc(want) >= io.batch;
        }
        if (__all_sync(0xffffffffu, ln.q < 0)) break;
        if (ln.q >= 0) {
            int st = S::pass_bu(p, l, ln, G, hg, w64, wft, has_cs);
            if (st < 0 && ln.iters >= p.max_iter) st = 1;
            if (st >= 0) {
                S::finish(p, l, ln, io, w64, st);
                ln.q = -1;
            } else {
                S::pass_f1(p, l, ln, G, hg, w64, wft, has_cs);
                S::pass_b2(p, l, ln, w64, wft);
                S::pass_f2(p, l, ln, G, hg, w64, wft);
                ln.iters += 1;
                ln.fresh = false;
            }
        }
    }
}
#endif

}  // namespace lbmpc
