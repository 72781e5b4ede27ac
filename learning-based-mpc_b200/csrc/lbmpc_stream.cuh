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
#ifdef __CUDACC__
#include "lbmpc_kernels.cuh"  // mbarrier / 1-D bulk copy helpers
#endif

namespace lbmpc {

template <int NX>
struct StreamLayout {
    static constexpr int NZ = NX + 1, NV = NX + 2, NVB = NX + 1, NH = NZ * (NZ + 1) / 2;
    // FP64 record: iterate x u s lambda, then the step [dx du] that the next pass BU applies (FP64 in every mode: the
    // iterate stays on the dynamics x+ = A x + B u + d exactly only if the applied step does)
    static constexpr int F_X = 0, F_U = NX, F_S = NVB, F_LB = 3 * NVB, F_DV = 5 * NVB, RS_IT = 6 * NVB, RS_IT0 = 5 * NVB;
    // FT record; order chosen so that what a pass reads of a stage is ONE contiguous range: BU [DA], F2 [DA .. KAP],
    // F1 [RL .. KAP], B2 [RL .. T2]
    static constexpr int D_DA = 0, D_RL = NVB, D_RI = D_RL + NZ, D_KAP = D_RI + 1, D_T2 = D_KAP + 1, RS_D = D_T2 + NVB;
    // second FP64 record: the sigma-independent part of the corrector right-hand side (a residual: FP64 in every mode)
    static constexpr int RS_Q = NV;
    static constexpr int kPolyChunk = 16;  // polytope rows fetched per pipeline item
    static constexpr int NJ = NX * 3;  // LTV: Jacobian of the learned term w.r.t. xi = [x1; x2; u] per stage
    int N, ng, o_it, o_q, o_cs, o_j, o_sg, o_lg, n64, nft;  // offsets / sizes in elements (per QP)
    int o_loop, qwin;  // fused closed loop: plant state x[NX], current plan [u (N); theta], data ring X[3 qwin], Y[NX qwin]
    LB_HD int lp_x() const { return o_loop; }
    LB_HD int lp_plan() const { return o_loop + NX; }
    LB_HD int lp_X() const { return o_loop + NX + N + 1; }
    LB_HD int lp_Y() const { return o_loop + NX + N + 1 + 3 * qwin; }
    LB_HD StreamLayout(int N_, int ng_, bool cs, bool ltv, int qwin_ = 0) {
        N = N_;
        ng = ng_;
        qwin = qwin_;
        int o = 0;
        o_it = o; o += RS_IT * (N + 1);
        o_q = o;  o += RS_Q * (N + 1);
        o_cs = o; o += cs ? NVB * (N + 1) : 0;  // [ex | eu] per stage
        o_j = o;  o += ltv ? NJ * N : 0;
        o_sg = o; o += ng;
        o_lg = o; o += ng;
        o_loop = o; o += qwin > 0 ? NX + N + 1 + (3 + NX) * qwin : 0;
        n64 = o;
        nft = RS_D * (N + 1);
    }
};

// per-lane state that lives across the passes of one QP (registers)
template <int NX, int CSTR = 1>
struct StreamLane {
    static constexpr int NZ = NX + 1;
    long long q;          // QP index, -1: idle
    int iters;
    bool fresh;           // no step to apply yet: the next pass BU initialises the rows
    double th, dtha, dth, alpha, sigmu, iptt, mu;
    // values that are touched at ONE stage of a pass only (the polytope stage, the reference stage) or at its end: parked in
    // shared memory on the device (one column per lane, stride CSTR) instead of holding 44 registers through every sweep
    static constexpr int C_GGL = 0, C_DG1 = NZ, C_DG2 = 2 * NZ, C_LIN = 3 * NZ, C_CCONST = 4 * NZ, C_OBJ = 4 * NZ + 1, C_N = 4 * NZ + 2;
    double* cold;
    LB_HD double& gGl(int a) const { return cold[(C_GGL + a) * CSTR]; }
    LB_HD double& dG1(int a) const { return cold[(C_DG1 + a) * CSTR]; }
    LB_HD double& dG2(int a) const { return cold[(C_DG2 + a) * CSTR]; }
    LB_HD double& lin(int a) const { return cold[(C_LIN + a) * CSTR]; }
    LB_HD double& cconst() const { return cold[C_CCONST * CSTR]; }
    LB_HD double& obj() const { return cold[C_OBJ * CSTR]; }
    // results of pass BU
    double rp, lam, hlam, rd, cert;
    bool piv_ok;
};

template <typename FT>
struct StreamIO {  // batch-major caller arrays (include/lbmpc.h)
    long long batch;
    const double *dx0, *dx_ref, *d_off, *warm, *cshift, *jac;
    int cs_stride;  // doubles per stage of cshift: NX (state shift) or NX + 1 (state and input shift)
    int row_shift;  // 0: cshift moves the COST; 1: it moves the ROWS (they act on x_k - ex_k, u_k - eu_k; first-order SQP)
    int qwin;       // fused closed loop: data window length (0 otherwise)
    // iteration budget of the mapping (0: none).  One iteration of a lane lasts ~0.5 ms x N/50, so the few QPs that need 20 - 40
    // iterations would hold the whole launch; a QP that has not reached its verdict after `evict_iters` iterations is handed
    // over (index appended to left_list) and solved from scratch by a shared-memory mapping right after this kernel.
    int evict_iters;
    long long* left_list;
    unsigned long long* left_count;
    double *uc, *theta, *xtraj, *obj;
    int *iters, *status;
    unsigned long long* queue;
    double* ws64;  // workspace: per warp n64 * 32 doubles ...
    FT* wsft;      // ... and nft * 32 FT
};

enum StreamPass : int { PASS_BU = 0, PASS_F1 = 1, PASS_B2 = 2, PASS_F2 = 3 };
constexpr int kStreamNone = -(1 << 30);  // pipeline item: nothing to fetch

// what one pass reads of a stage: the FP64 iterate record (all passes but B2) and ONE contiguous element range of the
// direction / factor record; cost shift in the two passes that evaluate the cost gradient
template <int NX>
struct StreamNeeds {
    using SL = StreamLayout<NX>;
    static LB_HD int it(int pass) { return pass == PASS_B2 ? 0 : (pass == PASS_BU ? SL::RS_IT : SL::RS_IT0); }  // leading elements of the FP64 record
    static LB_HD int d_lo(int pass) { return (pass == PASS_BU || pass == PASS_F2) ? SL::D_DA : SL::D_RL; }
    static LB_HD int d_hi(int pass) { return pass == PASS_BU ? SL::D_RL : (pass == PASS_B2 ? SL::RS_D : SL::D_T2); }
    static LB_HD bool cs(int pass, bool rows) { return pass == PASS_BU || pass == PASS_F1 || (rows && pass == PASS_F2); }
};

// read-only view of the pipeline item a pass is working on (stride LS between elements):
//   it[e]  element e of the stage's iterate record        d[e]  element e of its direction / factor record (pass range only)
//   cs[j]  cost shift of the stage                        jac[i] LTV Jacobian of the stage
//   sg[i], lg[i]  slack / multiplier of polytope row i (absolute row index, valid inside the fetched chunk)
template <typename P64, typename PFT>
struct StreamView {
    P64 it;
    PFT d;
    P64 q, cs, jac, sg, lg;
};
#ifdef __CUDACC__
// 32-bit shared-space addresses (ld.shared with immediate offsets instead of generic loads)
struct StreamSh64 { unsigned a; };
template <typename FT> struct StreamShFT { unsigned a; };
#endif

// Direct pipe: every read goes to the workspace itself (host emulation; LS = 1 or 32)
template <int NX, typename FT, int LS>
struct StreamPipeDirect {
    using SL = StreamLayout<NX>;
    const SL* l;
    const double* w64;
    const FT* wft;
    int ring[2], n_issue, n_wait;
    LB_HD StreamPipeDirect(const SL& l_, const double* w64_, const FT* wft_) : l(&l_), w64(w64_), wft(wft_), n_issue(0), n_wait(0) {}
    LB_HD void start(int) {}
    LB_HD void issue(int item) {
        if (item == kStreamNone) return;
        ring[n_issue++ & 1] = item;
    }
    using View = StreamView<const double*, const FT*>;
    LB_HD View acquire() {
        const int item = ring[n_wait++ & 1];
        View v;
        const int k = item >= 0 ? item : 0;
        v.it = w64 + (l->o_it + k * SL::RS_IT) * LS;
        v.d = wft + k * SL::RS_D * LS;
        v.q = w64 + (l->o_q + k * SL::RS_Q) * LS;
        v.cs = w64 + (l->o_cs + k * (NX + 1)) * LS;
        v.jac = w64 + (l->o_j + k * SL::NJ) * LS;
        v.sg = w64 + l->o_sg * LS;
        v.lg = w64 + l->o_lg * LS;
        return v;
    }
};

// fused closed loop: parameters and caller arrays (include/lbmpc.h lbmpc_closed_loop)
struct StreamLoopParams {
    int steps, chunk, warm_shift, use_oracle, use_w;
    long long nscen;
    double x_eq[4], u_eq, wbar[4], inv_h2, lambda;
    unsigned long long seed, scen0;
    const double* x_init;                      // NX x nscen, absolute plant states
    double *x_hist, *u_hist, *theta_hist;      // histories (any may be null)
    int *iters_hist, *status_hist;
    double* store;                             // nscen x store_len: scenario state between chunks
    long long* rq;                             // re-queued scenarios, entry = scenario + 1 (0: not produced yet)
    unsigned long long* rq_tail;
};

template <int NX, bool LTV, typename FT, int LS>
struct Stream {
    using P = Params<NX, 1, 1>;
    using SL = StreamLayout<NX>;
    using Lane = StreamLane<NX, LS>;
    using C = Core<NX, 1, 1>;
    static constexpr int NZ = NX + 1, NV = NX + 2, NVB = NX + 1, NH = SL::NH, CH = SL::kPolyChunk;

#ifdef __CUDACC__
    static __device__ __forceinline__ double ld(StreamSh64 p, int e) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(p.a + 8u * (unsigned)(e * LS)) : "memory");
        return v;
    }
    static __device__ __forceinline__ double ldf(StreamShFT<double> p, int e) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(p.a + 8u * (unsigned)(e * LS)) : "memory");
        return v;
    }
    static __device__ __forceinline__ double ldf(StreamShFT<float> p, int e) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(p.a + 4u * (unsigned)(e * LS)) : "memory");
        return (double)v;
    }
#endif
    static LB_HD double ld(const double* p, int e) { return p[e * LS]; }
    static LB_HD void st(double* p, int e, double v) { p[e * LS] = v; }
    static LB_HD double ldf(const FT* p, int e) { return (double)p[e * LS]; }
    static LB_HD void stf(FT* p, int e, double v) { p[e * LS] = (FT)v; }
    static LB_HD int sym(int a, int b) { return a <= b ? C::sym(a, b) : C::sym(b, a); }

    struct AB {  // dynamics of the stage being processed
        double A[NX * NX], B[NX];
    };
    // LTI: the constant bank; LTV: A_k = A + [J(:,1:2) 0 0], B_k = B + J(:,3), J = Jacobian of the stage (jac[c*3+b])
    template <typename PJ>
    static LB_HD void load_ab(const P& p, PJ jac, AB& ab) {
#pragma unroll
        for (int c = 0; c < NX; ++c) {
#pragma unroll
            for (int b = 0; b < NX; ++b) {
                double v = p.A[c * NX + b];
                if (LTV && b < 2) v += ld(jac, c * 3 + b);
                ab.A[c * NX + b] = v;
            }
            double v = p.B[c];
            if (LTV) v += ld(jac, c * 3 + 2);
            ab.B[c] = v;
        }
    }

    // cost gradient g = W_type(k) [x + e; theta; u] (u part dropped at the last stage), objective piece 0.5 v'g (+ lin'z at kT)
    template <typename PC>
    static LB_HD void cost_grad(const P& p, const Lane& ln, PC cs, int k, const double* v, bool has_cs, double* g,
                                double* Jacc) {
        const bool last = k >= p.N;
        double vv[NV];
#pragma unroll
        for (int j = 0; j < NX; ++j) vv[j] = v[j] + (has_cs ? ld(cs, j) : 0.0);
        vv[NX] = ln.th;
        vv[NZ] = last ? 0.0 : v[NX] + (has_cs ? ld(cs, NX) : 0.0);
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
                g[a] += ln.lin(a);
                J += ln.lin(a) * vv[a];
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
            ln.lin(a) = v;
            ln.gGl(a) = ln.dG1(a) = ln.dG2(a) = 0.0;
        }
        if (io.dx_ref) {
#pragma unroll
            for (int i = 0; i < NX; ++i)
#pragma unroll
                for (int j = 0; j < NX; ++j) cconst += io.dx_ref[q * NX + i] * p.Tm[i * NX + j] * io.dx_ref[q * NX + j];
        }
        ln.cconst() = cconst;
        if (LTV) {
            const double* jq = io.jac + q * (long long)(N * SL::NJ);
            for (int i = 0; i < N * SL::NJ; ++i) st(w64, l.o_j + i, jq[i]);
        }
        if (has_cs) {
            const double* cq = io.cshift + q * (long long)((N + 1) * io.cs_stride);
            for (int k = 0; k <= N; ++k)
#pragma unroll
                for (int j = 0; j < NVB; ++j) st(w64, l.o_cs + k * NVB + j, j < io.cs_stride ? cq[k * io.cs_stride + j] : 0.0);
        }
        double x[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) x[j] = io.dx0[q * NX + j];
        // the caller's warm start and offsets of KB stages are fetched before the (sequential) rollout over them: loaded stage by
        // stage, every stage waited for its own scattered loads — 5 % of the warp time of a closed-loop launch, where every QP
        // brings both arrays (same remedy as in finish())
        constexpr int KB = 8;
        for (int k0 = 0; k0 <= N; k0 += KB) {
            double wm[KB], dk[KB][NX];
#pragma unroll
            for (int i = 0; i < KB; ++i) {
                const int k = k0 + i < N ? k0 + i : N - 1;
                wm[i] = io.warm ? io.warm[q * (N + 1) + k] : 0.0;
#pragma unroll
                for (int a = 0; a < NX; ++a) dk[i][a] = io.d_off ? io.d_off[(q * N + k) * NX + a] : 0.0;
            }
#pragma unroll
            for (int i = 0; i < KB; ++i) {
                const int k = k0 + i;
                if (k > N) break;
                double* it = w64 + (l.o_it + k * SL::RS_IT) * LS;
#pragma unroll
                for (int j = 0; j < NX; ++j) st(it, SL::F_X + j, x[j]);
                if (k == N) break;
                AB ab;
                load_ab(p, w64 + (l.o_j + k * SL::NJ) * LS, ab);
                double u = wm[i];
#pragma unroll
                for (int j = 0; j < NX; ++j) u += p.Kinit[j] * x[j];
                st(it, SL::F_U, u);
                double xn[NX];
#pragma unroll
                for (int a = 0; a < NX; ++a) {
                    double v = dk[i][a];
#pragma unroll
                    for (int j = 0; j < NX; ++j) v += ab.A[a * NX + j] * x[j];
                    v += ab.B[a] * u;
                    xn[a] = v;
                }
#pragma unroll
                for (int a = 0; a < NX; ++a) x[a] = xn[a];
            }
        }
    }

    // next pipeline item after stage k's own record in a backward / forward pass
    static LB_HD int next_bwd(int k) { return k > 0 ? k - 1 : kStreamNone; }
    static LB_HD int next_fwd(int k, int N) { return k < N ? k + 1 : kStreamNone; }

    // ============================================================================================
    // pass BU.  Returns the verdict (LBMPC_ST_*) or -1 to continue.
    // ============================================================================================
    template <class Pipe>
    static LB_HD int pass_bu(const P& p, const SL& l, Lane& ln, const double* __restrict__ G, const double* __restrict__ hg,
                             double* w64, FT* wft, bool has_cs, bool rsh, Pipe& pp) {
        const int N = p.N;
        const int nch = (p.ng + CH - 1) / CH;
        const bool fresh = ln.fresh;
        const double alpha = fresh ? 0.0 : ln.alpha, sigmu = ln.sigmu;
        const double th_old = ln.th;
        ln.th = fresh ? th_old : th_old + alpha * ln.dth;
        const double th = ln.th;
        double Pm[NH], pv[NZ], pi[NZ], pc[NZ];
        double rpm = 0.0, sl = 0.0, lam = 0.0, hl = 0.0, rd = 0.0, ci = 0.0, yd = 0.0, J = 0.0;
        bool ok = true;
#pragma unroll
        for (int a = 0; a < NZ; ++a) ln.gGl(a) = 0.0;
        pp.start(PASS_BU);
        pp.issue(N);
        for (int k = N; k >= 0; --k) {
            const bool atkg = k == p.kg;
            // polytope sums: alive at the one stage that carries the block only (declared per stage so that they do not hold
            // 50 registers through the whole sweep)
            double HG[NH], gGl[NZ], dG[NZ];
#pragma unroll
            for (int a = 0; a < NH; ++a) HG[a] = 0.0;
#pragma unroll
            for (int a = 0; a < NZ; ++a) gGl[a] = dG[a] = 0.0;
            pp.issue((atkg && nch > 0) ? -1 : next_bwd(k));
            const auto vw = pp.acquire();
            double* it = w64 + (l.o_it + k * SL::RS_IT) * LS;
            FT* d = wft + k * SL::RS_D * LS;
            const unsigned rows = C::stage_rows(p, k);
            const bool last = k >= N;
            AB ab;
            load_ab(p, vw.jac, ab);
            // ---- apply the step / initialise the rows, then the predictor assembly at the new iterate ----
            double vo[NVB], vn[NVB], dva[NVB], dv[NVB];
#pragma unroll
            for (int j = 0; j < NVB; ++j) {
                const bool ex = j < NX || !last;
                vo[j] = ex ? ld(vw.it, j) : 0.0;
                dva[j] = (ex && !fresh) ? ldf(vw.d, SL::D_DA + j) : 0.0;
                dv[j] = (ex && !fresh) ? ld(vw.it, SL::F_DV + j) : 0.0;
                vn[j] = vo[j] + alpha * dv[j];
                if (ex) st(it, j, vn[j]);
            }
            double g[NV], es[NVB];  // es: row shift of the stage (rows act on v - es)
#pragma unroll
            for (int j = 0; j < NVB; ++j) es[j] = (has_cs && rsh && (j < NX || !last)) ? ld(vw.cs, j) : 0.0;
            cost_grad(p, ln, vw.cs, k, vn, has_cs && !rsh, g, &J);
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
                        S = ld(vw.it, SL::F_S + r);
                        Lm = ld(vw.it, SL::F_LB + r);
                    }
                    const double slack_o = side == 0 ? p.hi[j] - (vo[j] - es[j]) : (vo[j] - es[j]) - p.lo[j];
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
                    const double slack = side == 0 ? p.hi[j] - (vn[j] - es[j]) : (vn[j] - es[j]) - p.lo[j];
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
            // ---- polytope block (chunks of CH rows through the pipeline) ----
            if (atkg) {
                double zo[NZ], zn[NZ], dza[NZ], dz[NZ];
#pragma unroll
                for (int a = 0; a < NX; ++a) {
                    zo[a] = vo[a] - es[a]; zn[a] = vn[a] - es[a]; dza[a] = dva[a]; dz[a] = dv[a];
                }
                zo[NX] = th_old; zn[NX] = th; dza[NX] = ln.dtha; dz[NX] = ln.dth;
                for (int c = 0; c < nch; ++c) {
                    pp.issue(c + 1 < nch ? -(c + 2) : next_bwd(k));
                    const auto pw = pp.acquire();
                    const int i1 = (c + 1) * CH < p.ng ? (c + 1) * CH : p.ng;
                    for (int i = c * CH; i < i1; ++i) {
                        double gi[NZ], so = hg[i], sn = hg[i], adva = 0.0, adv = 0.0;
#pragma unroll
                        for (int a = 0; a < NZ; ++a) {
                            gi[a] = G[a * p.ngp + i];
                            so -= gi[a] * zo[a];
                            sn -= gi[a] * zn[a];
                            adva += gi[a] * dza[a];
                            adv += gi[a] * dz[a];
                        }
                        const double S = ld(pw.sg, i), Lm = ld(pw.lg, i);
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
#pragma unroll
                for (int a = 0; a < NZ; ++a) ln.gGl(a) = gGl[a];
            }
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
        ci += lb_abs(pc[NX]) * p.fk_th[0];
        yd += pc[NX] * th;
        ln.rp = rpm;
        ln.mu = sl * p.inv_m;
        ln.lam = lam;
        ln.hlam = hl + yd;
        ln.rd = rd;
        ln.cert = ci;
        ln.obj() = J + ln.cconst();
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
    template <class Pipe>
    static LB_HD void pass_f1(const P& p, const SL& l, Lane& ln, const double* __restrict__ G, const double* __restrict__ hg,
                              double* w64, FT* wft, bool has_cs, bool rsh, Pipe& pp) {
        const int N = p.N;
        const int nch = (p.ng + CH - 1) / CH;
        double dxa[NX], ratio = 0.0, s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int j = 0; j < NX; ++j) dxa[j] = 0.0;
#pragma unroll
        for (int a = 0; a < NZ; ++a) ln.dG1(a) = ln.dG2(a) = 0.0;
        pp.start(PASS_F1);
        pp.issue(0);
        for (int k = 0; k <= N; ++k) {
            const bool atkg = k == p.kg;
            pp.issue((atkg && nch > 0) ? -1 : next_fwd(k, N));
            const auto vw = pp.acquire();
            FT* d = wft + k * SL::RS_D * LS;
            const unsigned rows = C::stage_rows(p, k);
            const bool last = k >= N;
            AB ab;
            load_ab(p, vw.jac, ab);
            double v[NVB], dva[NVB];
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                v[j] = ld(vw.it, j);
                dva[j] = dxa[j];
            }
            v[NX] = last ? 0.0 : ld(vw.it, NX);
            dva[NX] = 0.0;
            if (!last) {
                double acc = ldf(vw.d, SL::D_RL + NX) * ln.dtha;
#pragma unroll
                for (int c = 0; c < NX; ++c) acc += ldf(vw.d, SL::D_RL + c) * dxa[c];
                dva[NX] = ldf(vw.d, SL::D_KAP) - acc;
            }
#pragma unroll
            for (int j = 0; j < NVB; ++j)
                if (j < NX || !last) stf(d, SL::D_DA + j, dva[j]);
            double g[NV], es[NVB];
#pragma unroll
            for (int j = 0; j < NVB; ++j) es[j] = (has_cs && rsh && (j < NX || !last)) ? ld(vw.cs, j) : 0.0;
            cost_grad(p, ln, vw.cs, k, v, has_cs && !rsh, g, nullptr);
            double* qr = w64 + (l.o_q + k * SL::RS_Q) * LS;
            st(qr, NX, g[NX]);
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
                        S = ld(vw.it, SL::F_S + r);
                        Lm = ld(vw.it, SL::F_LB + r);
                    }
                    const double slack = side == 0 ? p.hi[j] - (v[j] - es[j]) : (v[j] - es[j]) - p.lo[j];
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
                    st(qr, C::zidx(j), g[C::zidx(j)] + t1);
                    stf(d, SL::D_T2 + j, t2);
                }
            }
            if (atkg) {
                double dza[NZ], z[NZ];
#pragma unroll
                for (int a = 0; a < NX; ++a) {
                    dza[a] = dxa[a];
                    z[a] = v[a] - es[a];
                }
                dza[NX] = ln.dtha;
                z[NX] = ln.th;
                for (int c = 0; c < nch; ++c) {
                    pp.issue(c + 1 < nch ? -(c + 2) : next_fwd(k, N));
                    const auto pw = pp.acquire();
                    const int i1 = (c + 1) * CH < p.ng ? (c + 1) * CH : p.ng;
                    for (int i = c * CH; i < i1; ++i) {
                        double gi[NZ], slack = hg[i], adva = 0.0;
#pragma unroll
                        for (int a = 0; a < NZ; ++a) {
                            gi[a] = G[a * p.ngp + i];
                            slack -= gi[a] * z[a];
                            adva += gi[a] * dza[a];
                        }
                        const double S = ld(pw.sg, i), Lm = ld(pw.lg, i);
                        const double rp = S - slack, is = lb_rcp(S), w = Lm * is;
                        const double dsa = -rp - adva, dla = -Lm - w * dsa, rr = dsa * is;
                        ratio = lb_max(ratio, lb_max(-rr, 1.0 + rr));
                        s0 += S * Lm;
                        s1 += S * dla + Lm * dsa;
                        s2 += dsa * dla;
                        const double t1 = w * rp - dsa * dla * is - Lm;
#pragma unroll
                        for (int a = 0; a < NZ; ++a) {
                            ln.dG1(a) += gi[a] * t1;
                            ln.dG2(a) += gi[a] * is;
                        }
                    }
                }
            }
            if (!last) {
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
    template <class Pipe>
    static LB_HD void pass_b2(const P& p, const SL& l, Lane& ln, FT* wft, Pipe& pp) {
        const int N = p.N;
        const double sigmu = ln.sigmu;
        double pv[NZ], kgt[NZ];
#pragma unroll
        for (int a = 0; a < NZ; ++a) kgt[a] = ln.gGl(a) + ln.dG1(a) + sigmu * ln.dG2(a);
        pp.start(PASS_B2);
        pp.issue(N);
        for (int k = N; k >= 0; --k) {
            pp.issue(next_bwd(k));
            const auto vw = pp.acquire();
            if (k == N) {
#pragma unroll
                for (int a = 0; a < NZ; ++a)
                    pv[a] = ld(vw.q, a) + (a < NX ? sigmu * ldf(vw.d, SL::D_T2 + a) : 0.0) + (p.kg == N ? kgt[a] : 0.0);
                continue;
            }
            FT* d = wft + k * SL::RS_D * LS;
            AB ab;
            load_ab(p, vw.jac, ab);
            double rt = ld(vw.q, NZ) + sigmu * ldf(vw.d, SL::D_T2 + NX);
#pragma unroll
            for (int c = 0; c < NX; ++c) rt += ab.B[c] * pv[c];
            stf(d, SL::D_KAP, -ldf(vw.d, SL::D_RI) * rt);
            double pvn[NZ];
#pragma unroll
            for (int a = 0; a < NZ; ++a) {
                double v = ld(vw.q, a) + (a < NX ? sigmu * ldf(vw.d, SL::D_T2 + a) : 0.0);
                if (a < NX) {
#pragma unroll
                    for (int c = 0; c < NX; ++c) v += ab.A[c * NX + a] * pv[c];
                } else {
                    v += pv[NX];
                }
                v -= ldf(vw.d, SL::D_RL + a) * rt;
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
    template <class Pipe>
    static LB_HD void pass_f2(const P& p, const SL& l, Lane& ln, const double* __restrict__ G, const double* __restrict__ hg,
                              double* w64, FT* wft, bool has_cs, bool rsh, Pipe& pp) {
        const int N = p.N;
        const int nch = (p.ng + CH - 1) / CH;
        const double sigmu = ln.sigmu;
        double dx[NX], ratio = 0.0;
#pragma unroll
        for (int j = 0; j < NX; ++j) dx[j] = 0.0;
        pp.start(PASS_F2);
        pp.issue(0);
        for (int k = 0; k <= N; ++k) {
            const bool atkg = k == p.kg;
            pp.issue((atkg && nch > 0) ? -1 : next_fwd(k, N));
            const auto vw = pp.acquire();
            double* it = w64 + (l.o_it + k * SL::RS_IT) * LS;
            const unsigned rows = C::stage_rows(p, k);
            const bool last = k >= N;
            AB ab;
            load_ab(p, vw.jac, ab);
            double v[NVB], dva[NVB], dv[NVB];
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                v[j] = ld(vw.it, j);
                dva[j] = ldf(vw.d, SL::D_DA + j);
                dv[j] = dx[j];
            }
            v[NX] = last ? 0.0 : ld(vw.it, NX);
            dva[NX] = last ? 0.0 : ldf(vw.d, SL::D_DA + NX);
            dv[NX] = 0.0;
            if (!last) {
                double acc = ldf(vw.d, SL::D_RL + NX) * ln.dth;
#pragma unroll
                for (int c = 0; c < NX; ++c) acc += ldf(vw.d, SL::D_RL + c) * dx[c];
                dv[NX] = ldf(vw.d, SL::D_KAP) - acc;
            }
#pragma unroll
            for (int j = 0; j < NVB; ++j)
                if (j < NX || !last) st(it, SL::F_DV + j, dv[j]);
            double es[NVB];
#pragma unroll
            for (int j = 0; j < NVB; ++j) es[j] = (has_cs && rsh && (j < NX || !last)) ? ld(vw.cs, j) : 0.0;
#pragma unroll
            for (int j = 0; j < NVB; ++j) {
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    const int r = 2 * j + side;
                    const bool act = (rows >> r) & 1u;
                    const double sgn = side == 0 ? 1.0 : -1.0;
                    double S = 1.0, Lm = 1.0;
                    if (act) {
                        S = ld(vw.it, SL::F_S + r);
                        Lm = ld(vw.it, SL::F_LB + r);
                    }
                    const double slack = side == 0 ? p.hi[j] - (v[j] - es[j]) : (v[j] - es[j]) - p.lo[j];
                    const double rp = S - slack, is = lb_rcp(S), w = Lm * is;
                    const double dsa = -rp - sgn * dva[j], dla = -Lm - w * dsa;
                    const double ds = -rp - sgn * dv[j];
                    const double rc = S * Lm + dsa * dla - sigmu;
                    const double dl = (-rc - Lm * ds) * is;
                    ratio = lb_max(ratio, act ? lb_max(-ds * is, -dl * lb_rcp(Lm)) : 0.0);
                }
            }
            if (atkg) {
                double dza[NZ], dz[NZ], z[NZ];
#pragma unroll
                for (int a = 0; a < NX; ++a) {
                    dza[a] = dva[a];
                    dz[a] = dx[a];
                    z[a] = v[a] - es[a];
                }
                dza[NX] = ln.dtha;
                dz[NX] = ln.dth;
                z[NX] = ln.th;
                for (int c = 0; c < nch; ++c) {
                    pp.issue(c + 1 < nch ? -(c + 2) : next_fwd(k, N));
                    const auto pw = pp.acquire();
                    const int i1 = (c + 1) * CH < p.ng ? (c + 1) * CH : p.ng;
                    for (int i = c * CH; i < i1; ++i) {
                        double slack = hg[i], adva = 0.0, adv = 0.0;
#pragma unroll
                        for (int a = 0; a < NZ; ++a) {
                            const double gia = G[a * p.ngp + i];
                            slack -= gia * z[a];
                            adva += gia * dza[a];
                            adv += gia * dz[a];
                        }
                        const double S = ld(pw.sg, i), Lm = ld(pw.lg, i);
                        const double rp = S - slack, is = lb_rcp(S), w = Lm * is;
                        const double dsa = -rp - adva, dla = -Lm - w * dsa;
                        const double ds = -rp - adv;
                        const double rc = S * Lm + dsa * dla - sigmu;
                        const double dl = (-rc - Lm * ds) * is;
                        ratio = lb_max(ratio, lb_max(-ds * is, -dl * lb_rcp(Lm)));
                    }
                }
            }
            if (!last) {
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
    // Fused closed loop (ocpLBMPC.m:10-47, LBMPC_casadi.m:160-223): a lane owns a scenario for a CHUNK of
    // control steps and, between two solves, does everything the reference's loop body does besides the solve — first move
    // to the RK4 plant (DMS_tracking_LMPC_casadi.m:297-304), bounded uniform disturbance, data window (a ring: no O(q) shift),
    // warm-start shift, learned-model rollout with the L2NW oracle for the next QP's offsets — thread-locally, without
    // leaving the kernel.  Scenario state between chunks lives in a per-scenario store (any lane may continue a scenario).
    // ============================================================================================
    static LB_HD void mg_rhs_s(const double* x, double u, double* f) {  // trueModel.m:21-41
        f[0] = -x[1] + 1.0 + 3.0 * (x[0] / 2.0) - (x[0] * x[0] * x[0] / 2.0);
        f[1] = x[0] + 1.0 - x[2] * sqrt(x[1]);
        f[2] = x[3];
        f[3] = -1000.0 * x[2] - 2.0 * sqrt(500.0) * x[3] + 1000.0 * u;
    }
    static LB_HD double uniform_s(unsigned long long seed, unsigned long long scen, unsigned long long step, unsigned long long comp) {
        unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (scen * 0x100000001B3ULL + step * 8ULL + comp + 1ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z = z ^ (z >> 31);
        return (double)(z >> 11) * (1.0 / 9007199254740992.0);
    }
    struct LoopLane {      // scenario bookkeeping of a lane (registers)
        long long sc;      // scenario, -1: none
        long long ticket;  // queue ticket held while waiting for a re-queued scenario, -1: none
        int t, nd, head;   // control step, samples in the window, next ring slot
        bool cont;         // the scenario continues: set up the next QP at the refill point
    };
    // store <-> workspace; store record: nd head t | x[NX] plan[N+1] X[3 q] Y[NX q] (the loop region of the workspace, verbatim)
    static LB_HD int store_len(const SL& l) { return 3 + NX + l.N + 1 + (3 + NX) * l.qwin; }
    static LB_HD void loop_spill(const SL& l, const LoopLane& ll, const double* w64, double* st) {
        st[0] = (double)ll.nd; st[1] = (double)ll.head; st[2] = (double)ll.t;
        const int n = store_len(l) - 3;
        for (int i = 0; i < n; ++i) st[3 + i] = ld(w64, l.o_loop + i);
    }
    static LB_HD void loop_restore(const SL& l, LoopLane& ll, double* w64, const double* st) {
        ll.nd = (int)st[0]; ll.head = (int)st[1]; ll.t = (int)st[2];
        const int n = store_len(l) - 3;
        for (int i = 0; i < n; ++i) st_(w64, l.o_loop + i, st[3 + i]);
    }
    static LB_HD void st_(double* p, int e, double v) { p[e * LS] = v; }
    // a fresh scenario
    template <class LP>
    static LB_HD void loop_begin(const SL& l, const LP& lp, Lane& ln, LoopLane& ll, double* w64, long long sc) {
        ll.sc = sc; ll.t = 0; ll.nd = 0; ll.head = 0;
        (void)ln;
        for (int j = 0; j < NX; ++j) {
            const double v = lp.x_init[sc * NX + j];
            st_(w64, l.lp_x() + j, v);
            if (lp.x_hist) lp.x_hist[(sc * (lp.steps + 1)) * NX + j] = v;
        }
        for (int k = 0; k <= l.N; ++k) st_(w64, l.lp_plan() + k, 0.0);
    }
    // set up the QP of control step ll.t: oracle offsets along the learned rollout of the plan, initial rollout -> iterate record
    // lane state of the QP of control step ll.t (the iterate record is filled by loop_start_step / its warp-cooperative form)
    template <class LP>
    static LB_HD void loop_lane_reset(const P& p, const SL& l, const LP& lp, Lane& ln, const LoopLane& ll, const double* w64) {
        ln.q = ll.sc;
        ln.iters = 0;
        ln.fresh = true;
        ln.alpha = ln.sigmu = ln.dth = ln.dtha = 0.0;
        ln.cconst() = 0.0;
#pragma unroll
        for (int a = 0; a < NZ; ++a) ln.lin(a) = ln.gGl(a) = ln.dG1(a) = ln.dG2(a) = 0.0;
        const bool warm = lp.warm_shift && ll.t > 0;
        ln.th = warm ? ld(w64, l.lp_plan() + p.N) : 0.0;
    }
    template <class LP>
    static LB_HD void loop_start_step(const P& p, const SL& l, const LP& lp, Lane& ln, const LoopLane& ll, double* w64) {
        const int N = p.N;
        loop_lane_reset(p, l, lp, ln, ll, w64);
        const bool warm = lp.warm_shift && ll.t > 0;
        double x[NX], xo[NX];  // x: rollout of the QP's starting point; xo: learned rollout of the plan (oracle arguments)
#pragma unroll
        for (int j = 0; j < NX; ++j) x[j] = xo[j] = ld(w64, l.lp_x() + j) - lp.x_eq[j];
        const bool orc = lp.use_oracle && ll.nd > 0;
        for (int k = 0; k <= N; ++k) {
            double* it = w64 + (l.o_it + k * SL::RS_IT) * LS;
#pragma unroll
            for (int j = 0; j < NX; ++j) st_(it, SL::F_X + j, x[j]);
            if (k == N) break;
            const double pk = ld(w64, l.lp_plan() + k);
            double g[NX];
#pragma unroll
            for (int a = 0; a < NX; ++a) g[a] = 0.0;
            if (orc) {  // g([x1;x2;u]) = sum_i Y_i k_i / (lambda + sum_i k_i), oracleL2NW.m:26-36
                double sk = 0.0;
                for (int i = 0; i < ll.nd; ++i) {
                    const double d0 = ld(w64, l.lp_X() + 3 * i) - xo[0], d1 = ld(w64, l.lp_X() + 3 * i + 1) - xo[1],
                                 d2 = ld(w64, l.lp_X() + 3 * i + 2) - pk;
                    const double kv = exp(-(d0 * d0 + d1 * d1 + d2 * d2) * lp.inv_h2);
                    sk += kv;
#pragma unroll
                    for (int a = 0; a < NX; ++a) g[a] += ld(w64, l.lp_Y() + NX * i + a) * kv;
                }
                const double wn = 1.0 / (lp.lambda + sk);
#pragma unroll
                for (int a = 0; a < NX; ++a) g[a] *= wn;
            }
            const double u = warm ? pk : 0.0;
            st_(it, SL::F_U, u);
            double xn[NX], xon[NX];
#pragma unroll
            for (int a = 0; a < NX; ++a) {
                double v = g[a] + p.B[a] * u, w = g[a] + p.B[a] * pk;
#pragma unroll
                for (int j = 0; j < NX; ++j) {
                    v += p.A[a * NX + j] * x[j];
                    w += p.A[a * NX + j] * xo[j];
                }
                xn[a] = v;
                xon[a] = w;
            }
#pragma unroll
            for (int a = 0; a < NX; ++a) {
                x[a] = xn[a];
                xo[a] = xon[a];
            }
        }
    }
    // the QP of step ll.t has its verdict: apply the first move, advance plant / window / plan / histories
    template <class LP>
    static LB_HD void loop_advance(const P& p, const SL& l, const LP& lp, Lane& ln, LoopLane& ll, double* w64, int status) {
        const int N = p.N;
        const bool ok = status == 0;
        // no optimal solution: keep executing the last optimal plan (same rule as plant_kernel / lbo_closed_loop)
        const double du0 = ok ? ld(w64 + (l.o_it) * LS, SL::F_U) : ld(w64, l.lp_plan());
        double x[NX], dx[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            x[j] = ld(w64, l.lp_x() + j);
            dx[j] = x[j] - lp.x_eq[j];
        }
        const double u0 = du0 + lp.u_eq, delta = 0.01;
        double k1[NX], k2[NX], k3[NX], k4[NX], tt[NX], xn[NX];
        mg_rhs_s(x, u0, k1);
#pragma unroll
        for (int i = 0; i < NX; ++i) tt[i] = x[i] + delta / 2 * k1[i];
        mg_rhs_s(tt, u0, k2);
#pragma unroll
        for (int i = 0; i < NX; ++i) tt[i] = x[i] + delta / 2 * k2[i];
        mg_rhs_s(tt, u0, k3);
#pragma unroll
        for (int i = 0; i < NX; ++i) tt[i] = x[i] + delta * k3[i];
        mg_rhs_s(tt, u0, k4);
#pragma unroll
        for (int i = 0; i < NX; ++i) xn[i] = x[i] + delta / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
        if (lp.use_w) {
#pragma unroll
            for (int j = 0; j < NX; ++j)
                xn[j] += lp.wbar[j] * (2.0 * uniform_s(lp.seed, lp.scen0 + (unsigned long long)ll.sc, (unsigned long long)ll.t, (unsigned long long)j) - 1.0);
        }
        // data acquisition (ocpLBMPC.m:14-15): X = [dx1;dx2;du], Y = dx+ - (A dx + B du); ring slot `head`
        {
            const int slot = ll.head;
            st_(w64, l.lp_X() + 3 * slot, dx[0]);
            st_(w64, l.lp_X() + 3 * slot + 1, dx[1]);
            st_(w64, l.lp_X() + 3 * slot + 2, du0);
#pragma unroll
            for (int a = 0; a < NX; ++a) {
                double v = xn[a] - lp.x_eq[a] - p.B[a] * du0;
#pragma unroll
                for (int j = 0; j < NX; ++j) v -= p.A[a * NX + j] * dx[j];
                st_(w64, l.lp_Y() + NX * slot + a, v);
            }
            ll.head = slot + 1 == l.qwin ? 0 : slot + 1;
            ll.nd = ll.nd < l.qwin ? ll.nd + 1 : ll.nd;
        }
        // plan <- shift of (solution | previous plan), last input repeated
        if (ok) {
            for (int k = 0; k + 1 < N; ++k) st_(w64, l.lp_plan() + k, ld(w64 + (l.o_it + (k + 1) * SL::RS_IT) * LS, SL::F_U));
            st_(w64, l.lp_plan() + N - 1, ld(w64 + (l.o_it + (N - 1) * SL::RS_IT) * LS, SL::F_U));
            st_(w64, l.lp_plan() + N, ln.th);
        } else {
            for (int k = 0; k + 1 < N; ++k) st_(w64, l.lp_plan() + k, ld(w64, l.lp_plan() + k + 1));
        }
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            st_(w64, l.lp_x() + j, xn[j]);
            if (lp.x_hist) lp.x_hist[(ll.sc * (lp.steps + 1) + ll.t + 1) * NX + j] = xn[j];
        }
        if (lp.u_hist) lp.u_hist[ll.sc * lp.steps + ll.t] = u0;
        if (lp.theta_hist) lp.theta_hist[ll.sc * lp.steps + ll.t] = ld(w64, l.lp_plan() + N);
        if (lp.iters_hist) lp.iters_hist[ll.sc * lp.steps + ll.t] = ln.iters;
        if (lp.status_hist) lp.status_hist[ll.sc * lp.steps + ll.t] = status;
        ll.t += 1;
    }

    // ============================================================================================
    // results of a finished QP: c = u - Kout x (transitionNominal.m:12 undone), theta, objective
    // ============================================================================================
    static LB_HD void finish(const P& p, const SL& l, const Lane& ln, const StreamIO<FT>& io, const double* w64, int status) {
        const int N = p.N;
        const long long q = ln.q;
        // the iterate comes back from the workspace in groups of KB stages, every load of a group issued before the first
        // use: one stage at a time (5 dependent-latency loads, then the stores) cost ~2.5 k cycles per stage under a loaded
        // HBM — 10 % of the kernel's warp time with the other 31 lanes of the warp waiting
        constexpr int KB = 10;
        for (int k0 = 0; k0 <= N; k0 += KB) {
            double xs[KB][NX + 1];
#pragma unroll
            for (int i = 0; i < KB; ++i) {
                const int k = k0 + i <= N ? k0 + i : N;
                const double* it = w64 + (l.o_it + k * SL::RS_IT) * LS;
#pragma unroll
                for (int j = 0; j <= NX; ++j) xs[i][j] = ld(it, j);
            }
#pragma unroll
            for (int i = 0; i < KB; ++i) {
                const int k = k0 + i;
                if (k > N) break;
                if (io.xtraj) {
#pragma unroll
                    for (int j = 0; j < NX; ++j) io.xtraj[(q * (N + 1) + k) * NX + j] = xs[i][j];
                }
                if (k < N) {
                    double v = xs[i][NX];
#pragma unroll
                    for (int j = 0; j < NX; ++j) v -= p.Kout[j] * xs[i][j];
                    io.uc[q * N + k] = v;
                }
            }
        }
        io.theta[q] = ln.th;
        io.obj[q] = ln.obj();
        io.iters[q] = ln.iters;
        io.status[q] = status;
    }

    // one whole scenario by one lane, chunk by chunk through the store (tests/emul; the kernel adds the ticket queue)
    static LB_HD void loop_one(const P& p, const SL& l, const StreamLoopParams& lp, long long sc, const double* G, const double* hg,
                               double* w64, FT* wft, double* store_rec) {
        Lane ln;
        double coldbuf[Lane::C_N * LS];
        ln.cold = coldbuf;
        LoopLane ll;
        StreamPipeDirect<NX, FT, LS> pp(l, w64, wft);
        loop_begin(l, lp, ln, ll, w64, sc);
        while (ll.t < lp.steps) {
            loop_start_step(p, l, lp, ln, ll, w64);
            for (;;) {
                int st = pass_bu(p, l, ln, G, hg, w64, wft, false, false, pp);
                if (st < 0 && ln.iters >= p.max_iter) st = 1;
                if (st >= 0) {
                    loop_advance(p, l, lp, ln, ll, w64, st);
                    break;
                }
                pass_f1(p, l, ln, G, hg, w64, wft, false, false, pp);
                pass_b2(p, l, ln, wft, pp);
                pass_f2(p, l, ln, G, hg, w64, wft, false, false, pp);
                ln.iters += 1;
                ln.fresh = false;
            }
            if (ll.t < lp.steps && ll.t % lp.chunk == 0) {  // chunk boundary: the state travels through the store
                loop_spill(l, ll, w64, store_rec);
                for (int i = 0; i < store_len(l) - 3; ++i) st_(w64, l.o_loop + i, -7.0);
                ll.nd = ll.head = ll.t = -1;
                loop_restore(l, ll, w64, store_rec);
            }
        }
    }

    // one whole solve by one lane (tests/emul; the kernel interleaves the same calls with the queue logic)
    static LB_HD void solve_one(const P& p, const SL& l, const StreamIO<FT>& io, long long q, const double* G, const double* hg,
                                double* w64, FT* wft) {
        Lane ln;
        double coldbuf[Lane::C_N * LS];
        ln.cold = coldbuf;
        const bool has_cs = io.cshift != nullptr, rsh = LTV && io.row_shift != 0;  // row shifts come with LTV dynamics only (first-order SQP)
        StreamPipeDirect<NX, FT, LS> pp(l, w64, wft);
        init_qp(p, l, ln, io, q, w64, has_cs);
        for (;;) {
            int st = pass_bu(p, l, ln, G, hg, w64, wft, has_cs, rsh, pp);
            if (st < 0 && ln.iters >= p.max_iter) st = 1;
            if (st >= 0) {
                finish(p, l, ln, io, w64, st);
                return;
            }
            pass_f1(p, l, ln, G, hg, w64, wft, has_cs, rsh, pp);
            pass_b2(p, l, ln, wft, pp);
            pass_f2(p, l, ln, G, hg, w64, wft, has_cs, rsh, pp);
            ln.iters += 1;
            ln.fresh = false;
        }
    }
};

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// TMA pipe: the records a pass reads are fetched one pipeline item (a stage, or a chunk of polytope rows) AHEAD of the
// arithmetic with 1-D bulk copies (cp.async.bulk -> SASS UBLKCP) into a double buffer in shared memory private to the
// warp, completion on an mbarrier per buffer.  One elected lane issues the copies of an item (<= 4 contiguous ranges,
// 4.6 - 13 KB); every lane then reads its own column of the buffer.  This takes the HBM latency (~600 cycles per
// dependent access, and an in-order warp would pay it ~10 times per stage) off the critical path without spending
// registers on prefetched operands.
// ---------------------------------------------------------------------------------------------
template <int NX, typename FT>
struct StreamSmem {  // byte offsets inside one buffer
    using SL = StreamLayout<NX>;
    static constexpr int kLine = 32 * (int)sizeof(double), kLineF = 32 * (int)sizeof(FT);
    static constexpr int off_it = 0, off_d = SL::RS_IT * kLine;                      // B2 has no iterate record: its range starts at 0
    static constexpr int off_cs = off_d + (SL::D_T2 - SL::D_DA) * kLineF;             // after the widest range next to an iterate record (F2)
    static constexpr int off_q = (SL::RS_D - SL::D_RL) * kLineF;                      // B2: after its FT range
    static_assert(off_q + SL::RS_Q * kLine <= off_cs, "layout");
    static __host__ __device__ int off_j(bool cs) { return off_cs + (cs ? (NX + 1) * kLine : 0); }
    static __host__ __device__ int buf_bytes(bool cs, bool ltv) {
        int b = off_j(cs) + (ltv ? SL::NJ * kLine : 0);
        const int b2 = off_q + SL::RS_Q * kLine, poly = 2 * SL::kPolyChunk * kLine;
        b = b > b2 ? b : b2;
        b = b > poly ? b : poly;
        return (b + 127) & ~127;
    }
    static constexpr int kPark = StreamLane<NX, 32>::C_N * kLine;  // parked per-lane values (StreamLane::cold), one line per value
    static __host__ __device__ size_t warp_bytes(bool cs, bool ltv) { return 2 * (size_t)buf_bytes(cs, ltv) + 128 + kPark; }  // + two mbarriers + park
    // small polytope blocks (the 24-row robust set) are staged once per CTA behind the warps' regions: the row loops of three
    // passes read G and hg row by row, and with the L1 thrashed by the streamed records every one of those loads went to L2
    static __host__ __device__ size_t poly_bytes(int ngp) { return ngp <= 64 ? (size_t)(NX + 2) * ngp * sizeof(double) : 0; }
};

template <int NX, bool LTV, typename FT>
struct StreamPipeTma {
    using SL = StreamLayout<NX>;
    using SM = StreamSmem<NX, FT>;
    using ND = StreamNeeds<NX>;
    const SL* l;
    const double* w64g;   // warp base (lane 0) of the FP64 workspace
    const FT* wftg;
    unsigned char* buf;   // two buffers
    uint64_t* bar;        // two mbarriers
    int lane, pass, bufb, ring0, ring1;  // (two scalars: a dynamically indexed array would pin the whole struct to local memory)
    unsigned n_issue, n_wait;
    bool has_cs, rows;
    __device__ void start(int pass_) { pass = pass_; }
    __device__ void issue(int item) {
        if (item == kStreamNone) return;
        const unsigned b = n_issue & 1u;
        if (b) ring1 = item; else ring0 = item;
        ++n_issue;
        __syncwarp();  // every lane is done reading buffer b (item n_issue - 2) and has fenced its global stores
        if (lane == 0) {
            unsigned char* dst = buf + b * bufb;
            uint64_t* br = bar + b;
            if (item >= 0) {
                const int k = item, dlo = ND::d_lo(pass), dn = ND::d_hi(pass) - dlo;
                const int it = ND::it(pass);
                const bool cs = has_cs && ND::cs(pass, rows), jj = LTV && k < l->N;
                const bool qq = pass == PASS_B2;
                const uint32_t bytes = it * SM::kLine + dn * SM::kLineF + (cs ? (NX + 1) * SM::kLine : 0) + (jj ? SL::NJ * SM::kLine : 0) +
                                       (qq ? SL::RS_Q * SM::kLine : 0);
                mbar_expect_tx(br, bytes);
                if (qq) tma_bulk_g2s(dst + SM::off_q, w64g + (size_t)(l->o_q + k * SL::RS_Q) * 32, SL::RS_Q * SM::kLine, br);
                if (it) tma_bulk_g2s(dst + SM::off_it, w64g + (size_t)(l->o_it + k * SL::RS_IT) * 32, it * SM::kLine, br);
                tma_bulk_g2s(dst + (it ? SM::off_d : 0), wftg + (size_t)(k * SL::RS_D + dlo) * 32, dn * SM::kLineF, br);
                if (cs) tma_bulk_g2s(dst + SM::off_cs, w64g + (size_t)(l->o_cs + k * (NX + 1)) * 32, (NX + 1) * SM::kLine, br);
                if (jj) tma_bulk_g2s(dst + SM::off_j(has_cs), w64g + (size_t)(l->o_j + k * SL::NJ) * 32, SL::NJ * SM::kLine, br);
            } else {
                const int c = -(item + 1), i0 = c * SL::kPolyChunk;
                const int n = l->ng - i0 < SL::kPolyChunk ? l->ng - i0 : SL::kPolyChunk;
                mbar_expect_tx(br, 2u * n * SM::kLine);
                tma_bulk_g2s(dst, w64g + (size_t)(l->o_sg + i0) * 32, n * SM::kLine, br);
                tma_bulk_g2s(dst + SL::kPolyChunk * SM::kLine, w64g + (size_t)(l->o_lg + i0) * 32, n * SM::kLine, br);
            }
        }
    }
    using View = StreamView<StreamSh64, StreamShFT<FT>>;
    __device__ View acquire() {
        const unsigned b = n_wait & 1u, parity = (n_wait >> 1) & 1u;
        ++n_wait;
        mbar_wait(bar + b, parity);
        const int item = b ? ring1 : ring0;
        const unsigned src = smem_u32(buf + b * bufb) + 8u * (unsigned)lane;  // this lane's column (FP64 lines)
        View v;
        if (item >= 0) {
            const bool it = ND::it(pass) != 0;
            v.it.a = src + SM::off_it;
            v.d.a = smem_u32(buf + b * bufb) + (unsigned)sizeof(FT) * (unsigned)lane + (it ? SM::off_d : 0) - (unsigned)(ND::d_lo(pass) * SM::kLineF);
            v.q.a = src + SM::off_q;
            v.cs.a = src + SM::off_cs;
            v.jac.a = src + SM::off_j(has_cs);
            v.sg = v.lg = v.it;
        } else {
            const int i0 = -(item + 1) * SL::kPolyChunk;
            v.sg.a = src - (unsigned)(i0 * SM::kLine);
            v.lg.a = src + SL::kPolyChunk * SM::kLine - (unsigned)(i0 * SM::kLine);
            v.it = v.q = v.cs = v.jac = v.sg;
            v.d.a = src;
        }
        return v;
    }
};

// Warp-cooperative set-up of the next QP of ONE lane of the warp (fused closed loop).  The learned-model rollout evaluates
// the L2NW oracle N times over the q-sample window; done by the owning lane alone it is N q exponentials executed with 31
// idle lanes, and with ~4 of 32 lanes finishing a QP per iteration that would cost as much as the iteration itself.  Here
// the 32 lanes take the data points (like oracle_kernel) and reduce with shuffles; every lane carries the rollout, lane
// `b`'s workspace column receives the iterate.  Same arithmetic as Stream::loop_start_step up to the order of the sums.
template <int NX, typename S, typename LP>
__device__ __forceinline__ void stream_loop_start_step_coop(const Params<NX, 1, 1>& p, const StreamLayout<NX>& l, const LP& lp, int b,
                                                            int lane, bool warm, int nd, double* w64_lane0) {
    using SL = StreamLayout<NX>;
    const int N = p.N;
    double* const wb = w64_lane0 + b;  // lane b's column
    auto ldb = [&](int e) { return wb[e * 32]; };
    double x[NX], xo[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) x[j] = xo[j] = ldb(l.lp_x() + j) - lp.x_eq[j];
    const bool orc = lp.use_oracle && nd > 0;
    constexpr int PER = 16;  // q <= 512
    const int per = (nd + 31) / 32;
    for (int k = 0; k <= N; ++k) {
        double* it = wb + (l.o_it + k * SL::RS_IT) * 32;
        if (lane < NX) {
            double v = x[0];
#pragma unroll
            for (int j = 1; j < NX; ++j) v = lane == j ? x[j] : v;
            it[(SL::F_X + lane) * 32] = v;
        }
        if (k == N) break;
        const double pk = ldb(l.lp_plan() + k);
        double g[NX], sk = 0.0;
#pragma unroll
        for (int a = 0; a < NX; ++a) g[a] = 0.0;
        if (orc) {
            for (int r = 0; r < per && r < PER; ++r) {
                const int i = lane + 32 * r;
                if (i < nd) {
                    const double d0 = ldb(l.lp_X() + 3 * i) - xo[0], d1 = ldb(l.lp_X() + 3 * i + 1) - xo[1],
                                 d2 = ldb(l.lp_X() + 3 * i + 2) - pk;
                    const double kv = exp(-(d0 * d0 + d1 * d1 + d2 * d2) * lp.inv_h2);
                    sk += kv;
#pragma unroll
                    for (int a = 0; a < NX; ++a) g[a] += ldb(l.lp_Y() + NX * i + a) * kv;
                }
            }
            sk = warp_sum(sk);
            const double wn = 1.0 / (lp.lambda + sk);
#pragma unroll
            for (int a = 0; a < NX; ++a) g[a] = warp_sum(g[a]) * wn;
        }
        const double u = warm ? pk : 0.0;
        if (lane == 0) it[SL::F_U * 32] = u;
        double xn[NX], xon[NX];
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            double v = g[a] + p.B[a] * u, w = g[a] + p.B[a] * pk;
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                v += p.A[a * NX + j] * x[j];
                w += p.A[a * NX + j] * xo[j];
            }
            xn[a] = v;
            xon[a] = w;
        }
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            x[a] = xn[a];
            xo[a] = xon[a];
        }
    }
}

// generic-proxy stores to the workspace (this pass) -> async-proxy reads (bulk copies of the next pass)
__device__ __forceinline__ void stream_pass_fence() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// the kernel: persistent lanes, QPs from a global queue (warp-aggregated atomics).  Every lane of a warp runs every pass
// (the bulk-copy pipeline is warp-collective); a lane without a QP computes on whatever its workspace column holds and
// its results are never looked at.
// ---------------------------------------------------------------------------------------------
// Instruction cache: one iteration is ~4.5 k instructions per stage in four loop bodies (72 KB of SASS); warps that sit in
// different passes evict each other's lines (ncu, 8 free-running warps per SM: 1.7 of 6.6 stall cycles per issue are
// instruction fetches and two warps per scheduler issue no more than one).  Every warp executes every pass with the same
// trip counts, so a CTA-wide barrier in front of each pass keeps the warps of an SM on the same lines at almost no cost
// (cta_tick, lbmpc_kernels.cuh: it also counts the warps that still have work, so that the CTA leaves together).
// WARPS per CTA (one CTA per SM): 8 at 255 registers per thread, or 6 when the double buffers of 8 warps do not fit shared
// memory (LTV Jacobians + shift record).  12 warps at 168 registers measured 1.5x slower (spills in pass BU).
// LOOP: fused closed loop — the work items are (scenario, chunk of control steps) tickets; a lane sets up, solves and advances
// the control steps of its chunk back to back and hands the scenario over through the store + the re-queue ring.
template <int NX, bool LTV, typename FT, int WARPS, bool LOOP = false>
__global__ void __launch_bounds__(32 * WARPS, 1)
ipm_stream_kernel(const __grid_constant__ Params<NX, 1, 1> p, const StreamIO<FT> io, const double* G,
                  const double* hg, const StreamLoopParams lp = StreamLoopParams{}) {
    using S = Stream<NX, LTV, FT, 32>;
    using SM = StreamSmem<NX, FT>;
    extern __shared__ __align__(128) unsigned char stream_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    // row shifts exist only together with per-stage dynamics (first-order SQP): a compile-time false in the plain kernel, which
    // frees the shift registers of every pass there
    const bool has_cs = io.cshift != nullptr, rsh = LTV && io.row_shift != 0;
    const StreamLayout<NX> l(p.N, p.ng, has_cs, LTV, LOOP ? io.qwin : 0);
    double* const w64 = io.ws64 + warp * (long long)l.n64 * 32 + lane;
    FT* const wft = io.wsft + warp * (long long)l.nft * 32 + lane;
    if (SM::poly_bytes(p.ngp) > 0) {  // stage the polytope block; generic pointers, so the passes do not care where it lives
        double* Gs = reinterpret_cast<double*>(stream_smem + (blockDim.x >> 5) * SM::warp_bytes(has_cs, LTV));
        for (int i = threadIdx.x; i < (NX + 1) * p.ngp; i += blockDim.x) Gs[i] = G[i];
        for (int i = threadIdx.x; i < p.ngp; i += blockDim.x) Gs[(NX + 1) * p.ngp + i] = i < p.ng ? hg[i] : 0.0;
        __syncthreads();
        G = Gs;
        hg = Gs + (NX + 1) * p.ngp;
    }
    StreamPipeTma<NX, LTV, FT> pp;
    pp.l = &l;
    pp.w64g = w64 - lane;
    pp.wftg = wft - lane;
    pp.bufb = SM::buf_bytes(has_cs, LTV);
    pp.buf = stream_smem + wib * SM::warp_bytes(has_cs, LTV);
    pp.bar = reinterpret_cast<uint64_t*>(pp.buf + 2 * pp.bufb);
    pp.lane = lane;
    pp.pass = 0;
    pp.n_issue = pp.n_wait = 0;
    pp.has_cs = has_cs;
    pp.rows = rsh;
    if (lane == 0) {
        mbar_init(pp.bar, 1);
        mbar_init(pp.bar + 1, 1);
    }
    __syncwarp();
    typename S::Lane ln;
    ln.cold = reinterpret_cast<double*>(pp.buf + 2 * pp.bufb + 128) + lane;
    ln.q = -1;
    ln.iters = 0;
    ln.fresh = true;
    ln.th = ln.dth = ln.dtha = ln.alpha = ln.sigmu = ln.iptt = ln.mu = 0.0;
    bool drained = false;  // warp-uniform: the queue has run dry
    typename S::LoopLane ll;
    ll.sc = ll.ticket = -1;
    ll.t = ll.nd = ll.head = 0;
    ll.cont = false;
    bool lane_done = false;  // LOOP: this lane's ticket was beyond the last chunk
    const int nchunks = LOOP ? (lp.steps + lp.chunk - 1) / lp.chunk : 0;
    const long long n_tickets = LOOP ? lp.nscen * nchunks : 0;
    const int slen = LOOP ? S::store_len(l) : 0;
    for (;;) {
        bool working;
        if constexpr (LOOP) {
            const bool need = ln.q < 0 && !ll.cont && ll.ticket < 0 && !lane_done;
            const unsigned want = __ballot_sync(0xffffffffu, need);
            if (want) {
                unsigned long long base = 0;
                const int leader = __ffs(want) - 1;
                if (lane == leader) base = atomicAdd(io.queue, (unsigned long long)__popc(want));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (need) {
                    ll.ticket = (long long)base + __popc(want & ((1u << lane) - 1u));
                    if (ll.ticket >= n_tickets) {
                        lane_done = true;
                        ll.ticket = -1;
                    }
                }
            }
            if (ln.q < 0 && !ll.cont && ll.ticket >= 0) {
                if (ll.ticket < lp.nscen) {  // first chunk of a scenario
                    S::loop_begin(l, lp, ln, ll, w64, ll.ticket);
                    ll.cont = true;
                    ll.ticket = -1;
                } else {                     // a re-queued scenario: wait (without blocking the warp) until it has been produced
                    const long long e = *reinterpret_cast<volatile long long*>(lp.rq + (ll.ticket - lp.nscen));
                    if (e != 0) {
                        __threadfence();
                        ll.sc = e - 1;
                        S::loop_restore(l, ll, w64, lp.store + ll.sc * (long long)slen);
                        ll.cont = true;
                        ll.ticket = -1;
                    }
                }
            }
            {   // set up the next QP of every lane that continues its scenario: one lane at a time, the whole warp on its rollout
                unsigned todo = __ballot_sync(0xffffffffu, ln.q < 0 && ll.cont);
                while (todo) {
                    const int b = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int t_b = __shfl_sync(0xffffffffu, ll.t, b), nd_b = __shfl_sync(0xffffffffu, ll.nd, b);
                    stream_loop_start_step_coop<NX, S>(p, l, lp, b, lane, lp.warm_shift && t_b > 0, nd_b, w64 - lane);
                }
                __syncwarp();
                if (ln.q < 0 && ll.cont) {
                    S::loop_lane_reset(p, l, lp, ln, ll, w64);
                    ll.cont = false;
                }
            }
            stream_pass_fence();
            working = !__all_sync(0xffffffffu, ln.q < 0);
            const bool alive = !__all_sync(0xffffffffu, lane_done && ln.q < 0);
            if (cta_tick(alive) == 0) break;
        } else {
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
                stream_pass_fence();
            }
            working = !__all_sync(0xffffffffu, ln.q < 0);
            if (cta_tick(working) == 0) break;  // every warp of the CTA is out of work
        }
        if (working) {
            int st = S::pass_bu(p, l, ln, G, hg, w64, wft, has_cs, rsh, pp);
            stream_pass_fence();
            if (ln.q >= 0) {
                if (st < 0 && ln.iters >= p.max_iter) st = 1;
                if (st >= 0) {
                    if constexpr (LOOP) {
                        S::loop_advance(p, l, lp, ln, ll, w64, st);
                        if (ll.t < lp.steps && ll.t % lp.chunk == 0) {  // chunk done: hand the scenario over
                            S::loop_spill(l, ll, w64, lp.store + ll.sc * (long long)slen);
                            __threadfence();
                            const unsigned long long idx = atomicAdd(lp.rq_tail, 1ULL);
                            *reinterpret_cast<volatile long long*>(lp.rq + idx) = ll.sc + 1;
                        } else if (ll.t < lp.steps) {
                            ll.cont = true;
                        }
                    } else {
                        S::finish(p, l, ln, io, w64, st);
                    }
                    ln.q = -1;
                } else if (!LOOP && io.evict_iters > 0 && ln.iters >= io.evict_iters) {
                    io.left_list[atomicAdd(io.left_count, 1ULL)] = ln.q;
                    ln.q = -1;
                }
            }
            working = !__all_sync(0xffffffffu, ln.q < 0);  // nothing left to iterate on in this warp: refill or leave
        }
        __syncthreads();
        if (working) {
            S::pass_f1(p, l, ln, G, hg, w64, wft, has_cs, rsh, pp);
            stream_pass_fence();
        }
        __syncthreads();
        if (working) {
            S::pass_b2(p, l, ln, wft, pp);
            stream_pass_fence();
        }
        __syncthreads();
        if (working) {
            S::pass_f2(p, l, ln, G, hg, w64, wft, has_cs, rsh, pp);
            stream_pass_fence();
            if (ln.q >= 0) {
                ln.iters += 1;
                ln.fresh = false;
            }
        }
    }
}
#endif

}  // namespace lbmpc
