// emul_core.cpp — TEST-ONLY host harness for learning-based-mpc_b200/csrc/lbmpc_core.cuh.
// Runs the per-QP device functions (compiled for the host) in the same phase order the CUDA
// kernel uses, one QP at a time, so the arithmetic of every phase can be checked against the
// oracle on a machine without a GPU.  It is never linked into the product library.
#include <cstring>
#include <string>
#include <vector>

#include "../../learning-based-mpc_b200/csrc/lbmpc_problem.hpp"

using namespace lbmpc;

// Lane order of the emulated warp phases.  On the GPU the lanes of a phase run concurrently and only __syncwarp() / shuffles
// separate the phases, so a phase must not depend on the order in which its lanes execute.  The harness can run every lane
// loop backwards (emul_set_lane_order(1)): bit-identical results forwards and backwards are the CPU-side stand-in for
// compute-sanitizer's racecheck (closed on this GPU pool) — a read-after-write between lanes inside one phase would show.
static int g_reverse_lanes = 0;
extern "C" void emul_set_lane_order(int reverse) { g_reverse_lanes = reverse; }
#define LANES(var, n) for (int var##_i = 0, var = g_reverse_lanes ? (n) - 1 : 0; var##_i < (n); ++var##_i, var += g_reverse_lanes ? -1 : 1)

template <int NX, int NT, int NU>
static int solve_one(const HostProblem& hp, const double* dx0, const double* dx_ref, const double* d_off, const CShift csh,
                     const double* warm, double* uc, double* theta, double* xtraj, double* obj, int* iters,
                     int* status) {
    using C = Core<NX, NT, NU>;
    using L = Layout<NX, NT, NU>;
    constexpr int NZ = NX + NT, NH = L::NH;
    const Params<NX, NT, NU> p = to_params<NX, NT, NU>(hp);
    const L l(p.N, p.ngp);
    std::vector<double> buf(l.stride, 0.0);
    double* s = buf.data();
    double* m = s + l.o_misc;
    const double *G = hp.G.data(), *hg = hp.hg.data();
    const int N = p.N;
    // ---- load (kernel: warp-parallel) ----
    for (int j = 0; j < NX; ++j) s[l.i_x(j, 0)] = dx0[j];
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < NU; ++i) s[l.i_u(i, k)] = warm ? warm[k * NU + i] : 0.0;
        for (int j = 0; j < NX; ++j) s[l.i_x(j, k + 1)] = d_off ? d_off[k * NX + j] : 0.0;
    }
    for (int t = 0; t < NT; ++t) m[L::M_TH + t] = warm ? warm[N * NU + t] : 0.0;
    double cconst = 0.0;
    for (int a = 0; a < NZ; ++a) {
        double v = 0.0;
        if (dx_ref)
            for (int j = 0; j < NX; ++j) v += p.Lref[a * NX + j] * dx_ref[j];
        m[L::M_LIN + a] = v;
    }
    if (dx_ref)
        for (int i = 0; i < NX; ++i)
            for (int j = 0; j < NX; ++j) cconst += dx_ref[i] * p.Tm[i * NX + j] * dx_ref[j];
    m[L::M_CCONST] = cconst;
    C::rollout(p, l, s);
    for (int i = 0; i < p.ng; ++i) C::init_rows_gen(p, l, s, G, hg, i);
    double alpha = 0.0;
    std::vector<double> zero_rec(L::RS2, 0.0);
    // blocked substitution sweeps: lane-phases as loops (kernel: lanes of the QP's warp, __syncwarp between phases)
    auto solve_sweeps = [&](bool aff, bool with_T) {
        const int ntask = l.nb * (with_T ? NX + 1 : 1);
        LANES(t, ntask) C::bwd_p1(p, l, s, zero_rec.data(), t, with_T);
        C::bwd_p2(l, s, aff);
        std::vector<double> pin((size_t)l.nb * NZ);
        LANES(b, l.nb) C::bwd_p3_in(p, l, s, b, pin.data() + (size_t)b * NZ);
        LANES(b, l.nb) C::bwd_p3_fwd_p1(p, l, s, b, pin.data() + (size_t)b * NZ, aff);
        C::fwd_p2(l, s);
        LANES(b, l.nb) C::fwd_p3(p, l, s, b, aff);
    };
    int it = 0, st = 1;
    for (it = 0;; ++it) {
        // ---- phase A: predictor assembly (kernel: warp per QP, lanes over stages / rows) ----
        RedAsm ra{0, 0, 0, 0, 0};
        LANES(k, N + 1) {
            if (it == 0) C::init_assemble_stage(p, l, s, k, ra, csh);
            else C::update_assemble_stage(p, l, s, k, alpha, ra, csh);
        }
        double acc[NH + 2 * NZ];
        std::memset(acc, 0, sizeof acc);
        LANES(i, p.ng) C::assemble_gen_row(p, l, s, G, hg, i, acc, ra);
        for (int a = 0; a < NH; ++a) m[L::M_HG + a] = acc[a];
        for (int a = 0; a < NZ; ++a) { m[L::M_GGL + a] = acc[NH + a]; m[L::M_DG + a] = acc[NH + NZ + a]; }
        m[L::M_RP] = ra.rp; m[L::M_MU] = ra.sl * p.inv_m; m[L::M_LAM] = ra.lam; m[L::M_HLAM] = ra.hl;
        m[L::M_GTH] = ra.gth + acc[NH + NX];
        // ---- phase B: factorisation + adjoint (+ Farkas) recursions ----
        const bool cert = ra.lam >= p.inf_trigger;
        if constexpr (NT == 1 && NU == 1 && NX <= 4) {
            // 16-lane cooperative factorisation, lane-phases run as loops (kernel: __syncwarp between them)
            using CP = Coop<NX>;
            static thread_local typename CP::Lane ln[32];
            static thread_local double xch[CP::kXch];
            LANES(h, 32) { CP::lane_init(p, h, xch, ln[h]); CP::xch_init(h, xch); CP::begin(p, l, s, zero_rec.data(), ln[h]); }
            int type = C::stage_type(p, N);
            for (int k = N - 1; k >= 0; --k) {
                const int t = C::stage_type(p, k);
                if (t != type) { type = t; LANES(h, 32) CP::load_type(p, t, ln[h]); }
                double* rec = s + l.r2(k);
                if (k & 1) {
                    LANES(h, 32) CP::template st1<1>(ln[h], rec, k == p.kg);
                    LANES(h, 32) CP::template st2<1>(ln[h]);
                } else {
                    LANES(h, 32) CP::template st1<0>(ln[h], rec, k == p.kg);
                    LANES(h, 32) CP::template st2<0>(ln[h]);
                }
                double d1[32];
                for (int h = 0; h < 32; ++h) d1[h] = ln[h].d1;
                LANES(h, 32)          // kernel: three warp shuffles
                    CP::st3(ln[h], d1[ln[h].srcA], d1[ln[h].srcB], d1[CP::kFu]);
            }
            bool ok = true;
            double fin[32], rdm = 0.0;
            LANES(h, 32) fin[h] = CP::finish(ln[h]);
            for (int k = 0; k < N; ++k) CP::check_stage(l, s, k, ok, rdm);
            const double ptt = fin[NH - 1], iptt = 1.0 / ptt;
            m[L::M_PIV] = (ok && ptt > 0.0) ? 1.0 : 0.0;
            m[L::M_PTT] = iptt;
            m[L::M_DTHA] = -iptt * fin[CP::kPv + NX];
            m[L::M_RD] = lb_nanmax(rdm, lb_abs(m[L::M_GTH]));
        } else {
            C::factor_serial(p, l, s);
            C::adjoint_sweep(p, l, s, false);
        }
        if (cert) {
            if constexpr (NT == 1 && NU == 1 && NX <= 4) {   // blocked Farkas recursion (kernel: lanes = blocks)
                LANES(b, l.nb) C::farkas_p1(p, l, s, b);
                C::farkas_p2(p, l, s);
                double nrm = 0.0, yd = 0.0;
                for (int b = 0; b < l.nb; ++b) {
                    double n1 = 0.0, y1 = 0.0;
                    C::farkas_p3(p, l, s, b, n1, y1);
                    nrm += n1;
                    yd += y1;
                }
                m[L::M_CERT] = nrm;
                m[L::M_HLAM] += yd;
            } else {
                C::adjoint_sweep(p, l, s, true);
            }
        }
        // ---- phase B2: verdict, affine substitution sweeps ----
        const int v = C::verdict(p, m, cert);
        if (v >= 0) { st = v; break; }
        if (it >= p.max_iter) break;                     // the last allowed iterate has been tested: status 1
        if constexpr (NT == 1 && NU == 1 && NX <= 4) {   // affine backward substitution rode on the factor sweep
            for (int t = 0; t < l.nb * NX; ++t) C::bwd_p1_T(p, l, s, zero_rec.data(), t);
            LANES(b, l.nb) C::fwd_p1(p, l, s, b, true);
            C::fwd_p2(l, s);
            LANES(b, l.nb) C::fwd_p3(p, l, s, b, true);
        } else {
            solve_sweeps(true, true);
        }
        // ---- phase C: affine step length, sigma, corrector rhs ----
        RedStep rs{0, 0, 0, 0};
        LANES(k, N + 1) C::affine_stage(p, l, s, k, rs);
        double dg[2 * NZ];
        std::memset(dg, 0, sizeof dg);
        LANES(i, p.ng) C::affine_gen_row(p, l, s, G, hg, i, dg, rs);
        const double aaff = rs.ratio > 1.0 ? 1.0 / rs.ratio : 1.0;
        const double mu_aff = (rs.s0 + aaff * rs.s1 + aaff * aaff * rs.s2) * p.inv_m;
        const double sr = mu_aff / m[L::M_MU];
        const double sigmu = lb_max(sr * sr * m[L::M_MU], 0.1 * p.tol_mu);
        m[L::M_SIGMU] = sigmu;
        LANES(k, N + 1) C::corr_stage(p, l, s, k, sigmu);
        for (int a = 0; a < NZ; ++a) m[L::M_DG + a] = dg[a] + sigmu * dg[NZ + a];
        // ---- phase D: corrector substitution sweeps ----
        solve_sweeps(false, false);
        // ---- phase E: step length, update ----
        double ratio = 0.0;
        LANES(k, N + 1) ratio = lb_max(ratio, C::final_stage(p, l, s, k, sigmu));
        LANES(i, p.ng) ratio = lb_max(ratio, C::final_gen_row(p, l, s, G, hg, i, sigmu));
        alpha = ratio > 0.0 ? 0.99 / ratio : 1.0;
        if (alpha > 1.0) alpha = 1.0;
        LANES(i, p.ng) C::update_gen_row(p, l, s, G, hg, i, sigmu, alpha);
        for (int t = 0; t < NT; ++t) m[L::M_TH + t] += alpha * m[L::M_DTH + t];
        // the box rows / x / u step is applied at the top of the next pass (update_assemble_stage)
    }
    double J = m[L::M_CCONST];
    for (int k = 0; k <= N; ++k) J += C::objective_stage(p, l, s, k, csh);
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < NU; ++i) {
            double v = s[l.i_u(i, k)];
            for (int j = 0; j < NX; ++j) v -= p.Kout[i * NX + j] * s[l.i_x(j, k)];
            uc[k * NU + i] = v;
        }
    for (int t = 0; t < NT; ++t) theta[t] = m[L::M_TH + t];
    if (xtraj)
        for (int k = 0; k <= N; ++k)
            for (int j = 0; j < NX; ++j) xtraj[k * NX + j] = s[l.i_x(j, k)];
    *obj = J; *iters = it; *status = st;
    return 0;
}

static thread_local std::string g_err;

extern "C" const char* emul_last_error() { return g_err.c_str(); }

// same argument conventions as lbmpc_create + lbmpc_solve_batch (column-major, one column per QP)
extern "C" int emul_solve_batch(const lbmpc_model* mdl, const lbmpc_config* cfg, long batch, const double* dx0,
                                const double* dx_ref, const double* d_off, const double* cost_shift, int cs_stride, const double* warm,
                                double* uc, double* theta, double* xtraj, double* obj, int* iters, int* status) {
    HostProblem hp;
    int rc = build_problem(mdl, cfg, hp, g_err);
    if (rc) return rc;
    const int nx = hp.nx, nu = hp.nu, nt = hp.nt, N = hp.N;
    for (long b = 0; b < batch; ++b) {
        const double* x0 = dx0 + b * nx;
        const double* xr = dx_ref ? dx_ref + b * nx : nullptr;
        const double* dk = d_off ? d_off + b * (long)nx * N : nullptr;
        const CShift cs{cost_shift ? cost_shift + b * (long)cs_stride * (N + 1) : nullptr, cs_stride};
        const double* wm = warm ? warm + b * (long)(nu * N + nt) : nullptr;
        double* xt = xtraj ? xtraj + b * (long)nx * (N + 1) : nullptr;
        if (nx == 4 && nt == 1 && nu == 1)
            solve_one<4, 1, 1>(hp, x0, xr, dk, cs, wm, uc + b * (long)nu * N, theta + b * nt, xt, obj + b, iters + b, status + b);
        else if (nx == 2 && nt == 2 && nu == 2)
            solve_one<2, 2, 2>(hp, x0, xr, dk, cs, wm, uc + b * (long)nu * N, theta + b * nt, xt, obj + b, iters + b, status + b);
        else { g_err = "emul: unsupported dims"; return LBMPC_ESHAPE; }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// stream mapping (csrc/lbmpc_stream.cuh): one thread per QP, workspace in "HBM".  The per-lane passes
// compiled for the host with lane stride LS (1, or 32 with the QP in lane `lane`: the device layout).
// mode bit 0: float storage of the direction / factor records (mixed mode); jac: LTV Jacobians or NULL
// ------------------------------------------------------------------------------------------------
#include "../../learning-based-mpc_b200/csrc/lbmpc_stream.cuh"

template <bool LTV, typename FT, int LS>
static void stream_batch(const HostProblem& hp, StreamIO<FT> io, int lane) {
    using S = Stream<4, LTV, FT, LS>;
    const Params<4, 1, 1> p = to_params<4, 1, 1>(hp);
    const StreamLayout<4> l(p.N, p.ng, io.cshift != nullptr, LTV);
    std::vector<double> w64((size_t)l.n64 * LS, std::nan(""));   // uninitialised device memory: must never leak into results
    std::vector<FT> wft((size_t)l.nft * LS, (FT)std::nan(""));
    for (long long q = 0; q < io.batch; ++q)
        S::solve_one(p, l, io, q, hp.G.data(), hp.hg.data(), w64.data() + lane, wft.data() + lane);
}

extern "C" int emul_stream_solve_batch(const lbmpc_model* mdl, const lbmpc_config* cfg, long batch, int mode, int lane_stride,
                                       const double* dx0, const double* dx_ref, const double* d_off, const double* cost_shift,
                                       int cs_stride, int row_shift, const double* jac, const double* warm, double* uc, double* theta, double* xtraj,
                                       double* obj, int* iters, int* status) {
    HostProblem hp;
    int rc = build_problem(mdl, cfg, hp, g_err);
    if (rc) return rc;
    if (!(hp.nx == 4 && hp.nt == 1 && hp.nu == 1)) { g_err = "emul stream: (4,1,1) only"; return LBMPC_ESHAPE; }
    auto run = [&](auto ft, auto ltv) {
        using FT = decltype(ft);
        StreamIO<FT> io{};
        io.batch = batch; io.dx0 = dx0; io.dx_ref = dx_ref; io.d_off = d_off; io.warm = warm; io.cshift = cost_shift; io.cs_stride = cost_shift ? cs_stride : 0; io.row_shift = row_shift; io.jac = jac;
        io.uc = uc; io.theta = theta; io.xtraj = xtraj; io.obj = obj; io.iters = iters; io.status = status;
        if (lane_stride == 32) stream_batch<decltype(ltv)::value, FT, 32>(hp, io, 13);
        else stream_batch<decltype(ltv)::value, FT, 1>(hp, io, 0);
    };
    const bool f32 = mode & 1;
    if (jac) { if (f32) run(float(), std::true_type()); else run(double(), std::true_type()); }
    else     { if (f32) run(float(), std::false_type()); else run(double(), std::false_type()); }
    return 0;
}


// fused closed loop of the stream mapping, one scenario at a time (lane stride 1), chunked through the scenario store
extern "C" int emul_stream_closed_loop(const lbmpc_model* mdl, const lbmpc_config* cfg, long nscen, int steps, int q, int chunk,
                                       int use_oracle, int warm_shift, const double* x_eq, double u_eq, const double* x_init,
                                       const double* wbar, unsigned long long seed, unsigned long long scen0, double* x_hist,
                                       double* u_hist, double* theta_hist, int* iters_hist, int* status_hist) {
    HostProblem hp;
    int rc = build_problem(mdl, cfg, hp, g_err);
    if (rc) return rc;
    if (!(hp.nx == 4 && hp.nt == 1 && hp.nu == 1) || hp.form != LBMPC_FORM_C) { g_err = "emul loop: C-form (4,1,1)"; return LBMPC_ESHAPE; }
    using S = Stream<4, false, double, 1>;
    const Params<4, 1, 1> p = to_params<4, 1, 1>(hp);
    const StreamLayout<4> l(p.N, p.ng, false, false, q);
    StreamLoopParams lp{};
    lp.steps = steps; lp.chunk = chunk; lp.warm_shift = warm_shift; lp.use_oracle = use_oracle; lp.use_w = wbar != nullptr;
    lp.nscen = nscen; lp.u_eq = u_eq; lp.inv_h2 = 1.0 / 0.25; lp.lambda = 0.001; lp.seed = seed; lp.scen0 = scen0;
    for (int j = 0; j < 4; ++j) { lp.x_eq[j] = x_eq[j]; lp.wbar[j] = wbar ? wbar[j] : 0.0; }
    lp.x_init = x_init; lp.x_hist = x_hist; lp.u_hist = u_hist; lp.theta_hist = theta_hist; lp.iters_hist = iters_hist;
    lp.status_hist = status_hist;
    std::vector<double> w64((size_t)l.n64, std::nan("")), wft((size_t)l.nft, std::nan("")), rec((size_t)S::store_len(l));
    for (long sc = 0; sc < nscen; ++sc) S::loop_one(p, l, lp, sc, hp.G.data(), hp.hg.data(), w64.data(), wft.data(), rec.data());
    return 0;
}
