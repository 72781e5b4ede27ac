// lbmpc_kernel_cta.cuh — the LATENCY variant of the solver: one CTA of kCtaWarps warps per QP.
//
// ipm_kernel (lbmpc_kernels.cuh) gives every QP one warp, which is the right shape when there are many more
// QPs than warp slots.  At a few QPs per SM (BASELINE configs[1]: 1024 QPs on 148 SMs) a launch lasts as long
// as its slowest QP, i.e. (iterations of that QP) x (latency of one iteration on ONE warp).  Here the row
// phases of an iteration — everything that is independent per (stage, variable) or per polytope row — run
// with one THREAD per item on all warps of the CTA (four FP64 pipes instead of one, short per-thread
// chains), while the sequential sweeps stay on warp 0 exactly as in ipm_kernel (Coop<NX> factorisation with
// the fused affine backward substitution, time-blocked substitution sweeps, blocked Farkas recursion).
// Reductions: warp shuffles + one shared-memory combine; the polytope Hessian / rhs are summed with one
// thread per OUTPUT entry over a row buffer (no cross-thread reduction at all).
// Same arithmetic per row / stage as ipm_kernel up to the order of the sums; same verdict rule.
#pragma once

#include "lbmpc_kernels.cuh"

namespace lbmpc {

constexpr int kCtaMaxWarps = 4;  // CTA-per-QP kernel: warps per QP (template parameter W = 2 or 4)

template <int NX, int NT, int NU>
struct CtaPlan {
    using L = Layout<NX, NT, NU>;
    using C = Core<NX, NT, NU>;
    using CP = Coop<NX>;
    int xch_off, zero_off, ltab_off, rtab_off, rb_off, red_off, g_off, hg_off, ctl_off;  // doubles
    size_t bytes;
    // big: large polytope block (G stays in global memory / L2, polytope sums by per-thread accumulation + block reduction)
    __host__ __device__ CtaPlan(int N, int ngp, bool big) {
        const L l(N, ngp);
        int o = (l.stride + 1) & ~1;
        xch_off = o;   o += (CP::kXch + 1) & ~1;
        zero_off = o;  o += (L::RS2 + 1) & ~1;
        ltab_off = o;  o += (int)((sizeof(typename CP::LaneTab) + 15) / 16) * 2;
        rtab_off = o;  o += (int)((sizeof(typename C::RowTab) + 15) / 16) * 2;
        rb_off = o;    o += big ? 0 : 3 * ngp + (ngp & 1);  // polytope row buffer
        red_off = o;   o += 32 * kCtaMaxWarps;            // per-warp partial reductions
        g_off = o;     o += big ? 0 : (NX + NT) * ngp;     // staged polytope matrix (component-major)
        hg_off = o;    o += big ? 0 : ngp + (ngp & 1);
        ctl_off = o;   o += 4;                            // QP index (as double), spare
        bytes = (size_t)o * sizeof(double);
    }
};

// block-wide reductions of a few values: warp butterflies, lane 0 of every warp to shared memory, everybody combines
template <int W, int NSUM, int NMAX>
__device__ __forceinline__ void block_reduce(double (&sum)[NSUM], double (&mx)[NMAX], double* scratch, int warp, int lane) {
#pragma unroll
    for (int i = 0; i < NSUM; ++i) sum[i] = warp_sum(sum[i]);
#pragma unroll
    for (int i = 0; i < NMAX; ++i) mx[i] = warp_max_nan(mx[i]);
    __syncthreads();  // scratch may still be read from the previous use
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NSUM; ++i) scratch[warp * 32 + i] = sum[i];
#pragma unroll
        for (int i = 0; i < NMAX; ++i) scratch[warp * 32 + NSUM + i] = mx[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NSUM; ++i) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < W; ++w) v += scratch[w * 32 + i];
        sum[i] = v;
    }
#pragma unroll
    for (int i = 0; i < NMAX; ++i) {
        double v = scratch[NSUM + i];
#pragma unroll
        for (int w = 1; w < W; ++w) v = lb_nanmax(v, scratch[w * 32 + NSUM + i]);
        mx[i] = v;
    }
}

// block-wide sums of NV per-thread values; thread e < NV returns the total of element e (others: garbage)
template <int W, int NV>
__device__ __forceinline__ double block_sum_scatter(double (&v)[NV], double* scratch, int warp, int lane, int tid) {
    static_assert(NV <= 32, "scratch row");
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) scratch[warp * 32 + i] = v[i];
    }
    __syncthreads();
    double t = 0.0;
    if (tid < NV) {
#pragma unroll
        for (int w = 0; w < W; ++w) t += scratch[w * 32 + tid];
    }
    return t;
}

template <int NX, int NT, int NU, int W, bool BIG>
__global__ void __launch_bounds__(32 * W, W == 4 ? 4 : 7)
ipm_kernel_cta(const __grid_constant__ Params<NX, NT, NU> p, const BatchIO io, const double* __restrict__ Gglob,
               const double* __restrict__ hgglob) {
    static_assert(NT == 1 && NU == 1, "the CTA-per-QP kernel is built on the cooperative factorisation");
    using C = Core<NX, NT, NU>;
    using L = Layout<NX, NT, NU>;
    using CP = Coop<NX>;
    constexpr int NZ = NX + NT, NH = L::NH, NVB = NX + NU, NOUT = NH + 2 * NZ, kCtaThreads = 32 * W;
    static_assert(NOUT <= kCtaThreads && W <= kCtaMaxWarps, "reduction scratch");
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N;
    const L l(N, p.ngp);
    const CtaPlan<NX, NT, NU> plan(N, p.ngp, BIG);
    double* const slot = smem;
    double* const m = slot + l.o_misc;
    double* const xch = smem + plan.xch_off;
    const double* const zero_rec = smem + plan.zero_off;
    typename CP::LaneTab& ltab = *reinterpret_cast<typename CP::LaneTab*>(smem + plan.ltab_off);
    typename C::RowTab& tab = *reinterpret_cast<typename C::RowTab*>(smem + plan.rtab_off);
    double* const rb = smem + plan.rb_off;
    double* const red = smem + plan.red_off;
    const double* const Gs = BIG ? Gglob : smem + plan.g_off;
    const double* const hgs = BIG ? hgglob : smem + plan.hg_off;
    double* const ctl = smem + plan.ctl_off;

    // ---- one-time CTA setup ----
    if (!BIG) {
        for (int i = tid; i < NZ * p.ngp; i += kCtaThreads) smem[plan.g_off + i] = Gglob[i];
        for (int i = tid; i < p.ngp; i += kCtaThreads) smem[plan.hg_off + i] = hgglob[i];
    }
    for (int i = tid; i < L::RS2; i += kCtaThreads) smem[plan.zero_off + i] = 0.0;
    C::fill_rowtab(p, tab, tid, kCtaThreads);
    if (warp == 0) {
        typename CP::Lane ln;
        CP::lane_init(p, lane, xch, ln);
        CP::lane_store(ln, lane, ltab);
        CP::xch_init(lane, xch);
    }
    const int nitems = NVB * (N + 1);
    // The sequential sweeps of a QP run on ONE warp.  CTAs that share an SM use different warp indices for it, so
    // that their sweeps land on different SM sub-partitions (warp w of a CTA issues on sub-partition w % 4).
    unsigned nsm;
    asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
    const int sw = (int)((blockIdx.x / nsm) % W);
    const int slane = tid - 32 * sw;  // lane index within the sweep warp (only meaningful there)
    __syncthreads();

    // optional phase timing (diagnostic): thread 0 of CTA 0 accumulates clock64 deltas per phase
    const bool prof_on = io.prof != nullptr && blockIdx.x == 0 && tid == 0;
    long long t_prev = prof_on ? clock64() : 0;
#define LB_PROF(idx)                                           \
    if (prof_on) {                                             \
        const long long t_now = clock64();                     \
        io.prof[idx] += (unsigned long long)(t_now - t_prev);  \
        t_prev = t_now;                                        \
    }
    for (;;) {
        // ---- fetch the next QP and load its inputs ----
        if (tid == 0) {
            const unsigned long long t = atomicAdd(io.queue, 1ULL);
            const unsigned long long nq = io.qlist ? *io.qcount : (unsigned long long)io.batch;
            ctl[0] = t < nq ? (double)(io.qlist ? io.qlist[t] : (long long)t) : -1.0;
        }
        __syncthreads();
        const long long q = (long long)ctl[0];
        if (q < 0) break;
        if (tid < NX) slot[l.i_x(tid, 0)] = io.dx0[q * NX + tid];
        for (int k = tid; k < N; k += kCtaThreads) {
            slot[l.i_u(0, k)] = io.warm ? io.warm[q * (N + NT) + k] : 0.0;
#pragma unroll
            for (int j = 0; j < NX; ++j) slot[l.i_x(j, k + 1)] = io.d_off ? io.d_off[(q * N + k) * NX + j] : 0.0;
        }
        if (tid == 0) {
            m[L::M_TH] = io.warm ? io.warm[q * (N + NT) + N] : 0.0;
            double cconst = 0.0;
#pragma unroll
            for (int a = 0; a < NZ; ++a) {
                double v = 0.0;
                if (io.dx_ref) {
#pragma unroll
                    for (int j = 0; j < NX; ++j) v += p.Lref[a * NX + j] * io.dx_ref[q * NX + j];
                }
                m[L::M_LIN + a] = v;
            }
            if (io.dx_ref) {
#pragma unroll
                for (int i = 0; i < NX; ++i)
#pragma unroll
                    for (int j = 0; j < NX; ++j) cconst += io.dx_ref[q * NX + i] * p.Tm[i * NX + j] * io.dx_ref[q * NX + j];
            }
            m[L::M_CCONST] = cconst;
        }
        __syncthreads();
        if (slane == 0) C::rollout(p, l, slot);
        __syncthreads();

        int iters = 0, status = 1;  // LBMPC_ST_MAXITER unless a verdict is reached
        double alpha = 0.0;
        const CShift csh{io.cshift ? io.cshift + q * (long long)((N + 1) * io.cs_stride) : nullptr, io.cs_stride};
        LB_PROF(0)
        for (;;) {
            // ================= phase E+A: apply the previous step (or initialise the rows), predictor assembly =================
            if (iters > 0) {
                for (int it = tid; it < nitems; it += kCtaThreads) C::upd_item(p, l, slot, it / NVB, it % NVB, alpha);
            } else {
                for (int it = tid; it < nitems; it += kCtaThreads) C::init_item(p, tab, l, slot, it / NVB, it % NVB);
                for (int i = tid; i < p.ng; i += kCtaThreads) C::init_rows_gen(p, l, slot, Gs, hgs, i);
            }
            __syncthreads();  // the new iterate of a stage is read by all its items and by the polytope rows
            {
                RedAsm ra{0.0, 0.0, 0.0, 0.0, 0.0};
                for (int it = tid; it < nitems; it += kCtaThreads) C::asm_item(p, tab, l, slot, it / NVB, it % NVB, ra, csh);
                for (int k = tid; k <= N; k += kCtaThreads) C::asm_theta_item(p, tab, l, slot, N - k, ra, csh);  // from the other end: spreads the work
                double acc[NOUT];
                if (BIG) {
#pragma unroll
                    for (int a = 0; a < NOUT; ++a) acc[a] = 0.0;
                    for (int i = tid; i < p.ng; i += kCtaThreads) C::assemble_gen_row(p, l, slot, Gs, hgs, i, acc, ra);
                } else {
                    for (int i = tid; i < p.ng; i += kCtaThreads) C::gen_row_asm_scalars(p, l, slot, Gs, hgs, i, rb, ra);
                }
                double sm[3] = {ra.sl, ra.hl, ra.gth}, mx[2] = {ra.rp, ra.lam};
                block_reduce<W, 3, 2>(sm, mx, red, warp, lane);  // (also orders the row buffer before its readers)
                {
                    const double v = BIG ? block_sum_scatter<W, NOUT>(acc, red, warp, lane, tid)
                                         : (tid < NOUT ? C::gen_output_asm(p, Gs, rb, tid) : 0.0);
                    if (tid < NOUT) {
                        m[L::M_HG + tid] = v;  // HG, GGL, DG are contiguous
                        if (tid == NH + NX) m[L::M_GTH] = sm[2] + v;
                    }
                }
                if (tid == kCtaThreads - 1) {
                    m[L::M_RP] = mx[0];
                    m[L::M_MU] = sm[0] * p.inv_m;
                    m[L::M_LAM] = mx[1];
                    m[L::M_HLAM] = sm[1];
                }
                static_assert(L::M_GGL == L::M_HG + NH && L::M_DG == L::M_GGL + NZ, "misc layout");
            }
            __syncthreads();
            LB_PROF(1)

            // ================= phase B (warp 0): factorisation + dual residual + affine backward substitution; Farkas =================
            const bool cert = m[L::M_LAM] >= p.inf_trigger;
            if (warp == sw) {
                typename CP::Lane ln;
                CP::lane_load(ltab, lane, xch, ln);
                CP::begin(p, l, slot, zero_rec, ln);
                int type = C::stage_type(p, N), kseg = p.tseg[type];
                SA rec = sa_of(slot + l.r2(N - 1));
                int k = N - 1;
#define LB_STAGE(B)                                                          \
    {                                                                        \
        if (k < kseg) { /* crossed into the previous cost segment */         \
            type = C::stage_type(p, k);                                      \
            kseg = p.tseg[type];                                             \
            CP::load_type(p, type, ln);                                      \
        }                                                                    \
        CP::template st1<B>(ln, rec, k == p.kg);                             \
        __syncwarp();                                                        \
        CP::template st2<B>(ln);                                             \
        const double fa = __shfl_sync(kFull, ln.d1, ln.srcA);                \
        const double fb = __shfl_sync(kFull, ln.d1, ln.srcB);                \
        const double fuu = __shfl_sync(kFull, ln.d1, CP::kFu);               \
        CP::st3(ln, fa, fb, fuu);                                            \
        rec = sa_add(rec, -L::RS2);                                          \
        --k;                                                                 \
    }
                if (!(k & 1)) LB_STAGE(0)
                while (k >= 1) {  // k odd here: exchange buffer 1, then 0
                    LB_STAGE(1)
                    LB_STAGE(0)
                }
#undef LB_STAGE
                const double fin = CP::finish(ln);
                __syncwarp();
                bool okl = true;
                double rdl = 0.0;
                for (int kk = lane; kk < N; kk += 32) CP::check_stage(l, slot, kk, okl, rdl);
                const bool okall = __all_sync(kFull, okl);
                const double rdm = warp_max_nan(rdl);
                const double ptt = __shfl_sync(kFull, fin, NH - 1);
                const double pvth = __shfl_sync(kFull, fin, CP::kPv + NX);
                if (lane == 0) {
                    const double iptt = 1.0 / ptt;
                    m[L::M_PIV] = (okall && ptt > 0.0) ? 1.0 : 0.0;
                    m[L::M_PTT] = iptt;
                    m[L::M_DTHA] = -iptt * pvth;
                    m[L::M_RD] = lb_nanmax(rdm, lb_abs(m[L::M_GTH]));
                }
            }
            if (cert && warp == (sw + 1) % W) {  // Farkas recursion, blocked over the horizon (lanes = blocks), on ANOTHER warp:
                {                                  // it reads the assembly's output only and runs beside the factorisation
                    if (lane < l.nb) C::farkas_p1(p, l, slot, lane);
                    __syncwarp();
                    if (lane == 0) C::farkas_p2(p, l, slot);
                    __syncwarp();
                    double nrm = 0.0, yd = 0.0;
                    if (lane < l.nb) C::farkas_p3(p, l, slot, lane, nrm, yd);
                    nrm = warp_sum(nrm);
                    yd = warp_sum(yd);
                    if (lane == 0) {
                        m[L::M_CERT] = nrm;
                        m[L::M_HLAM] += yd;
                    }
                }
            }
            __syncthreads();
            LB_PROF(2)
            const int v = C::verdict(p, m, cert);
            if (v >= 0) {
                status = v;
                break;
            }
            if (iters >= p.max_iter) break;  // the last allowed iterate has been tested: LBMPC_ST_MAXITER
            // ================= phase B2: block transfer matrices (all warps), affine forward substitution (warp 0) =================
            for (int t = tid; t < l.nb * NX; t += kCtaThreads) C::bwd_p1_T(p, l, slot, zero_rec, t);
            if (tid >= kCtaThreads - l.nb) C::fwd_p1(p, l, slot, kCtaThreads - 1 - tid, true);  // the last warp: off the T tasks' lanes
            __syncthreads();
            if (slane == 0) C::fwd_p2(l, slot);
            __syncthreads();
            if (slane >= 0 && slane < l.nb) C::fwd_p3(p, l, slot, slane, true);
            __syncthreads();
            LB_PROF(3)

            // ================= phase C: affine step length, sigma, corrector rhs =================
            double sigmu;
            {
                RedStep rs{0.0, 0.0, 0.0, 0.0};
                for (int it = tid; it < nitems; it += kCtaThreads) C::aff_item(p, tab, l, slot, it / NVB, it % NVB, rs);
                double acc[2 * NZ];
                if (BIG) {
#pragma unroll
                    for (int a = 0; a < 2 * NZ; ++a) acc[a] = 0.0;
                    for (int i = tid; i < p.ng; i += kCtaThreads) C::affine_gen_row(p, l, slot, Gs, hgs, i, acc, rs);
                } else {
                    for (int i = tid; i < p.ng; i += kCtaThreads) C::gen_row_aff_scalars(p, l, slot, Gs, hgs, i, rb, rs);
                }
                double sm[3] = {rs.s0, rs.s1, rs.s2}, mx[1] = {rs.ratio};
                block_reduce<W, 3, 1>(sm, mx, red, warp, lane);
                const double aaff = mx[0] > 1.0 ? 1.0 / mx[0] : 1.0;
                const double mu = m[L::M_MU];
                const double mu_aff = (sm[0] + aaff * sm[1] + aaff * aaff * sm[2]) * p.inv_m;
                const double sr = mu_aff / mu;
                sigmu = fmax(sr * sr * mu, 0.1 * p.tol_mu);  // no centring below the target gap (round-off at degenerate vertices)
                for (int it = tid; it < nitems; it += kCtaThreads) C::corr_item(p, l, slot, it / NVB, it % NVB, sigmu);
                if (BIG) {
                    const double t = block_sum_scatter<W, 2 * NZ>(acc, red, warp, lane, tid);  // thread a: acc[a], thread NZ + a: acc[NZ + a]
                    if (tid >= NZ && tid < 2 * NZ) red[tid] = t;                               // (row 0 of the scratch has been consumed)
                    __syncthreads();
                    if (tid < NZ) m[L::M_DG + tid] = t + sigmu * red[NZ + tid];
                } else if (tid < NZ) {
                    m[L::M_DG + tid] = C::gen_output_aff(p, Gs, rb, tid, sigmu);
                }
            }
            __syncthreads();
            LB_PROF(4)

            // ================= phase D (warp 0): corrector backward / forward substitution =================
            if (warp == sw) solve_sweeps<NX, NT, NU>(p, l, slot, zero_rec, lane, false, false);
            __syncthreads();
            LB_PROF(5)

            // ================= phase E: final directions, step length, polytope rows and theta =================
            {
                double ratio = 0.0;
                for (int it = tid; it < nitems; it += kCtaThreads)
                    ratio = fmax(ratio, C::fin_item(p, tab, l, slot, it / NVB, it % NVB, sigmu));
                for (int i = tid; i < p.ng; i += kCtaThreads) ratio = fmax(ratio, C::final_gen_row(p, l, slot, Gs, hgs, i, sigmu));
                double sm[1] = {0.0}, mx[1] = {ratio};
                block_reduce<W, 1, 1>(sm, mx, red, warp, lane);
                alpha = mx[0] > 0.0 ? 0.99 / mx[0] : 1.0;
                alpha = alpha > 1.0 ? 1.0 : alpha;
                for (int i = tid; i < p.ng; i += kCtaThreads) C::update_gen_row(p, l, slot, Gs, hgs, i, sigmu, alpha);
                __syncthreads();  // the polytope rows read x_kg, theta of the old iterate
                if (tid == 0) m[L::M_TH] += alpha * m[L::M_DTH];
            }
            iters += 1;
            __syncthreads();
            LB_PROF(6)
            if (prof_on) io.prof[7] += 1;
        }

        // ---- results ----
        {
            double J = 0.0;
            for (int k = tid; k <= N; k += kCtaThreads) J += C::objective_stage(p, l, slot, k, csh);
            double sm[1] = {J}, mx[1] = {0.0};
            block_reduce<W, 1, 1>(sm, mx, red, warp, lane);
            for (int k = tid; k < N; k += kCtaThreads) {
                double v = slot[l.i_u(0, k)];
#pragma unroll
                for (int j = 0; j < NX; ++j) v -= p.Kout[j] * slot[l.i_x(j, k)];
                io.uc[q * N + k] = v;
            }
            if (io.xtraj) {
                for (int k = tid; k <= N; k += kCtaThreads)
#pragma unroll
                    for (int j = 0; j < NX; ++j) io.xtraj[(q * (N + 1) + k) * NX + j] = slot[l.i_x(j, k)];
            }
            if (tid == 0) {
                io.theta[q] = m[L::M_TH];
                io.obj[q] = sm[0] + m[L::M_CCONST];
                io.iters[q] = iters;
                io.status[q] = status;
            }
        }
        __syncthreads();
        LB_PROF(8)
    }
#undef LB_PROF
}

}  // namespace lbmpc
