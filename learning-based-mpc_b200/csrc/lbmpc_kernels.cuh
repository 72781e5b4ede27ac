// lbmpc_kernels.cuh — sm_100a kernels of the batched (LB)MPC interior-point engine.
//
//  ipm_kernel        persistent Mehrotra predictor-corrector solver.  One CTA keeps G QPs ("slots")
//                    resident in shared memory (whole iterate, factors and directions: no HBM
//                    traffic inside the solve).  Two thread mappings alternate, separated by
//                    __syncthreads():
//                      - stage/row phases : warp g works on slot g, lanes stride over the N+1
//                        stages and the polytope rows; reductions (duality gap, residual norms,
//                        step-length max-ratio, polytope Hessian) by warp shuffles;
//                      - sweep phases     : the Riccati backward / forward recursions run
//                        thread-local, lane g of warp 0 owns slot g, so one warp advances all G
//                        QPs of the CTA in lock step with no communication.
//                    Slots are refilled from a global atomic work queue as soon as their QP
//                    terminates, so QPs with different iteration counts never wait for each other.
//                    The polytope matrix (616 x 5 for the LMPC terminal set) is staged once per CTA
//                    with a 1-D TMA bulk copy (cp.async.bulk + mbarrier).
//  oracle_kernel     learned-model rollout x+ = A x + B u + g(xi) with the L2-regularised
//                    Nadaraya-Watson oracle (oracleL2NW.m:26-36 / casadiL2NW.m:14-28); warp per QP,
//                    lanes over the q data points.
//  plant_kernel      closed-loop step: RK4 Moore-Greitzer plant, disturbance, data window update,
//                    warm-start shift (LBMPC_casadi.m:191-208, get_data.m:3-9).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "lbmpc_core.cuh"

namespace lbmpc {

constexpr int kMaxSlots = 8;          // slots (= warps) per CTA; bounds registers per thread
constexpr unsigned kFull = 0xffffffffu;

struct BatchIO {
    long long batch;
    const double *dx0, *dx_ref, *d_off, *warm;
    double *uc, *theta, *xtraj, *obj;
    int *iters, *status;
    unsigned long long* queue;  // global work counter (zeroed before launch)
    unsigned long long* prof;   // optional (may be null): per-phase SM cycles of CTA 0, see lbmpc_debug_phase_cycles
    const double* cshift;       // optional cs_stride x (N+1) per QP: the cost is evaluated at [x_k + ex_k; u_k + eu_k] (CShift, lbmpc_core.cuh)
    int cs_stride;              // NX: state shift only; NX + NU: state and input shift
    int row_shift;              // stream mapping only: cshift moves the ROWS instead of the cost (first-order SQP)
    int lockstep;               // warp kernel: the warps of a CTA start every iteration together (see cta_tick)
    const long long* qlist;     // optional: the launch solves QPs qlist[0 .. *qcount) instead of 0 .. batch (QPs the stream mapping
    const unsigned long long* qcount;  // handed over after its iteration budget: lbmpc_b200.cu launch_ipm_stream)
};

// CTA-wide barrier that also counts the warps that still have a QP.  The warps of a CTA run independent QPs; at many
// QPs per SM they drift to different phases of a ~120 KB instruction stream and the SM's instruction cache thrashes
// (ncu, batch 16384: 26 % of the stall samples are instruction fetches, GPC instruction-cache requests at 87 % of
// peak).  One tick per iteration keeps them within a phase of each other, so they share the fetched lines
// (instruction-cache hit rate 70 % -> 94 %, fetch stalls 1.23 -> 0.10 per issue, 16.7 -> 14.3 ms at batch 65536).
// A tick per PHASE, or per half iteration, measured slower than one per iteration (16.3 / 16.1 ms): the warps then
// wait on each other's phase lengths.  Not used below ~3 QPs per warp slot, where the idle wait of a warp that
// starts a new QP mid-iteration costs more than the fetches (batch 1024: -3 %).
__device__ __forceinline__ unsigned cta_tick(bool working) {
    unsigned r;
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %1, 0; bar.red.popc.u32 %0, 1, p; }" : "=r"(r) : "r"((unsigned)working) : "memory");
    return r;
}

// ---------------------------------------------------------------------------------------------
// warp reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
// max that propagates NaN (fmax drops it): used for residual norms so that status 3 is seen
__device__ __forceinline__ double warp_max_nan(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(kFull, v, o);
        v = (v != v || w != w) ? (v != v ? v : w) : (v > w ? v : w);
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// the solver
// ---------------------------------------------------------------------------------------------
// shared-memory carve-up helper, identical on host and device
template <int NX, int NT, int NU>
struct SmemPlan {
    using L = Layout<NX, NT, NU>;
    static constexpr int kXchPerSlot = 104;  // >= Coop::kXch doubles per QP (even)
    int slots, stride, xch_off, zero_off, g_off, hg_off, meta_off;  // offsets in doubles
    size_t bytes;
    __host__ __device__ SmemPlan(int N, int ngp, int slots_, bool stage_g, bool poly_global = false) {
        const L l(N, ngp, poly_global);
        slots = slots_;
        stride = l.stride;
        int o = slots * stride;
        o = (o + 1) & ~1;  // 16-byte alignment
        xch_off = o;
        o += slots * kXchPerSlot;
        zero_off = o;      // all-zero record read by the homogeneous block recursions
        o += (L::RS2 + 1) & ~1;
        g_off = o;         // bulk copy destination (16-byte aligned)
        if (stage_g) o += (NX + NT) * ngp;
        hg_off = o;
        if (stage_g) o += ngp;
        meta_off = o;
        o += 2;  // mbarrier (8 B) + pad
        bytes = (size_t)o * sizeof(double);
    }
};

// blocked backward + forward substitution of one QP by its warp (lanes = blocks / block x vector tasks)
template <int NX, int NT, int NU>
__device__ __forceinline__ void solve_sweeps(const Params<NX, NT, NU>& p, const Layout<NX, NT, NU>& l, double* slot,
                                             const double* zero_rec, int lane, bool aff, bool with_T,
                                             unsigned long long* prof = nullptr) {
    using C = Core<NX, NT, NU>;
    constexpr int NZ = NX + NT;
    long long t0 = prof ? clock64() : 0;
#define LB_SW(i)                                             \
    if (prof) {                                              \
        const long long t1 = clock64();                      \
        prof[i] += (unsigned long long)(t1 - t0);            \
        t0 = t1;                                             \
    }
    const int ntask = l.nb * (with_T ? NX + 1 : 1);
    for (int t = lane; t < ntask; t += 32) C::bwd_p1(p, l, slot, zero_rec, t, with_T);
    __syncwarp();
    LB_SW(9)
    if (lane == 0) C::bwd_p2(l, slot, aff);
    __syncwarp();
    LB_SW(10)
    double pv[NZ];
    if (lane < l.nb) C::bwd_p3_in(p, l, slot, lane, pv);
    __syncwarp();
    if (lane < l.nb) C::bwd_p3_fwd_p1(p, l, slot, lane, pv, aff);
    __syncwarp();
    LB_SW(11)
    if (lane == 0) C::fwd_p2(l, slot);
    __syncwarp();
    LB_SW(12)
    if (lane < l.nb) C::fwd_p3(p, l, slot, lane, aff);
    __syncwarp();
    LB_SW(13)
#undef LB_SW
}

// predictor solve when the factor sweep has already run the affine backward substitution (kappa stored, d(theta)
// at M_DTHA): block transfer matrices (homogeneous backward recursion, once per factorisation) + forward substitution
template <int NX, int NT, int NU>
__device__ __forceinline__ void affine_forward(const Params<NX, NT, NU>& p, const Layout<NX, NT, NU>& l, double* slot,
                                               const double* zero_rec, int lane) {
    using C = Core<NX, NT, NU>;
    const int ntask = l.nb * NX;
    for (int t = lane; t < ntask; t += 32) C::bwd_p1_T(p, l, slot, zero_rec, t);
    if (lane < l.nb) C::fwd_p1(p, l, slot, lane, true);
    __syncwarp();
    if (lane == 0) C::fwd_p2(l, slot);
    __syncwarp();
    if (lane < l.nb) C::fwd_p3(p, l, slot, lane, true);
    __syncwarp();
}

// One warp per resident QP ("slot"), `slots` warps per CTA, no CTA-level synchronisation inside the
// solve: a warp fetches a QP from the global work queue, iterates it to its verdict with every phase
// mapped onto its 32 lanes, writes the result and fetches the next one.  The phases of different
// warps interleave freely on the SM sub-partitions (a single warp can issue only every other cycle,
// so two resident warps per scheduler is what saturates issue).
template <int NX, int NT, int NU>
__global__ void __launch_bounds__(32 * kMaxSlots, 1)
ipm_kernel(const __grid_constant__ Params<NX, NT, NU> p, const BatchIO io, const double* __restrict__ Gglob,
           const double* __restrict__ hgglob, const int slots, const int stage_g, double* gpoly = nullptr) {
    using C = Core<NX, NT, NU>;
    using L = Layout<NX, NT, NU>;
    constexpr int NZ = NX + NT, NH = L::NH, NACC = NH + 2 * NZ;
    constexpr bool kCoop = (NT == 1 && NU == 1 && NX >= 2 && NX <= 4);  // one-warp Riccati factorisation
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    L l(p.N, p.ngp, gpoly != nullptr);
    if (gpoly) l.gsg = gpoly + ((size_t)blockIdx.x * slots + warp) * 2 * p.ngp;  // this warp's slacks / multipliers of the polytope rows
    const SmemPlan<NX, NT, NU> plan(p.N, p.ngp, slots, stage_g != 0, gpoly != nullptr);
    double* const slot = smem + warp * l.stride;
    double* const m = slot + l.o_misc;
    uint64_t* const bar = reinterpret_cast<uint64_t*>(smem + plan.meta_off);
    const double* Gs = Gglob;
    const double* hgs = hgglob;
    const int N = p.N;

    // ---- one-time CTA setup: stage the polytope with a bulk TMA copy ----
    if (stage_g) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t gb = (uint32_t)(NZ * p.ngp * sizeof(double)), hb = (uint32_t)(p.ngp * sizeof(double));
            mbar_expect_tx(bar, gb + hb);
            tma_bulk_g2s(smem + plan.g_off, Gglob, gb, bar);
            tma_bulk_g2s(smem + plan.hg_off, hgglob, hb, bar);
        }
        Gs = smem + plan.g_off;
        hgs = smem + plan.hg_off;
    }
    // ---- cooperative factorisation: per-lane coefficient vector (registers), exchange buffer ----
    using CP = Coop<kCoop ? NX : 2>;
    typename CP::Lane ln;
    double* const xch = smem + plan.xch_off + warp * SmemPlan<NX, NT, NU>::kXchPerSlot;
    if constexpr (kCoop) {
        CP::lane_init(p, lane, xch, ln);
        CP::xch_init(lane, xch);
    }
    const double* const zero_rec = smem + plan.zero_off;
    for (int i = threadIdx.x; i < L::RS2; i += blockDim.x) smem[plan.zero_off + i] = 0.0;
    if (stage_g) mbar_wait(bar, 0);
    __syncthreads();

    // optional phase timing (diagnostic): lane 0 of warp 0 of CTA 0 accumulates clock64 deltas per phase
    const bool prof_on = io.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long t_prev = prof_on ? clock64() : 0;
    long long t_prev2 = t_prev;
#define LB_PROF(idx)                                           \
    if (prof_on) {                                             \
        const long long t_now = clock64();                     \
        io.prof[idx] += (unsigned long long)(t_now - t_prev);  \
        t_prev = t_now;                                        \
        t_prev2 = t_now;                                       \
    }
#define LB_PROF2(idx)                                           \
    if (prof_on) {                                              \
        const long long t_now = clock64();                      \
        io.prof[idx] += (unsigned long long)(t_now - t_prev2);  \
        t_prev2 = t_now;                                        \
    }

    for (;;) {
        // ---- fetch the next QP and load its inputs ----
        long long q = -1;
        if (lane == 0) {
            const unsigned long long t = atomicAdd(io.queue, 1ULL);
            const unsigned long long nq = io.qlist ? *io.qcount : (unsigned long long)io.batch;
            q = t < nq ? (io.qlist ? io.qlist[t] : (long long)t) : -1;
        }
        q = __shfl_sync(kFull, q, 0);
        if (q < 0) break;
        if (lane < NX) slot[l.i_x(lane, 0)] = io.dx0[q * NX + lane];
        for (int k = lane; k < N; k += 32) {
#pragma unroll
            for (int i = 0; i < NU; ++i) slot[l.i_u(i, k)] = io.warm ? io.warm[q * (NU * N + NT) + k * NU + i] : 0.0;
#pragma unroll
            for (int j = 0; j < NX; ++j) slot[l.i_x(j, k + 1)] = io.d_off ? io.d_off[(q * N + k) * NX + j] : 0.0;
        }
        if (lane == 0) {
#pragma unroll
            for (int t = 0; t < NT; ++t) m[L::M_TH + t] = io.warm ? io.warm[q * (NU * N + NT) + N * NU + t] : 0.0;
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
        __syncwarp();
        if (lane == 0) C::rollout(p, l, slot);
        __syncwarp();

        int iters = 0, status = 1;  // LBMPC_ST_MAXITER unless a verdict is reached
        double alpha = 0.0;
        const CShift csh{io.cshift ? io.cshift + q * (long long)((N + 1) * io.cs_stride) : nullptr, io.cs_stride};
        LB_PROF(0)
        for (;;) {
            if (io.lockstep) cta_tick(true);
            // =================================================================================
            // phase E+A: apply the step of the previous iteration (or initialise the rows of a fresh
            //            QP) fused with the predictor assembly of this iteration
            // =================================================================================
            {
                RedAsm ra{0.0, 0.0, 0.0, 0.0, 0.0};
                if (iters > 0) {
                    for (int k = lane; k <= N; k += 32) C::update_assemble_stage(p, l, slot, k, alpha, ra, csh);
                } else {
                    for (int k = lane; k <= N; k += 32) C::init_assemble_stage(p, l, slot, k, ra, csh);
                    for (int i = lane; i < p.ng; i += 32) C::init_rows_gen(p, l, slot, Gs, hgs, i);
                }
                __syncwarp();  // x_kg of the new iterate is read by the polytope rows
                double acc[NACC];
#pragma unroll
                for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
                for (int i = lane; i < p.ng; i += 32) C::assemble_gen_row(p, l, slot, Gs, hgs, i, acc, ra);
                LB_PROF2(14)
#pragma unroll
                for (int a = 0; a < NACC; ++a) acc[a] = warp_sum(acc[a]);
                const double rp = warp_max(ra.rp), sl = warp_sum(ra.sl), lam = warp_max(ra.lam), hl = warp_sum(ra.hl);
                const double gth = warp_sum(ra.gth);
                if (lane == 0) {
                    m[L::M_GTH] = gth + acc[NH + NX];
#pragma unroll
                    for (int a = 0; a < NH; ++a) m[L::M_HG + a] = acc[a];
#pragma unroll
                    for (int a = 0; a < NZ; ++a) {
                        m[L::M_GGL + a] = acc[NH + a];
                        m[L::M_DG + a] = acc[NH + NZ + a];
                    }
                    m[L::M_RP] = rp;
                    m[L::M_MU] = sl * p.inv_m;
                    m[L::M_LAM] = lam;
                    m[L::M_HLAM] = hl;
                }
                __syncwarp();
            }
            LB_PROF2(15)
            LB_PROF(1)

            // =================================================================================
            // phase B: Riccati factorisation + dual residual (one warp, one dot product per lane and
            //          stage); Farkas recursion when the multipliers are large
            // =================================================================================
            const bool cert = m[L::M_LAM] >= p.inf_trigger;
            if constexpr (kCoop) {
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
                if (cert) {  // Farkas recursion, blocked over the horizon (lanes = blocks)
                    __syncwarp();
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
            } else {
                if (lane == 0) C::factor_serial(p, l, slot);
                else if (lane == 1) C::adjoint_sweep(p, l, slot, false);
                else if (lane == 2 && cert) C::adjoint_sweep(p, l, slot, true);
            }
            __syncwarp();
            LB_PROF(2)

            // =================================================================================
            // phase B2: verdict, then the affine backward/forward substitution
            // =================================================================================
            const int v = C::verdict(p, m, cert);
            if (v >= 0) {
                status = v;
                break;
            }
            if (iters >= p.max_iter) break;  // the last allowed iterate has been tested: LBMPC_ST_MAXITER
            if constexpr (kCoop) affine_forward<NX, NT, NU>(p, l, slot, zero_rec, lane);
            else solve_sweeps<NX, NT, NU>(p, l, slot, zero_rec, lane, true, true);
            LB_PROF(3)

            // =================================================================================
            // phase C: affine step length, sigma, corrector rhs
            // =================================================================================
            double sigmu;
            {
                RedStep rs{0.0, 0.0, 0.0, 0.0};
                for (int k = lane; k <= N; k += 32) C::affine_stage(p, l, slot, k, rs);
                double acc[2 * NZ];
#pragma unroll
                for (int a = 0; a < 2 * NZ; ++a) acc[a] = 0.0;
                for (int i = lane; i < p.ng; i += 32) C::affine_gen_row(p, l, slot, Gs, hgs, i, acc, rs);
                const double ratio = warp_max(rs.ratio);
                const double s0 = warp_sum(rs.s0), s1 = warp_sum(rs.s1), s2 = warp_sum(rs.s2);
#pragma unroll
                for (int a = 0; a < 2 * NZ; ++a) acc[a] = warp_sum(acc[a]);
                const double aaff = ratio > 1.0 ? 1.0 / ratio : 1.0;
                const double mu = m[L::M_MU];
                const double mu_aff = (s0 + aaff * s1 + aaff * aaff * s2) * p.inv_m;
                const double sr = mu_aff / mu;
                sigmu = fmax(sr * sr * mu, 0.1 * p.tol_mu);  // no centring below the target gap (round-off at degenerate vertices)
                for (int k = lane; k <= N; k += 32) C::corr_stage(p, l, slot, k, sigmu);
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int a = 0; a < NZ; ++a) m[L::M_DG + a] = acc[a] + sigmu * acc[NZ + a];
                }
                __syncwarp();
            }
            LB_PROF(4)

            // =================================================================================
            // phase D: corrector backward/forward substitution
            // =================================================================================
            solve_sweeps<NX, NT, NU>(p, l, slot, zero_rec, lane, false, false, prof_on ? io.prof : nullptr);
            LB_PROF(5)

            // =================================================================================
            // phase E (first half): final directions of the rows, step length, polytope rows and theta
            // =================================================================================
            {
                double ratio = 0.0;
                for (int k = lane; k <= N; k += 32) ratio = fmax(ratio, C::final_stage(p, l, slot, k, sigmu));
                for (int i = lane; i < p.ng; i += 32) ratio = fmax(ratio, C::final_gen_row(p, l, slot, Gs, hgs, i, sigmu));
                ratio = warp_max(ratio);
                alpha = ratio > 0.0 ? 0.99 / ratio : 1.0;
                alpha = alpha > 1.0 ? 1.0 : alpha;
                for (int i = lane; i < p.ng; i += 32) C::update_gen_row(p, l, slot, Gs, hgs, i, sigmu, alpha);
                __syncwarp();  // the polytope rows read x_kg, theta of the old iterate
                if (lane == 0) {
#pragma unroll
                    for (int t = 0; t < NT; ++t) m[L::M_TH + t] += alpha * m[L::M_DTH + t];
                }
                __syncwarp();
            }
            iters += 1;
            LB_PROF(6)
            if (prof_on) io.prof[7] += 1;
        }

        // ---- results ----
        {
            double J = 0.0;
            for (int k = lane; k <= N; k += 32) J += C::objective_stage(p, l, slot, k, csh);
            J = warp_sum(J);
            for (int k = lane; k < N; k += 32) {
#pragma unroll
                for (int i = 0; i < NU; ++i) {
                    double v = slot[l.i_u(i, k)];
#pragma unroll
                    for (int j = 0; j < NX; ++j) v -= p.Kout[i * NX + j] * slot[l.i_x(j, k)];
                    io.uc[(q * N + k) * NU + i] = v;
                }
            }
            if (io.xtraj) {
                for (int k = lane; k <= N; k += 32)
#pragma unroll
                    for (int j = 0; j < NX; ++j) io.xtraj[(q * (N + 1) + k) * NX + j] = slot[l.i_x(j, k)];
            }
            if (lane == 0) {
#pragma unroll
                for (int t = 0; t < NT; ++t) io.theta[q * NT + t] = m[L::M_TH + t];
                io.obj[q] = J + m[L::M_CCONST];
                io.iters[q] = iters;
                io.status[q] = status;
            }
            __syncwarp();
        }
        LB_PROF(8)
    }
    if (io.lockstep) {  // keep arriving until every warp of the CTA has run out of work
        while (cta_tick(false) != 0) {}
    }
#undef LB_PROF
#undef LB_PROF2
}

// ---------------------------------------------------------------------------------------------
// learned-model rollout: d_k = g([x1;x2;u_k]) , x_{k+1} = A x_k + B u_k + d_k        (warp per QP)
// X: 3 x q x batch, Y: NX x q x batch (column-major, "one column per sample"), valid: q x batch
// ---------------------------------------------------------------------------------------------
constexpr int kOracleMaxPerLane = 16;  // q <= 512

// PER = data points held in registers per lane (4: q <= 128, the reference's q = 10/50/100; 16: q <= 512)
template <int NX, int NU, int PER>
__global__ void __launch_bounds__(128, 6)
oracle_kernel(const double* __restrict__ A, const double* __restrict__ B, const double* __restrict__ Kfb, int N, long long batch,
              int q, double inv_h2, double lambda, const double* __restrict__ dx0, const double* __restrict__ du,
              long long du_ld, const double* __restrict__ X, const double* __restrict__ Y, const double* __restrict__ valid,
              double* __restrict__ d_off) {
    constexpr int NI = 3;
    const int lane = threadIdx.x & 31;
    const long long qp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qp >= batch) return;
    double xs[PER][NI], ys[PER][NX], vs[PER];
    const int per = (q + 31) / 32;
#pragma unroll
    for (int r = 0; r < PER; ++r) {
        const int i = lane + 32 * r;
        const bool on = r < per && i < q;
#pragma unroll
        for (int a = 0; a < NI; ++a) xs[r][a] = on ? X[(qp * q + i) * NI + a] : 0.0;
#pragma unroll
        for (int a = 0; a < NX; ++a) ys[r][a] = on ? Y[(qp * q + i) * NX + a] : 0.0;
        vs[r] = on ? (valid ? valid[qp * q + i] : 1.0) : -1.0;  // -1: slot unused
        // a masked sample with Y = 0 (a window slot that has not been written yet) adds nothing to either sum: treat it as unused,
        // so that a partly filled window — the first q steps of every closed loop — costs only its own exponentials
        bool zero = vs[r] == 0.0;
#pragma unroll
        for (int a = 0; a < NX; ++a) zero = zero && ys[r][a] == 0.0;
        if (zero) vs[r] = -1.0;
    }
    double Am[NX * NX], Bm[NX * NU], Km[NU * NX], x[NX];
#pragma unroll
    for (int i = 0; i < NU * NX; ++i) Km[i] = Kfb ? Kfb[i] : 0.0;  // F-form: the sequence holds c, u = K x + c (transitionLearned.m:13)
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) Am[i] = A[i];
#pragma unroll
    for (int i = 0; i < NX * NU; ++i) Bm[i] = B[i];
#pragma unroll
    for (int j = 0; j < NX; ++j) x[j] = dx0[qp * NX + j];
    for (int k = 0; k < N; ++k) {
        double u[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            double v = du[qp * du_ld + k * NU + i];
#pragma unroll
            for (int j = 0; j < NX; ++j) v += Km[i * NX + j] * x[j];
            u[i] = v;
        }
        const double xi[NI] = {x[0], x[1], u[0]};
        double sk = 0.0, g[NX];
#pragma unroll
        for (int a = 0; a < NX; ++a) g[a] = 0.0;
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            if (r < per && vs[r] >= 0.0) {
                double d2 = 0.0;
#pragma unroll
                for (int a = 0; a < NI; ++a) {
                    const double d = xs[r][a] - xi[a];
                    d2 += d * d;
                }
                const double kv = exp(-d2 * inv_h2);
                sk += valid ? kv * vs[r] : kv;
#pragma unroll
                for (int a = 0; a < NX; ++a) g[a] += ys[r][a] * kv;
            }
        }
        sk = warp_sum(sk);
        const double wn = 1.0 / (lambda + sk);
        double xn[NX];
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            g[a] = warp_sum(g[a]) * wn;
            double v = g[a];
#pragma unroll
            for (int j = 0; j < NX; ++j) v += Am[a * NX + j] * x[j];
#pragma unroll
            for (int i = 0; i < NU; ++i) v += Bm[a * NU + i] * u[i];
            xn[a] = v;
        }
        if (lane < NX) {
            double gv = g[0];
#pragma unroll
            for (int a = 1; a < NX; ++a) gv = lane == a ? g[a] : gv;
            d_off[(qp * N + k) * NX + lane] = gv;
        }
#pragma unroll
        for (int a = 0; a < NX; ++a) x[a] = xn[a];
    }
}

// ---------------------------------------------------------------------------------------------
// first-order model of the learned dynamics along the learned rollout of the previous solution (warp per QP, lanes over the
// q data points):  g(xi) AND dg/dxi of the L2NW oracle (what CasADi's AD of casadiL2NW.m:14-28 hands IPOPT,
// DMS_LBMPC_casadi.m:252-319 / hybrid_LBMPC_casadi.m:250-267):
//     k_i = exp(-|X_i - xi|^2 / h^2), D = lambda + sum v_i k_i, g = sum Y_i k_i / D
//     dg/dxi = (sum Y_i dk_i - g sum v_i dk_i) / D,  dk_i = k_i 2 (X_i - xi) / h^2
// Outputs per stage: J_k (NX x 3), d_k = g(xibar_k) - J_k xibar_k, and (rs != nullptr) the gap between the learned and the
// nominal rollout of the same decision variables, records [e_k | K e_k] — the rows of the QP follow the nominal sequence.
// ---------------------------------------------------------------------------------------------
template <int NX, int PER>
__global__ void __launch_bounds__(128, 4)
oracle_jac_kernel(const double* __restrict__ A, const double* __restrict__ B, const double* __restrict__ Kfb, int N, long long batch,
                  int q, double inv_h2, double lambda, const double* __restrict__ dx0, const double* __restrict__ du,
                  long long du_ld, const double* __restrict__ X, const double* __restrict__ Y, const double* __restrict__ valid,
                  double* __restrict__ d_off, double* __restrict__ jac, double* __restrict__ rs) {
    constexpr int NI = 3;
    const int lane = threadIdx.x & 31;
    const long long qp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qp >= batch) return;
    double xs[PER][NI], ys[PER][NX], vs[PER];
    const int per = (q + 31) / 32;
#pragma unroll
    for (int r = 0; r < PER; ++r) {
        const int i = lane + 32 * r;
        const bool on = r < per && i < q;
#pragma unroll
        for (int a = 0; a < NI; ++a) xs[r][a] = on ? X[(qp * q + i) * NI + a] : 0.0;
#pragma unroll
        for (int a = 0; a < NX; ++a) ys[r][a] = on ? Y[(qp * q + i) * NX + a] : 0.0;
        vs[r] = on ? (valid ? valid[qp * q + i] : 1.0) : -1.0;
        bool zero = vs[r] == 0.0;  // unwritten window slot (mask 0, Y = 0): contributes to no sum, see oracle_kernel
#pragma unroll
        for (int a = 0; a < NX; ++a) zero = zero && ys[r][a] == 0.0;
        if (zero) vs[r] = -1.0;
    }
    double Am[NX * NX], Bm[NX], Km[NX], x[NX], xn[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) Km[i] = Kfb ? Kfb[i] : 0.0;
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) Am[i] = A[i];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        Bm[i] = B[i];
        x[i] = xn[i] = dx0[qp * NX + i];   // x: learned rollout, xn: nominal rollout
    }
    for (int k = 0; k <= N; ++k) {
        if (rs && lane <= NX) {  // gap record of stage k
            double eu = 0.0, ev = 0.0;
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                eu += Km[j] * (x[j] - xn[j]);
                ev = lane == j ? x[j] - xn[j] : ev;
            }
            rs[(qp * (N + 1) + k) * (NX + 1) + lane] = lane == NX ? eu : ev;
        }
        if (k == N) break;
        const double c = du[qp * du_ld + k];
        double u = c, un = c;
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            u += Km[j] * x[j];
            un += Km[j] * xn[j];
        }
        const double xi[NI] = {x[0], x[1], u};
        double sk = 0.0, g[NX], dD[NI], dn[NX][NI];
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            g[a] = 0.0;
#pragma unroll
            for (int cc = 0; cc < NI; ++cc) dn[a][cc] = 0.0;
        }
#pragma unroll
        for (int cc = 0; cc < NI; ++cc) dD[cc] = 0.0;
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            if (r < per && vs[r] >= 0.0) {
                double d2 = 0.0, dv[NI];
#pragma unroll
                for (int a = 0; a < NI; ++a) {
                    dv[a] = xs[r][a] - xi[a];
                    d2 += dv[a] * dv[a];
                }
                const double kv = exp(-d2 * inv_h2), vi = valid ? vs[r] : 1.0;
                sk += vi * kv;
#pragma unroll
                for (int cc = 0; cc < NI; ++cc) {
                    const double wk = kv * 2.0 * dv[cc] * inv_h2;
                    dD[cc] += vi * wk;
#pragma unroll
                    for (int a = 0; a < NX; ++a) dn[a][cc] += ys[r][a] * wk;
                }
#pragma unroll
                for (int a = 0; a < NX; ++a) g[a] += ys[r][a] * kv;
            }
        }
        sk = warp_sum(sk);
        const double wn = 1.0 / (lambda + sk);
#pragma unroll
        for (int cc = 0; cc < NI; ++cc) dD[cc] = warp_sum(dD[cc]);
        double Jv = 0.0, dvv = 0.0;  // lane a*3+c keeps J[a][c]; lane a keeps d[a]
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            g[a] = warp_sum(g[a]) * wn;
            double dk = g[a];
#pragma unroll
            for (int cc = 0; cc < NI; ++cc) {
                const double J = (warp_sum(dn[a][cc]) - g[a] * dD[cc]) * wn;
                dk -= J * xi[cc];
                Jv = lane == a * NI + cc ? J : Jv;
            }
            dvv = lane == a ? dk : dvv;
        }
        if (lane < NX * NI) jac[((qp * N + k) * NX) * NI + lane] = Jv;
        if (lane < NX) d_off[(qp * N + k) * NX + lane] = dvv;
        double xl[NX], xm[NX];
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            double v = g[a] + Bm[a] * u, w = Bm[a] * un;
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                v += Am[a * NX + j] * x[j];
                w += Am[a * NX + j] * xn[j];
            }
            xl[a] = v;
            xm[a] = w;
        }
#pragma unroll
        for (int a = 0; a < NX; ++a) {
            x[a] = xl[a];
            xn[a] = xm[a];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// closed loop plant step (thread per scenario): RK4 Moore-Greitzer, disturbance, data window,
// warm-start shift, history.  C-form conventions (LBMPC_casadi.m:184-208).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mg_rhs(const double* x, double u, double* f) {
    f[0] = -x[1] + 1.0 + 3.0 * (x[0] / 2.0) - (x[0] * x[0] * x[0] / 2.0);
    f[1] = x[0] + 1.0 - x[2] * sqrt(x[1]);
    f[2] = x[3];
    f[3] = -1000.0 * x[2] - 2.0 * sqrt(500.0) * x[3] + 1000.0 * u;
}
__device__ __forceinline__ double lb_uniform(unsigned long long seed, unsigned long long scen,
                                             unsigned long long step, unsigned long long comp) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (scen * 0x100000001B3ULL + step * 8ULL + comp + 1ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

struct LoopState {
    double* x;      // 4 x batch  absolute plant state
    double* dx0;    // 4 x batch  x - x_eq (solver input)
    double* X;      // 3 x q x batch
    double* Y;      // 4 x q x batch
    double* V;      // q x batch
    int* nd;        // batch      samples in the window
    double* warm;   // (N+1) x batch
    const double *uc, *theta;  // solver outputs of this step
    const int *iters, *status;
    double *x_hist, *u_hist, *theta_hist;  // histories (may be null)
    int *iters_hist, *status_hist;
};

__global__ void __launch_bounds__(128)
plant_kernel(LoopState S, const double* __restrict__ A, const double* __restrict__ B, long long batch, int N,
             int q, int step, int steps, double4 x_eq, double u_eq, double4 wbar, int use_w,
             unsigned long long seed, unsigned long long scen0, const double* __restrict__ Kfb = nullptr) {
    // Kfb != nullptr: F-form loop (ocpLBMPC.m:10-47 / ocpLMPC.m:11-40): the decision variables are c, the plant input is
    // u = K (x - x_wp) + c + u_wp (transitionTrue.m:11-12), the next solve starts from the UNSHIFTED opt_var (ocpLBMPC.m:31) and the
    // data window follows update_data.m:3-10 (appends while the 1-based iteration index is below q, then drops the oldest column:
    // with the zero column the scripts start from it holds at most q - 1 samples).
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const double xe[4] = {x_eq.x, x_eq.y, x_eq.z, x_eq.w}, wb[4] = {wbar.x, wbar.y, wbar.z, wbar.w};
    double x[4], dx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        x[j] = S.x[b * 4 + j];
        dx[j] = x[j] - xe[j];
    }
    // no optimal solution (infeasible / iteration cap): keep executing the last optimal plan = the shifted previous plan that
    // was this step's warm start (zero before any plan exists); the same rule as oracle/lbmpc_oracle.c lbo_closed_loop
    const bool ok = S.status[b] == 0;
    double* wm = S.warm + b * (N + 1);
    const double* plan = ok ? S.uc + b * N : wm;
    const double th_plan = ok ? S.theta[b] : wm[N];
    double du0 = plan[0];
    if (Kfb) {
#pragma unroll
        for (int j = 0; j < 4; ++j) du0 += Kfb[j] * dx[j];
    }
    const double u0 = du0 + u_eq;
    double k1[4], k2[4], k3[4], k4[4], t[4], xn[4];
    const double delta = 0.01;
    mg_rhs(x, u0, k1);
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = x[i] + delta / 2 * k1[i];
    mg_rhs(t, u0, k2);
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = x[i] + delta / 2 * k2[i];
    mg_rhs(t, u0, k3);
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = x[i] + delta * k3[i];
    mg_rhs(t, u0, k4);
#pragma unroll
    for (int i = 0; i < 4; ++i) xn[i] = x[i] + delta / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    if (use_w) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            xn[j] += wb[j] * (2.0 * lb_uniform(seed, scen0 + (unsigned long long)b, (unsigned long long)step,
                                               (unsigned long long)j) - 1.0);
    }
    // data acquisition: X = [dx1;dx2;du], Y = dx+ - (A dx + B du)
    double ys[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double v = xn[a] - xe[a];
#pragma unroll
        for (int j = 0; j < 4; ++j) v -= A[a * 4 + j] * dx[j];
        v -= B[a] * du0;
        ys[a] = v;
    }
    const double xs[3] = {dx[0], dx[1], du0};
    // the window is a RING: sample number nd goes to slot nd mod capacity (the Nadaraya-Watson sums do not depend on the order of the
    // samples, so dropping the oldest one — get_data.m:8 / update_data.m:8 — is overwriting it; no O(q) shift per step and scenario)
    const int cap = Kfb ? (q > 1 ? q - 1 : 1) : q;
    const int nd = S.nd[b], slot = nd % cap;
    double* Xb = S.X + b * 3 * q;
    double* Yb = S.Y + b * 4 * q;
#pragma unroll
    for (int a = 0; a < 3; ++a) Xb[slot * 3 + a] = xs[a];
#pragma unroll
    for (int a = 0; a < 4; ++a) Yb[slot * 4 + a] = ys[a];
    S.V[b * q + slot] = 1.0;
    S.nd[b] = nd + 1;
    // warm-start shift (in place when the plan is the previous warm start)
    if (Kfb) {  // F-form: opt_var is reused as it is
        if (ok)
            for (int k = 0; k < N; ++k) wm[k] = plan[k];
    } else {
        const double last = plan[N - 1];
        for (int k = 0; k + 1 < N; ++k) wm[k] = plan[k + 1];
        wm[N - 1] = last;
    }
    wm[N] = th_plan;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        S.x[b * 4 + j] = xn[j];
        S.dx0[b * 4 + j] = xn[j] - xe[j];
        if (S.x_hist) S.x_hist[(b * (steps + 1) + step + 1) * 4 + j] = xn[j];
    }
    if (S.u_hist) S.u_hist[b * steps + step] = u0;
    if (S.theta_hist) S.theta_hist[b * steps + step] = th_plan;
    if (S.iters_hist) S.iters_hist[b * steps + step] = S.iters[b];
    if (S.status_hist) S.status_hist[b * steps + step] = S.status[b];
}

__global__ void loop_init_kernel(LoopState S, const double* __restrict__ x_init, long long batch, int steps,
                                 double4 x_eq, int fform = 0, int q = 0) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const double xe[4] = {x_eq.x, x_eq.y, x_eq.z, x_eq.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double v = x_init[b * 4 + j];
        S.x[b * 4 + j] = v;
        S.dx0[b * 4 + j] = v - xe[j];
        if (S.x_hist) S.x_hist[(b * (steps + 1)) * 4 + j] = v;
    }
    S.nd[b] = fform ? 1 : 0;   // F-form scripts start from one all-zero sample (LBMPC_RunExample.m: data.X = zeros(3,1), data.Y = zeros(4,1))
    if (fform) S.V[b * q] = 1.0;
}

// ---------------------------------------------------------------------------------------------
// outer (SQP) iteration bookkeeping: step = |u - u_lin|inf, u_lin <- u, next warm start <- [u; theta]   (warp per QP)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
sqp_update_kernel(long long batch, int N, int nt, const double* __restrict__ uc, const double* __restrict__ theta,
                  double* __restrict__ ulin, double* __restrict__ warm2, double* __restrict__ step, int step_ld, int j) {
    const int lane = threadIdx.x & 31;
    const long long qp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qp >= batch) return;
    double mx = 0.0;
    for (int k = lane; k < N; k += 32) {
        const double u = uc[qp * N + k];
        mx = fmax(mx, fabs(u - ulin[qp * N + k]));
        ulin[qp * N + k] = u;
        warm2[qp * (N + nt) + k] = u;
    }
    if (lane < nt) warm2[qp * (N + nt) + N + lane] = theta[qp * nt + lane];
    mx = warp_max(mx);
    if (lane == 0 && step) step[qp * step_ld + j] = mx;
}

// ---------------------------------------------------------------------------------------------
// twin state sequences (DMS_LBMPC_casadi.m:283-319): gap between the learned and the nominal state sequence for frozen
// oracle corrections d_k:  e_0 = 0, e_{k+1} = A e_k + d_k   (thread per QP; feeds BatchIO::cshift)
// ---------------------------------------------------------------------------------------------
template <int NX>
__global__ void __launch_bounds__(128)
twin_shift_kernel(long long batch, int N, const double* __restrict__ A, const double* __restrict__ B, const double* __restrict__ K,
                  const double* __restrict__ d_off, double* __restrict__ csh) {
    // e_{k+1} = (A + B K) e_k + d_k, eu_k = K e_k; records [ex (NX) | eu (1)] (K = 0 for the C-form: plain A, eu = 0)
    const long long qp = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (qp >= batch) return;
    double a[NX * NX], kk[NX], e[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) kk[i] = K[i];
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
        for (int j = 0; j < NX; ++j) a[i * NX + j] = A[i * NX + j] + B[i] * kk[j];
#pragma unroll
    for (int i = 0; i < NX; ++i) e[i] = 0.0;
    const double* d = d_off + qp * (long long)N * NX;
    double* o = csh + qp * (long long)(N + 1) * (NX + 1);
    for (int k = 0;; ++k) {
        double eu = 0.0;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            o[k * (NX + 1) + i] = e[i];
            eu += kk[i] * e[i];
        }
        o[k * (NX + 1) + NX] = eu;
        if (k == N) break;
        double n[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double v = d[k * NX + i];
#pragma unroll
            for (int j = 0; j < NX; ++j) v += a[i * NX + j] * e[j];
            n[i] = v;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) e[i] = n[i];
    }
}

// ---------------------------------------------------------------------------------------------
// FP64-FMA peak microbenchmark (roofline denominator; MEASURED_PEAKS.json has no FP64 entry):
// 8 independent register-resident DFMA chains per thread.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
            c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
        }
    }
    const double s = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
    if (s == 123.456) out[0] = s;  // keep the chains alive
}

}  // namespace lbmpc
