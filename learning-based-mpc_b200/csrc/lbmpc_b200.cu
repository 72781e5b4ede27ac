// lbmpc_b200.cu — C ABI (include/lbmpc.h) over the sm_100a kernels.  One translation unit:
// nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC
//
// No CPU fallback: every entry point needs a CUDA device; errors come back as negative codes with
// a thread-local message (lbmpc_last_error).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "../../include/lbmpc.h"
#include "lbmpc_kernels.cuh"
#include "lbmpc_kernel_cta.cuh"
#include "lbmpc_problem.hpp"
#include "lbmpc_stream.cuh"

using namespace lbmpc;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CU_TRY(expr)                                                                                  \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail(LBMPC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));            \
    } while (0)

struct LoopScratch {
    int64_t batch = 0;
    int q = 0, steps = 0;
    double *x = nullptr, *dx0 = nullptr, *X = nullptr, *Y = nullptr, *V = nullptr, *warm = nullptr, *uc = nullptr,
           *theta = nullptr, *obj = nullptr, *doff = nullptr, *x_init = nullptr;
    int *nd = nullptr, *iters = nullptr, *status = nullptr;
    double *x_hist = nullptr, *u_hist = nullptr, *t_hist = nullptr;
    int *i_hist = nullptr, *s_hist = nullptr;
};

struct lbmpc_handle {
    int device = 0;
    HostProblem hp;
    int shape = 0;  // 0: <4,1,1>  1: <2,2,2>
    int max_slots = 0, stage_g = 0, num_sms = 0;
    bool poly_global = false;          // warp mapping: polytope slacks / multipliers in a global array (large sets)
    double* dpoly = nullptr;
    int cta_blocks_per_sm[2] = {0, 0};  // [0] > 0: the CTA-per-QP latency kernel (4 warps per QP) is available: resident CTAs per SM
    size_t cta_smem = 0;
    int cta_warps_force = 0;           // 2 / 4: warps per QP of the CTA kernel (LBMPC_CTA_WARPS, experiments)
    bool cta_big = false;              // large polytope block: G stays in global memory, sums by block reduction
    bool dev_ptrs = false;
    int64_t max_batch = 0;
    double *dG = nullptr, *dhg = nullptr, *dA = nullptr, *dB = nullptr, *dK = nullptr;  // dK: Kinit (zero for the C-form)
    unsigned long long* dqueue = nullptr;   // ring of kQueueRing work counters: every launch takes the next one, so a launch
                                            // that is still running never sees its counter reset by the following call
    int64_t queue_next = 0;
    int last_kernel = 0;
    unsigned long long* dprof = nullptr;  // 8 counters, enabled by lbmpc_debug_phase_cycles
    bool prof_on = false;
    // host-pointer staging
    double *s_dx0 = nullptr, *s_ref = nullptr, *s_doff = nullptr, *s_warm = nullptr, *s_uc = nullptr,
           *s_x = nullptr, *s_csh = nullptr;
    // theta | obj | iters | status of a call, packed back to back for the ACTUAL batch: one device-to-host copy into a
    // pinned host block instead of four small ones, scattered to the caller's arrays after the stream synchronises
    char *s_small = nullptr, *hs_small = nullptr;
    char* hs_bounce = nullptr;  // pinned + mapped: inputs / outputs of small host-pointer calls (kBounceBytes)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    bool timed = false;
    LoopScratch loop;
    // outer-iteration (SQP) scratch, grown on demand
    int64_t sqp_batch = 0;
    int sqp_iters = 0;
    double *q_ulin = nullptr, *q_warm = nullptr, *q_doff = nullptr, *q_step = nullptr, *q_csh = nullptr, *q_jac = nullptr;
    // staging of the host-pointer oracle / SQP calls (data windows, input sequences): grown on first use, then reused — no
    // allocation in the steady state of a closed loop
    double* og[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t og_cap[6] = {0, 0, 0, 0, 0, 0};
    // kernel choice (lbmpc_set_kernel; LBMPC_KERNEL / LBMPC_LOCKSTEP / LBMPC_STREAM_* are read ONCE, in lbmpc_create)
    int force_kernel = LBMPC_KERNEL_AUTO, force_lockstep = -1;
    // stream kernel (one thread per QP, iterate in HBM): workspace of the resident warps
    int st_ctas_per_sm = 0;            // resident 128-thread CTAs per SM (default layout)
    int st_warps_cap = 0;              // 6: use the 6-warp variant everywhere (LBMPC_STREAM_WARPS, experiments)
    int max_smem_optin = 0;
    bool st_evict_forced = false;      // LBMPC_STREAM_EVICT set: budget also when the mapping is forced (experiments)
    int st_evict_iters = 12;           // iteration budget of the stream mapping when the engine picks it for long horizons (0: none)
    long long* st_left = nullptr;      // [0]: count, [1..]: QPs handed over
    size_t st_left_cap = 0;
    int64_t st_min_batch_long = 0;     // long horizons (N > 100): stream + hand-over picked from this batch on
    bool st_spread = true;             // small stream launches use 2 / 4 warps per CTA so that every SM takes part (LBMPC_STREAM_SPREAD=0: always 8)
    bool loop_allow_stream = true;     // closed-loop steps may take the stream mapping (LBMPC_LOOP_ALLOW_STREAM=0: never)
    int64_t st_loop_min_batch = 0;     // fused closed loop picked automatically from this many scenarios (0: only when forced)
    int st_loop_chunk = 10;            // control steps a lane runs before it hands the scenario back to the queue
    double* lp_store = nullptr;        // fused closed loop: scenario state between chunks
    long long* lp_rq = nullptr;        // [0]: tail counter, [1..]: re-queued scenarios
    size_t lp_store_cap = 0, lp_rq_cap = 0;
    int64_t st_min_batch = 0;          // auto choice: batches at least this large take the stream kernel
    double* st_ws64 = nullptr;
    void* st_wsft = nullptr;
    size_t st_ws64_bytes = 0, st_wsft_bytes = 0;
};

constexpr int kQueueRing = 64;
static unsigned long long* next_queue(lbmpc_handle* h) { return h->dqueue + (h->queue_next++ % kQueueRing); }

template <int NX, int NT, int NU>
static int plan_slots(lbmpc_handle* h, size_t max_smem) {
    const HostProblem& hp = h->hp;
    // large polytope block: its slacks / multipliers (2 ngp doubles per QP) live in a global array, not in the slot
    h->poly_global = hp.ng > 64;
    if (const char* e = getenv("LBMPC_POLY_GLOBAL")) h->poly_global = atoi(e) != 0 && hp.ng > 0;
    for (int stage = 1; stage >= 0; --stage) {
        for (int s = kMaxSlots; s >= 1; --s) {
            const SmemPlan<NX, NT, NU> plan(hp.N, hp.ngp, s, stage != 0, h->poly_global);
            // staging the polytope must not cost more than one slot of residency
            if (plan.bytes <= max_smem) {
                if (stage == 1) {
                    const SmemPlan<NX, NT, NU> alt(hp.N, hp.ngp, std::min(s + 1, kMaxSlots), false, h->poly_global);
                    // one more resident QP beats a staged block: tiny blocks stay in L1, and for the 616-row set (slacks / multipliers
                    // in the global array) 8 QPs per SM reading G from L2 run 7.06 ms at batch 16384 against 7.52 ms with 7 QPs and G staged
                    if (s < kMaxSlots && alt.bytes <= max_smem) continue;
                }
                h->max_slots = s;
                h->stage_g = stage;
                return LBMPC_OK;
            }
        }
    }
    return fail(LBMPC_ESHAPE, "horizon too long: one QP does not fit in shared memory");
}

template <int NX, int NT, int NU>
static cudaError_t launch_ipm(lbmpc_handle* h, const BatchIO& io_in, cudaStream_t st) {
    const HostProblem& hp = h->hp;
    BatchIO io = io_in;
    const Params<NX, NT, NU> p = to_params<NX, NT, NU>(hp);
    // (hand-over launches — io.qlist — read their QP count on the device: sized for a fraction of the batch)
    const int64_t nq = io.qlist ? std::max<int64_t>(io.batch / 8, 1) : io.batch;
    int slots = (int)std::min<int64_t>(h->max_slots, (nq + h->num_sms - 1) / h->num_sms);
    slots = std::max(slots, 1);
    const int grid = (int)std::min<int64_t>(h->num_sms, (nq + slots - 1) / slots);
    const SmemPlan<NX, NT, NU> plan(hp.N, hp.ngp, slots, h->stage_g != 0, h->poly_global);
    if (h->poly_global && !h->dpoly) {  // one slab for the largest launch shape: num_sms CTAs x max_slots warps
        cudaError_t ea = cudaMalloc((void**)&h->dpoly, sizeof(double) * (size_t)h->num_sms * kMaxSlots * 2 * hp.ngp);
        if (ea != cudaSuccess) return ea;
    }
    // many QPs per warp slot: the warps of a CTA start their iterations together (shared instruction fetches, see cta_tick)
    io.lockstep = slots >= 4 && nq >= (int64_t)3 * grid * slots;
    if (h->force_lockstep >= 0) io.lockstep = h->force_lockstep;
    io.queue = next_queue(h);
    cudaError_t e = cudaMemsetAsync(io.queue, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    ipm_kernel<NX, NT, NU><<<grid, 32 * slots, plan.bytes, st>>>(p, io, h->dG, h->dhg, slots, h->stage_g, h->poly_global ? h->dpoly : nullptr);
    h->launches += 1;
    h->last_kernel = LBMPC_KERNEL_WARP;
    return cudaGetLastError();
}

// Few QPs per SM: the launch lasts as long as its slowest QP, so the latency variant (one CTA per QP) wins; many QPs
// per SM: the warp-per-QP kernel keeps more QPs resident.  LBMPC_KERNEL=warp|cta overrides (experiments).
// returns LBMPC_KERNEL_WARP / _CTA / _STREAM / _STREAM_MIXED
static int pick_kernel(const lbmpc_handle* h, int64_t batch, bool allow_stream = true) {
    const bool cta_ok = h->shape == 0 && h->cta_blocks_per_sm[0] > 0, stream_ok = allow_stream && h->shape == 0 && h->st_ctas_per_sm > 0;
    if (h->force_kernel == LBMPC_KERNEL_WARP) return LBMPC_KERNEL_WARP;
    if (h->force_kernel == LBMPC_KERNEL_CTA) return cta_ok ? LBMPC_KERNEL_CTA : LBMPC_KERNEL_WARP;
    if (h->force_kernel == LBMPC_KERNEL_STREAM || h->force_kernel == LBMPC_KERNEL_STREAM_MIXED)
        return stream_ok ? h->force_kernel : LBMPC_KERNEL_WARP;
    // very many QPs per SM, small polytope block: one thread per QP with the iterate streamed from HBM, an iteration budget and a
    // hand-over of the slow QPs to a shared-memory mapping (one iteration of a lane lasts ~0.5 ms x N/50, so the 20 - 40
    // iteration QPs would otherwise hold the launch).  Thread-local loops over a 616-row set lose to the CTA mapping.
    if (stream_ok && h->st_min_batch > 0 && batch >= h->st_min_batch && h->hp.ng <= 64 && h->hp.N <= 100) return LBMPC_KERNEL_STREAM;
    // long horizons: shared memory holds 1 - 2 QPs per SM, the stream mapping 256; its slow iterations (2 ms at N = 200) are
    // bounded by an iteration budget, the QPs beyond it are handed to the CTA mapping (launch_ipm_stream, evict)
    if (stream_ok && h->st_min_batch_long > 0 && batch >= h->st_min_batch_long && h->hp.ng <= 64 && h->hp.N > 100) return LBMPC_KERNEL_STREAM;
    // 616-row set: the thread-local row loops pay off at large batches only (warp mapping 26.6 / 39.5 ms vs stream 29.0 / 38.0 / 45.3 ms
    // at batch 65536 / 98304 / 131072)
    if (stream_ok && h->st_min_batch > 0 && batch >= 3 * h->st_min_batch && h->hp.ng > 64 && h->hp.N <= 100) return LBMPC_KERNEL_STREAM;
    if (!cta_ok) return LBMPC_KERNEL_WARP;
    // measured on B200 (C-form, N = 50).  24-row polytope (LBMPC): one CTA per QP wins while every QP is resident (4 CTAs per SM:
    // 1.24x at 1 QP/SM, 1.13x at 4); beyond that the QPs that queue behind the resident CTAs cost more than the faster iterations gain.  616-row
    // polytope (LMPC): the row phases dominate an iteration, the CTA kernel wins 1.5x at 1 QP/SM, 1.2x at 7, 1.1x at 24 and is
    // level from ~80 QPs/SM on, where the warp mapping is 2 - 4 % ahead.
    // 616-row set, with the polytope slacks / multipliers of the warp mapping in a global array (7 instead of 5 QPs per SM):
    // CTA 0.81 / 1.34 / 2.39 / 4.42 / 8.6 ms vs warp 0.88 / 1.35 / 2.28 / - / 7.06 ms at batch 1024 / 2048 / 4096 / 8192 / 16384
    if (h->cta_big) return batch >= (int64_t)h->num_sms * 20 ? LBMPC_KERNEL_WARP : LBMPC_KERNEL_CTA;
    if (batch <= (int64_t)h->num_sms * h->cta_blocks_per_sm[0]) return LBMPC_KERNEL_CTA;
    // (a two-warps-per-QP CTA variant exists — LBMPC_CTA_WARPS=2 — but 35 KB of shared memory per CTA keep it at 6 CTAs per SM,
    //  and even four warps per QP gain only 9 % over the warp mapping at 4 QPs per SM: not picked automatically)
    // long horizons: shared memory holds only 1-2 QPs per SM either way, so the four warps of a CTA are free (N = 200: 1.18x)
    if (h->max_slots <= 2 && h->cta_blocks_per_sm[0] >= h->max_slots) return LBMPC_KERNEL_CTA;
    return LBMPC_KERNEL_WARP;
}
static cudaError_t launch_ipm_cta(lbmpc_handle* h, const BatchIO& io_in, cudaStream_t st) {
    const Params<4, 1, 1> p = to_params<4, 1, 1>(h->hp);
    BatchIO io = io_in;
    // four warps per QP; two warps per QP only when forced (experiments)
    const bool two = !h->cta_big && h->cta_blocks_per_sm[1] > 0 && h->cta_warps_force == 2;
    const int grid = (int)std::min<int64_t>((int64_t)h->num_sms * h->cta_blocks_per_sm[two ? 1 : 0], io.batch);  // hand-over launches: CTAs without work leave at once
    io.queue = next_queue(h);
    cudaError_t e = cudaMemsetAsync(io.queue, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    if (h->cta_big) ipm_kernel_cta<4, 1, 1, 4, true><<<grid, 128, h->cta_smem, st>>>(p, io, h->dG, h->dhg);
    else if (two) ipm_kernel_cta<4, 1, 1, 2, false><<<grid, 64, h->cta_smem, st>>>(p, io, h->dG, h->dhg);
    else ipm_kernel_cta<4, 1, 1, 4, false><<<grid, 128, h->cta_smem, st>>>(p, io, h->dG, h->dhg);
    h->launches += 1;
    h->last_kernel = LBMPC_KERNEL_CTA;
    return cudaGetLastError();
}

// ---- stream kernel: one thread per QP, workspace of the resident warps in HBM ----
static cudaError_t stream_workspace(lbmpc_handle* h, int64_t warps, const StreamLayout<4>& l, bool mixed) {
    const size_t need64 = (size_t)warps * l.n64 * 32 * sizeof(double);
    const size_t needft = (size_t)warps * l.nft * 32 * (mixed ? sizeof(float) : sizeof(double));
    if (need64 > h->st_ws64_bytes) {  // grown only when a call brings a larger layout than the one sized at create
        cudaFree(h->st_ws64);
        h->st_ws64 = nullptr; h->st_ws64_bytes = 0;
        cudaError_t e = cudaMalloc((void**)&h->st_ws64, need64);
        if (e != cudaSuccess) return e;
        h->st_ws64_bytes = need64;
    }
    if (needft > h->st_wsft_bytes) {
        cudaFree(h->st_wsft);
        h->st_wsft = nullptr; h->st_wsft_bytes = 0;
        cudaError_t e = cudaMalloc(&h->st_wsft, needft);
        if (e != cudaSuccess) return e;
        h->st_wsft_bytes = needft;
    }
    return cudaSuccess;
}
static cudaError_t launch_ipm_cta(lbmpc_handle* h, const BatchIO& io_in, cudaStream_t st);
template <int NX, int NT, int NU>
static cudaError_t launch_ipm(lbmpc_handle* h, const BatchIO& io_in, cudaStream_t st);

template <bool LTV, typename FT, int WARPS>
static cudaError_t launch_stream_w(lbmpc_handle* h, const BatchIO& io, const double* jac, cudaStream_t st, size_t smem, bool evict = false) {
    const Params<4, 1, 1> p = to_params<4, 1, 1>(h->hp);
    const bool cs = io.cshift != nullptr;
    const StreamLayout<4> l(p.N, p.ng, cs, LTV);
    const int64_t ctas = std::max<int64_t>(1, std::min<int64_t>(h->num_sms, (io.batch + 32 * WARPS - 1) / (32 * WARPS)));
    cudaError_t e = stream_workspace(h, ctas * WARPS, l, sizeof(FT) == 4);
    if (e != cudaSuccess) return e;
    StreamIO<FT> s{};
    s.batch = io.batch; s.dx0 = io.dx0; s.dx_ref = io.dx_ref; s.d_off = io.d_off; s.warm = io.warm; s.cshift = io.cshift;
    s.cs_stride = io.cs_stride; s.row_shift = io.row_shift; s.jac = jac; s.uc = io.uc; s.theta = io.theta; s.xtraj = io.xtraj; s.obj = io.obj; s.iters = io.iters; s.status = io.status;
    s.queue = next_queue(h);
    s.ws64 = h->st_ws64;
    s.wsft = (FT*)h->st_wsft;
    e = cudaMemsetAsync(s.queue, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    if (evict) {  // iteration budget: the QPs still running after it go to a shared-memory mapping (below)
        if (h->st_left_cap < (size_t)io.batch + 1) {
            cudaFree(h->st_left); h->st_left = nullptr; h->st_left_cap = 0;
            e = cudaMalloc((void**)&h->st_left, sizeof(long long) * ((size_t)std::max<int64_t>(io.batch, h->max_batch) + 1));
            if (e != cudaSuccess) return e;
            h->st_left_cap = (size_t)std::max<int64_t>(io.batch, h->max_batch) + 1;
        }
        e = cudaMemsetAsync(h->st_left, 0, sizeof(long long), st);
        if (e != cudaSuccess) return e;
        s.evict_iters = h->st_evict_iters;
        s.left_count = reinterpret_cast<unsigned long long*>(h->st_left);
        s.left_list = h->st_left + 1;
    }
    ipm_stream_kernel<4, LTV, FT, WARPS><<<(unsigned)ctas, 32 * WARPS, smem, st>>>(p, s, h->dG, h->dhg);
    h->launches += 1;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (evict) {  // same inputs / outputs, QP indices from the hand-over list, count read on the device
        BatchIO lo = io;
        lo.qlist = h->st_left + 1;
        lo.qcount = reinterpret_cast<const unsigned long long*>(h->st_left);
        const bool cta = h->cta_blocks_per_sm[0] > 0 && (h->cta_big || (h->max_slots <= 2 && h->cta_blocks_per_sm[0] >= h->max_slots));
        e = cta ? launch_ipm_cta(h, lo, st) : launch_ipm<4, 1, 1>(h, lo, st);
        if (e != cudaSuccess) return e;
    }
    h->last_kernel = sizeof(FT) == 4 ? LBMPC_KERNEL_STREAM_MIXED : LBMPC_KERNEL_STREAM;
    return cudaSuccess;
}
// warps per CTA (= per SM): 8 (255 registers per thread) when their double buffers fit shared memory, else 6 (per-stage
// Jacobians and a shift record together make the buffers 15 KB).  More than 8 warps means three warps on some scheduler, i.e.
// at most 168 registers per thread: 10 and 12 warps both measured 1.3 - 1.5x SLOWER (1.6 - 2 KB of spills in pass BU).
template <bool LTV, typename FT>
static cudaError_t launch_stream_t(lbmpc_handle* h, const BatchIO& io, const double* jac, cudaStream_t st, bool evict) {
    const size_t per_warp = StreamSmem<4, FT>::warp_bytes(io.cshift != nullptr, LTV), poly = StreamSmem<4, FT>::poly_bytes(h->hp.ngp);
    const bool w8 = 8 * per_warp + poly <= (size_t)h->max_smem_optin && h->st_warps_cap != 6;
    // fewer QPs than resident lanes: spread them over ALL SMs with fewer warps per CTA instead of filling a few SMs with eight
    // (a lane's iteration is latency-bound; with 2 warps on an SM instead of 8 it runs ~2.5x faster)
    if (sizeof(FT) == 8 && h->st_spread && 4 * per_warp + poly <= (size_t)h->max_smem_optin) {
        const int64_t wps = (io.batch + 32 * (int64_t)h->num_sms - 1) / (32 * (int64_t)h->num_sms);
        if (wps <= 2) return launch_stream_w<LTV, FT, 2>(h, io, jac, st, 2 * per_warp + poly, evict);
        if (wps <= 4) return launch_stream_w<LTV, FT, 4>(h, io, jac, st, 4 * per_warp + poly, evict);
    }
    return w8 ? launch_stream_w<LTV, FT, 8>(h, io, jac, st, 8 * per_warp + poly, evict) : launch_stream_w<LTV, FT, 6>(h, io, jac, st, 6 * per_warp + poly, evict);
}
// evict: iteration budget + hand-over (only for plain QPs: no per-stage dynamics, no row shift — the shared-memory mappings
// have neither)
static cudaError_t launch_ipm_stream(lbmpc_handle* h, const BatchIO& io, const double* jac, cudaStream_t st, bool mixed, bool evict = false) {
    if (io.row_shift && !jac) return cudaErrorInvalidValue;  // row shifts are compiled into the LTV kernel only
    evict = evict && !jac && !io.row_shift && h->st_evict_iters > 0;
    if (mixed) return jac ? launch_stream_t<true, float>(h, io, jac, st, false) : launch_stream_t<false, float>(h, io, jac, st, false);
    return jac ? launch_stream_t<true, double>(h, io, jac, st, false) : launch_stream_t<false, double>(h, io, jac, st, evict);
}

static void launch_oracle(lbmpc_handle* h, cudaStream_t st, long long batch, int q, double inv_h2, double lambda,
                          const double* dx0, const double* du, long long du_ld, const double* X, const double* Y,
                          const double* valid, double* d_off) {
    const unsigned grid = (unsigned)((batch + 3) / 4);
    const double* Kfb = h->hp.form == LBMPC_FORM_F ? h->dK : nullptr;  // F-form sequences hold c: u = K x + c
    if (q <= 128)
        oracle_kernel<4, 1, 4><<<grid, 128, 0, st>>>(h->dA, h->dB, Kfb, h->hp.N, batch, q, inv_h2, lambda, dx0, du, du_ld, X,
                                                    Y, valid, d_off);
    else
        oracle_kernel<4, 1, kOracleMaxPerLane><<<grid, 128, 0, st>>>(h->dA, h->dB, Kfb, h->hp.N, batch, q, inv_h2, lambda,
                                                                    dx0, du, du_ld, X, Y, valid, d_off);
    h->launches += 1;
}

// allow_stream: closed-loop steps under disturbance hand a fifth of their QPs (infeasible or slow) over to the shared-memory
// mapping; measured at 125 k scenarios x 50 steps: 4.50 M QP/s with the stream mapping + hand-over, 3.73 M with the warp mapping,
// 3.11 M with the fused loop kernel (which has no iteration budget)
static cudaError_t launch_ipm_any(lbmpc_handle* h, const BatchIO& io, cudaStream_t st, const double* jac = nullptr, bool allow_stream = true) {
    if (h->shape == 0) {
        const int w = jac ? (h->force_kernel == LBMPC_KERNEL_STREAM_MIXED ? LBMPC_KERNEL_STREAM_MIXED : LBMPC_KERNEL_STREAM)
                          : pick_kernel(h, io.batch, allow_stream);  // per-stage dynamics (LTV) exist in the stream mapping only
        if (w == LBMPC_KERNEL_CTA) return launch_ipm_cta(h, io, st);
        if (w == LBMPC_KERNEL_STREAM || w == LBMPC_KERNEL_STREAM_MIXED)
            return launch_ipm_stream(h, io, jac, st, w == LBMPC_KERNEL_STREAM_MIXED,
                                     /*evict=*/h->force_kernel == LBMPC_KERNEL_AUTO || h->st_evict_forced);
    }
    return h->shape == 0 ? launch_ipm<4, 1, 1>(h, io, st) : launch_ipm<2, 2, 2>(h, io, st);
}

constexpr size_t kBounceBytes = 512 * 1024;

struct SmallOut {
    double *theta, *obj;
    int *iters, *status;
    size_t bytes;
};
static SmallOut small_out(char* base, size_t b, size_t nt) {
    SmallOut o;
    o.theta = reinterpret_cast<double*>(base);
    o.obj = o.theta + b * nt;
    o.iters = reinterpret_cast<int*>(o.obj + b);
    o.status = o.iters + b;
    o.bytes = sizeof(double) * b * (nt + 1) + sizeof(int) * 2 * b;
    return o;
}
static void scatter_small(const lbmpc_handle* h, size_t b, size_t nt, double* theta, double* obj, int32_t* iters, int32_t* status) {
    const SmallOut o = small_out(h->hs_small, b, nt);
    memcpy(theta, o.theta, sizeof(double) * b * nt);
    memcpy(obj, o.obj, sizeof(double) * b);
    memcpy(iters, o.iters, sizeof(int) * b);
    memcpy(status, o.status, sizeof(int) * b);
}

// Page-locked (pinned) caller arrays are mapped into the device's address space: the kernel then reads its 32-byte
// initial state / writes its results through PCIe itself, while other QPs are still being solved, instead of waiting for
// copy engines before and after the launch.  Returns the device alias of a pinned host pointer, nullptr for pageable memory.
template <typename T>
static T* pinned_alias(T* host) {
    if (!host) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? reinterpret_cast<T*>(a.devicePointer) : nullptr;
}

template <typename T>
static cudaError_t dmalloc(T** p, size_t n) {
    return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T));
}
// handle-owned staging buffer i with room for n doubles (grows, never shrinks)
static cudaError_t staging(lbmpc_handle* h, int i, size_t n, double** out) {
    if (h->og_cap[i] < n) {
        cudaFree(h->og[i]);
        h->og[i] = nullptr; h->og_cap[i] = 0;
        const size_t cap = std::max<size_t>(n, (size_t)h->max_batch * 4);
        cudaError_t e = cudaMalloc((void**)&h->og[i], cap * sizeof(double));
        if (e != cudaSuccess) return e;
        h->og_cap[i] = cap;
    }
    *out = h->og[i];
    return cudaSuccess;
}

extern "C" {

const char* lbmpc_last_error(void) { return g_err.c_str(); }
const char* lbmpc_version(void) { return "lbmpc_b200 0.1 (sm_100a)"; }

int lbmpc_create(const lbmpc_model* model, const lbmpc_config* cfg, int device, lbmpc_handle** out) {
    if (!out) return fail(LBMPC_EINVAL, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(LBMPC_ECUDA, "no CUDA device available (this engine has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(LBMPC_EINVAL, "bad device index");
    lbmpc_handle* h = new (std::nothrow) lbmpc_handle();
    if (!h) return fail(LBMPC_ENOMEM, "out of host memory");
    struct Guard {  // every early return below releases what has been allocated so far (lbmpc_destroy tolerates partial handles)
        lbmpc_handle* h;
        ~Guard() { if (h) lbmpc_destroy(h); }
    } guard{h};
    std::string err;
    int rc = build_problem(model, cfg, h->hp, err);
    if (rc != LBMPC_OK) return fail(rc, err);
    const HostProblem& hp = h->hp;
    if (hp.nx == 4 && hp.nt == 1 && hp.nu == 1) h->shape = 0;
    else if (hp.nx == 2 && hp.nt == 2 && hp.nu == 2) h->shape = 1;
    else return fail(LBMPC_ESHAPE, "compiled shapes: (nx,nt,nu) = (4,1,1) Moore-Greitzer, (2,2,2) double integrator");
    h->device = device;
    h->dev_ptrs = cfg->pointers_on_device != 0;
    h->max_batch = std::max<int64_t>(cfg->max_batch, 1);
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(LBMPC_ECUDA, "device is not sm_100 (Blackwell); the library carries sm_100a code only");
    h->num_sms = prop.multiProcessorCount;
    int max_smem = 0;
    CU_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    rc = h->shape == 0 ? plan_slots<4, 1, 1>(h, (size_t)max_smem) : plan_slots<2, 2, 2>(h, (size_t)max_smem);
    if (rc != LBMPC_OK) return rc;
    if (h->shape == 0)
        CU_TRY(cudaFuncSetAttribute(ipm_kernel<4, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    else
        CU_TRY(cudaFuncSetAttribute(ipm_kernel<2, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    if (h->shape == 0) {  // latency variant: one CTA per QP
        h->cta_big = hp.ng > 64;
        const CtaPlan<4, 1, 1> cplan(hp.N, hp.ngp, h->cta_big);
        if (cplan.bytes <= (size_t)max_smem) {
            if (h->cta_big) {
                CU_TRY(cudaFuncSetAttribute(ipm_kernel_cta<4, 1, 1, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cplan.bytes));
                CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->cta_blocks_per_sm[0], ipm_kernel_cta<4, 1, 1, 4, true>, 128, cplan.bytes));
            } else {
                CU_TRY(cudaFuncSetAttribute(ipm_kernel_cta<4, 1, 1, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cplan.bytes));
                CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->cta_blocks_per_sm[0], ipm_kernel_cta<4, 1, 1, 4, false>, 128, cplan.bytes));
                CU_TRY(cudaFuncSetAttribute(ipm_kernel_cta<4, 1, 1, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cplan.bytes));
                CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->cta_blocks_per_sm[1], ipm_kernel_cta<4, 1, 1, 2, false>, 64, cplan.bytes));
                if (const char* e = getenv("LBMPC_CTA_WARPS")) h->cta_warps_force = atoi(e);
            }
            h->cta_smem = cplan.bytes;
        }
    }
    const int nz = hp.nx + hp.nt;
    CU_TRY(dmalloc(&h->dG, (size_t)nz * hp.ngp));
    CU_TRY(dmalloc(&h->dhg, (size_t)hp.ngp));
    CU_TRY(dmalloc(&h->dA, (size_t)hp.nx * hp.nx));
    CU_TRY(dmalloc(&h->dB, (size_t)hp.nx * hp.nu));
    CU_TRY(dmalloc(&h->dK, (size_t)hp.nx * hp.nu));
    CU_TRY(cudaMemcpy(h->dK, hp.Kinit.data(), sizeof(double) * hp.nx * hp.nu, cudaMemcpyHostToDevice));
    CU_TRY(dmalloc(&h->dqueue, kQueueRing));
    CU_TRY(dmalloc(&h->dprof, 16));
    CU_TRY(cudaMemset(h->dprof, 0, 16 * sizeof(unsigned long long)));
    CU_TRY(cudaMemcpy(h->dG, hp.G.data(), sizeof(double) * nz * hp.ngp, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->dhg, hp.hg.data(), sizeof(double) * hp.ngp, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->dA, hp.A.data(), sizeof(double) * hp.nx * hp.nx, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->dB, hp.B.data(), sizeof(double) * hp.nx * hp.nu, cudaMemcpyHostToDevice));
    CU_TRY(cudaEventCreate(&h->ev0));
    CU_TRY(cudaEventCreate(&h->ev1));
    if (h->shape == 0) {  // stream mapping: resident CTAs per SM, workspace for the largest batch of this handle
        h->max_smem_optin = max_smem;
        const int smem8 = 8 * (int)StreamSmem<4, double>::warp_bytes(true, true);
        auto optin = [&](auto kern) { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, std::min(max_smem, std::max(smem8, max_smem))); };
        CU_TRY(optin(ipm_stream_kernel<4, false, double, 8>)); CU_TRY(optin(ipm_stream_kernel<4, true, double, 8>));
        CU_TRY(optin(ipm_stream_kernel<4, false, float, 8>));  CU_TRY(optin(ipm_stream_kernel<4, true, float, 8>));
        CU_TRY(optin(ipm_stream_kernel<4, false, double, 8, true>));
        if (const char* e = getenv("LBMPC_LOOP_CHUNK")) h->st_loop_chunk = std::max(1, atoi(e));
        if (const char* e = getenv("LBMPC_LOOP_MIN_BATCH")) h->st_loop_min_batch = atoll(e);
        CU_TRY(optin(ipm_stream_kernel<4, false, double, 6>)); CU_TRY(optin(ipm_stream_kernel<4, true, double, 6>));
        CU_TRY(optin(ipm_stream_kernel<4, false, double, 2>)); CU_TRY(optin(ipm_stream_kernel<4, false, double, 4>));
        CU_TRY(optin(ipm_stream_kernel<4, true, double, 2>));  CU_TRY(optin(ipm_stream_kernel<4, true, double, 4>));
        CU_TRY(optin(ipm_stream_kernel<4, false, float, 6>));  CU_TRY(optin(ipm_stream_kernel<4, true, float, 6>));
        h->st_ctas_per_sm = 1;
        // measured on B200 (C-form LBMPC, stream with iteration budget + hand-over vs the best shared-memory mapping):
        //   N = 50 : batch 24576 5.95 vs 5.58 ms, 32768 6.32 vs 7.39, 65536 11.1 vs 14.5, 131072 19.5 vs 28.7  -> from ~31 k QPs on
        //   N = 200: batch 16384 28.3 vs 26.5 ms, 24576 28.6 vs 39.5, 32768 30.2 vs 52.7, 49152 44.3 vs 78.6   -> from ~15 k QPs on (spread launches, below)
        //   616-row set, N = 50 (no budget): 65536 29.0 vs 26.6 ms, 98304 38.0 vs 39.5, 131072 45.3 vs ~52            -> from ~92 k QPs on
        //   (profiles/r2_threshold_sweep.log; budget at N = 200: 16 iterations — 14: +12 %, 22: +10 %, 32: +27 %)
        h->st_min_batch = (int64_t)h->num_sms * 208;
        h->st_evict_iters = hp.N > 100 ? 16 : 14;  // N = 50: 14 vs 12 -> closed loop 4.80 vs 4.50 M QP/s, batch 262144 34.5 vs 35.3 ms (profiles/r2_evict_sweep.log)
        if (const char* e = getenv("LBMPC_STREAM_WARPS")) h->st_warps_cap = atoi(e);
        if (const char* e = getenv("LBMPC_LOOP_ALLOW_STREAM")) h->loop_allow_stream = atoi(e) != 0;
        if (const char* e = getenv("LBMPC_STREAM_SPREAD")) h->st_spread = atoi(e) != 0;
        if (const char* e = getenv("LBMPC_STREAM_EVICT")) { h->st_evict_iters = atoi(e); h->st_evict_forced = true; }
        h->st_min_batch_long = (int64_t)h->num_sms * 104;  // with small launches spread over all SMs: 12288 22.6 vs 20.1 ms (CTA), 16384 23.3 vs 26.5 (profiles/r2_spread_sweep.log)
        if (const char* e = getenv("LBMPC_STREAM_MIN_BATCH_LONG")) h->st_min_batch_long = atoll(e);
        if (const char* e = getenv("LBMPC_STREAM_MIN_BATCH")) h->st_min_batch = atoll(e);
        if (h->max_batch >= h->st_min_batch) {  // workspace of the resident warps for the default layout; other layouts grow it on first use
            const StreamLayout<4> l(hp.N, hp.ng, false, false);
            const int64_t ctas = std::min<int64_t>(h->num_sms, (h->max_batch + 255) / 256);
            CU_TRY(stream_workspace(h, ctas * 8, l, false));
        }
    }
    // experiment / test overrides, read once
    if (const char* e = getenv("LBMPC_KERNEL"))
        h->force_kernel = e[0] == 'c' ? LBMPC_KERNEL_CTA : e[0] == 's' ? LBMPC_KERNEL_STREAM : e[0] == 'm' ? LBMPC_KERNEL_STREAM_MIXED
                        : e[0] == 'w' ? LBMPC_KERNEL_WARP : LBMPC_KERNEL_AUTO;
    if (const char* e = getenv("LBMPC_LOCKSTEP")) h->force_lockstep = atoi(e) != 0;
    if (!h->dev_ptrs) {
        const size_t b = (size_t)h->max_batch, N = hp.N, nx = hp.nx, nu = hp.nu, nt = hp.nt;
        CU_TRY(dmalloc(&h->s_dx0, b * nx));
        CU_TRY(dmalloc(&h->s_ref, b * nx));
        CU_TRY(dmalloc(&h->s_doff, b * nx * N));
        CU_TRY(dmalloc(&h->s_warm, b * (nu * N + nt)));
        CU_TRY(dmalloc(&h->s_uc, b * nu * N));
        CU_TRY(dmalloc(&h->s_x, b * nx * (N + 1)));
        const size_t small = small_out(nullptr, b, nt).bytes;
        CU_TRY(dmalloc(&h->s_small, small));
        CU_TRY(cudaHostAlloc((void**)&h->hs_small, std::max<size_t>(small, 16), cudaHostAllocDefault));
        CU_TRY(cudaHostAlloc((void**)&h->hs_bounce, kBounceBytes, cudaHostAllocMapped));
    }
    guard.h = nullptr;  // success: the caller owns the handle
    *out = h;
    return LBMPC_OK;
}

int lbmpc_solve_batch(lbmpc_handle* h, int64_t batch, const double* dx0, const double* dx_ref, const double* d_off,
                      const double* warm, double* u_or_c, double* theta, double* x_traj, double* obj, int32_t* iters,
                      int32_t* status, void* stream) {
    return lbmpc_solve_batch_shifted(h, batch, dx0, dx_ref, d_off, nullptr, warm, u_or_c, theta, x_traj, obj, iters, status, stream);
}

static int solve_batch_impl(lbmpc_handle* h, int64_t batch, const double* dx0, const double* dx_ref, const double* d_off,
                            const double* cost_shift, int cs_stride, const double* warm, double* u_or_c, double* theta, double* x_traj,
                            double* obj, int32_t* iters, int32_t* status, void* stream);

int lbmpc_solve_batch_shifted(lbmpc_handle* h, int64_t batch, const double* dx0, const double* dx_ref, const double* d_off,
                              const double* cost_shift, const double* warm, double* u_or_c, double* theta, double* x_traj,
                              double* obj, int32_t* iters, int32_t* status, void* stream) {
    return solve_batch_impl(h, batch, dx0, dx_ref, d_off, cost_shift, h ? h->hp.nx : 0, warm, u_or_c, theta, x_traj, obj, iters, status, stream);
}
int lbmpc_solve_batch_shifted_xu(lbmpc_handle* h, int64_t batch, const double* dx0, const double* dx_ref, const double* d_off,
                                 const double* cost_shift_xu, const double* warm, double* u_or_c, double* theta, double* x_traj,
                                 double* obj, int32_t* iters, int32_t* status, void* stream) {
    return solve_batch_impl(h, batch, dx0, dx_ref, d_off, cost_shift_xu, h ? h->hp.nx + h->hp.nu : 0, warm, u_or_c, theta, x_traj, obj, iters, status, stream);
}

}  // extern "C"

static int solve_batch_impl(lbmpc_handle* h, int64_t batch, const double* dx0, const double* dx_ref, const double* d_off,
                            const double* cost_shift, int cs_stride, const double* warm, double* u_or_c, double* theta, double* x_traj,
                            double* obj, int32_t* iters, int32_t* status, void* stream) {
    if (!h) return fail(LBMPC_EINVAL, "handle is NULL");
    if (batch < 0) return fail(LBMPC_EINVAL, "negative batch");
    if (batch == 0) return LBMPC_OK;
    if (!dx0 || !u_or_c || !theta || !obj || !iters || !status) return fail(LBMPC_EINVAL, "required array is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    CU_TRY(cudaSetDevice(h->device));
    const HostProblem& hp = h->hp;
    const size_t b = (size_t)batch, N = hp.N, nx = hp.nx, nu = hp.nu, nt = hp.nt;
    BatchIO io{};
    io.batch = batch;
    io.queue = h->dqueue;
    io.prof = h->prof_on ? h->dprof : nullptr;
    io.cs_stride = cost_shift ? cs_stride : 0;
    const size_t nb_cs = sizeof(double) * (size_t)batch * (size_t)cs_stride * (hp.N + 1);
    if (h->dev_ptrs) {
        io.dx0 = dx0; io.dx_ref = dx_ref; io.d_off = d_off; io.warm = warm; io.cshift = cost_shift;
        io.uc = u_or_c; io.theta = theta; io.xtraj = x_traj; io.obj = obj; io.iters = iters; io.status = status;
        CU_TRY(cudaEventRecord(h->ev0, st));
        CU_TRY(launch_ipm_any(h, io, st));
        CU_TRY(cudaEventRecord(h->ev1, st));
        h->timed = true;
        return LBMPC_OK;
    }
    if (batch > h->max_batch) return fail(LBMPC_EINVAL, "batch exceeds config.max_batch (host-pointer staging)");
    // Where the caller's arrays live decides how they travel:
    //  * page-locked arrays: the kernel reads / writes them in place (pinned_alias);
    //  * pageable arrays of a SMALL call (a closed-loop step, a handful of QPs): bounced through a pinned, mapped block of
    //    the handle — a host memcpy each way, no copy-engine operation at all, so the call costs one launch + one sync;
    //  * pageable arrays of a large call: device staging buffers and cudaMemcpyAsync (small outputs packed into one copy).
    const double* a_dx0 = pinned_alias(dx0);
    const double* a_ref = pinned_alias(dx_ref);
    double* a_uc = pinned_alias(u_or_c);
    double* a_x = pinned_alias(x_traj);
    double* a_th = pinned_alias(theta);
    double* a_obj = pinned_alias(obj);
    int32_t* a_it = pinned_alias(iters);
    int32_t* a_st = pinned_alias(status);
    const bool small_direct = a_th && a_obj && a_it && a_st;
    const size_t nb_x = sizeof(double) * b * nx, nb_doff = sizeof(double) * b * nx * N, nb_warm = sizeof(double) * b * (nu * N + nt),
                 nb_uc = sizeof(double) * b * nu * N, nb_xt = sizeof(double) * b * nx * (N + 1);
    const SmallOut so_sz = small_out(nullptr, b, nt);
    const size_t bounce_need = (a_dx0 ? 0 : nb_x) + ((dx_ref && !a_ref) ? nb_x : 0) + (d_off ? nb_doff : 0) + (warm ? nb_warm : 0) +
                               (a_uc ? 0 : nb_uc) + ((x_traj && !a_x) ? nb_xt : 0) + (small_direct ? 0 : so_sz.bytes) + 64;
    if (h->hs_bounce && bounce_need <= kBounceBytes) {
        char* cur = h->hs_bounce;  // mapped: the same address is valid on the device (unified addressing)
        auto in = [&](const double* src, size_t bytes) -> const double* {
            char* dst = cur;
            memcpy(dst, src, bytes);
            cur += (bytes + 15) & ~(size_t)15;
            return reinterpret_cast<const double*>(dst);
        };
        auto out = [&](size_t bytes) -> char* {
            char* dst = cur;
            cur += (bytes + 15) & ~(size_t)15;
            return dst;
        };
        io.dx0 = a_dx0 ? a_dx0 : in(dx0, nb_x);
        io.dx_ref = dx_ref ? (a_ref ? a_ref : in(dx_ref, nb_x)) : nullptr;
        io.d_off = d_off ? in(d_off, nb_doff) : nullptr;
        io.warm = warm ? in(warm, nb_warm) : nullptr;
        if (cost_shift) {  // read in every iteration: keep it in device memory
            if (!h->s_csh) CU_TRY(dmalloc(&h->s_csh, (size_t)h->max_batch * (nx + nu) * (N + 1)));
            CU_TRY(cudaMemcpyAsync(h->s_csh, cost_shift, nb_cs, cudaMemcpyHostToDevice, st));
            io.cshift = h->s_csh;
        }
        char* b_uc = a_uc ? nullptr : out(nb_uc);
        char* b_xt = (x_traj && !a_x) ? out(nb_xt) : nullptr;
        char* b_small = small_direct ? nullptr : out(so_sz.bytes);
        const SmallOut so = small_out(b_small, b, nt);
        io.uc = a_uc ? a_uc : reinterpret_cast<double*>(b_uc);
        io.xtraj = x_traj ? (a_x ? a_x : reinterpret_cast<double*>(b_xt)) : nullptr;
        io.theta = small_direct ? a_th : so.theta;
        io.obj = small_direct ? a_obj : so.obj;
        io.iters = small_direct ? a_it : so.iters;
        io.status = small_direct ? a_st : so.status;
        CU_TRY(cudaEventRecord(h->ev0, st));
        CU_TRY(launch_ipm_any(h, io, st));
        CU_TRY(cudaEventRecord(h->ev1, st));
        h->timed = true;
        CU_TRY(cudaStreamSynchronize(st));
        if (b_uc) memcpy(u_or_c, b_uc, nb_uc);
        if (b_xt) memcpy(x_traj, b_xt, nb_xt);
        if (b_small) {
            memcpy(theta, so.theta, sizeof(double) * b * nt);
            memcpy(obj, so.obj, sizeof(double) * b);
            memcpy(iters, so.iters, sizeof(int) * b);
            memcpy(status, so.status, sizeof(int) * b);
        }
        return LBMPC_OK;
    }
    if (!a_dx0) CU_TRY(cudaMemcpyAsync(h->s_dx0, dx0, nb_x, cudaMemcpyHostToDevice, st));
    if (dx_ref && !a_ref) CU_TRY(cudaMemcpyAsync(h->s_ref, dx_ref, nb_x, cudaMemcpyHostToDevice, st));
    if (d_off) CU_TRY(cudaMemcpyAsync(h->s_doff, d_off, nb_doff, cudaMemcpyHostToDevice, st));
    if (warm) CU_TRY(cudaMemcpyAsync(h->s_warm, warm, nb_warm, cudaMemcpyHostToDevice, st));
    if (cost_shift) {
        if (!h->s_csh) CU_TRY(dmalloc(&h->s_csh, (size_t)h->max_batch * (nx + nu) * (N + 1)));  // first use only
        CU_TRY(cudaMemcpyAsync(h->s_csh, cost_shift, nb_cs, cudaMemcpyHostToDevice, st));
        io.cshift = h->s_csh;
    }
    io.dx0 = a_dx0 ? a_dx0 : h->s_dx0;
    io.dx_ref = dx_ref ? (a_ref ? a_ref : h->s_ref) : nullptr;
    io.d_off = d_off ? h->s_doff : nullptr;
    io.warm = warm ? h->s_warm : nullptr;
    const SmallOut so = small_out(h->s_small, b, nt);
    io.uc = a_uc ? a_uc : h->s_uc;
    io.xtraj = x_traj ? (a_x ? a_x : h->s_x) : nullptr;
    io.theta = small_direct ? a_th : so.theta;
    io.obj = small_direct ? a_obj : so.obj;
    io.iters = small_direct ? a_it : so.iters;
    io.status = small_direct ? a_st : so.status;
    CU_TRY(cudaEventRecord(h->ev0, st));
    CU_TRY(launch_ipm_any(h, io, st));
    CU_TRY(cudaEventRecord(h->ev1, st));
    h->timed = true;
    if (!a_uc) CU_TRY(cudaMemcpyAsync(u_or_c, h->s_uc, nb_uc, cudaMemcpyDeviceToHost, st));
    if (!small_direct) CU_TRY(cudaMemcpyAsync(h->hs_small, h->s_small, so.bytes, cudaMemcpyDeviceToHost, st));
    if (x_traj && !a_x) CU_TRY(cudaMemcpyAsync(x_traj, h->s_x, nb_xt, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (!small_direct) scatter_small(h, b, nt, theta, obj, iters, status);
    return LBMPC_OK;
}

extern "C" {

int lbmpc_oracle_apply(lbmpc_handle* h, int64_t batch, int32_t q, double bandwidth, double lambda, const double* dx0,
                       const double* du, const double* X, const double* Y, const double* valid, double* d_off,
                       void* stream) {
    if (!h) return fail(LBMPC_EINVAL, "handle is NULL");
    if (batch <= 0) return batch == 0 ? LBMPC_OK : fail(LBMPC_EINVAL, "negative batch");
    if (!dx0 || !du || !X || !Y || !d_off) return fail(LBMPC_EINVAL, "required array is NULL");
    if (h->shape != 0) return fail(LBMPC_ESHAPE, "the L2NW oracle is defined on the 4-state model (xi=[x1;x2;u])");
    if (q < 1 || q > 32 * kOracleMaxPerLane) return fail(LBMPC_ESHAPE, "q must be in [1, 512]");
    if (!(bandwidth > 0)) return fail(LBMPC_EINVAL, "bandwidth must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    CU_TRY(cudaSetDevice(h->device));
    const HostProblem& hp = h->hp;
    const size_t b = (size_t)batch, N = hp.N;
    const double inv_h2 = 1.0 / (bandwidth * bandwidth);
    if (h->dev_ptrs) {
        launch_oracle(h, st, batch, q, inv_h2, lambda, dx0, du, (long long)N, X, Y, valid, d_off);
        CU_TRY(cudaGetLastError());
        return LBMPC_OK;
    }
    double *ddx0 = nullptr, *ddu = nullptr, *dX = nullptr, *dY = nullptr, *dV = nullptr, *dd = nullptr;
    CU_TRY(staging(h, 0, b * 4, &ddx0)); CU_TRY(staging(h, 1, b * N, &ddu)); CU_TRY(staging(h, 2, b * 3 * q, &dX));
    CU_TRY(staging(h, 3, b * 4 * q, &dY)); CU_TRY(staging(h, 4, b * 4 * N, &dd));
    if (valid) CU_TRY(staging(h, 5, b * q, &dV));
    CU_TRY(cudaMemcpyAsync(ddx0, dx0, 8 * b * 4, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(ddu, du, 8 * b * N, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(dX, X, 8 * b * 3 * q, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(dY, Y, 8 * b * 4 * q, cudaMemcpyHostToDevice, st));
    if (valid) CU_TRY(cudaMemcpyAsync(dV, valid, 8 * b * q, cudaMemcpyHostToDevice, st));
    launch_oracle(h, st, batch, q, inv_h2, lambda, ddx0, ddu, (long long)N, dX, dY, dV, dd);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(d_off, dd, 8 * b * 4 * N, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return LBMPC_OK;
}

int lbmpc_solve_sqp(lbmpc_handle* h, int64_t batch, int32_t sqp_iters, int32_t twin, int32_t q, double bandwidth, double lambda,
                    const double* dx0, const double* dx_ref, const double* X, const double* Y, const double* valid,
                    const double* warm, double* u, double* theta, double* x_traj, double* obj, int32_t* iters,
                    int32_t* status, double* du_step, void* stream) {
    return lbmpc_solve_sqp_ex(h, batch, sqp_iters, twin, 0, q, bandwidth, lambda, dx0, dx_ref, X, Y, valid, warm, u, theta, x_traj, obj,
                              iters, status, du_step, stream);
}

// device-pointer core of lbmpc_solve_sqp_ex (also the per-step solve of the F-form LBMPC closed loop): scratch sized on demand,
// then `sqp_iters` times { oracle (value, or value + Jacobian) along the previous solution -> one QP -> bookkeeping }
static int sqp_core(lbmpc_handle* h, int64_t batch, int32_t sqp_iters, int32_t twin, int32_t order, int32_t q, double bandwidth, double lambda,
                    const double* d_dx0, const double* d_ref, const double* d_X, const double* d_Y, const double* d_V, const double* d_warm,
                    double* d_u, double* d_th, double* d_xt, double* d_obj, int* d_it, int* d_st, bool du_step, cudaStream_t st) {
    const HostProblem& hp = h->hp;
    const size_t b = (size_t)batch, N = hp.N, nx = hp.nx, nt = hp.nt;
    if (h->sqp_batch < batch || h->sqp_iters < sqp_iters) {
        cudaFree(h->q_ulin); cudaFree(h->q_warm); cudaFree(h->q_doff); cudaFree(h->q_step);
        h->q_ulin = h->q_warm = h->q_doff = h->q_step = nullptr;
        CU_TRY(dmalloc(&h->q_ulin, b * N)); CU_TRY(dmalloc(&h->q_warm, b * (N + nt)));
        CU_TRY(dmalloc(&h->q_doff, b * nx * N)); CU_TRY(dmalloc(&h->q_step, b * (size_t)sqp_iters));
        cudaFree(h->q_csh); cudaFree(h->q_jac);
        h->q_csh = h->q_jac = nullptr;
        h->sqp_batch = batch; h->sqp_iters = sqp_iters;
    }
    if (order == 1 && !h->q_jac) CU_TRY(dmalloc(&h->q_jac, (size_t)h->sqp_batch * nx * 3 * N));
    if (twin && !h->q_csh) CU_TRY(dmalloc(&h->q_csh, (size_t)h->sqp_batch * (nx + 1) * (N + 1)));
    if (d_warm) CU_TRY(cudaMemcpy2DAsync(h->q_ulin, 8 * N, d_warm, 8 * (N + nt), 8 * N, b, cudaMemcpyDeviceToDevice, st));
    else CU_TRY(cudaMemsetAsync(h->q_ulin, 0, 8 * b * N, st));
    const double inv_h2 = 1.0 / (bandwidth * bandwidth);
    const unsigned ug = (unsigned)((batch + 3) / 4);
    for (int j = 0; j < sqp_iters; ++j) {
        if (order == 1) {  // value + Jacobian of the oracle along the learned rollout, gap to the nominal rollout (twin)
            const double* Kfb = hp.form == LBMPC_FORM_F ? h->dK : nullptr;
            const unsigned og = (unsigned)((batch + 3) / 4);
            if (q <= 128)
                oracle_jac_kernel<4, 4><<<og, 128, 0, st>>>(h->dA, h->dB, Kfb, hp.N, batch, q, inv_h2, lambda, d_dx0, h->q_ulin, (long long)N,
                                                           d_X, d_Y, d_V, h->q_doff, h->q_jac, twin ? h->q_csh : nullptr);
            else
                oracle_jac_kernel<4, kOracleMaxPerLane><<<og, 128, 0, st>>>(h->dA, h->dB, Kfb, hp.N, batch, q, inv_h2, lambda, d_dx0, h->q_ulin,
                                                                           (long long)N, d_X, d_Y, d_V, h->q_doff, h->q_jac,
                                                                           twin ? h->q_csh : nullptr);
            h->launches += 1;
            CU_TRY(cudaGetLastError());
            BatchIO io{};
            io.batch = batch; io.prof = nullptr;
            io.dx0 = d_dx0; io.dx_ref = d_ref; io.d_off = h->q_doff; io.warm = j == 0 ? d_warm : h->q_warm;
            if (twin) { io.cshift = h->q_csh; io.cs_stride = (int)nx + 1; io.row_shift = 1; }
            io.uc = d_u; io.theta = d_th; io.xtraj = d_xt; io.obj = d_obj; io.iters = d_it; io.status = d_st;
            CU_TRY(launch_ipm_any(h, io, st, h->q_jac));
            sqp_update_kernel<<<ug, 128, 0, st>>>(batch, hp.N, hp.nt, d_u, d_th, h->q_ulin, h->q_warm, du_step ? h->q_step : nullptr,
                                                  sqp_iters, j);
            h->launches += 1;
            CU_TRY(cudaGetLastError());
            continue;
        }
        launch_oracle(h, st, batch, q, inv_h2, lambda, d_dx0, h->q_ulin, (long long)N, d_X, d_Y, d_V, h->q_doff);
        CU_TRY(cudaGetLastError());
        BatchIO io{};
        io.batch = batch; io.queue = h->dqueue; io.prof = nullptr;
        io.dx0 = d_dx0; io.dx_ref = d_ref; io.d_off = h->q_doff; io.warm = j == 0 ? d_warm : h->q_warm;
        if (twin) {  // cost on the learned sequence x + e, rows and dynamics on the nominal one
            twin_shift_kernel<4><<<(unsigned)((batch + 127) / 128), 128, 0, st>>>(batch, hp.N, h->dA, h->dB, h->dK, h->q_doff, h->q_csh);
            h->launches += 1;
            CU_TRY(cudaGetLastError());
            io.d_off = nullptr;
            io.cshift = h->q_csh;
            io.cs_stride = (int)nx + 1;
        }
        io.uc = d_u; io.theta = d_th; io.xtraj = d_xt; io.obj = d_obj; io.iters = d_it; io.status = d_st;
        CU_TRY(launch_ipm_any(h, io, st));
        sqp_update_kernel<<<ug, 128, 0, st>>>(batch, hp.N, hp.nt, d_u, d_th, h->q_ulin, h->q_warm, du_step ? h->q_step : nullptr,
                                              sqp_iters, j);
        h->launches += 1;
        CU_TRY(cudaGetLastError());
    }
    return LBMPC_OK;
}

int lbmpc_solve_sqp_ex(lbmpc_handle* h, int64_t batch, int32_t sqp_iters, int32_t twin, int32_t order, int32_t q, double bandwidth,
                       double lambda, const double* dx0, const double* dx_ref, const double* X, const double* Y, const double* valid,
                       const double* warm, double* u, double* theta, double* x_traj, double* obj, int32_t* iters,
                       int32_t* status, double* du_step, void* stream) {
    if (!h) return fail(LBMPC_EINVAL, "handle is NULL");
    if (order != 0 && order != 1) return fail(LBMPC_EINVAL, "order must be 0 (frozen oracle value) or 1 (value and Jacobian)");
    if (batch < 0 || sqp_iters < 1) return fail(LBMPC_EINVAL, "batch must be >= 0 and sqp_iters >= 1");
    if (batch == 0) return LBMPC_OK;
    if (!dx0 || !X || !Y || !u || !theta || !obj || !iters || !status) return fail(LBMPC_EINVAL, "required array is NULL");
    if (h->shape != 0) return fail(LBMPC_ESHAPE, "solve_sqp: 4-state Moore-Greitzer model (the L2NW oracle is defined on xi = [x1;x2;u])");
    if (h->hp.form == LBMPC_FORM_F && !twin)
        return fail(LBMPC_ESHAPE, "solve_sqp: the F-form rolls the learned model in the cost only (costLBMPC.m:27 vs constraintsLBMPC.m:23): twin = 1");
    if (q < 1 || q > 32 * kOracleMaxPerLane) return fail(LBMPC_ESHAPE, "q must be in [1, 512]");
    if (!(bandwidth > 0)) return fail(LBMPC_EINVAL, "bandwidth must be positive");
    if (!h->dev_ptrs && batch > h->max_batch) return fail(LBMPC_EINVAL, "batch exceeds config.max_batch (host-pointer staging)");
    cudaStream_t st = (cudaStream_t)stream;
    CU_TRY(cudaSetDevice(h->device));
    const HostProblem& hp = h->hp;
    const size_t b = (size_t)batch, N = hp.N, nx = hp.nx, nt = hp.nt;
    // device views of the inputs / outputs
    const double *d_dx0 = dx0, *d_ref = dx_ref, *d_X = X, *d_Y = Y, *d_V = valid, *d_warm = warm;
    double *d_u = u, *d_th = theta, *d_xt = x_traj, *d_obj = obj;
    int *d_it = iters, *d_st = status;
    double *tX = nullptr, *tY = nullptr, *tV = nullptr;
    if (!h->dev_ptrs) {
        CU_TRY(staging(h, 2, b * 3 * q, &tX)); CU_TRY(staging(h, 3, b * 4 * q, &tY));
        if (valid) CU_TRY(staging(h, 5, b * q, &tV));
        CU_TRY(cudaMemcpyAsync(tX, X, 8 * b * 3 * q, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(tY, Y, 8 * b * 4 * q, cudaMemcpyHostToDevice, st));
        if (valid) CU_TRY(cudaMemcpyAsync(tV, valid, 8 * b * q, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(h->s_dx0, dx0, 8 * b * nx, cudaMemcpyHostToDevice, st));
        if (dx_ref) CU_TRY(cudaMemcpyAsync(h->s_ref, dx_ref, 8 * b * nx, cudaMemcpyHostToDevice, st));
        if (warm) CU_TRY(cudaMemcpyAsync(h->s_warm, warm, 8 * b * (N + nt), cudaMemcpyHostToDevice, st));
        d_dx0 = h->s_dx0; d_ref = dx_ref ? h->s_ref : nullptr; d_X = tX; d_Y = tY; d_V = tV; d_warm = warm ? h->s_warm : nullptr;
        const SmallOut so = small_out(h->s_small, b, nt);
        d_u = h->s_uc; d_th = so.theta; d_xt = x_traj ? h->s_x : nullptr; d_obj = so.obj; d_it = so.iters; d_st = so.status;
    }
    CU_TRY(cudaEventRecord(h->ev0, st));
    {
        const int rc = sqp_core(h, batch, sqp_iters, twin, order, q, bandwidth, lambda, d_dx0, d_ref, d_X, d_Y, d_V, d_warm, d_u, d_th, d_xt, d_obj,
                                d_it, d_st, du_step != nullptr, st);
        if (rc != LBMPC_OK) return rc;
    }
    CU_TRY(cudaEventRecord(h->ev1, st));
    h->timed = true;
    if (h->dev_ptrs) {
        if (du_step) CU_TRY(cudaMemcpyAsync(du_step, h->q_step, 8 * b * sqp_iters, cudaMemcpyDeviceToDevice, st));
        return LBMPC_OK;
    }
    CU_TRY(cudaMemcpyAsync(u, h->s_uc, 8 * b * N, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(h->hs_small, h->s_small, small_out(nullptr, b, nt).bytes, cudaMemcpyDeviceToHost, st));
    if (x_traj) CU_TRY(cudaMemcpyAsync(x_traj, h->s_x, 8 * b * nx * (N + 1), cudaMemcpyDeviceToHost, st));
    if (du_step) CU_TRY(cudaMemcpyAsync(du_step, h->q_step, 8 * b * sqp_iters, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    scatter_small(h, b, nt, theta, obj, iters, status);
    return LBMPC_OK;
}

static void free_loop(LoopScratch& L) {
    cudaFree(L.x); cudaFree(L.dx0); cudaFree(L.X); cudaFree(L.Y); cudaFree(L.V); cudaFree(L.warm); cudaFree(L.uc);
    cudaFree(L.theta); cudaFree(L.obj); cudaFree(L.doff); cudaFree(L.x_init); cudaFree(L.nd); cudaFree(L.iters);
    cudaFree(L.status); cudaFree(L.x_hist); cudaFree(L.u_hist); cudaFree(L.t_hist); cudaFree(L.i_hist);
    cudaFree(L.s_hist);
    L = LoopScratch();
}

int lbmpc_closed_loop(lbmpc_handle* h, int64_t batch, int32_t steps, int32_t q, int32_t use_oracle,
                      int32_t warm_shift, const double* x_eq, double u_eq, const double* x_init, const double* wbar,
                      uint64_t seed, uint64_t scenario0, double* x_hist, double* u_hist, double* theta_hist,
                      int32_t* iters_hist, int32_t* status_hist, void* stream) {
    if (!h) return fail(LBMPC_EINVAL, "handle is NULL");
    if (batch <= 0 || steps <= 0) return fail(LBMPC_EINVAL, "batch and steps must be positive");
    if (!x_eq || !x_init) return fail(LBMPC_EINVAL, "x_eq / x_init is NULL");
    if (h->shape != 0) return fail(LBMPC_ESHAPE, "closed loop: the 4-state Moore-Greitzer model");
    const bool fform = h->hp.form == LBMPC_FORM_F;
    if (q < 1 || q > 32 * kOracleMaxPerLane) return fail(LBMPC_ESHAPE, "q must be in [1, 512]");
    cudaStream_t st = (cudaStream_t)stream;
    CU_TRY(cudaSetDevice(h->device));
    const HostProblem& hp = h->hp;
    const size_t b = (size_t)batch, N = hp.N, S = (size_t)steps;
    LoopScratch& L = h->loop;
    // scratch is sized by (batch, q); the host-pointer history staging also by the number of steps: reallocated only when one of
    // them GROWS, so that a short warm-up run followed by the real one does not allocate inside the timed call
    if (L.batch < batch || L.q != q || (!h->dev_ptrs && L.steps < steps)) {
        free_loop(L);
        CU_TRY(dmalloc(&L.x, b * 4)); CU_TRY(dmalloc(&L.dx0, b * 4)); CU_TRY(dmalloc(&L.X, b * 3 * q));
        CU_TRY(dmalloc(&L.Y, b * 4 * q)); CU_TRY(dmalloc(&L.V, b * q)); CU_TRY(dmalloc(&L.warm, b * (N + 1)));
        CU_TRY(dmalloc(&L.uc, b * N)); CU_TRY(dmalloc(&L.theta, b)); CU_TRY(dmalloc(&L.obj, b));
        CU_TRY(dmalloc(&L.doff, b * 4 * N)); CU_TRY(dmalloc(&L.x_init, b * 4)); CU_TRY(dmalloc(&L.nd, b));
        CU_TRY(dmalloc(&L.iters, b)); CU_TRY(dmalloc(&L.status, b));
        if (!h->dev_ptrs) {
            CU_TRY(dmalloc(&L.x_hist, b * 4 * (S + 1))); CU_TRY(dmalloc(&L.u_hist, b * S));
            CU_TRY(dmalloc(&L.t_hist, b * S)); CU_TRY(dmalloc(&L.i_hist, b * S)); CU_TRY(dmalloc(&L.s_hist, b * S));
        }
        L.batch = batch; L.q = q; L.steps = steps;
    }
    CU_TRY(cudaMemsetAsync(L.X, 0, 8 * b * 3 * q, st));
    CU_TRY(cudaMemsetAsync(L.Y, 0, 8 * b * 4 * q, st));
    CU_TRY(cudaMemsetAsync(L.V, 0, 8 * b * q, st));
    CU_TRY(cudaMemsetAsync(L.warm, 0, 8 * b * (N + 1), st));
    const double* xi_dev = x_init;
    if (!h->dev_ptrs) {
        CU_TRY(cudaMemcpyAsync(L.x_init, x_init, 8 * b * 4, cudaMemcpyHostToDevice, st));
        xi_dev = L.x_init;
    }
    // x_eq / wbar are tiny parameter vectors: always host pointers
    const double4 xe = make_double4(x_eq[0], x_eq[1], x_eq[2], x_eq[3]);
    const double4 wb = wbar ? make_double4(wbar[0], wbar[1], wbar[2], wbar[3]) : make_double4(0, 0, 0, 0);
    LoopState Sx{};
    Sx.x = L.x; Sx.dx0 = L.dx0; Sx.X = L.X; Sx.Y = L.Y; Sx.V = L.V; Sx.nd = L.nd; Sx.warm = L.warm;
    Sx.uc = L.uc; Sx.theta = L.theta; Sx.iters = L.iters; Sx.status = L.status;
    if (h->dev_ptrs) {
        Sx.x_hist = x_hist; Sx.u_hist = u_hist; Sx.theta_hist = theta_hist; Sx.iters_hist = iters_hist;
        Sx.status_hist = status_hist;
    } else {
        Sx.x_hist = x_hist ? L.x_hist : nullptr; Sx.u_hist = u_hist ? L.u_hist : nullptr;
        Sx.theta_hist = theta_hist ? L.t_hist : nullptr; Sx.iters_hist = iters_hist ? L.i_hist : nullptr;
        Sx.status_hist = status_hist ? L.s_hist : nullptr;
    }
    // Fused closed loop (stream mapping): one persistent kernel runs every control step of every scenario.  Picked when
    // the stream mapping is forced, or automatically for scenario counts where it measures faster than one oracle / solve /
    // plant launch triple per step.
    const bool fused = !fform && h->shape == 0 && h->st_ctas_per_sm > 0 && hp.ng <= 64 && q <= 512 &&
                       (h->force_kernel == LBMPC_KERNEL_STREAM ||
                        (h->force_kernel == LBMPC_KERNEL_AUTO && h->st_loop_min_batch > 0 && batch >= h->st_loop_min_batch));
    if (fused) {
        constexpr int kW = 8;
        using SK = Stream<4, false, double, 32>;
        const Params<4, 1, 1> p = to_params<4, 1, 1>(hp);
        const StreamLayout<4> l(p.N, p.ng, false, false, q);
        const int chunk = std::max(1, std::min<int>(steps, h->st_loop_chunk));
        const int nchunks = (steps + chunk - 1) / chunk;
        const int64_t ctas = std::max<int64_t>(1, std::min<int64_t>(h->num_sms, (batch + 32 * kW - 1) / (32 * kW)));
        CU_TRY(stream_workspace(h, ctas * kW, l, false));
        const size_t slen = (size_t)SK::store_len(l), store_n = (size_t)batch * slen, rq_n = (size_t)batch * (size_t)std::max(nchunks - 1, 1);
        if (h->lp_store_cap < store_n) {
            cudaFree(h->lp_store); h->lp_store = nullptr; h->lp_store_cap = 0;
            CU_TRY(dmalloc(&h->lp_store, store_n));
            h->lp_store_cap = store_n;
        }
        if (h->lp_rq_cap < rq_n + 1) {
            cudaFree(h->lp_rq); h->lp_rq = nullptr; h->lp_rq_cap = 0;
            CU_TRY(dmalloc(&h->lp_rq, rq_n + 1));
            h->lp_rq_cap = rq_n + 1;
        }
        CU_TRY(cudaMemsetAsync(h->lp_rq, 0, sizeof(long long) * (rq_n + 1), st));
        StreamLoopParams lp{};
        lp.steps = steps; lp.chunk = chunk; lp.warm_shift = warm_shift; lp.use_oracle = use_oracle; lp.use_w = wbar != nullptr;
        lp.nscen = batch; lp.u_eq = u_eq; lp.inv_h2 = 1.0 / (0.5 * 0.5); lp.lambda = 0.001; lp.seed = seed; lp.scen0 = scenario0;
        for (int j = 0; j < 4; ++j) { lp.x_eq[j] = x_eq[j]; lp.wbar[j] = wbar ? wbar[j] : 0.0; }
        lp.x_init = xi_dev; lp.x_hist = Sx.x_hist; lp.u_hist = Sx.u_hist; lp.theta_hist = Sx.theta_hist;
        lp.iters_hist = Sx.iters_hist; lp.status_hist = Sx.status_hist;
        lp.store = h->lp_store; lp.rq = h->lp_rq + 1; lp.rq_tail = reinterpret_cast<unsigned long long*>(h->lp_rq);
        StreamIO<double> sio{};
        sio.batch = batch; sio.queue = next_queue(h); sio.ws64 = h->st_ws64; sio.wsft = (double*)h->st_wsft; sio.qwin = q;
        CU_TRY(cudaMemsetAsync(sio.queue, 0, sizeof(unsigned long long), st));
        const size_t smem = kW * StreamSmem<4, double>::warp_bytes(false, false) + StreamSmem<4, double>::poly_bytes(hp.ngp);
        CU_TRY(cudaEventRecord(h->ev0, st));
        ipm_stream_kernel<4, false, double, kW, true><<<(unsigned)ctas, 32 * kW, smem, st>>>(p, sio, h->dG, h->dhg, lp);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaEventRecord(h->ev1, st));
        h->timed = true;
        h->launches += 1;
        h->last_kernel = LBMPC_KERNEL_STREAM;
    }
    const unsigned tg = (unsigned)((batch + 127) / 128);
    if (!fused) loop_init_kernel<<<tg, 128, 0, st>>>(Sx, xi_dev, batch, steps, xe, fform ? 1 : 0, q);
    if (!fused) h->launches += 1;
    CU_TRY(cudaGetLastError());
    const double inv_h2 = 1.0 / (0.5 * 0.5);  // oracleL2NW.m:9 bandwidth = 0.5
    for (int it = 0; it < steps && !fused && fform; ++it) {
        // F-form loop (ocpLBMPC.m:10-47, ocpLMPC.m:11-40): every solve starts from the previous opt_var (zeros at the first step);
        // LBMPC with the oracle: the learned term acts on the cost only (costLBMPC.m:27 vs constraintsLBMPC.m:23) -> two first-order
        // SQP iterations on the data window (what lbmpc_b200.ocpLBMPC does per scenario); otherwise one exact QP
        if (use_oracle && hp.variant == LBMPC_VARIANT_LBMPC) {
            const int rc = sqp_core(h, batch, 2, /*twin=*/1, /*order=*/1, q, 0.5, 0.001, L.dx0, nullptr, L.X, L.Y, L.V, L.warm, L.uc, L.theta,
                                    nullptr, L.obj, L.iters, L.status, false, st);
            if (rc != LBMPC_OK) return rc;
        } else {
            BatchIO io{};
            io.batch = batch; io.queue = h->dqueue; io.prof = nullptr;
            io.dx0 = L.dx0; io.dx_ref = nullptr; io.d_off = nullptr; io.warm = L.warm;
            io.uc = L.uc; io.theta = L.theta; io.xtraj = nullptr; io.obj = L.obj; io.iters = L.iters; io.status = L.status;
            CU_TRY(launch_ipm_any(h, io, st));
        }
        plant_kernel<<<tg, 128, 0, st>>>(Sx, h->dA, h->dB, batch, hp.N, q, it, steps, xe, u_eq, wb, wbar != nullptr, seed, scenario0, h->dK);
        h->launches += 1;
        CU_TRY(cudaGetLastError());
    }
    for (int it = 0; it < steps && !fused && !fform; ++it) {
        const bool have = it > 0;
        if (use_oracle && have) {
            launch_oracle(h, st, batch, q, inv_h2, 0.001, L.dx0, L.warm, (long long)(N + 1), L.X, L.Y, L.V, L.doff);
            CU_TRY(cudaGetLastError());
        }
        BatchIO io{};
        io.batch = batch; io.queue = h->dqueue; io.prof = nullptr;
        io.dx0 = L.dx0; io.dx_ref = nullptr; io.d_off = (use_oracle && have) ? L.doff : nullptr;
        io.warm = (warm_shift && have) ? L.warm : nullptr;
        io.uc = L.uc; io.theta = L.theta; io.xtraj = nullptr; io.obj = L.obj; io.iters = L.iters; io.status = L.status;
        CU_TRY(launch_ipm_any(h, io, st, nullptr, /*allow_stream=*/h->loop_allow_stream));
        plant_kernel<<<tg, 128, 0, st>>>(Sx, h->dA, h->dB, batch, hp.N, q, it, steps, xe, u_eq, wb, wbar != nullptr,
                                         seed, scenario0);
        h->launches += 1;
        CU_TRY(cudaGetLastError());
    }
    if (!h->dev_ptrs) {
        if (x_hist) CU_TRY(cudaMemcpyAsync(x_hist, L.x_hist, 8 * b * 4 * (S + 1), cudaMemcpyDeviceToHost, st));
        if (u_hist) CU_TRY(cudaMemcpyAsync(u_hist, L.u_hist, 8 * b * S, cudaMemcpyDeviceToHost, st));
        if (theta_hist) CU_TRY(cudaMemcpyAsync(theta_hist, L.t_hist, 8 * b * S, cudaMemcpyDeviceToHost, st));
        if (iters_hist) CU_TRY(cudaMemcpyAsync(iters_hist, L.i_hist, 4 * b * S, cudaMemcpyDeviceToHost, st));
        if (status_hist) CU_TRY(cudaMemcpyAsync(status_hist, L.s_hist, 4 * b * S, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
    }
    return LBMPC_OK;
}

int lbmpc_set_kernel(lbmpc_handle* h, int32_t kernel, int32_t lockstep) {
    if (!h) return fail(LBMPC_EINVAL, "handle is NULL");
    if (kernel < LBMPC_KERNEL_AUTO || kernel > LBMPC_KERNEL_STREAM_MIXED) return fail(LBMPC_EINVAL, "kernel must be one of LBMPC_KERNEL_*");
    if ((kernel == LBMPC_KERNEL_CTA || kernel >= LBMPC_KERNEL_STREAM) && h->shape != 0)
        return fail(LBMPC_ESHAPE, "the CTA and stream mappings are compiled for the (4,1,1) Moore-Greitzer shape");
    h->force_kernel = kernel;
    h->force_lockstep = lockstep < 0 ? -1 : (lockstep != 0);
    return LBMPC_OK;
}
int lbmpc_last_kernel(const lbmpc_handle* h) { return h ? h->last_kernel : 0; }

int lbmpc_num_rows(const lbmpc_handle* h) { return h ? h->hp.m_rows : 0; }
int lbmpc_slots_per_cta(const lbmpc_handle* h) { return h ? h->max_slots : 0; }
int64_t lbmpc_kernel_launches(const lbmpc_handle* h) { return h ? h->launches : 0; }

int lbmpc_debug_phase_cycles(lbmpc_handle* h, int enable, uint64_t* out8) {
    if (!h) return fail(LBMPC_EINVAL, "handle is NULL");
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaDeviceSynchronize());
    if (out8) CU_TRY(cudaMemcpy(out8, h->dprof, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemset(h->dprof, 0, 16 * sizeof(unsigned long long)));
    h->prof_on = enable != 0;
    return LBMPC_OK;
}

float lbmpc_last_kernel_ms(lbmpc_handle* h) {
    if (!h || !h->timed) return -1.0f;
    float ms = -1.0f;
    if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0f;
    return ms;
}

int lbmpc_measure_fp64_peak(int device, double* tflops) {
    if (!tflops) return fail(LBMPC_EINVAL, "tflops is NULL");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return fail(LBMPC_ECUDA, "no usable CUDA device");
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    double* d = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    struct Guard {
        double*& d; cudaEvent_t &e0, &e1;
        ~Guard() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); cudaFree(d); }
    } guard{d, e0, e1};
    CU_TRY(dmalloc(&d, 1));
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    const int grid = prop.multiProcessorCount * 2, threads = 1024, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CU_TRY(cudaEventRecord(e0));
        dfma_peak_kernel<<<grid, threads>>>(d, iters, 1.0000001, 1e-9);
        CU_TRY(cudaEventRecord(e1));
        CU_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * (double)iters * (double)grid * (double)threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) * 1e-12);
    }
    *tflops = best;
    return LBMPC_OK;
}

void lbmpc_destroy(lbmpc_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    free_loop(h->loop);
    cudaFree(h->dpoly); cudaFree(h->dG); cudaFree(h->dhg); cudaFree(h->dA); cudaFree(h->dB); cudaFree(h->dK); cudaFree(h->dqueue); cudaFree(h->dprof);
    cudaFree(h->s_dx0); cudaFree(h->s_ref); cudaFree(h->s_doff); cudaFree(h->s_warm); cudaFree(h->s_uc);
    cudaFree(h->s_x); cudaFree(h->s_small); cudaFree(h->s_csh);
    if (h->hs_small) cudaFreeHost(h->hs_small);
    if (h->hs_bounce) cudaFreeHost(h->hs_bounce);
    cudaFree(h->q_ulin); cudaFree(h->q_warm); cudaFree(h->q_doff); cudaFree(h->q_step); cudaFree(h->q_csh); cudaFree(h->q_jac);
    cudaFree(h->st_ws64); cudaFree(h->st_wsft);
    for (int i = 0; i < 6; ++i) cudaFree(h->og[i]);
    cudaFree(h->lp_store); cudaFree(h->lp_rq); cudaFree(h->st_left);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
}

}  // extern "C"
