// Isolated timing of the 16-lane cooperative Riccati stage (Coop<4>) — one warp, variants that drop
// one ingredient at a time, to see what the per-stage dependency chain really costs on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../learning-based-mpc_b200/csrc/lbmpc_core.cuh"
using namespace lbmpc;
using CP = Coop<4>;
using P4 = Params<4, 1, 1>;
using L4 = Layout<4, 1, 1>;
constexpr unsigned kFull = 0xffffffffu;

template <int MODE>
__global__ void k_factor(const __grid_constant__ P4 p, long long* cyc, double* out, int reps) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, hl = lane & 15, half = lane >> 4;
    const L4 l(p.N, p.ngp);
    double* sl = smem + half * l.stride;
    double* xch = smem + 2 * l.stride + 2 + half * 24;
    for (int i = lane; i < 2 * l.stride + 64; i += 32) smem[i] = 0.01 * (i % 7) + 0.5;
    __syncwarp();
    typename CP::Lane ln;
    CP::lane_init(p, hl, ln);
    CP::xch_init(hl, xch);
    __syncwarp();
    const int N = p.N, hb = lane & 16;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        CP::terminal(p, l, sl, ln);
        for (int k = N - 1; k >= 0; --k) {
            CP::st1(p, l, sl, k, hl, ln, xch);
            __syncwarp();
            CP::st2(l, sl, k + 1, hl, ln, xch, MODE != 1);
            double fa, fb, fuu;
            if (MODE == 2) { fa = ln.pub; fb = ln.pub * 0.5; fuu = ln.pub * 0.25; }
            else {
                fa = __shfl_sync(kFull, ln.pub, hb | ln.sa);
                fb = __shfl_sync(kFull, ln.pub, hb | ln.sb);
                fuu = __shfl_sync(kFull, ln.pub, hb | 4);
            }
            CP::st3(ln, fa, fb, fuu);
            if (MODE == 3) ln.pt = ln.pt * 1e-3 + 1.0;  // keep values tame
            __syncwarp();
        }
        CP::finish(l, sl, hl, ln, true);
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = ln.pt;
}

#ifndef REPS
#define REPS 20
#endif
int main() {
    P4 p{};
    p.N = 50; p.ng = 24; p.ngp = 24; p.kg = 1; p.kT = 50; p.kx0 = 1; p.kx1 = 50; p.ku0 = 0; p.ku1 = 49; p.ntypes = 2;
    p.tseg[0] = 0; p.tseg[1] = 50; p.tseg[2] = p.tseg[3] = p.tseg[4] = 51; p.rowmask = 0x3ff; p.max_iter = 60; p.m_rows = 524;
    for (int i = 0; i < 16; ++i) p.A[i] = (i % 5 == 0) ? 0.9 : 0.01 * i;
    for (int i = 0; i < 4; ++i) p.B[i] = 0.1 * (i + 1);
    for (int t = 0; t < 4; ++t) for (int i = 0; i < 36; ++i) p.W[t][i] = (i % 7 == 0) ? 2.0 : 0.01;
    long long* cyc; double* out;
    cudaMalloc(&cyc, 8); cudaMalloc(&out, 32 * 8);
    const L4 l(p.N, p.ngp);
    size_t smem = (2 * l.stride + 128) * 8;
    const char* names[4] = {"full stage", "no factor stores", "no shuffles (local values)", "full + damping"};
    for (int rep = 0; rep < 2; ++rep) {
#define RUN(M) { cudaFuncSetAttribute(k_factor<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        k_factor<M><<<1, 32, smem>>>(p, cyc, out, REPS); long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
        printf("%-30s %8.1f cycles/stage  (%s)\n", names[M], (double)h / ((double)REPS * 50), cudaGetErrorString(cudaGetLastError())); }
        RUN(0) RUN(1) RUN(2) RUN(3)
    }
    return 0;
}
