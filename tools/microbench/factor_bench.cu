// Isolated timing of the one-warp cooperative Riccati stage (Coop<4>) for 1..8 resident warps per SM:
// what the per-stage dependency chain costs on sm_100a and how warps sharing a scheduler interact.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../learning-based-mpc_b200/csrc/lbmpc_core.cuh"
using namespace lbmpc;
using CP = Coop<4>;
using P4 = Params<4, 1, 1>;
using L4 = Layout<4, 1, 1>;
constexpr unsigned kFull = 0xffffffffu;
#ifndef REPS
#define REPS 20
#endif

template <int MODE>
__global__ void k_factor(const __grid_constant__ P4 p, long long* cyc, double* out, int reps) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const L4 l(p.N, p.ngp);
    double* sl = smem;  // all warps read the same slot (stores suppressed in MODE 1)
    double* xch = smem + ((l.stride + 3) & ~1) + warp * 40;
    for (int i = threadIdx.x; i < l.stride + 2; i += blockDim.x) smem[i] = 0.01 * (i % 7) + 0.5;
    __syncthreads();
    typename CP::Lane ln;
    CP::lane_init(p, lane, ln);
    CP::xch_init(lane, xch);
    __syncthreads();
    const int N = p.N;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        CP::terminal(p, l, sl, ln);
        for (int k = N - 1; k >= 0; --k) {
            CP::st1(p, l, sl, k, lane, ln, xch);
            __syncwarp();
            CP::st2(l, sl, k + 1, lane, ln, xch);
            const double fa = __shfl_sync(kFull, ln.pub, CP::kFz + ln.a);
            const double fb = __shfl_sync(kFull, ln.pub, CP::kFz + ln.b);
            const double fuu = __shfl_sync(kFull, ln.pub, CP::kFu);
            CP::st3(ln, fa, fb, fuu);
            if (MODE == 1) ln.val = ln.val * 1e-3 + 1.0;  // keep values tame
            __syncwarp();
        }
        CP::finish(l, sl, lane, ln);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = ln.val;
}

int main() {
    P4 p{};
    p.N = 50; p.ng = 24; p.ngp = 24; p.kg = 1; p.kT = 50; p.kx0 = 1; p.kx1 = 50; p.ku0 = 0; p.ku1 = 49; p.ntypes = 2;
    p.tseg[0] = 0; p.tseg[1] = 50; p.tseg[2] = p.tseg[3] = p.tseg[4] = 51; p.rowmask = 0x3ff; p.max_iter = 60; p.m_rows = 524;
    for (int i = 0; i < 16; ++i) p.A[i] = (i % 5 == 0) ? 0.9 : 0.01 * i;
    for (int i = 0; i < 4; ++i) p.B[i] = 0.1 * (i + 1);
    for (int t = 0; t < 4; ++t) for (int i = 0; i < 36; ++i) p.W[t][i] = (i % 7 == 0) ? 2.0 : 0.01;
    long long* cyc; double* out;
    cudaMalloc(&cyc, 8); cudaMalloc(&out, 256 * 8);
    const L4 l(p.N, p.ngp);
    size_t smem = (l.stride + 8 * 40 + 16) * 8;
    cudaFuncSetAttribute(k_factor<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_factor<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep)
        for (int warps = 1; warps <= 8; warps *= 2) {
            k_factor<0><<<1, 32 * warps, smem>>>(p, cyc, out, REPS);
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("one-warp stage, %d warps/CTA : %8.1f cycles/stage  (%s)\n", warps, (double)h / ((double)REPS * 50), cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
