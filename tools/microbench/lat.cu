// Latency / issue-rate microbenchmarks for the instruction mix of the IPM kernel (sm_100a).
// One warp per measurement, clock64 around an unrolled dependent chain.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k_dfma_lat(double* out, long long* cyc, double a, double b) {
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dadd_lat(double* out, long long* cyc, double a) {
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) x = x + a;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
__global__ void k_dfma_ilp(double* out, long long* cyc, double a, double b) {
    double x[ILP];
    for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x + j;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < ILP; ++j) s += x[j];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds_lat(double* out, long long* cyc) {
    __shared__ int nxt[256];
    for (int i = threadIdx.x; i < 256; i += 32) nxt[i] = (i + 33) & 255;
    __syncwarp();
    int p = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) p = nxt[p];
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds64_fma(double* out, long long* cyc, double a) {
    // chain: LDS.64 (address independent) feeding a DFMA chain through the loaded value: x = fma(x, a, s[i])
    __shared__ double s[64];
    for (int i = threadIdx.x; i < 64; i += 32) s[i] = i * 1e-3;
    __syncwarp();
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) x = fma(x, a, s[(i + threadIdx.x) & 63]);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rcp_lat(double* out, long long* cyc) {
    double x = 1.5 + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) {
        double y;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        x = y;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_div_lat(double* out, long long* cyc, double a) {
    double x = 1.5 + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = a / x;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl_lat(double* out, long long* cyc) {
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_xchg_lat(double* out, long long* cyc) {
    // STS -> __syncwarp -> LDS of a neighbour's value, dependent round trip
    __shared__ double s[32];
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        s[threadIdx.x] = x;
        __syncwarp();
        x = s[(threadIdx.x + 1) & 31];
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_xchg1_lat(double* out, long long* cyc) {
    // double-buffered: one __syncwarp per round trip
    __shared__ double s[64];
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        s[(i & 1) * 32 + threadIdx.x] = x;
        __syncwarp();
        x = s[(i & 1) * 32 + ((threadIdx.x + 1) & 31)];
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_syncthreads(double* out, long long* cyc) {
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    long long h;
#define RUN(name, launch) launch; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-34s %8.2f cycles/op\n", name, (double)h / N);
    for (int rep = 0; rep < 2; ++rep) {
        RUN("DFMA dependent", (k_dfma_lat<<<1, 32>>>(out, cyc, 1.0000001, 1e-9)))
        RUN("DADD dependent", (k_dadd_lat<<<1, 32>>>(out, cyc, 1e-9)))
        RUN("DFMA ILP2 (per group of 2)", (k_dfma_ilp<2><<<1, 32>>>(out, cyc, 1.0000001, 1e-9)))
        RUN("DFMA ILP4 (per group of 4)", (k_dfma_ilp<4><<<1, 32>>>(out, cyc, 1.0000001, 1e-9)))
        RUN("DFMA ILP8 (per group of 8)", (k_dfma_ilp<8><<<1, 32>>>(out, cyc, 1.0000001, 1e-9)))
        RUN("DFMA ILP8, 1 active lane", (k_dfma_ilp<8><<<1, 1>>>(out, cyc, 1.0000001, 1e-9)))
        RUN("DFMA ILP8, 4 warps (1/SMSP)", (k_dfma_ilp<8><<<1, 128>>>(out, cyc, 1.0000001, 1e-9)))
        RUN("DFMA ILP8, 8 warps (2/SMSP)", (k_dfma_ilp<8><<<1, 256>>>(out, cyc, 1.0000001, 1e-9)))
        RUN("LDS.32 pointer chase", (k_lds_lat<<<1, 32>>>(out, cyc)))
        RUN("DFMA with LDS.64 operand", (k_lds64_fma<<<1, 32>>>(out, cyc, 1.0000001)))
        RUN("MUFU.RCP64H dependent", (k_rcp_lat<<<1, 32>>>(out, cyc)))
        RUN("double division dependent", (k_div_lat<<<1, 32>>>(out, cyc, 1.0000001)))
        RUN("SHFL (double = 2x) dependent", (k_shfl_lat<<<1, 32>>>(out, cyc)))
        RUN("STS+syncwarp+LDS+syncwarp", (k_xchg_lat<<<1, 32>>>(out, cyc)))
        RUN("STS+syncwarp+LDS (dbl buffer)", (k_xchg1_lat<<<1, 32>>>(out, cyc)))
        RUN("__syncthreads 256 thr", (k_syncthreads<<<1, 256>>>(out, cyc)))
    }
    return 0;
}
