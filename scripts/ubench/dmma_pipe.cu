// Is the FP64 tensor path (DMMA, mma.sync.m8n8k4.f64) a second FP64 pipe next to DFMA on sm_100a?
// The chain-batched Legendre contraction F[ring][col] += lambda[ring][l] * c[l][col] is the GEMM the north_star asks about;
// it can only pay if DMMA work overlaps the DFMA recurrence that generates lambda.  Three loops per variant:
//   dfma  : 16 independent DFMA chains per thread
//   dmma  : NACC independent m8n8k4 accumulators per warp
//   mixed : per loop trip NACC DMMAs and NF DFMAs (independent of each other)
// If the two instruction kinds went to separate pipes, t(mixed) ~ max(t(dfma part), t(dmma part)); on one shared pipe
// t(mixed) ~ t(dfma part) + t(dmma part).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_pipe dmma_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC, int NF>
__global__ void __launch_bounds__(128) k(double* out, int iters, double a, double b)
{
    double c0[NACC ? NACC : 1], c1[NACC ? NACC : 1], v[NF ? NF : 1];
    const double fa = a + 1e-9 * threadIdx.x, fb = b + 1e-9 * (threadIdx.x & 7);
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c0[i] = i; c1[i] = -i; }
#pragma unroll
    for (int i = 0; i < NF; ++i) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (NACC > NF ? NACC : NF); ++i) {
            if (i < NACC) dmma(c0[i], c1[i], fa, fb);
            if (i < NF) v[i] = fma(v[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += v[i];
    if (s == 123.456) out[0] = s;
}

template <int NACC, int NF>
float run(int ctas_per_sm, int iters)
{
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        k<NACC, NF><<<grid, 128>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    cudaFree(d);
    return best;
}

int main()
{
    const int iters = 1 << 13;
    printf("sm_100a: DFMA vs DMMA (m8n8k4.f64) pipes; 128-thread CTAs, iters %d; TFLOP/s counts 2 flop per DFMA lane and 512 flop per DMMA\n", iters);
    printf("%-10s %10s %10s %10s %12s %12s %14s\n", "CTAs/SM", "dfma16 ms", "dmma8 ms", "mixed ms", "dfma TF/s", "dmma TF/s", "mixed/(a+b)");
    for (int c : {1, 2, 4, 8}) {
        const float tf = run<0, 16>(c, iters), tm = run<8, 0>(c, iters), tx = run<8, 16>(c, iters);
        const double warps = 148.0 * c * 4;
        const double ff = 2.0 * 16 * iters * warps * 32 / (tf * 1e-3) * 1e-12;
        const double fm = 512.0 * 8 * iters * warps / (tm * 1e-3) * 1e-12;
        printf("%-10d %10.3f %10.3f %10.3f %12.2f %12.2f %14.3f\n", c, tf, tm, tx, ff, fm, tx / (tf + tm));
    }
    printf("mixed/(a+b) ~ 1: one shared FP64 pipe (no overlap);  ~ max(a,b)/(a+b): two pipes\n");
    // ratio sweep: DMMA count fixed, DFMA count varied
    printf("\n8 CTAs/SM, 8 DMMA + NF DFMA per trip:\n%-6s %10s %10s\n", "NF", "ms", "sum-model");
    const float tm = run<8, 0>(8, iters);
    const float t16 = run<0, 16>(8, iters);
    printf("%-6d %10.3f %10.3f\n", 4, run<8, 4>(8, iters), tm + t16 * 4 / 16);
    printf("%-6d %10.3f %10.3f\n", 8, run<8, 8>(8, iters), tm + t16 * 8 / 16);
    printf("%-6d %10.3f %10.3f\n", 16, run<8, 16>(8, iters), tm + t16);
    printf("%-6d %10.3f %10.3f\n", 32, run<8, 32>(8, iters), tm + t16 * 2);
    return 0;
}
