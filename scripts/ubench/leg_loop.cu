// The fast loop of leg_synth_kernel<2, R> in isolation: coefficients (e, b, r) of NL multipoles sit in shared memory, every
// thread runs the spin-2 recurrence + 8 accumulations per ring pair and l for R ring pairs, no barriers, no global traffic
// inside the loop.  Answers: how much of the FP64 pipe can this instruction mix reach at a given number of resident warps?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o leg_loop leg_loop.cu
#include <cstdio>
#include <cuda_runtime.h>

#define NL 256

template <int R, int NC, int VAR>   // VAR 0: as the kernel; 1: coefficients loaded once (no LDS in the loop); 2: snake order of the accumulations;
                                     // 3: loads of iteration i + 1 issued before the arithmetic of iteration i (register double buffer); 4: per-lane (vector) addresses
__global__ void __launch_bounds__(128) k(double* out, int reps, double seed)
{
    __shared__ double2 sE[NC][NL], sB[NC][NL], sR[NL];
    for (int i = threadIdx.x; i < NL; i += 128) {
        for (int c = 0; c < NC; ++c) {
            sE[c][i] = make_double2(1e-3 * i + c, 1.0 - 1e-3 * i);
            sB[c][i] = make_double2(0.5 + 1e-4 * i, -0.25 + c);
        }
        sR[i] = make_double2(2.0 - 1e-6 * i, 1e-3);
    }
    __syncthreads();
    double x[R], pc[R], pp[R], mc[R], mp[R], a[R][NC][8];
#pragma unroll
    for (int j = 0; j < R; ++j) {
        x[j] = 0.3 + 1e-3 * threadIdx.x + 0.1 * j + seed; pc[j] = 1e-3; pp[j] = 0; mc[j] = 2e-3; mp[j] = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int q = 0; q < 8; ++q) a[j][c][q] = 0;
    }
    for (int rep = 0; rep < reps; ++rep) {
        double2 nr0 = sR[0], nr1 = sR[1], ne0[NC], nb0[NC], ne1[NC], nb1[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) { ne0[c] = sE[c][0]; nb0[c] = sB[c][0]; ne1[c] = sE[c][1]; nb1[c] = sB[c][1]; }
        const int lz = VAR == 4 ? (int)(threadIdx.x >> 10) : 0;   // always 0, but a per-lane value as far as the compiler knows
#pragma unroll 2
        for (int i = 0; i < NL; i += 2) {
            const int ii = VAR == 1 ? 0 : i + lz;
            double2 r0, r1, e0[NC], b0[NC], e1[NC], b1[NC];
            if (VAR == 3) {
                r0 = nr0; r1 = nr1;
#pragma unroll
                for (int c = 0; c < NC; ++c) { e0[c] = ne0[c]; b0[c] = nb0[c]; e1[c] = ne1[c]; b1[c] = nb1[c]; }
                const int in = (i + 2) & (NL - 1);
                nr0 = sR[in]; nr1 = sR[in + 1];
#pragma unroll
                for (int c = 0; c < NC; ++c) { ne0[c] = sE[c][in]; nb0[c] = sB[c][in]; ne1[c] = sE[c][in + 1]; nb1[c] = sB[c][in + 1]; }
            } else {
                r0 = sR[ii]; r1 = sR[ii + 1];
#pragma unroll
                for (int c = 0; c < NC; ++c) { e0[c] = sE[c][ii]; b0[c] = sB[c][ii]; e1[c] = sE[c][ii + 1]; b1[c] = sB[c][ii + 1]; }
            }
#pragma unroll
            for (int j = 0; j < R; ++j) {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    double* A = a[j][c];
                    if (VAR == 2) {
                        A[0] = fma(pc[j], e0[c].x, A[0]); A[5] = fma(pc[j], e0[c].y, A[5]); A[2] = fma(pc[j], b0[c].x, A[2]); A[7] = fma(pc[j], b0[c].y, A[7]);
                        A[3] = fma(mc[j], b0[c].y, A[3]); A[6] = fma(mc[j], b0[c].x, A[6]); A[1] = fma(mc[j], e0[c].y, A[1]); A[4] = fma(mc[j], e0[c].x, A[4]);
                    } else {
                    A[0] = fma(pc[j], e0[c].x, A[0]); A[1] = fma(mc[j], e0[c].y, A[1]); A[2] = fma(pc[j], b0[c].x, A[2]); A[3] = fma(mc[j], b0[c].y, A[3]);
                    A[4] = fma(mc[j], e0[c].x, A[4]); A[5] = fma(pc[j], e0[c].y, A[5]); A[6] = fma(mc[j], b0[c].x, A[6]); A[7] = fma(pc[j], b0[c].y, A[7]);
                    }
                }
                {
                    const double tp = fma(r0.x, x[j], r0.y), tm = fma(r0.x, x[j], -r0.y);
                    const double np = fma(tp, pc[j], -pp[j]), nm = fma(tm, mc[j], -mp[j]);
                    pp[j] = pc[j]; pc[j] = np; mp[j] = mc[j]; mc[j] = nm;
                }
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    double* A = a[j][c];
                    A[0] = fma(pc[j], e1[c].x, A[0]); A[1] = fma(mc[j], e1[c].y, A[1]); A[2] = fma(pc[j], b1[c].x, A[2]); A[3] = fma(mc[j], b1[c].y, A[3]);
                    A[4] = fma(-mc[j], e1[c].x, A[4]); A[5] = fma(-pc[j], e1[c].y, A[5]); A[6] = fma(-mc[j], b1[c].x, A[6]); A[7] = fma(-pc[j], b1[c].y, A[7]);
                }
                {
                    const double tp = fma(r1.x, x[j], r1.y), tm = fma(r1.x, x[j], -r1.y);
                    const double np = fma(tp, pc[j], -pp[j]), nm = fma(tm, mc[j], -mp[j]);
                    pp[j] = pc[j]; pc[j] = np; mp[j] = mc[j]; mc[j] = nm;
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int q = 0; q < 8; ++q) s += a[j][c][q];
    if (s == 123.456) out[0] = s;
}

template <int R, int NC, int VAR>
void run(const char* name)
{
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 64;
    int maxb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, k<R, NC, VAR>, 128, 0);
    printf("%-18s max CTAs/SM %d :", name, maxb);
    for (int c : {1, 2, 3, 4, 6, 8}) {
        if (c > maxb) break;
        float best = 1e30f;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0);
            k<R, NC, VAR><<<148 * c, 128>>>(d, reps, 0.0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r && ms < best) best = ms;
        }
        const double dfma = (double)reps * NL * R * (4.0 + 8.0 * NC) * 148.0 * c * 128.0;
        printf("  %d: %6.2f TF/s", c, 2.0 * dfma / (best * 1e-3) * 1e-12);
    }
    printf("\n");
    cudaFree(d);
}

int main()
{
    printf("leg_synth fast loop in isolation (2 flop per executed DFMA); columns = CTAs (of 4 warps) per SM\n");
    run<2, 1, 0>("R=2 NC=1");
    run<4, 1, 0>("R=4 NC=1");
    run<2, 2, 0>("R=2 NC=2");
    run<2, 1, 1>("R=2 NC=1 no LDS");
    run<2, 2, 1>("R=2 NC=2 no LDS");
    run<2, 1, 2>("R=2 NC=1 snake");
    run<2, 1, 3>("R=2 NC=1 prefetch");
    run<4, 1, 3>("R=4 NC=1 prefetch");
    run<2, 2, 3>("R=2 NC=2 prefetch");
    run<2, 1, 4>("R=2 NC=1 vec addr");
    return 0;
}
