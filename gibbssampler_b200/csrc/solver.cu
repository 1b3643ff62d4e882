// Constrained-realization solves.
//
//   Q x = b,  Q = C^-1 + B A^T N^-1 A B   (CenteredGibbs.py:448-491 + the forked qcinv's
//   multigrid_chain.sample / opfilt_pp.fwd_op / diag_cl preconditioner / cd_solve, not vendored)
//
// Everything lives in the reference's real alm layout so that dot products are plain sums
// (= qcinv's (2 - delta_m0)-weighted complex dot).  The mat-vec is the spin-2 SHT pair of
// legendre.cu / ringfft.cu with b_l fused into the Legendre staging on both sides and N^-1 fused
// into the ring-analysis load; the vector updates are three fused HBM-bound kernels whose
// reductions finish on the device (last-block pattern, fixed summation order => bit-reproducible),
// so alpha/beta and the convergence flag never travel to the host inside the loop; once the flag
// is set every later kernel (SHT stages included) returns immediately.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "gs_internal.h"
#include "rng.cuh"

#define SV_NT 256
#define SV_GRID (148 * 4)

struct PcgState {
    double delta, pq, rr, rz, d0, alpha, beta, eps2;
    int iter, itermax, done;
    unsigned counter;
};

struct gs_pcg_ws {
    double *r[2], *p[2], *q[2], *invc[2], *pre[2];
    // r02: the solver regenerates C^-1_l and the preconditioner M_l per coefficient from per-l tables (L1-resident) and a 2-byte
    // multipole index per coefficient instead of streaming two expanded 8-byte arrays: 12 instead of 15 array passes per iteration
    const unsigned short* lof;   // [n] multipole of every coefficient of the (local) real layout
    const double* ic_l[2];       // [L+1] 1 / C_l (E, B)
    const double* pre_l[2];      // [L+1] M_l
    double* partials;  // SV_GRID * 2
    double* fuse_partials;  // per-block partials of the fused analysis-finish + <p, q> kernel (unsharded plans)
    double* fuse_out;       // its result
    double* red;       // sharded plans: local sums in, all-reduced sums out (NULL on one GPU)
    PcgState* state;
    PcgState* host_state;  // pinned, two slots (the graph path of a solve keeps two polls in flight)
};

// block-level sum of NR values, then the last block to finish adds the per-block partials in a
// fixed order.  Returns true (in every thread of that last block) with the totals in tot[].
template <int NR>
__device__ bool grid_reduce(double (&v)[NR], double* partials, unsigned* counter, double (&tot)[NR])
{
    __shared__ double sm[NR][SV_NT / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        double s = v[k];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sm[k][w] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            double s = 0.0;
            for (int i = 0; i < SV_NT / 32; ++i) s += sm[k][i];
            partials[blockIdx.x * NR + k] = s;
        }
        __threadfence();
        const unsigned t = atomicInc(counter, gridDim.x - 1);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        double s = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += partials[i * NR + k];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        __syncthreads();
        if (lane == 0) sm[k][w] = s;
        __syncthreads();
        double t = 0.0;
        for (int i = 0; i < SV_NT / 32; ++i) t += sm[k][i];
        tot[k] = t;
    }
    return true;
}

// diag_cl preconditioner of qcinv's opfilt_pp: M_l = 1 / (1/C_l + b_l^2 sum(N^-1)/(4 pi))
__global__ void precond_kernel(const double* __restrict__ dl, const double* __restrict__ bl, double ninv, int L,
                               double* __restrict__ invc_l, double* __restrict__ pre_l)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l > L) return;
    double c = dl[l];
    if (l) c = c * 2.0 * 3.14159265358979323846 / ((double)l * (double)(l + 1));
    const double ic = c != 0.0 ? 1.0 / c : 0.0;
    const double d = ic + bl[l] * bl[l] * ninv;
    invc_l[l] = ic;
    pre_l[l] = d != 0.0 ? 1.0 / d : 0.0;
}

__device__ __forceinline__ void pick(int64_t i, int64_t n, int& c, int64_t& j)
{
    c = i >= n;
    j = c ? i - n : i;
}

// scalar updates that follow each reduction.  One GPU: run by the last block of the reducing kernel.
// Sharded: that block stores its local sums in W.red, the host enqueues the all-reduce and then
// pcg_scalar_kernel, so every rank applies the same update to the same numbers.
__device__ __forceinline__ void fin_init(PcgState* s, double rr, double rz)
{
    s->rr = rr; s->d0 = rr; s->delta = rz; s->iter = 0;
    s->done = (rr <= 0.0 || s->itermax <= 0) ? 1 : 0;
}
__device__ __forceinline__ void fin_apq(PcgState* s, double pq)
{
    s->pq = pq;
    s->alpha = pq != 0.0 ? s->delta / pq : 0.0;
}
__device__ __forceinline__ void fin_update(PcgState* s, double rr, double rz)
{
    s->rr = rr; s->rz = rz;
    s->beta = s->delta != 0.0 ? rz / s->delta : 0.0;
    s->delta = rz;
    s->iter += 1;
    // qcinv cd_monitors.monitor_basic: stop when <r,r> <= eps^2 <r0,r0> or iter >= iter_max
    if (rr <= s->eps2 * s->d0 || s->iter >= s->itermax) s->done = 1;
}
// after the fused analysis-finish kernel: alpha = delta / <p, q>
__global__ void pcg_apq_scalar_kernel(gs_pcg_ws W)
{
    if (W.state->done) return;
    fin_apq(W.state, W.fuse_out[0]);
}
__global__ void pcg_scalar_kernel(gs_pcg_ws W, int stage)
{
    PcgState* s = W.state;
    if (stage == 0) fin_init(s, W.red[0], W.red[1]);
    else if (s->done) return;
    else if (stage == 1) fin_apq(s, W.red[0]);
    else fin_update(s, W.red[0], W.red[1]);
}

// r = b - q (q = Q x0, or q absent for x0 = 0), p = M r, delta = <r, M r>, d0 = rr = <r, r>
__global__ void __launch_bounds__(SV_NT)
pcg_init_kernel(gs_pcg_ws W, const double* bE, const double* bB, int have_q, int64_t n, int nc)
{
    double v[2] = {0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nc * n; i += (int64_t)gridDim.x * blockDim.x) {
        int c; int64_t j;
        pick(i, n, c, j);
        double r = (c ? bB : bE)[j];
        if (have_q) r -= (c ? W.q[1] : W.q[0])[j];
        const double z = (c ? W.pre_l[1] : W.pre_l[0])[W.lof[j]] * r;
        (c ? W.r[1] : W.r[0])[j] = r;
        (c ? W.p[1] : W.p[0])[j] = z;
        v[0] += r * r;
        v[1] += r * z;
    }
    double tot[2];
    if (grid_reduce<2>(v, W.partials, &W.state->counter, tot) && threadIdx.x == 0) {
        if (W.red) { W.red[0] = tot[0]; W.red[1] = tot[1]; }
        else fin_init(W.state, tot[0], tot[1]);
    }
}

// The vector kernels walk the E and B arrays with a grid stride, SV_U elements per thread and step with all loads issued
// before the first store (the arrays may alias as far as the compiler knows): enough bytes in flight to approach HBM speed.
#define SV_U 4

// q += C^-1 p ; pq = <p, q> ; alpha = delta / pq
__global__ void __launch_bounds__(SV_NT) pcg_apq_kernel(gs_pcg_ws W, int64_t n, int nc)
{
    if (W.state->done) return;
    double v[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, total = nc * n;
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < total; i0 += SV_U * stride) {
        double p[SV_U], q[SV_U], ic[SV_U];
#pragma unroll
        for (int k = 0; k < SV_U; ++k) {
            const int64_t i = i0 + k * stride;
            int c; int64_t j;
            pick(i, n, c, j);
            const bool ok = i < total;
            p[k] = ok ? (c ? W.p[1] : W.p[0])[j] : 0.0; q[k] = ok ? (c ? W.q[1] : W.q[0])[j] : 0.0;
            ic[k] = ok ? __ldg((c ? W.ic_l[1] : W.ic_l[0]) + W.lof[j]) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < SV_U; ++k) {
            const int64_t i = i0 + k * stride;
            if (i >= total) break;
            int c; int64_t j;
            pick(i, n, c, j);
            const double qq = fma(ic[k], p[k], q[k]);
            (c ? W.q[1] : W.q[0])[j] = qq;
            v[0] += p[k] * qq;
        }
    }
    double tot[1];
    if (grid_reduce<1>(v, W.partials, &W.state->counter, tot) && threadIdx.x == 0) {
        if (W.red) W.red[0] = tot[0];
        else fin_apq(W.state, tot[0]);
    }
}

// x += alpha p ; r -= alpha q ; rr = <r,r> ; rz = <r, M r> ; beta = rz/delta ; convergence test
__global__ void __launch_bounds__(SV_NT) pcg_update_kernel(gs_pcg_ws W, double* xE, double* xB, int64_t n, int nc)
{
    if (W.state->done) return;
    const double alpha = W.state->alpha;
    double v[2] = {0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, total = nc * n;
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < total; i0 += SV_U * stride) {
        double x[SV_U], p[SV_U], q[SV_U], r[SV_U], pre[SV_U];
#pragma unroll
        for (int k = 0; k < SV_U; ++k) {
            const int64_t i = i0 + k * stride;
            int c; int64_t j;
            pick(i, n, c, j);
            const bool ok = i < total;
            x[k] = ok ? (c ? xB : xE)[j] : 0.0; p[k] = ok ? (c ? W.p[1] : W.p[0])[j] : 0.0; q[k] = ok ? (c ? W.q[1] : W.q[0])[j] : 0.0;
            r[k] = ok ? (c ? W.r[1] : W.r[0])[j] : 0.0; pre[k] = ok ? __ldg((c ? W.pre_l[1] : W.pre_l[0]) + W.lof[j]) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < SV_U; ++k) {
            const int64_t i = i0 + k * stride;
            if (i >= total) break;
            int c; int64_t j;
            pick(i, n, c, j);
            (c ? xB : xE)[j] = fma(alpha, p[k], x[k]);
            const double rr = fma(-alpha, q[k], r[k]);
            (c ? W.r[1] : W.r[0])[j] = rr;
            v[0] += rr * rr;
            v[1] += rr * rr * pre[k];
        }
    }
    double tot[2];
    if (grid_reduce<2>(v, W.partials, &W.state->counter, tot) && threadIdx.x == 0) {
        if (W.red) { W.red[0] = tot[0]; W.red[1] = tot[1]; }
        else fin_update(W.state, tot[0], tot[1]);
    }
}

// p = M r + beta p
__global__ void __launch_bounds__(SV_NT) pcg_dir_kernel(gs_pcg_ws W, int64_t n, int nc)
{
    if (W.state->done) return;
    const double beta = W.state->beta;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, total = nc * n;
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < total; i0 += SV_U * stride) {
        double p[SV_U], r[SV_U], pre[SV_U];
#pragma unroll
        for (int k = 0; k < SV_U; ++k) {
            const int64_t i = i0 + k * stride;
            int c; int64_t j;
            pick(i, n, c, j);
            const bool ok = i < total;
            p[k] = ok ? (c ? W.p[1] : W.p[0])[j] : 0.0; r[k] = ok ? (c ? W.r[1] : W.r[0])[j] : 0.0;
            pre[k] = ok ? __ldg((c ? W.pre_l[1] : W.pre_l[0]) + W.lof[j]) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < SV_U; ++k) {
            const int64_t i = i0 + k * stride;
            if (i >= total) break;
            int c; int64_t j;
            pick(i, n, c, j);
            (c ? W.p[1] : W.p[0])[j] = fma(beta, p[k], pre[k] * r[k]);
        }
    }
}

__global__ void zero2_kernel(double* a, double* b, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        a[i] = 0.0;
        b[i] = 0.0;
    }
}

// out = a + x * y   (per coefficient, real layout)
__global__ void axy_kernel(const double* a, const double* __restrict__ x, const double* __restrict__ y, double* out, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = fma(x[i], y[i], a ? a[i] : 0.0);
}

__global__ void scale_map_kernel(const double* __restrict__ a, const double* __restrict__ w, double* __restrict__ out, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = a[i] * w[i];
}

__global__ void rhs_combine_kernel(double* rhs, const double* sic, const double* xi, const double* bdata, double resc, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        rhs[i] = fma(resc, rhs[i], fma(sic[i], xi[i], bdata[i]));
}

// multipole of every coefficient of the real layout (unsharded: index algebra of utils.py:49-76; sharded: the plan's l_of_loc)
__global__ void lof_build_kernel(unsigned short* __restrict__ lof, const int* __restrict__ l_of_loc, int L, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        lof[i] = (unsigned short)(l_of_loc ? l_of_loc[i] : l_of_real(i, L));
}
static bool build_lof(gs_plan* p, const unsigned short** out)
{
    void* d = nullptr;
    const int64_t n = p->nreal_loc;
    if (cudaMalloc(&d, (size_t)n * sizeof(unsigned short)) != cudaSuccess) return false;
    p->owned.push_back(d);
    lof_build_kernel<<<SV_GRID, SV_NT>>>((unsigned short*)d, p->world > 1 ? p->d.sh.l_of_loc : nullptr, p->d.lmax, n);
    if (cudaDeviceSynchronize() != cudaSuccess) return false;
    *out = (const unsigned short*)d;
    return true;
}

// ------------------------------------------------------------------ workspace
// The workspace belongs to the plan (gs_plan::pcg_ws): created on the first solve, released by gs_plan_destroy through
// gs_pcg_ws_free.  A plan is used by one host thread at a time (include/gibbs_b200.h), so no lock is taken.
static gs_pcg_ws* get_ws(gs_plan* p)
{
    if (p->pcg_ws) return (gs_pcg_ws*)p->pcg_ws;
    gs_pcg_ws* w = new gs_pcg_ws();
    const size_t n = (size_t)p->nreal_loc;
    w->red = p->world > 1 ? p->red_loc : nullptr;
    auto alloc = [&](double** q, size_t cnt) { void* d = nullptr; if (cudaMalloc(&d, cnt * sizeof(double)) != cudaSuccess) return false; p->owned.push_back(d); *q = (double*)d; return true; };
    bool ok = true;
    for (int c = 0; c < 2 && ok; ++c) ok = alloc(&w->r[c], n) && alloc(&w->p[c], n) && alloc(&w->q[c], n) && alloc(&w->invc[c], n) && alloc(&w->pre[c], n);
    ok = ok && alloc(&w->partials, SV_GRID * 2 + 4 * (size_t)(p->d.lmax + 1));
    ok = ok && build_lof(p, &w->lof);
    if (ok) {
        double* tl = w->partials + SV_GRID * 2;   // [icE | preE | icB | preB], filled by precond_kernel at the start of a solve
        const int L1 = p->d.lmax + 1;
        w->ic_l[0] = tl; w->pre_l[0] = tl + L1; w->ic_l[1] = tl + 2 * L1; w->pre_l[1] = tl + 3 * L1;
    }
    ok = ok && alloc(&w->fuse_partials, (size_t)((p->d.lmax + 256) / 256) * (p->d.lmax + 1));
    ok = ok && alloc(&w->fuse_out, 4);
    void* d = nullptr;
    ok = ok && cudaMalloc(&d, sizeof(PcgState)) == cudaSuccess;
    if (ok) { p->owned.push_back(d); w->state = (PcgState*)d; }
    ok = ok && cudaMallocHost((void**)&w->host_state, 2 * sizeof(PcgState)) == cudaSuccess;
    if (!ok) { gs_set_error("PCG workspace allocation failed: %s", cudaGetErrorString(cudaGetLastError())); delete w; return nullptr; }
    p->pcg_ws = w;
    return w;
}

// Chain batch (gs_cr_pcg_pol_batch): two workspaces whose p and q vectors lie one batch stride (2 n doubles: E then B of a chain)
// apart, so that the chain-batched mat-vec reads / writes them as one strided array.  Allocated on the first batched solve.
static gs_pcg_ws* get_ws_batch(gs_plan* p)
{
    if (p->pcg_ws_batch) return (gs_pcg_ws*)p->pcg_ws_batch;
    gs_pcg_ws* w = new gs_pcg_ws[2]();
    const size_t n = (size_t)p->nreal_loc;
    auto alloc = [&](double** q, size_t cnt) { void* d = nullptr; if (cudaMalloc(&d, cnt * sizeof(double)) != cudaSuccess) return false; p->owned.push_back(d); *q = (double*)d; return true; };
    double *P = nullptr, *Q = nullptr;
    bool ok = alloc(&P, 4 * n) && alloc(&Q, 4 * n);
    for (int k = 0; k < 2 && ok; ++k) {
        w[k].red = nullptr;
        for (int c = 0; c < 2 && ok; ++c) {
            w[k].p[c] = P + (2 * k + c) * n;
            w[k].q[c] = Q + (2 * k + c) * n;
            ok = alloc(&w[k].r[c], n) && alloc(&w[k].invc[c], n) && alloc(&w[k].pre[c], n);
        }
        ok = ok && alloc(&w[k].partials, SV_GRID * 2 + 4 * (size_t)(p->d.lmax + 1));
        if (ok) {
            if (k == 0) ok = build_lof(p, &w[0].lof); else w[1].lof = w[0].lof;
            double* tl = w[k].partials + SV_GRID * 2;
            const int L1 = p->d.lmax + 1;
            w[k].ic_l[0] = tl; w[k].pre_l[0] = tl + L1; w[k].ic_l[1] = tl + 2 * L1; w[k].pre_l[1] = tl + 3 * L1;
        }
        w[k].fuse_partials = nullptr;
        ok = ok && alloc(&w[k].fuse_out, 4);
        void* d = nullptr;
        ok = ok && cudaMalloc(&d, sizeof(PcgState)) == cudaSuccess;
        if (ok) { p->owned.push_back(d); w[k].state = (PcgState*)d; }
        ok = ok && cudaMallocHost((void**)&w[k].host_state, 2 * sizeof(PcgState)) == cudaSuccess;
    }
    void* d = nullptr;
    ok = ok && cudaMalloc(&d, sizeof(int)) == cudaSuccess;
    if (ok) { p->owned.push_back(d); p->pcg_alldone = (int*)d; }
    if (!ok) { gs_set_error("batched PCG workspace allocation failed: %s", cudaGetErrorString(cudaGetLastError())); delete[] w; return nullptr; }
    p->pcg_ws_batch = w;
    return w;
}

void gs_pcg_ws_free(gs_plan* p)
{
    gs_pcg_ws* w = (gs_pcg_ws*)p->pcg_ws;
    if (w) {
        if (w->host_state) cudaFreeHost(w->host_state);   // the device buffers are in p->owned
        delete w;
        p->pcg_ws = nullptr;
    }
    gs_pcg_ws* b = (gs_pcg_ws*)p->pcg_ws_batch;
    if (b) {
        for (int k = 0; k < 2; ++k) if (b[k].host_state) cudaFreeHost(b[k].host_state);
        delete[] b;
        p->pcg_ws_batch = nullptr;
    }
}

int g_gs_fuse_apq = 0;   // 1: fused analysis-finish + (q += C^-1 p, <p, q>) kernel in the unsharded PCG (measured 0.4 % SLOWER
                         // than the separate pass at NSIDE 512: 5125 block partials + a scalar kernel; kept as an option)
extern "C" int gs_set_fuse_apq(int on) { const int old = g_gs_fuse_apq; g_gs_fuse_apq = on ? 1 : 0; return old; }
int g_gs_pcg_graph = 1;   // 1: the iterations between two convergence polls of an unsharded PCG solve are replayed from a CUDA graph
extern "C" int gs_set_pcg_graph(int on) { const int old = g_gs_pcg_graph; g_gs_pcg_graph = on ? 1 : 0; return old; }
int g_gs_ring_fused = 1;
extern "C" int gs_set_ring_fused(int fused) { const int old = g_gs_ring_fused; g_gs_ring_fused = fused ? 1 : 0; return old; }

// Rings whose N^-1 is identically zero (inside the mask) add nothing to A^T N^-1 A: the mat-vec walks the rings that carry
// weight only (gs_active_rings_build; lists rebuilt from the weight map of each solve, on the device).
int g_gs_ring_skip = 1;
// Rings on which N^-1 is constant (isotropic noise, ring not cut by the mask edge): DFT^H diag(w) DFT = n w on the alias-folded
// spectrum, so the fused ring stage of the mat-vec needs no transform there (ring_apply_kernel; flags from gs_active_rings_build).
int g_gs_ring_const = 1;
extern "C" int gs_set_ring_const(int on) { const int old = g_gs_ring_const; g_gs_ring_const = on ? 1 : 0; return old; }
extern "C" int gs_set_ring_skip(int on) { const int old = g_gs_ring_skip; g_gs_ring_skip = on ? 1 : 0; return old; }

struct ActiveRings {   // scope guard: the plan's launchers use the active lists while one of these lives
    gs_plan* p;
    bool on;
    ActiveRings(gs_plan* p_) : p(p_), on(false) {}
    int begin(const double* pixw, cudaStream_t st)
    {
        if (!g_gs_ring_skip || !pixw) return GS_OK;
        int rc = gs_active_rings_build(p, pixw, st);
        if (rc == GS_OK) { p->use_act = true; on = true; }
        return rc;
    }
    ~ActiveRings() { if (on) p->use_act = false; }
};

// q = B A^T N^-1 A B v   (C^-1 v is added by pcg_apq_kernel / the caller)
// nc = 2 (chain batch): chain c reads vE/vB + c stride and writes qE/qB + c stride; one recurrence serves both chains
static int apply_noise_op(gs_plan* p, const double* vE, const double* vB, const double* bl, const double* inv_noise,
                          double* qE, double* qB, cudaStream_t st, const int* skip, int spin = 2, const FinishFuse* fuse = nullptr,
                          int nc = 1, int64_t stride = 0)
{
    int rc;
    if ((rc = gs_leg_synth(p, spin, vE, vB, GS_ALM_REAL, bl, st, skip, nullptr, nc, stride))) return rc;
    if (g_gs_ring_fused || nc > 1) {
        if ((rc = gs_ring_apply(p, spin, inv_noise, st, skip, nc))) return rc;
    } else {
        if ((rc = gs_ring_synth(p, spin, p->mapQ_tmp, p->mapU_tmp, st, skip))) return rc;
        if ((rc = gs_ring_anal(p, spin, p->mapQ_tmp, p->mapU_tmp, inv_noise, st, skip))) return rc;
    }
    return gs_leg_anal(p, spin, qE, qB, GS_ALM_REAL, bl, 1.0, 0, st, skip, fuse, nc, stride);
}

// spin 2: (E, B) system; spin 0: temperature (dl_BB, rhs_B, x_B unused)
static int cr_pcg_impl(gs_plan* p, int spin, const double* dl_EE, const double* dl_BB, const double* bl,
                       const double* inv_noise, double ninv_sum_over_4pi, const double* rhs_E,
                       const double* rhs_B, double* x_E, double* x_B, int warm_start, double eps, int itermax,
                       int check_every, int* n_iter_out, double* resid_out, void* stream)
{
    const int nc = spin ? 2 : 1;
    GS_REQUIRE(eps > 0.0 && itermax >= 0, "eps must be > 0 and itermax >= 0");
    GS_CHECK_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    gs_pcg_ws* w = get_ws(p);
    if (!w) return GS_E_NOMEM;
    const int L = p->d.lmax;
    const int64_t n = p->nreal_loc;
    if (check_every < 1) check_every = 8;

    // per-l C^-1 and preconditioner, expanded to the real layout
    double* tmp_l = w->partials + SV_GRID * 2;  // 4 (L+1) doubles of scratch
    const int lb = (L + 256) / 256;
    precond_kernel<<<lb, 256, 0, st>>>(dl_EE, bl, ninv_sum_over_4pi, L, tmp_l, tmp_l + (L + 1));
    if (nc == 2) precond_kernel<<<lb, 256, 0, st>>>(dl_BB, bl, ninv_sum_over_4pi, L, tmp_l + 2 * (L + 1), tmp_l + 3 * (L + 1));
    GS_CHECK_LAUNCH();
    int rc;
    // (the per-l tables are used as they are: the vector kernels index them with the multipole of each coefficient, w->lof)
    const bool dist = p->world > 1;
    ActiveRings act(p);
    if ((rc = act.begin(inv_noise, st))) return rc;

    PcgState h;
    memset(&h, 0, sizeof(h));
    h.eps2 = eps * eps;
    h.itermax = itermax;
    *w->host_state = h;
    GS_CHECK_CUDA(cudaMemcpyAsync(w->state, w->host_state, sizeof(PcgState), cudaMemcpyHostToDevice, st));
    if (warm_start) {  // r0 = b - Q x0
        if ((rc = gs_plan_expand_per_l(p, tmp_l, 0, w->invc[0], st))) return rc;
        if (nc == 2 && (rc = gs_plan_expand_per_l(p, tmp_l + 2 * (L + 1), 0, w->invc[1], st))) return rc;
        if ((rc = apply_noise_op(p, x_E, x_B, bl, inv_noise, w->q[0], w->q[1], st, nullptr, spin))) return rc;
        axy_kernel<<<SV_GRID, SV_NT, 0, st>>>(w->q[0], w->invc[0], x_E, w->q[0], n);
        if (nc == 2) axy_kernel<<<SV_GRID, SV_NT, 0, st>>>(w->q[1], w->invc[1], x_B, w->q[1], n);
    } else {
        zero2_kernel<<<SV_GRID, SV_NT, 0, st>>>(x_E, nc == 2 ? x_B : x_E, n);
    }
    pcg_init_kernel<<<SV_GRID, SV_NT, 0, st>>>(*w, rhs_E, rhs_B, warm_start ? 1 : 0, n, nc);
    GS_CHECK_LAUNCH();
    if (dist) {
        if ((rc = gs_shard_allreduce(p, w->red, 2, st))) return rc;
        pcg_scalar_kernel<<<1, 1, 0, st>>>(*w, 0);
    }

    const int* done = &w->state->done;
    // optional (gs_set_fuse_apq): q += C^-1 p and <p, q> ride on the last kernel of the analysis
    FinishFuse ff;
    ff.pE = w->p[0]; ff.pB = w->p[1]; ff.icE = tmp_l; ff.icB = tmp_l + 2 * (L + 1);
    ff.partials = w->fuse_partials; ff.counter = &w->state->counter; ff.out = w->fuse_out;
    const bool fused_apq = !dist && g_gs_fuse_apq;
    int launched = 0;
    bool finished = false;
    // one PCG iteration: 7 launches whose arguments do not change during the solve (alpha, beta and the done flag live on the device)
    auto iteration = [&](cudaStream_t s) -> int {
        int r;
        if ((r = apply_noise_op(p, w->p[0], w->p[1], bl, inv_noise, w->q[0], w->q[1], s, done, spin, fused_apq ? &ff : nullptr))) return r;
        if (fused_apq) pcg_apq_scalar_kernel<<<1, 1, 0, s>>>(*w);
        else pcg_apq_kernel<<<SV_GRID, SV_NT, 0, s>>>(*w, n, nc);
        if (dist) {
            if ((r = gs_shard_allreduce(p, w->red, 1, s))) return r;
            pcg_scalar_kernel<<<1, 1, 0, s>>>(*w, 1);
        }
        pcg_update_kernel<<<SV_GRID, SV_NT, 0, s>>>(*w, x_E, x_B, n, nc);
        if (dist) {
            if ((r = gs_shard_allreduce(p, w->red, 2, s))) return r;
            pcg_scalar_kernel<<<1, 1, 0, s>>>(*w, 2);
        }
        pcg_dir_kernel<<<SV_GRID, SV_NT, 0, s>>>(*w, n, nc);
        GS_CHECK_LAUNCH();
        g_gs_launches += 3;
        return GS_OK;
    };
    // Unsharded plans: the `check_every` iterations between two polls of the convergence flag are captured ONCE per solve in a
    // CUDA graph on a stream of the plan (capture is not allowed on the legacy default stream) and replayed: one graph launch
    // instead of 7 check_every kernel launches per poll (the host side of a solve was ~2300 launches; SCALE_r01 lost 1.8 % at 8
    // GPUs to 8 processes doing that on shared cores).  gs_set_pcg_graph(0) restores plain launches.
    cudaGraphExec_t gexec = nullptr;
    long long launches_per_graph = 0;
    static const bool env_graph_off = [] { const char* e = getenv("GS_PCG_GRAPH"); return e && e[0] == '0'; }();   // GS_PCG_GRAPH=0: plain launches
    if (g_gs_pcg_graph && !env_graph_off && !dist && itermax >= check_every) {
        if (!p->work_stream) {
            cudaStream_t ws = nullptr;
            GS_CHECK_CUDA(cudaStreamCreateWithFlags(&ws, cudaStreamNonBlocking));
            p->work_stream = ws;
            cudaEvent_t ev = nullptr;
            GS_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            p->work_event = ev;
        }
        cudaStream_t ws = (cudaStream_t)p->work_stream;
        GS_CHECK_CUDA(cudaEventRecord((cudaEvent_t)p->work_event, st));
        GS_CHECK_CUDA(cudaStreamWaitEvent(ws, (cudaEvent_t)p->work_event, 0));
        st = ws;   // the rest of the solve runs here; every poll (and the return) synchronises it with the host
        cudaGraph_t graph = nullptr;
        const long long l0 = g_gs_launches;
        GS_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        rc = GS_OK;
        for (int k = 0; k < check_every && rc == GS_OK; ++k) rc = iteration(st);
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        launches_per_graph = g_gs_launches - l0;
        g_gs_launches = l0;   // nothing ran yet: counted per replay below
        if (rc != GS_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) { gs_set_error("PCG graph capture: %s", cudaGetErrorString(ce)); return GS_E_CUDA; }
        ce = cudaGraphInstantiate(&gexec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) { gs_set_error("PCG graph instantiate: %s", cudaGetErrorString(ce)); return GS_E_CUDA; }
    }
    if (gexec) {
        // Two batches in flight: the host enqueues batch k + 1 (graph replay + copy of the solver state + event) BEFORE it waits for the
        // state that follows batch k, so the wake-up after a poll and the next graph launch are hidden behind 8 iterations of GPU work
        // (43 polls per solve at NSIDE 512; with 8 ranks on shared host cores this was the remaining loss of the chains mode).  Once
        // the done flag is set every kernel of a later batch returns at once: at most one batch of no-ops per solve, same iterations.
        PcgState* hs = w->host_state;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        cudaError_t ce = cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
        int cur = 0, inflight = 0;
        while (ce == cudaSuccess && !finished) {
            while (ce == cudaSuccess && inflight < 2 && launched + check_every <= itermax) {
                const int sl = (cur + inflight) & 1;
                ce = cudaGraphLaunch(gexec, st);
                if (ce == cudaSuccess) ce = cudaMemcpyAsync(hs + sl, w->state, sizeof(PcgState), cudaMemcpyDeviceToHost, st);
                if (ce == cudaSuccess) ce = cudaEventRecord(ev[sl], st);
                launched += check_every;
                g_gs_launches += launches_per_graph;
                ++inflight;
            }
            if (ce != cudaSuccess || inflight == 0) break;   // fewer than check_every iterations left: the loop below runs them
            ce = cudaEventSynchronize(ev[cur]);
            if (ce != cudaSuccess) break;
            --inflight;
            if (hs[cur].done || launched >= itermax) {
                if (inflight) ce = cudaStreamSynchronize(st);   // the batch behind it: no-ops after `done`, or the last iterations
                if (inflight && ce == cudaSuccess) cur ^= 1;
                hs[0] = hs[cur];
                finished = hs[0].done || launched >= itermax;
                inflight = 0;
                if (!finished) cur = 0;
                continue;
            }
            cur ^= 1;
        }
        if (ev[0]) cudaEventDestroy(ev[0]);
        if (ev[1]) cudaEventDestroy(ev[1]);
        if (ce != cudaSuccess) { cudaGraphExecDestroy(gexec); gs_set_error("PCG graph loop: %s", cudaGetErrorString(ce)); return GS_E_CUDA; }
    }
    while (!finished) {
        for (int k = 0; k < check_every && launched < itermax; ++k, ++launched)
            if ((rc = iteration(st))) { if (gexec) cudaGraphExecDestroy(gexec); return rc; }
        cudaError_t ce = cudaMemcpyAsync(w->host_state, w->state, sizeof(PcgState), cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) { if (gexec) cudaGraphExecDestroy(gexec); gs_set_error("PCG poll: %s", cudaGetErrorString(ce)); return GS_E_CUDA; }
        finished = w->host_state->done || launched >= itermax;
    }
    if (gexec) cudaGraphExecDestroy(gexec);
    const PcgState& s = *w->host_state;
    if (n_iter_out) *n_iter_out = s.iter;
    if (resid_out) *resid_out = s.d0 > 0.0 ? sqrt(s.rr / s.d0) : 0.0;
    if (s.d0 > 0.0 && s.rr > s.eps2 * s.d0) {
        gs_set_error("PCG stopped at iter_max = %d with |r|/|r0| = %.3e > eps = %.1e", itermax, sqrt(s.rr / s.d0), eps);
        return GS_E_NOTCONVERGED;
    }
    return GS_OK;
}

extern "C" int gs_cr_pcg_pol(gs_plan* p, const double* dl_EE, const double* dl_BB, const double* bl,
                             const double* inv_noise, double ninv_sum_over_4pi, const double* rhs_E,
                             const double* rhs_B, double* x_E, double* x_B, int warm_start, double eps, int itermax,
                             int check_every, int* n_iter_out, double* resid_out, void* stream)
{
    if (!p) { gs_set_error("null plan"); return GS_E_BADARG; }
    GS_REQUIRE(dl_EE && dl_BB && bl && inv_noise && rhs_E && rhs_B && x_E && x_B, "null pointer argument");
    return cr_pcg_impl(p, 2, dl_EE, dl_BB, bl, inv_noise, ninv_sum_over_4pi, rhs_E, rhs_B, x_E, x_B, warm_start, eps, itermax,
                       check_every, n_iter_out, resid_out, stream);
}

__global__ void pcg_alldone_kernel(const PcgState* a, const PcgState* b, int* out) { *out = (a->done && b->done) ? 1 : 0; }

// Two independent chains (same data, beam and N^-1; their own D_l, right-hand sides and solutions, chain c at + c stride
// doubles, dl at + c (L+1)) solved side by side: while both are iterating, every mat-vec is ONE chain-batched launch per stage
// (leg_synth / ring_apply / leg_anal with two right-hand sides sharing the Legendre recurrence); the vector kernels, alpha,
// beta and the stopping rule stay per chain, so each chain performs exactly the iterations of its own gs_cr_pcg_pol call.
// When one chain has converged (seen at the next poll) the other continues on the single-chain kernels.
extern "C" int gs_cr_pcg_pol_batch(gs_plan* p, int n_chain, const double* dl_EE, const double* dl_BB, const double* bl,
                                   const double* inv_noise, double ninv_sum_over_4pi, const double* rhs_E,
                                   const double* rhs_B, double* x_E, double* x_B, int64_t stride, double eps, int itermax,
                                   int check_every, int* n_iter_out, double* resid_out, void* stream)
{
    if (!p) { gs_set_error("null plan"); return GS_E_BADARG; }
    GS_REQUIRE(dl_EE && dl_BB && bl && inv_noise && rhs_E && rhs_B && x_E && x_B, "null pointer argument");
    GS_REQUIRE(n_chain == 1 || n_chain == 2, "n_chain must be 1 or 2");
    const int L = p->d.lmax;
    if (n_chain == 1)
        return cr_pcg_impl(p, 2, dl_EE, dl_BB, bl, inv_noise, ninv_sum_over_4pi, rhs_E, rhs_B, x_E, x_B, 0, eps, itermax, check_every,
                           n_iter_out, resid_out, stream);
    GS_REQUIRE(p->world == 1, "chain batches need an unsharded plan");
    GS_REQUIRE(eps > 0.0 && itermax >= 0, "eps must be > 0 and itermax >= 0");
    GS_CHECK_CUDA(cudaSetDevice(p->device));
    int rc;
    if ((rc = gs_plan_reserve_chains(p, 2))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    gs_pcg_ws* W = get_ws_batch(p);
    if (!W) return GS_E_NOMEM;
    const int64_t n = p->nreal_loc;
    if (check_every < 1) check_every = 8;
    const int lb = (L + 256) / 256;
    ActiveRings act(p);
    if ((rc = act.begin(inv_noise, st))) return rc;
    for (int k = 0; k < 2; ++k) {
        gs_pcg_ws* w = &W[k];
        double* tmp_l = w->partials + SV_GRID * 2;
        precond_kernel<<<lb, 256, 0, st>>>(dl_EE + k * (L + 1), bl, ninv_sum_over_4pi, L, tmp_l, tmp_l + (L + 1));
        precond_kernel<<<lb, 256, 0, st>>>(dl_BB + k * (L + 1), bl, ninv_sum_over_4pi, L, tmp_l + 2 * (L + 1), tmp_l + 3 * (L + 1));
        GS_CHECK_LAUNCH();
        PcgState h;
        memset(&h, 0, sizeof(h));
        h.eps2 = eps * eps;
        h.itermax = itermax;
        *w->host_state = h;
        GS_CHECK_CUDA(cudaMemcpyAsync(w->state, w->host_state, sizeof(PcgState), cudaMemcpyHostToDevice, st));
        zero2_kernel<<<SV_GRID, SV_NT, 0, st>>>(x_E + k * stride, x_B + k * stride, n);
        pcg_init_kernel<<<SV_GRID, SV_NT, 0, st>>>(*w, rhs_E + k * stride, rhs_B + k * stride, 0, n, 2);
        GS_CHECK_LAUNCH();
    }
    pcg_alldone_kernel<<<1, 1, 0, st>>>(W[0].state, W[1].state, p->pcg_alldone);
    const int64_t bstride = 2 * n;   // W[1].p[c] = W[0].p[c] + 2 n (get_ws_batch)
    bool done[2] = {false, false};
    int launched = 0;
    while (!(done[0] && done[1]) && launched < itermax) {
        for (int it = 0; it < check_every && launched < itermax; ++it, ++launched) {
            if (!done[0] && !done[1]) {
                if ((rc = apply_noise_op(p, W[0].p[0], W[0].p[1], bl, inv_noise, W[0].q[0], W[0].q[1], st, p->pcg_alldone, 2, nullptr, 2, bstride))) return rc;
            } else {
                gs_pcg_ws* w = &W[done[0] ? 1 : 0];
                if ((rc = apply_noise_op(p, w->p[0], w->p[1], bl, inv_noise, w->q[0], w->q[1], st, &w->state->done, 2))) return rc;
            }
            for (int k = 0; k < 2; ++k) {
                if (done[k]) continue;
                pcg_apq_kernel<<<SV_GRID, SV_NT, 0, st>>>(W[k], n, 2);
                pcg_update_kernel<<<SV_GRID, SV_NT, 0, st>>>(W[k], x_E + k * stride, x_B + k * stride, n, 2);
                pcg_dir_kernel<<<SV_GRID, SV_NT, 0, st>>>(W[k], n, 2);
                g_gs_launches += 3;
            }
            if (!done[0] && !done[1]) { pcg_alldone_kernel<<<1, 1, 0, st>>>(W[0].state, W[1].state, p->pcg_alldone); g_gs_launches += 1; }
            GS_CHECK_LAUNCH();
        }
        for (int k = 0; k < 2; ++k) GS_CHECK_CUDA(cudaMemcpyAsync(W[k].host_state, W[k].state, sizeof(PcgState), cudaMemcpyDeviceToHost, st));
        GS_CHECK_CUDA(cudaStreamSynchronize(st));
        for (int k = 0; k < 2; ++k) done[k] = W[k].host_state->done != 0;
    }
    int ret = GS_OK;
    for (int k = 0; k < 2; ++k) {
        const PcgState& sdone = *W[k].host_state;
        if (n_iter_out) n_iter_out[k] = sdone.iter;
        if (resid_out) resid_out[k] = sdone.d0 > 0.0 ? sqrt(sdone.rr / sdone.d0) : 0.0;
        if (sdone.d0 > 0.0 && sdone.rr > sdone.eps2 * sdone.d0) {
            gs_set_error("PCG (chain %d of the batch) stopped at iter_max = %d with |r|/|r0| = %.3e > eps = %.1e", k, itermax, sqrt(sdone.rr / sdone.d0), eps);
            ret = GS_E_NOTCONVERGED;
        }
    }
    return ret;
}

// Temperature twin (qcinv opfilt_tt chain of ConstrainedRealization.py:40-41, run at CenteredGibbs.py:141-165 and
// NonCenteredGibbs.py:57-72): Q = C^-1 + B A^T N^-1 A B on one spin-0 field.
extern "C" int gs_cr_pcg_tt(gs_plan* p, const double* dl_TT, const double* bl, const double* inv_noise,
                            double ninv_sum_over_4pi, const double* rhs, double* x, int warm_start, double eps, int itermax,
                            int check_every, int* n_iter_out, double* resid_out, void* stream)
{
    if (!p) { gs_set_error("null plan"); return GS_E_BADARG; }
    GS_REQUIRE(dl_TT && bl && inv_noise && rhs && x, "null pointer argument");
    return cr_pcg_impl(p, 0, dl_TT, nullptr, bl, inv_noise, ninv_sum_over_4pi, rhs, nullptr, x, nullptr, warm_start, eps, itermax,
                       check_every, n_iter_out, resid_out, stream);
}

extern "C" int gs_cr_apply_q_tt(gs_plan* p, const double* dl_TT, const double* bl, const double* inv_noise, const double* x,
                                double* y, void* stream)
{
    if (!p) { gs_set_error("null plan"); return GS_E_BADARG; }
    GS_REQUIRE(dl_TT && bl && inv_noise && x && y, "null pointer argument");
    GS_CHECK_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    gs_pcg_ws* w = get_ws(p);
    if (!w) return GS_E_NOMEM;
    int rc;
    ActiveRings act(p);
    if ((rc = act.begin(inv_noise, st))) return rc;
    if ((rc = gs_plan_expand_per_l(p, dl_TT, 2, w->invc[0], st))) return rc;
    if ((rc = apply_noise_op(p, x, nullptr, bl, inv_noise, w->q[0], nullptr, st, nullptr, 0))) return rc;
    axy_kernel<<<SV_GRID, SV_NT, 0, st>>>(w->q[0], w->invc[0], x, y, p->nreal_loc);
    GS_CHECK_LAUNCH();
    return GS_OK;
}

// Right-hand side of the temperature system (CenteredGibbs.py:145-147, NonCenteredGibbs.py:61-63; RNG order xi_alm, xi_pix):
//   b = bdata + C^-1/2 xi_alm + B (Npix/4pi) map2alm_{iter}(N^-1/2 xi_pix),  bdata = B A^T N^-1 d (what qcinv adds in
//   chain.sample) passed precomputed or computed from d.
extern "C" int gs_cr_rhs_tt(gs_plan* p, const double* dl_TT, const double* bl, const double* inv_noise,
                            const double* sqrt_inv_noise, const double* bdata, const double* d, const double* xi_alm,
                            const double* xi_pix, int fluct_iter, double* rhs, void* stream)
{
    if (!p) { gs_set_error("null plan"); return GS_E_BADARG; }
    GS_REQUIRE(dl_TT && bl && inv_noise && sqrt_inv_noise && xi_alm && xi_pix && rhs, "null pointer argument");
    GS_REQUIRE(bdata || d, "need either the precomputed data term or the data map");
    GS_REQUIRE(fluct_iter >= 0, "fluct_iter must be >= 0");
    GS_CHECK_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    gs_pcg_ws* w = get_ws(p);
    if (!w) return GS_E_NOMEM;
    int rc;
    if ((rc = gs_map2alm_spin0(p, xi_pix, sqrt_inv_noise, fluct_iter, 0, bl, rhs, GS_ALM_REAL, stream))) return rc;
    const double resc = (double)p->d.npix / (4.0 * 3.14159265358979323846);
    if ((rc = gs_plan_expand_per_l(p, dl_TT, 4, w->r[0], st))) return rc;
    const double* bd = bdata;
    if (!bd) {
        if ((rc = gs_map2alm_spin0(p, d, inv_noise, 0, 1, bl, w->p[0], GS_ALM_REAL, stream))) return rc;
        bd = w->p[0];
    }
    rhs_combine_kernel<<<SV_GRID, SV_NT, 0, st>>>(rhs, w->r[0], xi_alm, bd, resc, p->nreal_loc);
    GS_CHECK_LAUNCH();
    return GS_OK;
}

// y = Q x = C^-1 x + B A^T N^-1 A B x   (qcinv opfilt_pp.fwd_op; CenteredGibbs.py:629, 651-656)
extern "C" int gs_cr_apply_q_pol(gs_plan* p, const double* dl_EE, const double* dl_BB, const double* bl,
                                 const double* inv_noise, const double* x_E, const double* x_B, double* y_E,
                                 double* y_B, void* stream)
{
    if (!p) { gs_set_error("null plan"); return GS_E_BADARG; }
    GS_REQUIRE(dl_EE && dl_BB && bl && inv_noise && x_E && x_B && y_E && y_B, "null pointer argument");
    GS_CHECK_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    gs_pcg_ws* w = get_ws(p);
    if (!w) return GS_E_NOMEM;
    const int64_t n = p->nreal_loc;
    int rc;
    ActiveRings act(p);
    if ((rc = act.begin(inv_noise, st))) return rc;
    if ((rc = gs_plan_expand_per_l(p, dl_EE, 2, w->invc[0], st))) return rc;
    if ((rc = gs_plan_expand_per_l(p, dl_BB, 2, w->invc[1], st))) return rc;
    if ((rc = apply_noise_op(p, x_E, x_B, bl, inv_noise, w->q[0], w->q[1], st, nullptr))) return rc;
    axy_kernel<<<SV_GRID, SV_NT, 0, st>>>(w->q[0], w->invc[0], x_E, y_E, n);
    axy_kernel<<<SV_GRID, SV_NT, 0, st>>>(w->q[1], w->invc[1], x_B, y_B, n);
    GS_CHECK_LAUNCH();
    return GS_OK;
}

// Right-hand side of the masked-sky system (CenteredGibbs.py:469-483, RNG order xi_Q, xi_U, xi_E, xi_B):
//   b = B A^T N^-1 d  +  B (Npix/4pi) map2alm_{iter}(N^-1/2 xi_pix)  +  C^-1/2 xi_alm
// The first term is what qcinv's calc_prep adds inside chain.sample; it equals the reference's
// second_part_grad (CenteredGibbs.py:298-308) and may be passed precomputed (bdata_*).
extern "C" int gs_cr_rhs_pol(gs_plan* p, const double* dl_EE, const double* dl_BB, const double* bl,
                             const double* inv_noise, const double* sqrt_inv_noise, const double* bdata_E,
                             const double* bdata_B, const double* d_Q, const double* d_U, const double* xi_Q,
                             const double* xi_U, const double* xi_E, const double* xi_B, int fluct_iter,
                             double* rhs_E, double* rhs_B, void* stream)
{
    if (!p) { gs_set_error("null plan"); return GS_E_BADARG; }
    GS_REQUIRE(dl_EE && dl_BB && bl && inv_noise && sqrt_inv_noise && xi_Q && xi_U && xi_E && xi_B && rhs_E && rhs_B,
               "null pointer argument");
    GS_REQUIRE((bdata_E && bdata_B) || (d_Q && d_U), "need either the precomputed data term or the data maps");
    GS_REQUIRE(fluct_iter >= 0, "fluct_iter must be >= 0");
    GS_CHECK_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    gs_pcg_ws* w = get_ws(p);
    if (!w) return GS_E_NOMEM;
    const int64_t n = p->nreal_loc;
    int rc;
    // fluctuation term 1: utils.adjoint_synthesis_hp (utils.py:79-111) = bl * (Npix/4pi) * map2alm(iter=3)
    if ((rc = gs_map2alm_spin2(p, xi_Q, xi_U, sqrt_inv_noise, fluct_iter, 0, bl, rhs_E, rhs_B, GS_ALM_REAL, stream))) return rc;
    const double resc = (double)p->d.npix / (4.0 * 3.14159265358979323846);
    // fluctuation term 2 and data term; use r[] / p[] of the PCG workspace as scratch
    if ((rc = gs_plan_expand_per_l(p, dl_EE, 4, w->r[0], st))) return rc;
    if ((rc = gs_plan_expand_per_l(p, dl_BB, 4, w->r[1], st))) return rc;
    const double* bdE = bdata_E;
    const double* bdB = bdata_B;
    if (!bdE) {
        if ((rc = gs_map2alm_spin2(p, d_Q, d_U, inv_noise, 0, 1, bl, w->p[0], w->p[1], GS_ALM_REAL, stream))) return rc;
        bdE = w->p[0];
        bdB = w->p[1];
    }
    // rhs = resc * rhs + sqrt(C^-1) xi + bdata
    rhs_combine_kernel<<<SV_GRID, SV_NT, 0, st>>>(rhs_E, w->r[0], xi_E, bdE, resc, n);
    rhs_combine_kernel<<<SV_GRID, SV_NT, 0, st>>>(rhs_B, w->r[1], xi_B, bdB, resc, n);
    GS_CHECK_LAUNCH();
    return GS_OK;
}


// ------------------------------------------------------------------ measurement helpers (bench.py)
long long g_gs_launches = 0;
extern "C" long long gs_launch_count(void) { return g_gs_launches; }

// Times the four kernels of one PCG mat-vec (Legendre synthesis, ring synthesis, ring analysis with
// fused N^-1, Legendre analysis + finish) with CUDA events on the launching stream; ms_out[0..3]
// receive the average duration of each stage over nrep repetitions.
extern "C" int gs_profile_matvec(gs_plan* p, const double* x_E, const double* x_B, const double* bl,
                                 const double* inv_noise, double* y_E, double* y_B, int nrep, float* ms_out,
                                 void* stream)
{
    if (!p) { gs_set_error("null plan"); return GS_E_BADARG; }
    GS_REQUIRE(x_E && x_B && bl && inv_noise && y_E && y_B && ms_out && nrep >= 1, "bad arguments");
    GS_CHECK_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t ev[5];
    for (int i = 0; i < 5; ++i) GS_CHECK_CUDA(cudaEventCreate(&ev[i]));
    double acc[4] = {0, 0, 0, 0};
    int rc = GS_OK;
    ActiveRings act(p);   // as in the PCG: rings without weight are skipped unless gs_set_ring_skip(0)
    if ((rc = act.begin(inv_noise, st))) { for (int i = 0; i < 5; ++i) cudaEventDestroy(ev[i]); return rc; }
    for (int r = 0; r < nrep && rc == GS_OK; ++r) {
        cudaEventRecord(ev[0], st);
        rc = gs_leg_synth(p, 2, x_E, x_B, GS_ALM_REAL, bl, st);
        cudaEventRecord(ev[1], st);
        // fused ring stage (the PCG's own path): reported as stage 1, stage 2 = 0
        if (!rc) rc = g_gs_ring_fused ? gs_ring_apply(p, 2, inv_noise, st) : gs_ring_synth(p, 2, p->mapQ_tmp, p->mapU_tmp, st);
        cudaEventRecord(ev[2], st);
        if (!rc && !g_gs_ring_fused) rc = gs_ring_anal(p, 2, p->mapQ_tmp, p->mapU_tmp, inv_noise, st);
        cudaEventRecord(ev[3], st);
        if (!rc) rc = gs_leg_anal(p, 2, y_E, y_B, GS_ALM_REAL, bl, 1.0, 0, st);
        cudaEventRecord(ev[4], st);
        if (cudaEventSynchronize(ev[4]) != cudaSuccess) { gs_set_error("gs_profile_matvec: %s", cudaGetErrorString(cudaGetLastError())); rc = GS_E_CUDA; break; }
        for (int i = 0; i < 4; ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); acc[i] += ms; }
    }
    for (int i = 0; i < 4; ++i) ms_out[i] = (float)(acc[i] / nrep);
    if (g_gs_ring_fused) ms_out[2] = 0.0f;
    for (int i = 0; i < 5; ++i) cudaEventDestroy(ev[i]);
    return rc;
}

// gs_profile_matvec for a chain batch: n_chain (1 or 2) right-hand sides at x / y + c stride (doubles) per launch; the fused ring
// stage only.  ms_out[0] Legendre synthesis, [1] fused ring stage, [2] Legendre analysis + finish, each for the WHOLE batch.
extern "C" int gs_profile_matvec_batch(gs_plan* p, int n_chain, const double* x_E, const double* x_B, int64_t stride,
                                       const double* bl, const double* inv_noise, double* y_E, double* y_B, int nrep,
                                       float* ms_out, void* stream)
{
    if (!p) { gs_set_error("null plan"); return GS_E_BADARG; }
    GS_REQUIRE(x_E && x_B && bl && inv_noise && y_E && y_B && ms_out && nrep >= 1 && (n_chain == 1 || n_chain == 2), "bad arguments");
    GS_CHECK_CUDA(cudaSetDevice(p->device));
    int rc;
    if (n_chain > 1 && (rc = gs_plan_reserve_chains(p, n_chain))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t ev[4];
    for (int i = 0; i < 4; ++i) GS_CHECK_CUDA(cudaEventCreate(&ev[i]));
    double acc[3] = {0, 0, 0};
    rc = GS_OK;
    {
        ActiveRings act(p);
        rc = act.begin(inv_noise, st);
        for (int r = 0; r < nrep && rc == GS_OK; ++r) {
            cudaEventRecord(ev[0], st);
            rc = gs_leg_synth(p, 2, x_E, x_B, GS_ALM_REAL, bl, st, nullptr, nullptr, n_chain, stride);
            cudaEventRecord(ev[1], st);
            if (!rc) rc = gs_ring_apply(p, 2, inv_noise, st, nullptr, n_chain);
            cudaEventRecord(ev[2], st);
            if (!rc) rc = gs_leg_anal(p, 2, y_E, y_B, GS_ALM_REAL, bl, 1.0, 0, st, nullptr, nullptr, n_chain, stride);
            cudaEventRecord(ev[3], st);
            if (cudaEventSynchronize(ev[3]) != cudaSuccess) { gs_set_error("gs_profile_matvec_batch: %s", cudaGetErrorString(cudaGetLastError())); rc = GS_E_CUDA; break; }
            for (int i = 0; i < 3; ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); acc[i] += ms; }
        }
    }
    for (int i = 0; i < 3; ++i) ms_out[i] = (float)(acc[i] / nrep);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(ev[i]);
    return rc;
}

// Times the three vector kernels of one PCG iteration (q += C^-1 p with <p,q>; x, r update with <r,r>, <r,Mr>; p = M r + beta p)
// on the plan's own workspace, zero-filled, with CUDA events on `stream`: ms_out[0..2] = average of nrep launches each.  The
// reductions run as in the solver but their results go to a scratch slot, so no solver state changes.
extern "C" int gs_profile_pcg_vectors(gs_plan* p, int spin, int nrep, float* ms_out, void* stream)
{
    if (!p) { gs_set_error("null plan"); return GS_E_BADARG; }
    GS_REQUIRE(ms_out && nrep >= 1 && (spin == 0 || spin == 2), "bad arguments");
    GS_CHECK_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    gs_pcg_ws* w = get_ws(p);
    if (!w) return GS_E_NOMEM;
    const int nc = spin ? 2 : 1;
    const int64_t n = p->nreal_loc;
    for (int c = 0; c < 2; ++c) {
        double* arrs[5] = {w->r[c], w->p[c], w->q[c], w->invc[c], w->pre[c]};
        for (double* a : arrs) GS_CHECK_CUDA(cudaMemsetAsync(a, 0, (size_t)n * sizeof(double), st));
    }
    GS_CHECK_CUDA(cudaMemsetAsync(p->almE_tmp, 0, (size_t)n * sizeof(double), st));
    GS_CHECK_CUDA(cudaMemsetAsync(p->almB_tmp, 0, (size_t)n * sizeof(double), st));
    GS_CHECK_CUDA(cudaMemsetAsync(w->state, 0, sizeof(PcgState), st));
    GS_CHECK_CUDA(cudaMemsetAsync(w->partials + SV_GRID * 2, 0, (size_t)4 * (p->d.lmax + 1) * sizeof(double), st));   // per-l tables
    gs_pcg_ws v = *w;
    v.red = w->fuse_out;   // reductions land here: the solver state (done flag, alpha, beta = 0) stays untouched
    cudaEvent_t e0, e1;
    GS_CHECK_CUDA(cudaEventCreate(&e0));
    GS_CHECK_CUDA(cudaEventCreate(&e1));
    // between launches the analysis partials (> L2 at the bench size) are overwritten, as the Legendre / ring kernels do
    // between the vector kernels of a real iteration: the arrays come from HBM, not from the 126 MB L2
    const size_t flush_bytes = (size_t)p->anal_chunks * (size_t)(p->world > 1 ? p->d.sh.nalm_loc : p->d.nalm) * 4 * sizeof(double);
    for (int k = 0; k < 3; ++k) {
        double acc = 0.0;
        for (int rep = -2; rep < nrep; ++rep) {   // two warm-up launches
            if (cudaMemsetAsync(p->partial, 0, flush_bytes, st) != cudaSuccess) {
                gs_set_error("gs_profile_pcg_vectors: %s", cudaGetErrorString(cudaGetLastError()));
                cudaEventDestroy(e0); cudaEventDestroy(e1);
                return GS_E_CUDA;
            }
            cudaEventRecord(e0, st);
            if (k == 0) pcg_apq_kernel<<<SV_GRID, SV_NT, 0, st>>>(v, n, nc);
            else if (k == 1) pcg_update_kernel<<<SV_GRID, SV_NT, 0, st>>>(v, p->almE_tmp, p->almB_tmp, n, nc);
            else pcg_dir_kernel<<<SV_GRID, SV_NT, 0, st>>>(v, n, nc);
            cudaEventRecord(e1, st);
            if (cudaEventSynchronize(e1) != cudaSuccess) {
                gs_set_error("gs_profile_pcg_vectors: %s", cudaGetErrorString(cudaGetLastError()));
                cudaEventDestroy(e0); cudaEventDestroy(e1);
                return GS_E_CUDA;
            }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 0) acc += ms;
        }
        ms_out[k] = (float)(acc / nrep);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    GS_CHECK_LAUNCH();
    return GS_OK;
}

// FP64 FMA throughput of the device (MEASURED_PEAKS.json holds no FP64 figure).  16 independent sums per thread in the
// operand pattern of the Legendre kernels, sum += x_k * y_j with register operands that change during the loop (three
// distinct register pairs per DFMA; scripts/ubench/dfma_operands.cu compares patterns: this one is the highest the pipe
// delivers, 36.9 TFLOP/s on B200 = 99 % of 148 x 64 x 2 x 1.965 GHz); 148 x 8 CTAs of 128 threads; TFLOP/s, best of 5.
__global__ void __launch_bounds__(128) dfma_peak_kernel(double* out, int iters, double a, double b)
{
    double v[16], w[4], u[4];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = threadIdx.x * 1e-3 + k;
#pragma unroll
    for (int k = 0; k < 4; ++k) { w[k] = a + 1e-9 * k; u[k] = b * (k + 1); }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = fma(w[k & 3], u[k >> 2], v[k]);
        w[i & 3] = fma(w[i & 3], a, b * 1e-30);   // keeps the multiplicands live (1 of 17 DFMAs, counted)
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += v[k];
    if (s == 123.456) out[0] = s;
}

extern "C" int gs_measure_fp64_peak(double* tflops_out, void* stream)
{
    GS_REQUIRE(tflops_out, "null output");
    cudaStream_t st = (cudaStream_t)stream;
    double* d = nullptr;
    GS_CHECK_CUDA(cudaMalloc(&d, 8));
    cudaEvent_t e0, e1;
    GS_CHECK_CUDA(cudaEventCreate(&e0));
    GS_CHECK_CUDA(cudaEventCreate(&e1));
    const int iters = 1 << 14, grid = 148 * 8;
    double best = 0.0;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0, st);
        dfma_peak_kernel<<<grid, 128, 0, st>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1, st);
        GS_CHECK_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 17.0 * iters * (double)grid * 128.0 / (ms * 1e-3) * 1e-12;
        if (r > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops_out = best;
    return GS_OK;
}

extern "C" int gs_constant_rings(gs_plan* p, int* count_out)
{
    if (!p) { gs_set_error("null plan"); return GS_E_BADARG; }
    GS_REQUIRE(count_out, "null output");
    *count_out = 0;
    if (!p->ring_wconst) return GS_OK;   // sharded plans do not flag constant rings
    GS_CHECK_CUDA(cudaSetDevice(p->device));
    std::vector<double> w(p->d.nring);
    GS_CHECK_CUDA(cudaMemcpy(w.data(), p->ring_wconst, w.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (double v : w) *count_out += (v == v);
    return GS_OK;
}

extern "C" int gs_active_ring_pairs(gs_plan* p, int* active_out, int* total_out)
{
    if (!p) { gs_set_error("null plan"); return GS_E_BADARG; }
    GS_REQUIRE(active_out && total_out, "null output");
    GS_CHECK_CUDA(cudaSetDevice(p->device));
    GS_CHECK_CUDA(cudaMemcpy(active_out, p->act_count, sizeof(int), cudaMemcpyDeviceToHost));
    *total_out = p->d.npair;
    return GS_OK;
}
