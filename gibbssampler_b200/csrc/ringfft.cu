// Ring (azimuthal) stage of the spherical-harmonic transforms: per HEALPix ring, the sum over m
// of F_m(theta) exp(i m phi) on the ring's n equidistant pixels (synthesis) and its transpose
// (analysis).  Replaces the per-ring FFTs of libsharp behind hp.alm2map / hp.map2alm.
//
// One CTA per job; a job is ONE complex DFT of length n that carries TWO real ring sequences
// (spin 2: Q and U of a ring; spin 0: a north ring and its southern mirror), so no real-FFT
// post-processing pass is needed.  m > n/2 is alias-folded in shared memory, the phase
// exp(i m phi_0) of the ring's first pixel is applied on load.  The DFT itself runs in shared
// memory in FP64: power-of-two rings (the 2 nside + 1 equatorial-belt rings when nside is a power
// of two) use an in-place radix-4 decimation-in-time transform on bit-reversed input; every other
// length (polar-cap rings have 4 i pixels, i < nside) uses Bluestein's chirp-z algorithm on top of
// the same kernel: DIF forward -> pointwise product with a precomputed table stored in DIF
// (bit-reversed) order -> DIT inverse, so no permutation pass is ever executed.
#include <math.h>

#include <algorithm>
#include <map>

#include "gs_internal.h"

#define RF_NT 256

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a conj(b)
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// ---- shared-memory FFT core -------------------------------------------------------------------
// Twiddles: quarter table twq[k] = exp(-2 pi i k / twn), k <= twn/4, staged in shared memory by the
// caller; W^(k + twn/4) = -i W^k gives the second quadrant (all the radix-4 butterflies need).
// Every pass loads up to RF_U radix-4 butterflies per thread before computing any of them, so the
// shared-memory round trips of the butterflies overlap instead of serialising (the compiler cannot
// hoist loads above stores to the same buffer by itself).
#define RF_U 2

__device__ __forceinline__ double2 tw_get(const double2* twq, int idx, int quarter)
{
    if (idx < quarter) return twq[idx];
    const double2 w = twq[idx - quarter];
    return make_double2(w.y, -w.x);  // -i w
}

// one radix-4 DIF pass (sub-transform size N) of an M-point transform; POST: multiply outputs by post[index]
template <bool POST>
__device__ __forceinline__ void dif4_pass(double2* buf, int M, int N, const double2* twq, int twn,
                                          const double2* __restrict__ post)
{
    const int q = N >> 2, lq = 31 - __clz(q), ts = twn / N, quarter = twn >> 2, nq = M >> 2;
    for (int t0 = 0; t0 < nq; t0 += RF_NT * RF_U) {
        double2 a[RF_U][4], w1[RF_U], w2[RF_U];
        int i0[RF_U];
#pragma unroll
        for (int u = 0; u < RF_U; ++u) {
            const int t = t0 + u * RF_NT + threadIdx.x;
            i0[u] = -1;
            if (t < nq) {
                const int g = t >> lq, j = t & (q - 1);
                i0[u] = g * N + j;
                a[u][0] = buf[i0[u]]; a[u][1] = buf[i0[u] + q]; a[u][2] = buf[i0[u] + 2 * q]; a[u][3] = buf[i0[u] + 3 * q];
                w1[u] = twq[j * ts];
                w2[u] = tw_get(twq, 2 * j * ts, quarter);
            }
        }
#pragma unroll
        for (int u = 0; u < RF_U; ++u) {
            if (i0[u] < 0) continue;
            const double2 b0 = cadd(a[u][0], a[u][2]), b2 = cmul(csub(a[u][0], a[u][2]), w1[u]), b1 = cadd(a[u][1], a[u][3]);
            const double2 d13 = csub(a[u][1], a[u][3]);
            const double2 b3 = cmul(make_double2(d13.y, -d13.x), w1[u]);  // (a1 - a3) (-i) W^j
            double2 c0 = cadd(b0, b1), c1 = cmul(csub(b0, b1), w2[u]), c2 = cadd(b2, b3), c3 = cmul(csub(b2, b3), w2[u]);
            if (POST) {
                c0 = cmul(c0, __ldg(&post[i0[u]])); c1 = cmul(c1, __ldg(&post[i0[u] + q]));
                c2 = cmul(c2, __ldg(&post[i0[u] + 2 * q])); c3 = cmul(c3, __ldg(&post[i0[u] + 3 * q]));
            }
            buf[i0[u]] = c0; buf[i0[u] + q] = c1; buf[i0[u] + 2 * q] = c2; buf[i0[u] + 3 * q] = c3;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void dit4_pass(double2* buf, int M, int N, const double2* twq, int twn)
{
    const int q = N >> 2, lq = 31 - __clz(q), ts = twn / N, quarter = twn >> 2, nq = M >> 2;
    for (int t0 = 0; t0 < nq; t0 += RF_NT * RF_U) {
        double2 c[RF_U][4], w1[RF_U], w2[RF_U];
        int i0[RF_U];
#pragma unroll
        for (int u = 0; u < RF_U; ++u) {
            const int t = t0 + u * RF_NT + threadIdx.x;
            i0[u] = -1;
            if (t < nq) {
                const int g = t >> lq, j = t & (q - 1);
                i0[u] = g * N + j;
                c[u][0] = buf[i0[u]]; c[u][1] = buf[i0[u] + q]; c[u][2] = buf[i0[u] + 2 * q]; c[u][3] = buf[i0[u] + 3 * q];
                w1[u] = twq[j * ts];
                w2[u] = tw_get(twq, 2 * j * ts, quarter);
            }
        }
#pragma unroll
        for (int u = 0; u < RF_U; ++u) {
            if (i0[u] < 0) continue;
            const double2 t1 = cmulc(c[u][1], w2[u]), t3 = cmulc(c[u][3], w2[u]);
            const double2 b0 = cadd(c[u][0], t1), b1 = csub(c[u][0], t1), b2 = cadd(c[u][2], t3), b3 = csub(c[u][2], t3);
            const double2 t2 = cmulc(b2, w1[u]), v3 = cmulc(b3, w1[u]);
            const double2 u3 = make_double2(-v3.y, v3.x);  // (+i) conj(W)^j b3
            buf[i0[u]] = cadd(b0, t2); buf[i0[u] + 2 * q] = csub(b0, t2);
            buf[i0[u] + q] = cadd(b1, u3); buf[i0[u] + 3 * q] = csub(b1, u3);
        }
    }
    __syncthreads();
}

template <bool POST>
__device__ __forceinline__ void r2_pass(double2* buf, int M, const double2* __restrict__ post)
{
    for (int t = threadIdx.x; t < (M >> 1); t += RF_NT) {
        const double2 a = buf[2 * t], b = buf[2 * t + 1];
        double2 s = cadd(a, b), d = csub(a, b);
        if (POST) { s = cmul(s, __ldg(&post[2 * t])); d = cmul(d, __ldg(&post[2 * t + 1])); }
        buf[2 * t] = s; buf[2 * t + 1] = d;
    }
    __syncthreads();
}

// In-place forward DIF FFT (kernel exp(-2 pi i jk/M)), natural order in, bit-reversed order out;
// POST: the bit-reversed-order output is multiplied by post[] inside the last pass.
template <bool POST>
__device__ void fft_dif(double2* buf, int M, const double2* twq, int twn, const double2* __restrict__ post)
{
    const int lg = 31 - __clz(M);
    int N = M;
    while (N >= 4) {
        const bool last = !(lg & 1) && N == 4;
        if (POST && last) dif4_pass<true>(buf, M, N, twq, twn, post);
        else dif4_pass<false>(buf, M, N, twq, twn, post);
        N >>= 2;
    }
    if (N == 2) r2_pass<POST>(buf, M, post);
}

// In-place inverse DIT FFT (kernel exp(+2 pi i jk/M), unnormalised), bit-reversed in, natural out.
__device__ void fft_dit_inv(double2* buf, int M, const double2* twq, int twn)
{
    const int lg = 31 - __clz(M);
    int N = 4;
    if (lg & 1) { r2_pass<false>(buf, M, nullptr); N = 8; }
    while (N <= M) { dit4_pass(buf, M, N, twq, twn); N <<= 2; }
}

__device__ __forceinline__ void load_twq(const PlanDev& P, double2* twq)
{
    const int nq = (P.tw_n >> 2) + 1;
    for (int k = threadIdx.x; k < nq; k += RF_NT) twq[k] = __ldg(&P.tw[k]);
}

__device__ __forceinline__ double2 chirp_val(int t, int n)
{  // exp(i pi t^2 / n)
    const int r = (int)(((long long)t * t) % (2 * n));
    double s, c;
    sincospi((double)r / (double)n, &s, &c);
    return make_double2(c, s);
}

// z_j = sum_{k<n} Z_k exp(+2 pi i jk / n), in place in buf[0..n).
// Power-of-two n (bsi < 0): the caller stored Z_k at the bit-reversed position of k.
// Otherwise (Bluestein): the caller stored Z_k * chirp[k] at k < n and zeros at n <= k < M; the result
// still has to be multiplied by chirp[j] by the caller (fused into its output pass).
// All threads of the CTA must call; begins and ends with a barrier.
__device__ void ring_idft(const PlanDev& P, double2* buf, const double2* twq, int n, int bsi)
{
    __syncthreads();
    if (bsi < 0) { fft_dit_inv(buf, n, twq, P.tw_n); return; }
    const BluesteinDesc d = P.bs[bsi];
    fft_dif<true>(buf, d.M, twq, P.tw_n, P.bs_tab + d.bhat_off);
    fft_dit_inv(buf, d.M, twq, P.tw_n);
}

__global__ void __launch_bounds__(RF_NT) bluestein_setup_kernel(PlanDev P, double2* tab, int nbs)
{
    extern __shared__ double2 smem_all[];
    double2* twq = smem_all;
    double2* smem = smem_all + (P.tw_n >> 2) + 1;
    load_twq(P, twq);
    const BluesteinDesc d = P.bs[blockIdx.x];
    double2* chirp = tab + d.chirp_off;
    double2* bhat = tab + d.bhat_off;
    for (int k = threadIdx.x; k < d.M; k += blockDim.x) smem[k] = make_double2(0.0, 0.0);
    __syncthreads();
    for (int t = threadIdx.x; t < d.n; t += blockDim.x) {
        const double2 c = chirp_val(t, d.n);
        chirp[t] = c;
        const double2 cc = make_double2(c.x, -c.y);
        smem[t] = cc;
        if (t > 0) smem[d.M - t] = cc;
    }
    __syncthreads();
    fft_dif<false>(smem, d.M, twq, P.tw_n, nullptr);
    const double inv = 1.0 / (double)d.M;
    for (int k = threadIdx.x; k < d.M; k += blockDim.x) bhat[k] = make_double2(smem[k].x * inv, smem[k].y * inv);
    (void)nbs;
}

__device__ __forceinline__ double2 ring_phase(const PlanDev& P, int ring, int m)
{  // exp(i m phi0), phi0 = pi q / den
    const int q = P.ring_phq[ring], den = P.ring_phden[ring];
    const int r = (int)(((long long)m * q) % (2 * den));
    double s, c;
    sincospi((double)r / (double)den, &s, &c);
    return make_double2(c, s);
}

__global__ void __launch_bounds__(RF_NT, 2)
ring_synth_kernel(PlanDev P, const RingJob* __restrict__ jobs, const double2* __restrict__ Fm, double* __restrict__ mapQ,
                  double* __restrict__ mapU, const int* __restrict__ skip)
{
    if (skip && *skip) return;
    extern __shared__ double2 smem[];
    const int L = P.lmax, nm = L + 1;
    const RingJob job = jobs[blockIdx.x];
    const int n = P.ring_nphi[job.ringA], bsi = P.ring_bs[job.ringA];
    double2* twq = smem;
    double2* stA = twq + (P.tw_n >> 2) + 1;
    double2* stB = stA + nm;
    double2* buf = stB + nm;
    load_twq(P, twq);
    const double2* FA = Fm + ((int64_t)job.compA * P.nring + job.ringA) * nm;
    const double2* FB = job.ringB >= 0 ? Fm + ((int64_t)job.compB * P.nring + job.ringB) * nm : nullptr;
    const bool same_phase = job.ringB == job.ringA ||
                            (job.ringB >= 0 && P.ring_phq[job.ringB] == P.ring_phq[job.ringA] && P.ring_phden[job.ringB] == P.ring_phden[job.ringA]);
    for (int m = threadIdx.x; m <= L; m += RF_NT) {
        const double w = m ? 1.0 : 0.5;  // (2 - delta_m0) / 2
        double2 ph = ring_phase(P, job.ringA, m);
        ph.x *= w; ph.y *= w;
        stA[m] = cmul(FA[m], ph);
        if (FB) {
            double2 phb = ph;
            if (!same_phase) { phb = ring_phase(P, job.ringB, m); phb.x *= w; phb.y *= w; }
            stB[m] = cmul(FB[m], phb);
        } else stB[m] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    const int lg = 31 - __clz(n);
    const double2* chirp = bsi >= 0 ? P.bs_tab + P.bs[bsi].chirp_off : nullptr;
    const int M = bsi >= 0 ? P.bs[bsi].M : n;
    for (int k = threadIdx.x; k < M; k += RF_NT) {
        if (k >= n) { buf[k] = make_double2(0.0, 0.0); continue; }
        const int kk = (n - k) % n;
        double2 ga = make_double2(0.0, 0.0), gb = ga, ha = ga, hb = ga;
        for (int m = k; m <= L; m += n) { ga = cadd(ga, stA[m]); gb = cadd(gb, stB[m]); }
        for (int m = kk; m <= L; m += n) { ha = cadd(ha, stA[m]); hb = cadd(hb, stB[m]); }
        // X = G[k] + conj G[n-k] (the 1/2 is in w);  Z = Xa + i Xb
        const double2 xa = make_double2(ga.x + ha.x, ga.y - ha.y), xb = make_double2(gb.x + hb.x, gb.y - hb.y);
        const double2 z = make_double2(xa.x - xb.y, xa.y + xb.x);
        if (bsi < 0) buf[(int)(__brev((unsigned)k) >> (32 - lg))] = z;
        else buf[k] = cmul(z, __ldg(&chirp[k]));
    }
    ring_idft(P, buf, twq, n, bsi);
    double* oa = (job.compA ? mapU : mapQ) + P.ring_start[job.ringA];
    double* ob = job.ringB >= 0 ? (job.compB ? mapU : mapQ) + P.ring_start[job.ringB] : nullptr;
    for (int j = threadIdx.x; j < n; j += RF_NT) {
        double2 z = buf[j];
        if (bsi >= 0) z = cmul(z, __ldg(&chirp[j]));
        oa[j] = z.x;
        if (ob) ob[j] = z.y;
    }
}

__global__ void __launch_bounds__(RF_NT, 2)
ring_anal_kernel(PlanDev P, const RingJob* __restrict__ jobs, const double* __restrict__ mapQ, const double* __restrict__ mapU,
                 const double* __restrict__ pixw, double2* __restrict__ Fm, const int* __restrict__ skip)
{
    if (skip && *skip) return;
    extern __shared__ double2 smem[];
    const int L = P.lmax, nm = L + 1;
    const RingJob job = jobs[blockIdx.x];
    const int n = P.ring_nphi[job.ringA], bsi = P.ring_bs[job.ringA];
    double2* twq = smem;
    double2* buf = twq + (P.tw_n >> 2) + 1;
    load_twq(P, twq);
    const int64_t sa = P.ring_start[job.ringA], sb = job.ringB >= 0 ? P.ring_start[job.ringB] : 0;
    const double* ia = (job.compA ? mapU : mapQ) + sa;
    const double* ib = job.ringB >= 0 ? (job.compB ? mapU : mapQ) + sb : nullptr;
    const int lg = 31 - __clz(n);
    const double2* chirp = bsi >= 0 ? P.bs_tab + P.bs[bsi].chirp_off : nullptr;
    const int M = bsi >= 0 ? P.bs[bsi].M : n;
    // Z[k] = sum_j z_j exp(-2 pi i jk/n) = conj( idft( conj z ) )
    for (int j = threadIdx.x; j < M; j += RF_NT) {
        if (j >= n) { buf[j] = make_double2(0.0, 0.0); continue; }
        double a = ia[j], b = ib ? ib[j] : 0.0;
        if (pixw) { a *= pixw[sa + j]; if (ib) b *= pixw[sb + j]; }
        const double2 z = make_double2(a, -b);
        if (bsi < 0) buf[(int)(__brev((unsigned)j) >> (32 - lg))] = z;
        else buf[j] = cmul(z, __ldg(&chirp[j]));
    }
    ring_idft(P, buf, twq, n, bsi);
    if (bsi >= 0) {  // finish Bluestein in place: every m below reads two entries
        for (int k = threadIdx.x; k < n; k += RF_NT) buf[k] = cmul(buf[k], __ldg(&chirp[k]));
        __syncthreads();
    }
    double2* FA = Fm + ((int64_t)job.compA * P.nring + job.ringA) * nm;
    double2* FB = job.ringB >= 0 ? Fm + ((int64_t)job.compB * P.nring + job.ringB) * nm : nullptr;
    const bool same_phase = job.ringB == job.ringA ||
                            (job.ringB >= 0 && P.ring_phq[job.ringB] == P.ring_phq[job.ringA] && P.ring_phden[job.ringB] == P.ring_phden[job.ringA]);
    for (int m = threadIdx.x; m <= L; m += RF_NT) {
        const int k = m % n, kk = (n - k) % n;
        const double2 c1 = buf[k], c2 = buf[kk];
        const double2 z1 = make_double2(c1.x, -c1.y);  // Z[k]
        const double2 z2c = c2;                         // conj Z[n-k]
        const double2 xa = make_double2(0.5 * (z1.x + z2c.x), 0.5 * (z1.y + z2c.y));
        const double2 d = csub(z1, z2c);
        const double2 xb = make_double2(0.5 * d.y, -0.5 * d.x);
        const double2 pa = ring_phase(P, job.ringA, m);
        FA[m] = cmulc(xa, pa);
        if (FB) { const double2 pb = same_phase ? pa : ring_phase(P, job.ringB, m); FB[m] = cmulc(xb, pb); }
    }
}

// ------------------------------------------------------------------ host side
static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

int gs_ring_setup(gs_plan* p)
{
    const int nside = p->d.nside, nring = p->d.nring, npair = p->d.npair, L = p->d.lmax;
    std::vector<int> rn(nring);
    for (int r = 0; r < nring; ++r) { int i = std::min(r + 1, 4 * nside - (r + 1)); rn[r] = i < nside ? 4 * i : 4 * nside; }
    std::map<int, int> n2bs;
    std::vector<BluesteinDesc> descs;
    int64_t off = 0;
    int maxM = 4;
    for (int r = 0; r < npair; ++r) {
        const int n = rn[r];
        if ((n & (n - 1)) == 0) { maxM = std::max(maxM, n); continue; }
        if (n2bs.count(n)) continue;
        BluesteinDesc d;
        d.n = n; d.M = next_pow2(2 * n - 1); d.chirp_off = off; off += n; d.bhat_off = off; off += d.M;
        n2bs[n] = (int)descs.size();
        descs.push_back(d);
        maxM = std::max(maxM, d.M);
    }
    std::vector<int> rbs(nring);
    for (int r = 0; r < nring; ++r) rbs[r] = n2bs.count(rn[r]) ? n2bs[rn[r]] : -1;
    p->d.max_M = maxM;
    p->d.tw_n = maxM;
    p->ring_smem = (size_t)(2 * (L + 1) + maxM + maxM / 4 + 1) * sizeof(double2);
    if (p->ring_smem > 227 * 1024) {
        gs_set_error("ring FFT needs %zu bytes of shared memory (> 227 KB): nside/lmax too large for this build", p->ring_smem);
        return GS_E_BADARG;
    }
    std::vector<double2> tw(maxM);
    for (int k = 0; k < maxM; ++k) {
        long double a = 2.0L * 3.14159265358979323846264338327950288L * k / maxM;
        tw[k] = make_double2((double)cosl(a), (double)(-sinl(a)));
    }
    void* d = nullptr;
    GS_CHECK_CUDA(cudaMalloc(&d, tw.size() * sizeof(double2))); p->owned.push_back(d);
    GS_CHECK_CUDA(cudaMemcpy(d, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice));
    p->d.tw = (const double2*)d;
    GS_CHECK_CUDA(cudaMalloc(&d, std::max<size_t>(1, nring) * sizeof(int))); p->owned.push_back(d);
    GS_CHECK_CUDA(cudaMemcpy(d, rbs.data(), nring * sizeof(int), cudaMemcpyHostToDevice));
    p->d.ring_bs = (const int*)d;
    GS_CHECK_CUDA(cudaMalloc(&d, std::max<size_t>(1, descs.size()) * sizeof(BluesteinDesc))); p->owned.push_back(d);
    if (!descs.empty()) GS_CHECK_CUDA(cudaMemcpy(d, descs.data(), descs.size() * sizeof(BluesteinDesc), cudaMemcpyHostToDevice));
    p->d.bs = (const BluesteinDesc*)d;
    GS_CHECK_CUDA(cudaMalloc(&d, std::max<int64_t>(1, off) * sizeof(double2))); p->owned.push_back(d);
    p->d.bs_tab = (const double2*)d;

    GS_CHECK_CUDA(cudaFuncSetAttribute(bluestein_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    GS_CHECK_CUDA(cudaFuncSetAttribute(ring_synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    GS_CHECK_CUDA(cudaFuncSetAttribute(ring_anal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (!descs.empty()) {
        bluestein_setup_kernel<<<(int)descs.size(), RF_NT, p->ring_smem>>>(p->d, (double2*)d, (int)descs.size());
        GS_CHECK_LAUNCH();
    }

    // job lists, heaviest transforms first
    auto cost = [&](int ring) { int n = rn[ring]; return rbs[ring] < 0 ? n : 3 * descs[rbs[ring]].M; };
    std::vector<RingJob> j2(nring), j0;
    for (int r = 0; r < nring; ++r) j2[r] = RingJob{r, 0, r, 1};
    for (int r = 0; r < npair; ++r) { int rs = nring - 1 - r; j0.push_back(RingJob{r, 0, rs != r ? rs : -1, 0}); }
    std::stable_sort(j2.begin(), j2.end(), [&](const RingJob& a, const RingJob& b) { return cost(a.ringA) > cost(b.ringA); });
    std::stable_sort(j0.begin(), j0.end(), [&](const RingJob& a, const RingJob& b) { return cost(a.ringA) > cost(b.ringA); });
    GS_CHECK_CUDA(cudaMalloc(&d, j2.size() * sizeof(RingJob))); p->owned.push_back(d);
    GS_CHECK_CUDA(cudaMemcpy(d, j2.data(), j2.size() * sizeof(RingJob), cudaMemcpyHostToDevice));
    p->jobs2 = (RingJob*)d; p->njobs2 = (int)j2.size();
    GS_CHECK_CUDA(cudaMalloc(&d, j0.size() * sizeof(RingJob))); p->owned.push_back(d);
    GS_CHECK_CUDA(cudaMemcpy(d, j0.data(), j0.size() * sizeof(RingJob), cudaMemcpyHostToDevice));
    p->jobs0 = (RingJob*)d; p->njobs0 = (int)j0.size();
    return GS_OK;
}

int gs_ring_synth(gs_plan* p, int spin, double* mapQ, double* mapU, cudaStream_t st, const int* skip)
{
    if (spin == 0) ring_synth_kernel<<<p->njobs0, RF_NT, p->ring_smem, st>>>(p->d, p->jobs0, p->Fm, mapQ, mapQ, skip);
    else ring_synth_kernel<<<p->njobs2, RF_NT, p->ring_smem, st>>>(p->d, p->jobs2, p->Fm, mapQ, mapU, skip);
    GS_CHECK_LAUNCH();
    g_gs_launches += 1;
    return GS_OK;
}

int gs_ring_anal(gs_plan* p, int spin, const double* mapQ, const double* mapU, const double* pixw, cudaStream_t st,
                 const int* skip)
{
    if (spin == 0) ring_anal_kernel<<<p->njobs0, RF_NT, p->ring_smem, st>>>(p->d, p->jobs0, mapQ, mapQ, pixw, p->Fm, skip);
    else ring_anal_kernel<<<p->njobs2, RF_NT, p->ring_smem, st>>>(p->d, p->jobs2, mapQ, mapU, pixw, p->Fm, skip);
    GS_CHECK_LAUNCH();
    g_gs_launches += 1;
    return GS_OK;
}
