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
#include <stdlib.h>

#include <algorithm>
#include <map>

#include "gs_internal.h"

#define RF_NT 256

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a conj(b)
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// ---- shared-memory FFT core -------------------------------------------------------------------
// Layout: element i of the transform lives at buf[PADI(i)], PADI(i) = i + (i >> 4): one 16-byte pad per
// 16 elements makes the stride-16 accesses of the final radix-16 pass bank-conflict free and leaves the
// unit-stride passes (consecutive lanes -> consecutive elements) conflict free as well.
// Passes: a radix-16 butterfly (= two radix-4 levels) is done entirely in registers, so a 4096-point
// transform makes 3 trips through shared memory (6 with radix-4); leftover stages (log2 M mod 4) are a
// radix-2 and/or radix-4 pass at the top (DIF) / bottom (DIT).
// Twiddles: quarter table twq[k] = exp(-2 pi i k / twn), k <= twn/4, in shared memory;
// W^(k + twn/4) = -i W^k covers the second quadrant (indices stay below twn/2).
#define PADI(i) ((i) + ((i) >> 4))

__device__ __forceinline__ double2 tw_get(const double2* twq, int idx, int quarter)
{
    if (idx < quarter) return twq[idx];
    const double2 w = twq[idx - quarter];
    return make_double2(w.y, -w.x);  // -i w
}

// radix-4 DIF butterfly in registers: twiddles w1 = W^x, w2 = W^2x (forward kernel exp(-2 pi i ..))
__device__ __forceinline__ void bf4_dif(double2& a0, double2& a1, double2& a2, double2& a3, double2 w1, double2 w2)
{
    const double2 b0 = cadd(a0, a2), b2 = cmul(csub(a0, a2), w1), b1 = cadd(a1, a3);
    const double2 d13 = csub(a1, a3);
    const double2 b3 = cmul(make_double2(d13.y, -d13.x), w1);  // (a1 - a3) (-i) W^x
    a0 = cadd(b0, b1); a1 = cmul(csub(b0, b1), w2); a2 = cadd(b2, b3); a3 = cmul(csub(b2, b3), w2);
}
// its transpose-conjugate (inverse DIT butterfly, kernel exp(+2 pi i ..))
__device__ __forceinline__ void bf4_dit(double2& c0, double2& c1, double2& c2, double2& c3, double2 w1, double2 w2)
{
    const double2 t1 = cmulc(c1, w2), t3 = cmulc(c3, w2);
    const double2 b0 = cadd(c0, t1), b1 = csub(c0, t1), b2 = cadd(c2, t3), b3 = csub(c2, t3);
    const double2 t2 = cmulc(b2, w1), v3 = cmulc(b3, w1);
    const double2 u3 = make_double2(-v3.y, v3.x);  // (+i) conj(W)^x b3
    c0 = cadd(b0, t2); c2 = csub(b0, t2); c1 = cadd(b1, u3); c3 = csub(b1, u3);
}

// exp(-2 pi i a / 16) and exp(-2 pi i a / 8), a = 0..3
__device__ __forceinline__ double2 w16c(int a)
{
    const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
    return a == 0 ? make_double2(1.0, 0.0) : a == 1 ? make_double2(c1, -s1) : a == 2 ? make_double2(h, -h) : make_double2(s1, -c1);
}
__device__ __forceinline__ double2 w8c(int a)
{
    const double h = 0.70710678118654752440;
    return a == 0 ? make_double2(1.0, 0.0) : a == 1 ? make_double2(h, -h) : a == 2 ? make_double2(0.0, -1.0) : make_double2(-h, -h);
}

// radix-16 pass over sub-transforms of size N (N >= 16).  INV = false: DIF (forward), true: DIT (inverse).
// POST (DIF only): outputs are multiplied by post[position].
template <bool INV, bool POST>
__device__ __forceinline__ void pass16(double2* buf, int M, int N, const double2* twq, int twn, const double2* __restrict__ post)
{
    const int s = N >> 4, ls = 31 - __clz(s), ts = twn / N, quarter = twn >> 2, ng = M >> 4;
    for (int t = threadIdx.x; t < ng; t += RF_NT) {
        const int g = t >> ls, j = t & (s - 1), i0 = g * N + j;
        double2 x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = buf[PADI(i0 + k * s)];
        double2 wj = make_double2(1.0, 0.0), w2j = wj, w4j = wj, w8j = wj;
        if (s > 1) {
            wj = twq[j * ts];
            w2j = tw_get(twq, 2 * j * ts, quarter);
            w4j = tw_get(twq, 4 * j * ts, quarter);
            w8j = tw_get(twq, 8 * j * ts, quarter);
        }
        if (!INV) {
#pragma unroll
            for (int a = 0; a < 4; ++a) bf4_dif(x[a], x[a + 4], x[a + 8], x[a + 12], cmul(wj, w16c(a)), cmul(w2j, w8c(a)));
#pragma unroll
            for (int c = 0; c < 4; ++c) bf4_dif(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3], w4j, w8j);
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) bf4_dit(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3], w4j, w8j);
#pragma unroll
            for (int a = 0; a < 4; ++a) bf4_dit(x[a], x[a + 4], x[a + 8], x[a + 12], cmul(wj, w16c(a)), cmul(w2j, w8c(a)));
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            double2 v = x[k];
            if (POST) v = cmul(v, __ldg(&post[i0 + k * s]));
            buf[PADI(i0 + k * s)] = v;
        }
    }
    __syncthreads();
}

// radix-4 pass over sub-transforms of size N (N >= 4)
template <bool INV>
__device__ __forceinline__ void pass4(double2* buf, int M, int N, const double2* twq, int twn)
{
    const int q = N >> 2, lq = 31 - __clz(q), ts = twn / N, quarter = twn >> 2, nq = M >> 2;
    for (int t = threadIdx.x; t < nq; t += RF_NT) {
        const int g = t >> lq, j = t & (q - 1), i0 = g * N + j;
        double2 a0 = buf[PADI(i0)], a1 = buf[PADI(i0 + q)], a2 = buf[PADI(i0 + 2 * q)], a3 = buf[PADI(i0 + 3 * q)];
        const double2 w1 = twq[j * ts], w2 = tw_get(twq, 2 * j * ts, quarter);
        if (!INV) bf4_dif(a0, a1, a2, a3, w1, w2); else bf4_dit(a0, a1, a2, a3, w1, w2);
        buf[PADI(i0)] = a0; buf[PADI(i0 + q)] = a1; buf[PADI(i0 + 2 * q)] = a2; buf[PADI(i0 + 3 * q)] = a3;
    }
    __syncthreads();
}

// radix-2 pass over sub-transforms of size N (N >= 2)
template <bool INV>
__device__ __forceinline__ void pass2(double2* buf, int M, int N, const double2* twq, int twn)
{
    const int h = N >> 1, lh = 31 - __clz(h), ts = twn / N;
    for (int t = threadIdx.x; t < (M >> 1); t += RF_NT) {
        const int g = t >> lh, j = t & (h - 1), i0 = g * N + j;
        const double2 a = buf[PADI(i0)], b = buf[PADI(i0 + h)];
        const double2 w = tw_get(twq, j * ts, twn >> 2);
        if (!INV) { buf[PADI(i0)] = cadd(a, b); buf[PADI(i0 + h)] = cmul(csub(a, b), w); }
        else { const double2 tb = cmulc(b, w); buf[PADI(i0)] = cadd(a, tb); buf[PADI(i0 + h)] = csub(a, tb); }
    }
    __syncthreads();
}

// In-place forward DIF FFT (kernel exp(-2 pi i jk/M)), natural order in, bit-reversed order out;
// POST: the bit-reversed-order output is multiplied by post[] inside the last pass (needs M >= 16).
template <bool POST>
__device__ void fft_dif(double2* buf, int M, const double2* twq, int twn, const double2* __restrict__ post)
{
    const int lg = 31 - __clz(M), r = lg & 3;
    int N = M;
    if (r & 1) { pass2<false>(buf, M, N, twq, twn); N >>= 1; }
    if (r & 2) { pass4<false>(buf, M, N, twq, twn); N >>= 2; }
    while (N >= 16) {
        if (POST && N == 16) pass16<false, true>(buf, M, N, twq, twn, post);
        else pass16<false, false>(buf, M, N, twq, twn, post);
        N >>= 4;
    }
}

// In-place inverse DIT FFT (kernel exp(+2 pi i jk/M), unnormalised), bit-reversed in, natural out.
__device__ void fft_dit_inv(double2* buf, int M, const double2* twq, int twn)
{
    const int lg = 31 - __clz(M), r = lg & 3;
    const int Ntop = M >> r;  // largest radix-16 sub-transform size
    for (int N = 16; N <= Ntop; N <<= 4) pass16<true, false>(buf, M, N, twq, twn, nullptr);
    int N = Ntop;
    if (r & 2) { N <<= 2; pass4<true>(buf, M, N, twq, twn); }
    if (r & 1) { N <<= 1; pass2<true>(buf, M, N, twq, twn); }
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
    double2* buf = smem_all + (P.tw_n >> 2) + 1;
    load_twq(P, twq);
    const BluesteinDesc d = P.bs[blockIdx.x];
    double2* chirp = tab + d.chirp_off;
    double2* bhat = tab + d.bhat_off;
    for (int k = threadIdx.x; k < d.M; k += blockDim.x) buf[PADI(k)] = make_double2(0.0, 0.0);
    __syncthreads();
    for (int t = threadIdx.x; t < d.n; t += blockDim.x) {
        const double2 c = chirp_val(t, d.n);
        chirp[t] = c;
        const double2 cc = make_double2(c.x, -c.y);
        buf[PADI(t)] = cc;
        if (t > 0) buf[PADI(d.M - t)] = cc;
    }
    __syncthreads();
    fft_dif<false>(buf, d.M, twq, P.tw_n, nullptr);
    const double inv = 1.0 / (double)d.M;
    for (int k = threadIdx.x; k < d.M; k += blockDim.x) { const double2 v = buf[PADI(k)]; bhat[k] = make_double2(v.x * inv, v.y * inv); }
    (void)nbs;
}

__device__ __forceinline__ double2 ring_phase(const PlanDev& P, int ring, int m)
{  // exp(i m phi0), phi0 = pi q / den
    const int q = P.ring_phq[ring], den = P.ring_phden[ring];
    const unsigned r = ((unsigned)m * (unsigned)q) % (2u * (unsigned)den);  // m q < 2^31 for every supported size
    double s, c;
    sincospi((double)r / (double)den, &s, &c);
    return make_double2(c, s);
}

// Position of F_m(ring) in the ring-spectra buffer as seen from the ring-sharded side (see ShardDev).
template <bool SH>
__device__ __forceinline__ int64_t fm_ring_index(const PlanDev& P, int comp, int ring, int m)
{
    if (SH) return (((int64_t)P.sh.m_owner[m] * 2 + comp) * P.sh.RL + P.sh.ring_loc[ring]) * P.sh.ML + P.sh.m_loc[m];
    return ((int64_t)comp * P.nring + ring) * (P.lmax + 1) + m;
}
template <bool SH>
__device__ __forceinline__ int64_t ring_first_pixel(const PlanDev& P, int ring)
{
    return SH ? P.sh.ring_start_loc[ring] : P.ring_start[ring];
}

// Z[k] = Xa[k] + i Xb[k], X[k] = G[k] + conj G[n-k], G[k] = sum_{m = k mod n} w_m F_m e^{i m phi0} (alias fold), for the two
// real sequences of a job, written to buf ready for ring_idft: bit-reversed (power-of-two n) or chirp-multiplied and
// zero-padded (Bluestein).  F_m is taken as zero for m > mtop.
//  * n > lmax (every belt ring and the longer cap rings): each k aliases at most one m on either side, so the
//    spectrum is read straight from global memory, both components in one sweep (e^{i (n-k) phi0} = e^{i n phi0}
//    conj e^{i k phi0}, phases by recurrence over the thread's k).
//  * shorter rings: the phased spectrum is staged in st[0..lmax] one component at a time and folded from there.
// All threads of the CTA must call; the caller synchronises before the transform (ring_idft does).
template <bool SH>
__device__ __forceinline__ void ring_build_Z(const PlanDev& P, const RingJob& job, const double2* __restrict__ Fm, int mtop,
                                             double2* st, double2* buf, int n, int bsi, int M, const double2* __restrict__ chirp)
{
    const int L = P.lmax, nm = L + 1, lg = 31 - __clz(n);
    const bool same_phase = job.ringB < 0 || job.ringB == job.ringA ||
                            (P.ring_phq[job.ringB] == P.ring_phq[job.ringA] && P.ring_phden[job.ringB] == P.ring_phden[job.ringA]);
    if (n > L) {
        const double2* FA = Fm + ((int64_t)job.compA * P.nring + job.ringA) * nm;
        const double2* FB = job.ringB >= 0 ? Fm + ((int64_t)job.compB * P.nring + job.ringB) * nm : nullptr;
        double2 pa = ring_phase(P, job.ringA, threadIdx.x);
        const double2 stepa = ring_phase(P, job.ringA, RF_NT), phna = ring_phase(P, job.ringA, n);
        double2 pb = pa, stepb = stepa, phnb = phna;
        if (!same_phase) { pb = ring_phase(P, job.ringB, threadIdx.x); stepb = ring_phase(P, job.ringB, RF_NT); phnb = ring_phase(P, job.ringB, n); }
        for (int k = threadIdx.x; k < M; k += RF_NT) {
            const int pos = PADI(bsi < 0 ? (int)(__brev((unsigned)k) >> (32 - lg)) : k);
            if (k >= n) { buf[pos] = make_double2(0.0, 0.0); continue; }
            const int kk = k ? n - k : 0;
            const double wk = k ? 1.0 : 0.5;   // (2 - delta_m0) / 2; kk = 0 only when k = 0
            double2 fa = make_double2(0.0, 0.0), fa2 = fa, fb = fa, fb2 = fa;
            if (k <= mtop) {
                fa = SH ? Fm[fm_ring_index<true>(P, job.compA, job.ringA, k)] : FA[k];
                if (job.ringB >= 0) fb = SH ? Fm[fm_ring_index<true>(P, job.compB, job.ringB, k)] : FB[k];
            }
            if (kk <= mtop) {
                fa2 = SH ? Fm[fm_ring_index<true>(P, job.compA, job.ringA, kk)] : FA[kk];
                if (job.ringB >= 0) fb2 = SH ? Fm[fm_ring_index<true>(P, job.compB, job.ringB, kk)] : FB[kk];
            }
            const double2 pka = make_double2(pa.x * wk, pa.y * wk), pkb = make_double2(pb.x * wk, pb.y * wk);
            const double2 qa = k ? cmulc(phna, pa) : make_double2(0.5, 0.0), qb = k ? cmulc(phnb, pb) : make_double2(0.5, 0.0);
            const double2 ga = cmul(fa, pka), ha = cmul(fa2, qa), gb = cmul(fb, pkb), hb = cmul(fb2, qb);
            const double2 xa = make_double2(ga.x + ha.x, ga.y - ha.y), xb = make_double2(gb.x + hb.x, gb.y - hb.y);
            double2 z = make_double2(xa.x - xb.y, xa.y + xb.x);   // Xa + i Xb
            if (bsi >= 0) z = cmul(z, __ldg(&chirp[k]));
            buf[pos] = z;
            pa = cmul(pa, stepa);
            pb = cmul(pb, stepb);
        }
        return;
    }
    for (int comp = 0; comp < 2; ++comp) {
        const int ring = comp ? job.ringB : job.ringA;
        if (comp) __syncthreads();
        if (ring >= 0) {
            const int cc = comp ? job.compB : job.compA;
            const double2* F = Fm + ((int64_t)cc * P.nring + ring) * nm;
            // e^{i m phi0} for m = tid, tid + 256, ...: one sincospi, then multiply by e^{i 256 phi0}
            const int pr = (comp && !same_phase) ? job.ringB : job.ringA;
            double2 ph = ring_phase(P, pr, threadIdx.x);
            const double2 step = ring_phase(P, pr, RF_NT);
            for (int m = threadIdx.x; m <= L; m += RF_NT) {
                const double w = m ? 1.0 : 0.5;  // (2 - delta_m0) / 2
                double2 f = make_double2(0.0, 0.0);
                if (m <= mtop) f = SH ? Fm[fm_ring_index<true>(P, cc, ring, m)] : F[m];
                st[m] = cmul(f, make_double2(ph.x * w, ph.y * w));
                ph = cmul(ph, step);
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < M; k += RF_NT) {
            const int pos = PADI(bsi < 0 ? (int)(__brev((unsigned)k) >> (32 - lg)) : k);
            if (k >= n) { if (!comp) buf[pos] = make_double2(0.0, 0.0); continue; }
            double2 x = make_double2(0.0, 0.0);
            if (ring >= 0) {
                const int kk = (n - k) % n;
                double2 g = make_double2(0.0, 0.0), h = g;
                for (int m = k; m <= L; m += n) g = cadd(g, st[m]);
                for (int m = kk; m <= L; m += n) h = cadd(h, st[m]);
                x = make_double2(g.x + h.x, g.y - h.y);
            }
            if (!comp) buf[pos] = x;
            else {
                double2 z = buf[pos];
                z = make_double2(z.x - x.y, z.y + x.x);   // + i Xb
                if (bsi >= 0) z = cmul(z, __ldg(&chirp[k]));
                buf[pos] = z;
            }
        }
    }
}

// F_m of the two real sequences of a job from the length-n DFT held in buf, m = 0..lmax:
//   Xa[k] = (Z[k] + conj Z[n-k]) / 2, Xb[k] = (Z[k] - conj Z[n-k]) / (2i), F_m = X[m mod n] e^{-i m phi0}.
// BR = false: buf[PADI(k)] = conj Z[k] (transform done as conj(idft(conj z))); BR = true: buf holds Z[k] at the
// bit-reversed position of k (forward DIF transform of a power-of-two ring).
template <bool SH>
__device__ __forceinline__ void ring_unpack_F(const PlanDev& P, const RingJob& job, const double2* buf, int n, bool br,
                                              double2* __restrict__ Fm)
{
    const int L = P.lmax, nm = L + 1, lg = 31 - __clz(n);
    double2* FA = Fm + ((int64_t)job.compA * P.nring + job.ringA) * nm;
    double2* FB = job.ringB >= 0 ? Fm + ((int64_t)job.compB * P.nring + job.ringB) * nm : nullptr;
    const bool same_phase = job.ringB < 0 || job.ringB == job.ringA ||
                            (P.ring_phq[job.ringB] == P.ring_phq[job.ringA] && P.ring_phden[job.ringB] == P.ring_phden[job.ringA]);
    double2 pa = ring_phase(P, job.ringA, threadIdx.x), pb = same_phase ? pa : ring_phase(P, job.ringB, threadIdx.x);
    const double2 stepa = ring_phase(P, job.ringA, RF_NT), stepb = same_phase ? stepa : ring_phase(P, job.ringB, RF_NT);
    for (int m = threadIdx.x; m <= L; m += RF_NT) {
        const int k = m % n, kk = (n - k) % n;
        double2 z1, z2c;
        if (br) {
            const double2 c1 = buf[PADI((int)(__brev((unsigned)k) >> (32 - lg)))], c2 = buf[PADI((int)(__brev((unsigned)kk) >> (32 - lg)))];
            z1 = c1; z2c = make_double2(c2.x, -c2.y);
        } else {
            const double2 c1 = buf[PADI(k)], c2 = buf[PADI(kk)];
            z1 = make_double2(c1.x, -c1.y); z2c = c2;
        }
        const double2 xa = make_double2(0.5 * (z1.x + z2c.x), 0.5 * (z1.y + z2c.y));
        const double2 d = csub(z1, z2c);
        const double2 xb = make_double2(0.5 * d.y, -0.5 * d.x);
        if (SH) {
            Fm[fm_ring_index<true>(P, job.compA, job.ringA, m)] = cmulc(xa, pa);
            if (FB) Fm[fm_ring_index<true>(P, job.compB, job.ringB, m)] = cmulc(xb, pb);
        } else {
            FA[m] = cmulc(xa, pa);
            if (FB) FB[m] = cmulc(xb, pb);
        }
        pa = cmul(pa, stepa);
        pb = cmul(pb, stepb);
    }
}

template <bool SH>
__global__ void __launch_bounds__(RF_NT, 2)
ring_synth_kernel(PlanDev P, const RingJob* __restrict__ jobs, const double2* __restrict__ Fm, double* __restrict__ mapQ,
                  double* __restrict__ mapU, const int* __restrict__ skip, int64_t f_stride, int64_t map_stride,
                  const int* __restrict__ mmax)
{
    if (skip && *skip) return;
    extern __shared__ double2 smem[];
    // batched use (blockIdx.y = transform index): spectra at Fm + y f_stride hold m <= mmax[y] only, maps at + y map_stride
    Fm += blockIdx.y * f_stride;
    mapQ += blockIdx.y * map_stride;
    mapU += blockIdx.y * map_stride;
    const int L = P.lmax, nm = L + 1, mtop = mmax ? min(mmax[blockIdx.y], L) : L;
    const RingJob job = jobs[blockIdx.x];
    const int n = P.ring_nphi[job.ringA], bsi = P.ring_bs[job.ringA];
    double2* twq = smem;
    double2* st = twq + (P.tw_n >> 2) + 1;   // one component of the phased ring spectrum at a time (rings with n <= lmax)
    double2* buf = st + nm;
    load_twq(P, twq);
    const double2* chirp = bsi >= 0 ? P.bs_tab + P.bs[bsi].chirp_off : nullptr;
    const int M = bsi >= 0 ? P.bs[bsi].M : n;
    ring_build_Z<SH>(P, job, Fm, mtop, st, buf, n, bsi, M, chirp);
    ring_idft(P, buf, twq, n, bsi);
    double* oa = (job.compA ? mapU : mapQ) + ring_first_pixel<SH>(P, job.ringA);
    double* ob = job.ringB >= 0 ? (job.compB ? mapU : mapQ) + ring_first_pixel<SH>(P, job.ringB) : nullptr;
    for (int j = threadIdx.x; j < n; j += RF_NT) {
        double2 z = buf[PADI(j)];
        if (bsi >= 0) z = cmul(z, __ldg(&chirp[j]));
        oa[j] = z.x;
        if (ob) ob[j] = z.y;
    }
}

template <bool SH>
__global__ void __launch_bounds__(RF_NT, 2)
ring_anal_kernel(PlanDev P, const RingJob* __restrict__ jobs, const double* __restrict__ mapQ, const double* __restrict__ mapU,
                 const double* __restrict__ pixw, double2* __restrict__ Fm, const int* __restrict__ skip)
{
    if (skip && *skip) return;
    extern __shared__ double2 smem[];
    const RingJob job = jobs[blockIdx.x];
    const int n = P.ring_nphi[job.ringA], bsi = P.ring_bs[job.ringA];
    double2* twq = smem;
    double2* buf = twq + (P.tw_n >> 2) + 1;
    load_twq(P, twq);
    const int64_t sa = ring_first_pixel<SH>(P, job.ringA), sb = job.ringB >= 0 ? ring_first_pixel<SH>(P, job.ringB) : 0;
    const double* ia = (job.compA ? mapU : mapQ) + sa;
    const double* ib = job.ringB >= 0 ? (job.compB ? mapU : mapQ) + sb : nullptr;
    const int lg = 31 - __clz(n);
    const double2* chirp = bsi >= 0 ? P.bs_tab + P.bs[bsi].chirp_off : nullptr;
    const int M = bsi >= 0 ? P.bs[bsi].M : n;
    // Z[k] = sum_j z_j exp(-2 pi i jk/n) = conj( idft( conj z ) )
    for (int j = threadIdx.x; j < M; j += RF_NT) {
        const int pos = PADI(bsi < 0 ? (int)(__brev((unsigned)j) >> (32 - lg)) : j);
        if (j >= n) { buf[pos] = make_double2(0.0, 0.0); continue; }
        double a = ia[j], b = ib ? ib[j] : 0.0;
        if (pixw) { a *= pixw[sa + j]; if (ib) b *= pixw[sb + j]; }
        double2 z = make_double2(a, -b);
        if (bsi >= 0) z = cmul(z, __ldg(&chirp[j]));
        buf[pos] = z;
    }
    ring_idft(P, buf, twq, n, bsi);
    if (bsi >= 0) {  // finish Bluestein in place: every m below reads two entries
        for (int k = threadIdx.x; k < n; k += RF_NT) buf[PADI(k)] = cmul(buf[PADI(k)], __ldg(&chirp[k]));
        __syncthreads();
    }
    ring_unpack_F<SH>(P, job, buf, n, false, Fm);
}

// Ring stage of the PCG mat-vec A^T N^-1 A in ONE kernel: F_m(ring) -> pixels of the ring (kept in shared memory) ->
// times the pixel weights -> F'_m(ring), written over F_m.  Equivalent to ring_synth_kernel + ring_anal_kernel(pixw)
// without the map round trip through global memory, the second table load and the second launch.
template <bool SH>
__global__ void __launch_bounds__(RF_NT, 2)
ring_apply_kernel(PlanDev P, const RingJob* __restrict__ jobs, double2* __restrict__ Fm, const double* __restrict__ pixw,
                  const int* __restrict__ skip)
{
    if (skip && *skip) return;
    extern __shared__ double2 smem[];
    const int L = P.lmax, nm = L + 1;
    const RingJob job = jobs[blockIdx.x];
    const int n = P.ring_nphi[job.ringA], bsi = P.ring_bs[job.ringA];
    double2* twq = smem;
    double2* st = twq + (P.tw_n >> 2) + 1;
    double2* buf = st + nm;
    load_twq(P, twq);
    const double2* chirp = bsi >= 0 ? P.bs_tab + P.bs[bsi].chirp_off : nullptr;
    const int M = bsi >= 0 ? P.bs[bsi].M : n;
    const double* wa = pixw + ring_first_pixel<SH>(P, job.ringA);
    const double* wb = job.ringB >= 0 ? pixw + ring_first_pixel<SH>(P, job.ringB) : wa;
    ring_build_Z<SH>(P, job, Fm, L, st, buf, n, bsi, M, chirp);
    ring_idft(P, buf, twq, n, bsi);
    if (bsi < 0) {
        // pixels z_j = a_j + i b_j in natural order -> weighted -> forward DIF transform (bit-reversed output)
        for (int j = threadIdx.x; j < n; j += RF_NT) {
            const double2 z = buf[PADI(j)];
            buf[PADI(j)] = make_double2(z.x * wa[j], z.y * wb[j]);
        }
        __syncthreads();
        fft_dif<false>(buf, n, twq, P.tw_n, nullptr);
        ring_unpack_F<SH>(P, job, buf, n, true, Fm);
    } else {
        // Bluestein both ways: Z = conj(idft(conj z)); the chirp of the synthesis output and of the analysis input fuse
        for (int j = threadIdx.x; j < M; j += RF_NT) {
            if (j >= n) { buf[PADI(j)] = make_double2(0.0, 0.0); continue; }
            const double2 c = __ldg(&chirp[j]);
            const double2 z = cmul(buf[PADI(j)], c);
            buf[PADI(j)] = cmul(make_double2(z.x * wa[j], -z.y * wb[j]), c);
        }
        ring_idft(P, buf, twq, n, bsi);
        for (int k = threadIdx.x; k < n; k += RF_NT) buf[PADI(k)] = cmul(buf[PADI(k)], __ldg(&chirp[k]));
        __syncthreads();
        ring_unpack_F<SH>(P, job, buf, n, false, Fm);
    }
}

// ------------------------------------------------------------------ split path (rings longer than one CTA can hold)
// n = 4 n2.  Synthesis (z_j = sum_k Z_k e^{+2 pi i jk/n}, j = j2 + n2 q, k = 4a + b):
//   z_{j2 + n2 q} = sum_b i^{qb} e^{2 pi i j2 b/n} Y_b[j2],   Y_b[j2] = sum_a Z_{4a+b} e^{2 pi i j2 a/n2}
// CTA (job, b) computes Y_b with the shared-memory transform of length n2 and stores it in the scratch;
// ring_synth_combine_kernel applies the twiddles and the radix-4 butterfly and writes the pixels.
// Analysis is the transpose: CTA (job, b) loads v_b[j2] = e^{2 pi i j2 b/n} sum_q i^{qb} c_{j2 + n2 q}, transforms
// it to C_{4a+b}, and ring_anal_finish_kernel unpacks the two real sequences into F_m.
template <bool SH>
__device__ __forceinline__ double2 fold_spectrum(const PlanDev& P, const double2* __restrict__ Fm, int comp, int ring, int k, int n)
{  // X_k = G_k + conj G_{n-k},  G_k = sum_{m = k mod n} w_m F_m e^{i m phi0}
    const int L = P.lmax, kk = (n - k) % n;
    double2 g = make_double2(0.0, 0.0), h = g;
    for (int m = k; m <= L; m += n) {
        const double w = m ? 1.0 : 0.5;
        const double2 ph = ring_phase(P, ring, m);
        g = cadd(g, cmul(Fm[fm_ring_index<SH>(P, comp, ring, m)], make_double2(ph.x * w, ph.y * w)));
    }
    for (int m = kk; m <= L; m += n) {
        const double w = m ? 1.0 : 0.5;
        const double2 ph = ring_phase(P, ring, m);
        h = cadd(h, cmul(Fm[fm_ring_index<SH>(P, comp, ring, m)], make_double2(ph.x * w, ph.y * w)));
    }
    return make_double2(g.x + h.x, g.y - h.y);
}

__device__ __forceinline__ double2 mul_ipow(double2 v, int p)
{  // v * i^p
    p &= 3;
    return p == 0 ? v : p == 1 ? make_double2(-v.y, v.x) : p == 2 ? make_double2(-v.x, -v.y) : make_double2(v.y, -v.x);
}

__device__ __forceinline__ double2 unit_root(int num, int n)
{  // exp(2 pi i num / n), 0 <= num
    const int r = (int)((2LL * num) % (2LL * n));
    double s, c;
    sincospi((double)r / (double)n, &s, &c);
    return make_double2(c, s);
}

template <bool SH>
__global__ void __launch_bounds__(RF_NT, 2)
ring_synth_split_kernel(PlanDev P, const SplitJob* __restrict__ jobs, const double2* __restrict__ Fm, double2* __restrict__ scratch,
                        const int* __restrict__ skip)
{
    if (skip && *skip) return;
    extern __shared__ double2 smem[];
    const SplitJob job = jobs[blockIdx.x >> 2];
    const int b = blockIdx.x & 3;
    const int n = P.ring_nphi[job.ringA], n2 = n >> 2, bsi = job.bs2;
    double2* twq = smem;
    double2* buf = twq + (P.tw_n >> 2) + 1;
    load_twq(P, twq);
    const int lg = 31 - __clz(n2);
    const double2* chirp = bsi >= 0 ? P.bs_tab + P.bs[bsi].chirp_off : nullptr;
    const int M = bsi >= 0 ? P.bs[bsi].M : n2;
    for (int a = threadIdx.x; a < M; a += RF_NT) {
        const int pos = PADI(bsi < 0 ? (int)(__brev((unsigned)a) >> (32 - lg)) : a);
        if (a >= n2) { buf[pos] = make_double2(0.0, 0.0); continue; }
        const int k = 4 * a + b;
        const double2 xa = fold_spectrum<SH>(P, Fm, job.compA, job.ringA, k, n);
        double2 z = xa;
        if (job.ringB >= 0) {
            const double2 xb = fold_spectrum<SH>(P, Fm, job.compB, job.ringB, k, n);
            z = make_double2(xa.x - xb.y, xa.y + xb.x);  // + i Xb
        }
        if (bsi >= 0) z = cmul(z, __ldg(&chirp[a]));
        buf[pos] = z;
    }
    ring_idft(P, buf, twq, n2, bsi);
    double2* out = scratch + job.off + (int64_t)b * n2;
    for (int j = threadIdx.x; j < n2; j += RF_NT) {
        double2 z = buf[PADI(j)];
        if (bsi >= 0) z = cmul(z, __ldg(&chirp[j]));
        out[j] = z;
    }
}

template <bool SH>
__global__ void __launch_bounds__(RF_NT)
ring_synth_combine_kernel(PlanDev P, const SplitJob* __restrict__ jobs, const double2* __restrict__ scratch, double* __restrict__ mapQ,
                          double* __restrict__ mapU, const int* __restrict__ skip)
{
    if (skip && *skip) return;
    const SplitJob job = jobs[blockIdx.x];
    const int n = P.ring_nphi[job.ringA], n2 = n >> 2;
    const double2* Y = scratch + job.off;
    double* oa = (job.compA ? mapU : mapQ) + ring_first_pixel<SH>(P, job.ringA);
    double* ob = job.ringB >= 0 ? (job.compB ? mapU : mapQ) + ring_first_pixel<SH>(P, job.ringB) : nullptr;
    for (int j = threadIdx.x; j < n2; j += RF_NT) {
        const double2 t1 = unit_root(j, n), t2 = cmul(t1, t1), t3 = cmul(t2, t1);
        const double2 y0 = Y[j], y1 = cmul(Y[n2 + j], t1), y2 = cmul(Y[2 * n2 + j], t2), y3 = cmul(Y[3 * n2 + j], t3);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double2 z = cadd(cadd(y0, mul_ipow(y1, q)), cadd(mul_ipow(y2, 2 * q), mul_ipow(y3, 3 * q)));
            oa[j + q * n2] = z.x;
            if (ob) ob[j + q * n2] = z.y;
        }
    }
}

template <bool SH>
__global__ void __launch_bounds__(RF_NT, 2)
ring_anal_split_kernel(PlanDev P, const SplitJob* __restrict__ jobs, const double* __restrict__ mapQ, const double* __restrict__ mapU,
                       const double* __restrict__ pixw, double2* __restrict__ scratch, const int* __restrict__ skip)
{
    if (skip && *skip) return;
    extern __shared__ double2 smem[];
    const SplitJob job = jobs[blockIdx.x >> 2];
    const int b = blockIdx.x & 3;
    const int n = P.ring_nphi[job.ringA], n2 = n >> 2, bsi = job.bs2;
    double2* twq = smem;
    double2* buf = twq + (P.tw_n >> 2) + 1;
    load_twq(P, twq);
    const int64_t sa = ring_first_pixel<SH>(P, job.ringA), sb = job.ringB >= 0 ? ring_first_pixel<SH>(P, job.ringB) : 0;
    const double* ia = (job.compA ? mapU : mapQ) + sa;
    const double* ib = job.ringB >= 0 ? (job.compB ? mapU : mapQ) + sb : nullptr;
    const int lg = 31 - __clz(n2);
    const double2* chirp = bsi >= 0 ? P.bs_tab + P.bs[bsi].chirp_off : nullptr;
    const int M = bsi >= 0 ? P.bs[bsi].M : n2;
    for (int j = threadIdx.x; j < M; j += RF_NT) {
        const int pos = PADI(bsi < 0 ? (int)(__brev((unsigned)j) >> (32 - lg)) : j);
        if (j >= n2) { buf[pos] = make_double2(0.0, 0.0); continue; }
        double2 v = make_double2(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int jj = j + q * n2;
            double x = ia[jj], y = ib ? ib[jj] : 0.0;
            if (pixw) { x *= pixw[sa + jj]; if (ib) y *= pixw[sb + jj]; }
            v = cadd(v, mul_ipow(make_double2(x, -y), q * b));   // c_j = conj(z_j)
        }
        if (b) v = cmul(v, unit_root(j * b, n));
        if (bsi >= 0) v = cmul(v, __ldg(&chirp[j]));
        buf[pos] = v;
    }
    ring_idft(P, buf, twq, n2, bsi);
    double2* out = scratch + job.off + (int64_t)b * n2;   // C_{4a+b} at [b][a]
    for (int a = threadIdx.x; a < n2; a += RF_NT) {
        double2 z = buf[PADI(a)];
        if (bsi >= 0) z = cmul(z, __ldg(&chirp[a]));
        out[a] = z;
    }
}

template <bool SH>
__global__ void __launch_bounds__(RF_NT)
ring_anal_finish_kernel(PlanDev P, const SplitJob* __restrict__ jobs, const double2* __restrict__ scratch, double2* __restrict__ Fm,
                        const int* __restrict__ skip)
{
    if (skip && *skip) return;
    const SplitJob job = jobs[blockIdx.x];
    const int L = P.lmax, n = P.ring_nphi[job.ringA], n2 = n >> 2;
    const double2* Cs = scratch + job.off;
    for (int m = threadIdx.x; m <= L; m += RF_NT) {
        const int k = m % n, kk = (n - k) % n;
        const double2 c1 = Cs[(int64_t)(k & 3) * n2 + (k >> 2)], c2 = Cs[(int64_t)(kk & 3) * n2 + (kk >> 2)];
        const double2 z1 = make_double2(c1.x, -c1.y);  // Z[k]
        const double2 z2c = c2;                         // conj Z[n-k]
        const double2 xa = make_double2(0.5 * (z1.x + z2c.x), 0.5 * (z1.y + z2c.y));
        const double2 d = csub(z1, z2c);
        const double2 xb = make_double2(0.5 * d.y, -0.5 * d.x);
        Fm[fm_ring_index<SH>(P, job.compA, job.ringA, m)] = cmulc(xa, ring_phase(P, job.ringA, m));
        if (job.ringB >= 0) Fm[fm_ring_index<SH>(P, job.compB, job.ringB, m)] = cmulc(xb, ring_phase(P, job.ringB, m));
    }
}

// ------------------------------------------------------------------ host side
static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

static int ring_M(int n) { return (n & (n - 1)) == 0 ? n : next_pow2(2 * n - 1); }

int gs_ring_setup(gs_plan* p)
{
    const int nside = p->d.nside, nring = p->d.nring, npair = p->d.npair, L = p->d.lmax;
    const size_t smem_max = 227 * 1024;
    std::vector<int> rn(nring);
    for (int r = 0; r < nring; ++r) { int i = std::min(r + 1, 4 * nside - (r + 1)); rn[r] = i < nside ? 4 * i : 4 * nside; }
    // direct path: staging (L+1) + transform buffer + quarter twiddle table in one CTA; Mcap = largest
    // power-of-two transform length for which that fits.  Longer rings take the split path (n/4 per CTA).
    auto direct_smem = [&](int M, int twn) { return (size_t)((L + 1) + (M + M / 16 + 2) + twn / 4 + 1) * sizeof(double2); };
    int Mcap = 4;
    while (direct_smem(2 * Mcap, 2 * Mcap) <= smem_max) Mcap *= 2;
    if (const char* e = getenv("GS_RING_MCAP")) {  // test knob: exercise the split path at small nside
        const int v = atoi(e);
        if (v >= 4 && v < Mcap && (v & (v - 1)) == 0) Mcap = v;
    }
    auto is_split = [&](int ring) { return ring_M(rn[ring]) > Mcap; };
    std::map<int, int> n2bs;
    std::vector<BluesteinDesc> descs;
    int64_t off = 0;
    int maxMd = 4, maxMs = 4;
    auto need_len = [&](int n, int& maxM) {
        maxM = std::max(maxM, ring_M(n));
        if ((n & (n - 1)) == 0 || n2bs.count(n)) return;
        BluesteinDesc d;
        d.n = n; d.M = next_pow2(2 * n - 1); d.chirp_off = off; off += n; d.bhat_off = off; off += d.M;
        n2bs[n] = (int)descs.size();
        descs.push_back(d);
    };
    for (int r = 0; r < npair; ++r) {
        if (!is_split(r)) need_len(rn[r], maxMd);
        else {
            if (ring_M(rn[r] / 4) > Mcap) {
                gs_set_error("ring FFT: nside %d / lmax %d needs a deeper ring split than this build provides", nside, L);
                return GS_E_BADARG;
            }
            need_len(rn[r] / 4, maxMs);
        }
    }
    const int maxM = std::max(maxMd, maxMs);
    std::vector<int> rbs(nring);
    for (int r = 0; r < nring; ++r) rbs[r] = (!is_split(r) && n2bs.count(rn[r])) ? n2bs[rn[r]] : -1;
    p->d.max_M = maxM;
    p->d.tw_n = maxM;
    p->ring_smem = direct_smem(maxMd, maxM);
    p->split_smem = (size_t)((maxMs + maxMs / 16 + 2) + maxM / 4 + 1) * sizeof(double2);
    if (p->ring_smem > smem_max || p->split_smem > smem_max) {
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
    GS_CHECK_CUDA(cudaFuncSetAttribute(ring_synth_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    GS_CHECK_CUDA(cudaFuncSetAttribute(ring_anal_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    GS_CHECK_CUDA(cudaFuncSetAttribute(ring_synth_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    GS_CHECK_CUDA(cudaFuncSetAttribute(ring_anal_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    GS_CHECK_CUDA(cudaFuncSetAttribute(ring_apply_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    GS_CHECK_CUDA(cudaFuncSetAttribute(ring_apply_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    GS_CHECK_CUDA(cudaFuncSetAttribute(ring_synth_split_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    GS_CHECK_CUDA(cudaFuncSetAttribute(ring_anal_split_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    GS_CHECK_CUDA(cudaFuncSetAttribute(ring_synth_split_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    GS_CHECK_CUDA(cudaFuncSetAttribute(ring_anal_split_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (!descs.empty()) {
        const size_t bsm = (size_t)((maxM + maxM / 16 + 2) + maxM / 4 + 1) * sizeof(double2);
        bluestein_setup_kernel<<<(int)descs.size(), RF_NT, bsm>>>(p->d, (double2*)d, (int)descs.size());
        GS_CHECK_LAUNCH();
    }

    // job lists, heaviest transforms first; sharded plans transform only the rings of the owned ring pairs
    auto cost = [&](int ring) { int n = rn[ring]; return rbs[ring] < 0 ? n : 3 * descs[rbs[ring]].M; };
    auto owned = [&](int ring) { return p->world <= 1 || std::min(ring, nring - 1 - ring) % p->world == p->rank; };
    std::vector<RingJob> j2, j0;
    std::vector<SplitJob> s2, s0;
    int64_t so2 = 0, so0 = 0;
    auto bs2_of = [&](int ring) { const int n2 = rn[ring] / 4; return n2bs.count(n2) ? n2bs[n2] : -1; };
    for (int r = 0; r < nring; ++r) {
        if (!owned(r)) continue;
        if (!is_split(r)) j2.push_back(RingJob{r, 0, r, 1});
        else { s2.push_back(SplitJob{r, 0, r, 1, bs2_of(r), 0, so2}); so2 += rn[r]; }
    }
    for (int r = 0; r < npair; ++r) {
        const int rs = nring - 1 - r;
        if (!owned(r)) continue;
        if (!is_split(r)) j0.push_back(RingJob{r, 0, rs != r ? rs : -1, 0});
        else { s0.push_back(SplitJob{r, 0, rs != r ? rs : -1, 0, bs2_of(r), 0, so0}); so0 += rn[r]; }
    }
    std::stable_sort(j2.begin(), j2.end(), [&](const RingJob& a, const RingJob& b) { return cost(a.ringA) > cost(b.ringA); });
    std::stable_sort(j0.begin(), j0.end(), [&](const RingJob& a, const RingJob& b) { return cost(a.ringA) > cost(b.ringA); });
    auto put = [&](const void* h, size_t bytes, void** out) -> int {
        void* q = nullptr;
        GS_CHECK_CUDA(cudaMalloc(&q, std::max<size_t>(16, bytes)));
        p->owned.push_back(q);
        if (bytes) GS_CHECK_CUDA(cudaMemcpy(q, h, bytes, cudaMemcpyHostToDevice));
        *out = q;
        return GS_OK;
    };
    int rc;
    if ((rc = put(j2.data(), j2.size() * sizeof(RingJob), (void**)&p->jobs2))) return rc;
    if ((rc = put(j0.data(), j0.size() * sizeof(RingJob), (void**)&p->jobs0))) return rc;
    if ((rc = put(s2.data(), s2.size() * sizeof(SplitJob), (void**)&p->sjobs2))) return rc;
    if ((rc = put(s0.data(), s0.size() * sizeof(SplitJob), (void**)&p->sjobs0))) return rc;
    p->njobs2 = (int)j2.size(); p->njobs0 = (int)j0.size();
    p->nsjobs2 = (int)s2.size(); p->nsjobs0 = (int)s0.size();
    if (so2 || so0) {
        GS_CHECK_CUDA(cudaMalloc(&d, (size_t)std::max(so2, so0) * sizeof(double2)));
        p->owned.push_back(d);
        p->ring_scratch = (double2*)d;
    }
    return GS_OK;
}

int gs_ring_synth(gs_plan* p, int spin, double* mapQ, double* mapU, cudaStream_t st, const int* skip)
{
    const int nj = spin == 0 ? p->njobs0 : p->njobs2, ns = spin == 0 ? p->nsjobs0 : p->nsjobs2;
    const RingJob* jobs = spin == 0 ? p->jobs0 : p->jobs2;
    const SplitJob* sj = spin == 0 ? p->sjobs0 : p->sjobs2;
    double* mu = spin == 0 ? mapQ : mapU;
    const bool sh = p->world > 1;
    const double2* F = sh ? p->Fx : p->Fm;
    if (ns > 0) {  // long rings first: they are the heavy ones
        if (sh) {
            ring_synth_split_kernel<true><<<4 * ns, RF_NT, p->split_smem, st>>>(p->d, sj, F, p->ring_scratch, skip);
            ring_synth_combine_kernel<true><<<ns, RF_NT, 0, st>>>(p->d, sj, p->ring_scratch, mapQ, mu, skip);
        } else {
            ring_synth_split_kernel<false><<<4 * ns, RF_NT, p->split_smem, st>>>(p->d, sj, F, p->ring_scratch, skip);
            ring_synth_combine_kernel<false><<<ns, RF_NT, 0, st>>>(p->d, sj, p->ring_scratch, mapQ, mu, skip);
        }
        GS_CHECK_LAUNCH();
        g_gs_launches += 2;
    }
    if (nj > 0) {
        if (sh) ring_synth_kernel<true><<<nj, RF_NT, p->ring_smem, st>>>(p->d, jobs, F, mapQ, mu, skip, 0, 0, nullptr);
        else ring_synth_kernel<false><<<nj, RF_NT, p->ring_smem, st>>>(p->d, jobs, F, mapQ, mu, skip, 0, 0, nullptr);
        GS_CHECK_LAUNCH();
    }
    g_gs_launches += 1;
    return GS_OK;
}

int gs_ring_anal(gs_plan* p, int spin, const double* mapQ, const double* mapU, const double* pixw, cudaStream_t st,
                 const int* skip)
{
    const int nj = spin == 0 ? p->njobs0 : p->njobs2, ns = spin == 0 ? p->nsjobs0 : p->nsjobs2;
    const RingJob* jobs = spin == 0 ? p->jobs0 : p->jobs2;
    const SplitJob* sj = spin == 0 ? p->sjobs0 : p->sjobs2;
    const double* mu = spin == 0 ? mapQ : mapU;
    const bool sh = p->world > 1;
    double2* F = sh ? p->Fx : p->Fm;
    if (ns > 0) {
        if (sh) {
            ring_anal_split_kernel<true><<<4 * ns, RF_NT, p->split_smem, st>>>(p->d, sj, mapQ, mu, pixw, p->ring_scratch, skip);
            ring_anal_finish_kernel<true><<<ns, RF_NT, 0, st>>>(p->d, sj, p->ring_scratch, F, skip);
        } else {
            ring_anal_split_kernel<false><<<4 * ns, RF_NT, p->split_smem, st>>>(p->d, sj, mapQ, mu, pixw, p->ring_scratch, skip);
            ring_anal_finish_kernel<false><<<ns, RF_NT, 0, st>>>(p->d, sj, p->ring_scratch, F, skip);
        }
        GS_CHECK_LAUNCH();
        g_gs_launches += 2;
    }
    if (nj > 0) {
        if (sh) ring_anal_kernel<true><<<nj, RF_NT, p->ring_smem, st>>>(p->d, jobs, mapQ, mu, pixw, F, skip);
        else ring_anal_kernel<false><<<nj, RF_NT, p->ring_smem, st>>>(p->d, jobs, mapQ, mu, pixw, F, skip);
        GS_CHECK_LAUNCH();
    }
    g_gs_launches += 1;
    return GS_OK;
}

// ring stage of A^T diag(pixw) A on the ring spectra in place; falls back to synthesis + weighted analysis through the
// plan's scratch maps when some rings take the split path
int gs_ring_apply(gs_plan* p, int spin, const double* pixw, cudaStream_t st, const int* skip)
{
    const int nj = spin == 0 ? p->njobs0 : p->njobs2, ns = spin == 0 ? p->nsjobs0 : p->nsjobs2;
    if (ns > 0 || !pixw) {
        int rc = gs_ring_synth(p, spin, p->mapQ_tmp, p->mapU_tmp, st, skip);
        if (rc) return rc;
        return gs_ring_anal(p, spin, p->mapQ_tmp, p->mapU_tmp, pixw, st, skip);
    }
    if (nj <= 0) return GS_OK;
    const RingJob* jobs = spin == 0 ? p->jobs0 : p->jobs2;
    if (p->world > 1) ring_apply_kernel<true><<<nj, RF_NT, p->ring_smem, st>>>(p->d, jobs, p->Fx, pixw, skip);
    else ring_apply_kernel<false><<<nj, RF_NT, p->ring_smem, st>>>(p->d, jobs, p->Fm, pixw, skip);
    GS_CHECK_LAUNCH();
    g_gs_launches += 1;
    return GS_OK;
}

// nb spin-2 ring syntheses in one launch: spectra F + k f_stride (m <= mmax[k]) -> maps Q/U + k map_stride
int gs_ring_synth_batch(gs_plan* p, const double2* F, int64_t f_stride, const int* mmax, double* mapQ, double* mapU,
                        int64_t map_stride, int nb, cudaStream_t st)
{
    if (p->world > 1 || p->nsjobs2 > 0) { gs_set_error("batched ring synthesis needs an unsharded plan without split rings"); return GS_E_BADARG; }
    if (nb <= 0 || p->njobs2 <= 0) return GS_OK;
    ring_synth_kernel<false><<<dim3(p->njobs2, nb), RF_NT, p->ring_smem, st>>>(p->d, p->jobs2, F, mapQ, mapU, nullptr, f_stride, map_stride, mmax);
    GS_CHECK_LAUNCH();
    g_gs_launches += 1;
    return GS_OK;
}
