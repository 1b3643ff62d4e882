/* oracle/sht_fast_hot.h -- hot loops of sht_fast.c (TEST / BENCH INFRASTRUCTURE ONLY).  Included three times under
 * different `#pragma GCC target` settings with SUF = _avx512 / _avx2 / _base; sht_fast.c picks one set at load time
 * from the host's CPUID, so the shared object is not tied to the machine it was built on. */
#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
#define FN(name) CAT(name, SUF)

/* ------------------------------------------------------------------ FFT (4 sequences per vector) */
/* Stockham autosort, forward sign: out[k] = sum_j x[j] exp(-2 pi i j k / n).  Radix-4 passes, one radix-2 pass
 * when log2(n) is odd.  Input in (xr, xi), scratch (yr, yi); the result buffer is returned in outr/outi. */
static void FN(fft_pow2)(int n, v4d *xr, v4d *xi, v4d *yr, v4d *yi, const double *twr, const double *twi, int twn,
                         v4d **outr, v4d **outi)
{
    int Ns = 1;
    int lg = 0; while ((1 << lg) < n) ++lg;
    if (lg & 1) {   /* radix-2 pass */
        const int h = n / 2;
        for (int j = 0; j < h; ++j) {
            v4d ar = xr[j], ai = xi[j], br = xr[j + h], bi = xi[j + h];
            yr[2 * j] = ar + br; yi[2 * j] = ai + bi;
            yr[2 * j + 1] = ar - br; yi[2 * j + 1] = ai - bi;
        }
        { v4d *t = xr; xr = yr; yr = t; t = xi; xi = yi; yi = t; }
        Ns = 2;
    }
    const int q4 = n / 4;
    for (; Ns < n; Ns *= 4) {
        const int tstep = twn / (4 * Ns);
        for (int j = 0; j < q4; ++j) {
            const int k = j & (Ns - 1);
            const int j0 = ((j - k) << 2) + k;
            const int t1 = k * tstep;
            const double w1r = twr[t1], w1i = twi[t1], w2r = twr[2 * t1], w2i = twi[2 * t1], w3r = twr[3 * t1], w3i = twi[3 * t1];
            v4d a0r = xr[j], a0i = xi[j];
            v4d b1r = xr[j + q4], b1i = xi[j + q4], b2r = xr[j + 2 * q4], b2i = xi[j + 2 * q4], b3r = xr[j + 3 * q4], b3i = xi[j + 3 * q4];
            v4d a1r = b1r * w1r - b1i * w1i, a1i = b1r * w1i + b1i * w1r;
            v4d a2r = b2r * w2r - b2i * w2i, a2i = b2r * w2i + b2i * w2r;
            v4d a3r = b3r * w3r - b3i * w3i, a3i = b3r * w3i + b3i * w3r;
            v4d s02r = a0r + a2r, s02i = a0i + a2i, d02r = a0r - a2r, d02i = a0i - a2i;
            v4d s13r = a1r + a3r, s13i = a1i + a3i, d13r = a1r - a3r, d13i = a1i - a3i;
            /* -i (a1 - a3) = (d13i, -d13r) */
            yr[j0] = s02r + s13r;          yi[j0] = s02i + s13i;
            yr[j0 + Ns] = d02r + d13i;     yi[j0 + Ns] = d02i - d13r;
            yr[j0 + 2 * Ns] = s02r - s13r; yi[j0 + 2 * Ns] = s02i - s13i;
            yr[j0 + 3 * Ns] = d02r - d13i; yi[j0 + 3 * Ns] = d02i + d13r;
        }
        { v4d *t = xr; xr = yr; yr = t; t = xi; xi = yi; yi = t; }
    }
    *outr = xr; *outi = xi;
}

/* forward DFT of length b->n of 4 sequences held in (xr, xi)[0..n); work arrays of 2 x max(n, M) v4d each.
 * Result pointer returned (somewhere inside the work arrays). */
static void FN(ring_fft)(const fplan *p, const blue_t *b, v4d *xr, v4d *xi, v4d *yr, v4d *yi, v4d **outr, v4d **outi)
{
    const int n = b->n, M = b->M;
    if (!M) { FN(fft_pow2)(n, xr, xi, yr, yi, p->twr, p->twi, p->twn, outr, outi); return; }
    for (int j = 0; j < n; ++j) {
        v4d ar = xr[j], ai = xi[j];
        const double cr = b->cr[j], ci = b->ci[j];
        xr[j] = ar * cr - ai * ci; xi[j] = ar * ci + ai * cr;
    }
    for (int j = n; j < M; ++j) { xr[j] = (v4d){0, 0, 0, 0}; xi[j] = (v4d){0, 0, 0, 0}; }
    v4d *fr, *fi;
    FN(fft_pow2)(M, xr, xi, yr, yi, p->twr, p->twi, p->twn, &fr, &fi);
    v4d *gr = (fr == xr) ? yr : xr, *gi = (fi == xi) ? yi : xi;
    /* conj(A * Bhat): the inverse transform is conj(FFT(conj(.))) (the 1/M is folded into Bhat) */
    for (int j = 0; j < M; ++j) {
        v4d ar = fr[j], ai = fi[j];
        const double br = b->br[j], bi = b->bi[j];
        gr[j] = ar * br - ai * bi; gi[j] = -(ar * bi + ai * br);
    }
    v4d *hr, *hi;
    FN(fft_pow2)(M, gr, gi, fr, fi, p->twr, p->twi, p->twn, &hr, &hi);
    v4d *orr = (hr == gr) ? fr : gr, *oii = (hi == gi) ? fi : gi;
    for (int k = 0; k < n; ++k) {
        v4d ar = hr[k], ai = -hi[k];
        const double cr = b->cr[k], ci = b->ci[k];
        orr[k] = ar * cr - ai * ci; oii[k] = ar * ci + ai * cr;
    }
    *outr = orr; *outi = oii;
}

/* ------------------------------------------------------------------ Legendre stage */

/* Range extension: a lane whose seed underflows carries an integer scale (true value = v * 2^(-SC_K * scale)) and
 * stays out of the sums (weight 0) until the scale reaches zero.  The check runs every CHK multipoles in scalar
 * code: between checks a scaled value grows by less than 2^(6 * CHK), far inside the exponent range, and a lane
 * that crosses the threshold inside a block joins the sums at most CHK - 1 multipoles late with terms below
 * 2^-500 of the result. */
static inline __attribute__((always_inline)) int FN(rescale_lanes)(v8d *pc, v8d *pp, v8d *mc, v8d *mp, long long *sc, v8d *w)
{
    const double BIG = 0x1p-644, DOWN = 0x1p-256;
    int pending = 0;
    for (int i = 0; i < VL; ++i) {
        while (sc[i] > 0 && (fabs((*pc)[i]) > BIG || fabs((*mc)[i]) > BIG)) {
            (*pc)[i] *= DOWN; (*pp)[i] *= DOWN; (*mc)[i] *= DOWN; (*mp)[i] *= DOWN;
            --sc[i];
        }
        (*w)[i] = sc[i] == 0 ? 1.0 : 0.0;
        pending |= sc[i] != 0;
    }
    return pending;
}

/* ring stage of one pair, synthesis: alias-fold the records of the pair (phases cr + i ci = exp(i m phi0)) into
 * 4 spectra (QN, UN, QS, US), one DFT, real parts to the maps */
static void FN(ring_synth_pair)(const fplan *p, int q, const double *cr, const double *ci, v4d *w, size_t wlen,
                                double *mapQ, double *mapU)
{
    const int L = p->lmax;
    const int n = p->nphi[q];
    v4d *xr = w, *xi = w + wlen, *yr = w + 2 * wlen, *yi = w + 3 * wlen;
    for (int k = 0; k < n; ++k) { xr[k] = (v4d){0, 0, 0, 0}; xi[k] = (v4d){0, 0, 0, 0}; }
    for (int m = 0, k = 0; m <= L; ++m, k = (k + 1 == n) ? 0 : k + 1) {
        const double *f = p->F + ((size_t)q * (L + 1) + m) * 8;
        v4d fr = {f[0], f[2], f[4], f[6]}, fi = {f[1], f[3], f[5], f[7]};
        const double wgt = m ? 2.0 : 1.0, pr = wgt * cr[m], pi = wgt * ci[m];
        /* conj(z): map = Re(sum_k z_k e^{+2 pi i jk/n}) = Re(FFT_fwd(conj z)) */
        xr[k] += fr * pr - fi * pi;
        xi[k] -= fr * pi + fi * pr;
    }
    v4d *orr, *oii;
    FN(ring_fft)(p, &p->blue[q], xr, xi, yr, yi, &orr, &oii);
    double *qn = mapQ + p->startN[q], *un = mapU + p->startN[q];
    for (int j = 0; j < n; ++j) { qn[j] = orr[j][0]; un[j] = orr[j][1]; }
    if (q != p->npair - 1) {
        double *qs = mapQ + p->startS[q], *us = mapU + p->startS[q];
        for (int j = 0; j < n; ++j) { qs[j] = orr[j][2]; us[j] = orr[j][3]; }
    }
}

/* ring stage of one pair, analysis: DFT of the 4 pixel sequences, phase-shifted bins to the records of the pair */
static void FN(ring_anal_pair)(const fplan *p, int q, const double *cr, const double *ci, v4d *w, size_t wlen,
                               const double *mapQ, const double *mapU, const double *pixw, double weight)
{
    const int L = p->lmax;
    const int n = p->nphi[q];
    v4d *xr = w, *xi = w + wlen, *yr = w + 2 * wlen, *yi = w + 3 * wlen;
    const double *qn = mapQ + p->startN[q], *un = mapU + p->startN[q];
    const double *qs = mapQ + p->startS[q], *us = mapU + p->startS[q];
    const int eq = (q == p->npair - 1);
    if (pixw) {   /* per-pixel weight (N^-1) applied while the ring is read */
        const double *wn = pixw + p->startN[q], *ws = pixw + p->startS[q];
        for (int j = 0; j < n; ++j) {
            xr[j] = (v4d){qn[j] * wn[j], un[j] * wn[j], eq ? 0.0 : qs[j] * ws[j], eq ? 0.0 : us[j] * ws[j]};
            xi[j] = (v4d){0, 0, 0, 0};
        }
    } else {
        for (int j = 0; j < n; ++j) {
            xr[j] = (v4d){qn[j], un[j], eq ? 0.0 : qs[j], eq ? 0.0 : us[j]};
            xi[j] = (v4d){0, 0, 0, 0};
        }
    }
    v4d *orr, *oii;
    FN(ring_fft)(p, &p->blue[q], xr, xi, yr, yi, &orr, &oii);
    for (int m = 0, k = 0; m <= L; ++m, k = (k + 1 == n) ? 0 : k + 1) {
        v4d gr = weight * (orr[k] * cr[m] + oii[k] * ci[m]), gi = weight * (oii[k] * cr[m] - orr[k] * ci[m]);
        double *f = p->F + ((size_t)q * (L + 1) + m) * 8;
        f[0] = gr[0]; f[1] = gi[0]; f[2] = gr[1]; f[3] = gi[1];
        f[4] = gr[2]; f[5] = gi[2]; f[6] = gr[3]; f[7] = gi[3];
    }
}

/* DFMA throughput of one core with this vector ISA: 16 independent chains v = v * a + b, `iters` rounds; returns the
 * sum of the chains so the loop cannot be removed.  flops = iters * 16 * 8 * 2. */
static double FN(fma_peak_loop)(long iters, double a0, double b0)
{
    v8d a = {a0, a0, a0, a0, a0, a0, a0, a0}, b = {b0, b0, b0, b0, b0, b0, b0, b0};
    v8d v0 = b, v1 = b + 1.0, v2 = b + 2.0, v3 = b + 3.0, v4 = b + 4.0, v5 = b + 5.0, v6 = b + 6.0, v7 = b + 7.0;
    v8d v8 = b + 8.0, v9 = b + 9.0, v10 = b + 10.0, v11 = b + 11.0, v12 = b + 12.0, v13 = b + 13.0, v14 = b + 14.0, v15 = b + 15.0;
    for (long i = 0; i < iters; ++i) {
        v0 = v0 * a + b; v1 = v1 * a + b; v2 = v2 * a + b; v3 = v3 * a + b; v4 = v4 * a + b; v5 = v5 * a + b; v6 = v6 * a + b; v7 = v7 * a + b;
        v8 = v8 * a + b; v9 = v9 * a + b; v10 = v10 * a + b; v11 = v11 * a + b; v12 = v12 * a + b; v13 = v13 * a + b; v14 = v14 * a + b; v15 = v15 * a + b;
    }
    v8d s = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7 + v8 + v9 + v10 + v11 + v12 + v13 + v14 + v15;
    double r = 0;
    for (int i = 0; i < VL; ++i) r += s[i];
    return r;
}

#define x_of(p, q0) (*(const v8d *)((p)->cth + (q0)))

/* initial state of one vector of ring pairs for multipole m */
static inline __attribute__((always_inline)) int FN(seed_vec)(const fplan *p, int m, int q0, v8d *pc, v8d *mc, long long *sc, v8d *w)
{
    const size_t o = (size_t)m * p->npad + q0;
    int pending = 0;
    for (int i = 0; i < VL; ++i) {
        int ex = p->seede[o + i], s = 0;
        if (ex < SC_LO) { s = (SC_LO - ex + SC_K - 1) / SC_K; ex += SC_K * s; }
        (*pc)[i] = ldexp(p->seedp[o + i], ex);
        (*mc)[i] = ldexp(p->seedm[o + i], ex);
        sc[i] = s;
        (*w)[i] = s == 0 ? 1.0 : 0.0;
        pending |= s != 0;
    }
    return pending;
}

/* Ring spectra live in one 64-byte record per (ring pair, m):
 *   F[(pair * (L+1) + m) * 8 + c],  c = QN.re QN.im UN.re UN.im QS.re QS.im US.re US.im,
 * so the ring stage of a pair streams through (L+1) consecutive cache lines; the Legendre stage transposes its
 * 8 x 8 block of (lane, c) values when it stores / loads them. */

/* synthesis for one m: coefficient stream cf[l - l0] (alpha, -1/2, parity signs folded in), writes the records of m */
static void FN(leg_synth_m)(const fplan *p, int m, const coef_t *cf)
{
    const int L = p->lmax, npad = p->npad, l0 = m > 2 ? m : 2;
    const int q_first = p->pmin[m] & ~(VL - 1);
    const size_t ps = (size_t)(L + 1) * 8;   /* doubles per ring pair */
    for (int q = 0; q < q_first && q < npad; ++q) memset(p->F + (size_t)q * ps + (size_t)m * 8, 0, sizeof(double) * 8);
    for (int q0 = q_first; q0 < npad; q0 += VL) {
        v8d x = *(const v8d *)(p->cth + q0);
        v8d pc, mc, pp = {0}, mp = {0}, w;
        long long sc[VL];
        int pending = FN(seed_vec)(p, m, q0, &pc, &mc, sc, &w);
        v8d T0 = {0}, T1 = {0}, T2 = {0}, T3 = {0}, T4 = {0}, T5 = {0}, T6 = {0}, T7 = {0};
        int l = l0;
        /* phase 1: some lanes still carry a scale */
        while (pending && l <= L) {
            const int lend = l + CHK <= L + 1 ? l + CHK : L + 1;
            int active = 0;
            for (int i = 0; i < VL; ++i) active |= sc[i] == 0;
            if (active) {
                for (; l < lend; ++l) {
                    const coef_t *c = &cf[l - l0];
                    v8d lp = pc * w, lm = mc * w;
                    T0 += c->c[0] * lp; T1 += c->c[1] * lp; T2 += c->c[2] * lm; T3 += c->c[3] * lm;
                    T4 += c->c[4] * lm; T5 += c->c[5] * lm; T6 += c->c[6] * lp; T7 += c->c[7] * lp;
                    v8d np = (c->a * x + c->b) * pc - pp, nm = (c->a * x - c->b) * mc - mp;
                    pp = pc; pc = np; mp = mc; mc = nm;
                }
            } else {
                for (; l < lend; ++l) {
                    const coef_t *c = &cf[l - l0];
                    v8d np = (c->a * x + c->b) * pc - pp, nm = (c->a * x - c->b) * mc - mp;
                    pp = pc; pc = np; mp = mc; mc = nm;
                }
            }
            pending = FN(rescale_lanes)(&pc, &pp, &mc, &mp, sc, &w);
        }
        /* phase 2: every lane is in range */
        for (; l <= L; ++l) {
            const coef_t *c = &cf[l - l0];
            T0 += c->c[0] * pc; T1 += c->c[1] * pc; T2 += c->c[2] * mc; T3 += c->c[3] * mc;
            T4 += c->c[4] * mc; T5 += c->c[5] * mc; T6 += c->c[6] * pc; T7 += c->c[7] * pc;
            v8d np = (c->a * x + c->b) * pc - pp, nm = (c->a * x - c->b) * mc - mp;
            pp = pc; pc = np; mp = mc; mc = nm;
        }
        /* T0/1 = sum c+ lam+, T2/3 = sum c- lam-, T4/5 = sum s c+ lam-, T6/7 = sum s c- lam+  (s = (-1)^(l+m))
         * north: Q = Sp + Sm, U = -i (Sp - Sm) with Sp = T01, Sm = T23; south: Sp = T45, Sm = T67 */
        v8d rec[8];
        rec[0] = T0 + T2; rec[1] = T1 + T3; rec[2] = T1 - T3; rec[3] = T2 - T0;
        rec[4] = T4 + T6; rec[5] = T5 + T7; rec[6] = T5 - T7; rec[7] = T6 - T4;
        double *dst = p->F + (size_t)q0 * ps + (size_t)m * 8;
        for (int i = 0; i < VL; ++i)
            for (int c = 0; c < 8; ++c) dst[(size_t)i * ps + c] = rec[c][i];
    }
}

/* analysis for one m: reads the records of m (ring spectra of the maps) and accumulates
 *   X_l = sum_pairs lam+ W+_N + s lam- W+_S,  Y_l = sum_pairs lam- W-_N + s lam+ W-_S   (W+- = Q +- iU)
 * into acc[(l - l0) * 4 + {X.re, X.im, Y.re, Y.im}] (vectors; the caller sums the lanes).  The l range is walked in
 * tiles of TL multipoles for ALL vectors of ring pairs, so the accumulators of a tile stay in L1; the recurrence state
 * of every vector is parked in `st` (13 vectors each) between tiles. */
#define TL 64
static void FN(leg_anal_m)(const fplan *p, int m, const double *rab /* (a,b) per l */, v8d *acc, v8d *st, int *lst)
{
    const int L = p->lmax, npad = p->npad, l0 = m > 2 ? m : 2;
    const int nl = L - l0 + 1;
    for (int i = 0; i < 4 * nl; ++i) acc[i] = (v8d){0};
    const int q_first = p->pmin[m] & ~(VL - 1);
    const int nvec = (npad - q_first) / VL;
    /* pass 1: seeds, W, and the (rare) scaled start-up phase of every vector */
    for (int v = 0; v < nvec; ++v) {
        const int q0 = q_first + v * VL;
        v8d pc, mc, pp = {0}, mp = {0}, w;
        long long sc[VL];
        int pending = FN(seed_vec)(p, m, q0, &pc, &mc, sc, &w);
        v8d rec[8];
        {
            const size_t ps = (size_t)(L + 1) * 8;
            const double *src = p->F + (size_t)q0 * ps + (size_t)m * 8;
            for (int i = 0; i < VL; ++i)
                for (int c = 0; c < 8; ++c) rec[c][i] = src[(size_t)i * ps + c];
        }
        v8d qnr = rec[0], qni = rec[1], unr = rec[2], uni = rec[3], qsr = rec[4], qsi = rec[5], usr = rec[6], usi = rec[7];
        v8d wpnr = qnr - uni, wpni = qni + unr, wmnr = qnr + uni, wmni = qni - unr;
        v8d wpsr = qsr - usi, wpsi = qsi + usr, wmsr = qsr + usi, wmsi = qsi - usr;
        int l = l0;
        while (pending && l <= L) {
            const int lend = l + CHK <= L + 1 ? l + CHK : L + 1;
            int active = 0;
            for (int i = 0; i < VL; ++i) active |= sc[i] == 0;
            for (; l < lend; ++l) {
                if (active) {
                    const double s = ((l + m) & 1) ? -1.0 : 1.0;
                    v8d lp = pc * w, lm = (mc * w) * s;
                    v8d lps = lp * s, lmn = mc * w;
                    v8d *a = acc + 4 * (l - l0);
                    a[0] += lp * wpnr + lm * wpsr; a[1] += lp * wpni + lm * wpsi;
                    a[2] += lmn * wmnr + lps * wmsr; a[3] += lmn * wmni + lps * wmsi;
                }
                const double ca = rab[2 * (l - l0)], cb = rab[2 * (l - l0) + 1];
                v8d np = (ca * x_of(p, q0) + cb) * pc - pp, nm = (ca * x_of(p, q0) - cb) * mc - mp;
                pp = pc; pc = np; mp = mc; mc = nm;
            }
            pending = FN(rescale_lanes)(&pc, &pp, &mc, &mp, sc, &w);
        }
        v8d *s = st + 12 * (size_t)v;
        s[0] = pc; s[1] = pp; s[2] = mc; s[3] = mp;
        s[4] = wpnr; s[5] = wpni; s[6] = wmnr; s[7] = wmni; s[8] = wpsr; s[9] = wpsi; s[10] = wmsr; s[11] = wmsi;
        lst[v] = l;   /* first multipole of the unscaled phase (L + 1: the vector never contributes) */
    }
    /* pass 2: tiles of TL multipoles */
    for (int t0 = l0; t0 <= L; t0 += TL) {
        const int t1 = t0 + TL <= L + 1 ? t0 + TL : L + 1;
        for (int v = 0; v < nvec; ++v) {
            int l = lst[v] > t0 ? lst[v] : t0;
            if (l >= t1) continue;
            const v8d x = x_of(p, q_first + v * VL);
            v8d *s = st + 12 * (size_t)v;
            v8d pc = s[0], pp = s[1], mc = s[2], mp = s[3];
            const v8d wpnr = s[4], wpni = s[5], wmnr = s[6], wmni = s[7];
            v8d wpsr = s[8], wpsi = s[9], wmsr = s[10], wmsi = s[11];
            if ((l + m) & 1) { wpsr = -wpsr; wpsi = -wpsi; wmsr = -wmsr; wmsi = -wmsi; }   /* sign of the south terms at l */
            for (; l + 1 < t1; l += 2) {
                v8d *a = acc + 4 * (l - l0);
                const double ca = rab[2 * (l - l0)], cb = rab[2 * (l - l0) + 1];
                const double da = rab[2 * (l - l0) + 2], db = rab[2 * (l - l0) + 3];
                a[0] += pc * wpnr + mc * wpsr; a[1] += pc * wpni + mc * wpsi;
                a[2] += mc * wmnr + pc * wmsr; a[3] += mc * wmni + pc * wmsi;
                v8d np = (ca * x + cb) * pc - pp, nm = (ca * x - cb) * mc - mp;
                a[4] += np * wpnr - nm * wpsr; a[5] += np * wpni - nm * wpsi;
                a[6] += nm * wmnr - np * wmsr; a[7] += nm * wmni - np * wmsi;
                pp = np; mp = nm;
                np = (da * x + db) * np - pc; nm = (da * x - db) * nm - mc;
                pc = np; mc = nm;
            }
            if (l < t1) {
                v8d *a = acc + 4 * (l - l0);
                const double ca = rab[2 * (l - l0)], cb = rab[2 * (l - l0) + 1];
                a[0] += pc * wpnr + mc * wpsr; a[1] += pc * wpni + mc * wpsi;
                a[2] += mc * wmnr + pc * wmsr; a[3] += mc * wmni + pc * wmsi;
                v8d np = (ca * x + cb) * pc - pp, nm = (ca * x - cb) * mc - mp;
                pp = pc; pc = np; mp = mc; mc = nm;
            }
            s[0] = pc; s[1] = pp; s[2] = mc; s[3] = mp;
        }
    }
}
#undef TL

#undef FN
#undef CAT
#undef CAT2
