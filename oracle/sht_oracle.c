/*
 * oracle/sht_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the third-party arithmetic the reference's hot path executes
 * through healpy (not vendored in /root/reference; no version pin -- SURVEY.md 8c):
 *   hp.alm2map / hp.alm2map(pol=True)        call sites CenteredGibbs.py:505-508,698-699
 *                                            NonCenteredGibbs.py:350-351
 *   hp.map2alm(iter=0, use_weights=False)    call sites CenteredGibbs.py:298-299,513,717-719
 *                                            utils.py:89,104 (iter=3 is built in Python on top)
 * following the published HEALPix definitions (Gorski et al. 2005; HEALPix primer):
 *   RING pixelisation geometry, Condon-Shortley Y_lm,
 *   Q +- iU = sum_lm -(E_lm +- i B_lm) (+-2)Y_lm,
 *   map2alm(iter=0): a_lm = (4 pi / Npix) sum_p Y*_lm(p) f(p).
 * PARITY UNPINNED against healpy itself (healpy is not installable here); the
 * restatement is pinned instead by analytic known answers, scipy's sph_harm_y and
 * sympy's Wigner-d (tests/test_oracle_sht.py).
 *
 * Algorithm: plain three-term recurrence in l of the normalised Wigner-d functions
 *   lam^{m'}_{lm}(theta) = sqrt((2l+1)/4pi) d^l_{m,m'}(theta),  m' in {0,+2,-2},
 * evaluated in ORACLE_REAL (long double by default: 64-bit mantissa and 15-bit
 * exponent, so it is both more accurate and wider-ranged than the FP64 device code),
 * north/south ring pairs share one recurrence through the parity of d^l, ring
 * Fourier sums by a recursive mixed-radix FFT. The same file compiled with
 * -DORACLE_REAL=double is the "port" CPU baseline timed by bench.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef ORACLE_REAL
#define ORACLE_REAL long double
#endif
typedef ORACLE_REAL real;

#ifdef ORACLE_IS_DOUBLE
#define R_SQRT sqrt
#define R_FABS fabs
#define R_LDEXP ldexp
#define R_FREXP frexp
#define SC_LO (-900)
#else
#define R_SQRT sqrtl
#define R_FABS fabsl
#define R_LDEXP ldexpl
#define R_FREXP frexpl
#define SC_LO (-16000)
#endif
#define SC_K 256

static const long double PI_L = 3.14159265358979323846264338327950288L;

/* ---------------------------------------------------------------- geometry */
/* HEALPix RING scheme, ring index i = 1 .. 4*nside-1 from the north pole. */
typedef struct { long double z, sth, phi0; int nphi; int64_t start; } ring_t;

static void ring_geom(int nside, int i, ring_t *r)
{
    int64_t ns = nside, npix = 12 * ns * ns, ncap = 2 * ns * (ns - 1);
    int north = i, south = 0;
    if (i > 2 * nside) { north = 4 * nside - i; south = 1; }
    long double z, omz; /* omz = 1 - |z| computed without cancellation */
    if (north < nside) {
        omz = (long double)north * north / (3.0L * ns * ns);
        z = 1.0L - omz;
        r->nphi = 4 * north;
        r->phi0 = PI_L / (4.0L * north);
        r->start = 2 * (int64_t)north * (north - 1);
        r->sth = sqrtl(omz * (1.0L + z));
    } else {
        z = 4.0L / 3.0L - 2.0L * north / (3.0L * ns);
        r->nphi = 4 * nside;
        /* HEALPix pix2ang_ring: phi = (j - fodd) pi / (2 nside), j = 1.., fodd = 1/2 when (ring + nside) is even (the
         * half-pixel-shifted rings, starting with ring nside) and 1 otherwise: unshifted belt rings START AT phi = 0
         * (pixel 4 of nside 1 sits at (pi/2, 0)). */
        int s = (north - nside + 1) & 1;
        r->phi0 = s * PI_L / (4.0L * ns);
        r->start = ncap + (int64_t)(north - nside) * 4 * ns;
        r->sth = sqrtl((1.0L - z) * (1.0L + z));
    }
    if (south) {
        z = -z;
        r->start = npix - r->start - r->nphi;
    }
    r->z = z;
}

int orc_ring_info(int nside, int ring, double *z, double *sth, double *phi0, int *nphi, int64_t *start)
{
    if (nside < 1 || ring < 1 || ring > 4 * nside - 1) return -1;
    ring_t r; ring_geom(nside, ring, &r);
    *z = (double)r.z; *sth = (double)r.sth; *phi0 = (double)r.phi0; *nphi = r.nphi; *start = r.start;
    return 0;
}

/* ------------------------------------------------- Wigner-d seeds/recurrence */
/* Seed lam^{mp}_{l0,m}(theta), l0 = max(m,|mp|), as mant * 2^ex (m >= 0).
 *   d^m_{m,mp}   = (-1)^(m-mp) sqrt(C(2m,m+mp)) cos(t/2)^(m+mp) sin(t/2)^(m-mp)   (m >= |mp|)
 *   d^2_{m,+2}   = sqrt(C(4,2+m)) cos(t/2)^(2+m) sin(t/2)^(2-m)                    (m < 2)
 *   d^2_{m,-2}   = (-1)^m sqrt(C(4,2-m)) cos(t/2)^(2-m) sin(t/2)^(2+m)             (m < 2)
 * with sqrt(C(2m,m)) (sin t/2 cos t/2)^m = prod_{k<=m} sqrt((2k-1)/(2k)) sin(t)^m. */
/* x^n as mant * 2^ex by binary exponentiation with exponent tracking */
static long double pow_scaled(long double x, int n, int *ex)
{
    long double r = 1.0L, b = x; int er = 0, eb = 0, t;
    b = frexpl(b, &eb);
    while (n > 0) {
        if (n & 1) { r *= b; er += eb; r = frexpl(r, &t); er += t; }
        b *= b; eb *= 2; b = frexpl(b, &t); eb += t;
        n >>= 1;
    }
    *ex = er; return r;
}
/* prod_{k=1..m} sqrt((2k-1)/(2k)) as mant * 2^ex */
static long double mfac_scaled(int m, int *ex)
{
    long double v = 1.0L; int e = 0, t;
    for (int k = 1; k <= m; ++k) { v *= sqrtl((2.0L * k - 1.0L) / (2.0L * k)); v = frexpl(v, &t); e += t; }
    *ex = e; return v;
}
static void lam_seed(int m, int mp, long double z, long double sth, long double mf, int mfe, real *mant, int *ex)
{
    int amp = mp < 0 ? -mp : mp;
    long double c2 = 0.5L * (1.0L + z), s2 = 0.5L * (1.0L - z); /* cos^2(t/2), sin^2(t/2) */
    if (z > 0.5L) s2 = 0.5L * sth * sth / (1.0L + z);
    if (z < -0.5L) c2 = 0.5L * sth * sth / (1.0L - z);
    int l0 = m > amp ? m : amp;
    long double v; int e = 0;
    if (m >= amp) {
        int e2;
        v = mf * pow_scaled(sth, m, &e2); e = mfe + e2;
        if (amp == 2) {
            v *= sqrtl((long double)m * (m - 1) / ((long double)(m + 1) * (m + 2)));
            v *= (mp > 0) ? c2 / s2 : s2 / c2;
        }
        if (m & 1) v = -v; /* (-1)^(m-mp), mp even */
    } else { /* m in {0,1}, |mp| = 2 */
        long double c = sqrtl(c2), s = sqrtl(s2);
        long double binom = (m == 0) ? 6.0L : 4.0L;
        if (mp > 0) v = sqrtl(binom) * powl(c, 2 + m) * powl(s, 2 - m);
        else        v = ((m & 1) ? -1.0L : 1.0L) * sqrtl(binom) * powl(c, 2 - m) * powl(s, 2 + m);
    }
    v *= sqrtl((2.0L * l0 + 1.0L) / (4.0L * PI_L));
    int ee; v = frexpl(v, &ee); e += ee;
    *mant = (real)v; *ex = e;
}

/* lam_{l+1} = A_l (x - B_l) lam_l - C_l lam_{l-1}; coefficients for l = l0 .. lmax-1 */
typedef struct { real A, B, C; } rec_t;
static void rec_coef(int lmax, int m, int mp, rec_t *rc)
{
    int amp = mp < 0 ? -mp : mp;
    int l0 = m > amp ? m : amp;
    for (int l = l0; l < lmax; ++l) {
        long double L1 = l + 1.0L, ll = l;
        long double den = sqrtl((L1 * L1 - (long double)m * m) * (L1 * L1 - (long double)mp * mp));
        long double f = L1 * (2.0L * ll + 1.0L) / den;
        rec_t *c = &rc[l - l0];
        c->A = (real)(sqrtl((2.0L * ll + 3.0L) / (2.0L * ll + 1.0L)) * f);
        c->B = (l == 0) ? 0 : (real)((long double)m * mp / (ll * L1));
        if (l == l0) c->C = 0;
        else c->C = (real)(sqrtl((2.0L * ll + 3.0L) / (2.0L * ll - 1.0L)) * f *
                           sqrtl((ll * ll - (long double)m * m) * (ll * ll - (long double)mp * mp)) /
                           (ll * (2.0L * ll + 1.0L)));
    }
}

/* scaled recurrence state: true value = v * 2^(-SC_K*scale) */
typedef struct { real cur, prev; int scale; } lam_state;
static inline void lam_init(lam_state *s, real mant, int ex)
{
    s->prev = 0;
    if (ex >= SC_LO) { s->scale = 0; s->cur = R_LDEXP(mant, ex); }
    else {
        s->scale = (SC_LO - ex + SC_K - 1) / SC_K;
        s->cur = R_LDEXP(mant, ex + SC_K * s->scale);
    }
}
static inline void lam_step(lam_state *s, const rec_t *c, real x)
{
    real nw = c->A * (x - c->B) * s->cur - c->C * s->prev;
    s->prev = s->cur; s->cur = nw;
    if (s->scale > 0 && R_FABS(nw) > R_LDEXP((real)1, SC_LO + SC_K)) {
        s->cur = R_LDEXP(s->cur, -SC_K); s->prev = R_LDEXP(s->prev, -SC_K); s->scale--;
    }
}

/* lam^{mp}_{lm}(theta) for l = 0..lmax (zero below l0) at arbitrary z = cos(theta); for tests */
int orc_lambda(int lmax, int m, int mp, double z_in, double *out)
{
    if (m < 0 || m > lmax || (mp != 0 && mp != 2 && mp != -2)) return -1;
    long double z = z_in, sth = sqrtl((1.0L - z) * (1.0L + z));
    int amp = mp < 0 ? -mp : mp, l0 = m > amp ? m : amp;
    for (int l = 0; l <= lmax; ++l) out[l] = 0.0;
    if (l0 > lmax) return 0;
    rec_t *rc = (rec_t *)malloc(sizeof(rec_t) * (size_t)(lmax - l0 + 1));
    rec_coef(lmax, m, mp, rc);
    int mfe; long double mf = mfac_scaled(m, &mfe);
    real mant; int ex; lam_seed(m, mp, z, sth, mf, mfe, &mant, &ex);
    lam_state s; lam_init(&s, mant, ex);
    for (int l = l0; l <= lmax; ++l) {
        out[l] = (s.scale == 0) ? (double)s.cur : 0.0;
        if (l < lmax) lam_step(&s, &rc[l - l0], (real)z);
    }
    free(rc);
    return 0;
}

/* Same with cos(theta) and sin(theta) both given as doubles -- what libsharp's Legendre loop sees (cth, sth of a double
 * theta).  Next to the poles of a large map the rounding of cos(theta) to a double alone moves lambda_lm by
 * ~ l ulp(1) / sin(theta) (5e-10 at ring 1 of nside 2048, l = 4096): a check of an FP64 transform at such rings has to feed
 * the same rounded cos(theta), with sin(theta) from 1 - |z| (not from the rounded z). */
int orc_lambda_zs(int lmax, int m, int mp, double z_in, double sth_in, double *out)
{
    if (m < 0 || m > lmax || (mp != 0 && mp != 2 && mp != -2)) return -1;
    long double z = z_in, sth = sth_in;
    int amp = mp < 0 ? -mp : mp, l0 = m > amp ? m : amp;
    for (int l = 0; l <= lmax; ++l) out[l] = 0.0;
    if (l0 > lmax) return 0;
    rec_t *rc = (rec_t *)malloc(sizeof(rec_t) * (size_t)(lmax - l0 + 1));
    rec_coef(lmax, m, mp, rc);
    int mfe; long double mf = mfac_scaled(m, &mfe);
    real mant; int ex; lam_seed(m, mp, z, sth, mf, mfe, &mant, &ex);
    lam_state s; lam_init(&s, mant, ex);
    for (int l = l0; l <= lmax; ++l) {
        out[l] = (s.scale == 0) ? (double)s.cur : 0.0;
        if (l < lmax) lam_step(&s, &rc[l - l0], (real)z);
    }
    free(rc);
    return 0;
}

/* ---------------------------------------------------- mixed-radix complex FFT */
/* out[k] = sum_j in[j*stride] exp(sign*2*pi*i*j*k/n); recursive decimation in time on the
 * smallest prime factor, direct DFT for prime lengths. tw = exp(sign*2*pi*i*t/N0), t<N0. */
static void fft_rec(int n, const real *inr, const real *ini, int stride, real *outr, real *outi,
                    const real *twr, const real *twi, int N0)
{
    if (n == 1) { outr[0] = inr[0]; outi[0] = ini[0]; return; }
    int p = 0;
    for (int q = 2; q * q <= n; ++q) if (n % q == 0) { p = q; break; }
    if (!p) p = n;
    int mlen = n / p, tstep = N0 / n;
    if (p == n) { /* direct DFT of prime length */
        for (int k = 0; k < n; ++k) {
            real sr = 0, si = 0;
            for (int j = 0; j < n; ++j) {
                int t = (int)(((int64_t)j * k) % n) * tstep;
                real ar = inr[(size_t)j * stride], ai = ini[(size_t)j * stride];
                sr += ar * twr[t] - ai * twi[t];
                si += ar * twi[t] + ai * twr[t];
            }
            outr[k] = sr; outi[k] = si;
        }
        return;
    }
    for (int q = 0; q < p; ++q)
        fft_rec(mlen, inr + (size_t)q * stride, ini + (size_t)q * stride, stride * p,
                outr + (size_t)q * mlen, outi + (size_t)q * mlen, twr, twi, N0);
    real *tr = (real *)malloc(sizeof(real) * 2 * (size_t)p), *ti = tr + p;
    real *sr = (real *)malloc(sizeof(real) * 2 * (size_t)n), *si = sr + n;
    for (int k = 0; k < mlen; ++k) {
        for (int q = 0; q < p; ++q) {
            int t = (q * k) * tstep;
            real ar = outr[(size_t)q * mlen + k], ai = outi[(size_t)q * mlen + k];
            tr[q] = ar * twr[t] - ai * twi[t];
            ti[q] = ar * twi[t] + ai * twr[t];
        }
        for (int r = 0; r < p; ++r) {
            real ar = 0, ai = 0;
            for (int q = 0; q < p; ++q) {
                int t = (int)(((int64_t)q * r) % p) * (N0 / p);
                ar += tr[q] * twr[t] - ti[q] * twi[t];
                ai += tr[q] * twi[t] + ti[q] * twr[t];
            }
            sr[k + (size_t)r * mlen] = ar; si[k + (size_t)r * mlen] = ai;
        }
    }
    memcpy(outr, sr, sizeof(real) * (size_t)n); memcpy(outi, si, sizeof(real) * (size_t)n);
    free(tr); free(sr);
}

static void make_twiddles(int n, int sign, real *twr, real *twi)
{
    for (int t = 0; t < n; ++t) {
        long double a = 2.0L * PI_L * t / n;
        twr[t] = (real)cosl(a); twi[t] = (real)(sign * sinl(a));
    }
}

/* ------------------------------------------------------------ transforms */
static inline int64_t alm_idx(int lmax, int l, int m) { return (int64_t)m * (2 * lmax + 1 - m) / 2 + l; }

/* Legendre synthesis for one m: fills Fq/Fu (re,im) for every ring (index ring-1).
 * spin 0: F = sum_l a_lm lam_lm.  spin 2 (HEALPix sign convention):
 *   Q_m = -sum_l [E F1 + i B F2],  U_m = -sum_l [B F1 - i E F2],
 *   F1 = (lam+ + lam-)/2, F2 = (lam+ - lam-)/2, lam+- = sqrt((2l+1)/4pi) d^l_{m,-+2}. */
static void leg_synth_m(int nside, int lmax, int spin, int m, const double *almE, const double *almB,
                        real *Fqr, real *Fqi, real *Fur, real *Fui, int nring)
{
    int l0 = m > spin ? m : spin;
    if (l0 > lmax) return;
    int ncoef = lmax - l0 + 1;
    rec_t *rc = (rec_t *)malloc(sizeof(rec_t) * (size_t)ncoef);
    rec_coef(lmax, m, spin ? -2 : 0, rc); /* lam+ uses m' = -2; lam- flips the sign of B */
    int mfe; long double mf = mfac_scaled(m, &mfe);
    int npair = 2 * nside; /* north rings 1..2nside; ring 2nside is the equator (its own mirror) */
    for (int ir = 1; ir <= npair; ++ir) {
        ring_t rg; ring_geom(nside, ir, &rg);
        int is = 4 * nside - ir; /* mirror ring */
        real x = (real)rg.z;
        real mant; int ex;
        lam_state sp, sm;
        lam_seed(m, spin ? -2 : 0, rg.z, rg.sth, mf, mfe, &mant, &ex); lam_init(&sp, mant, ex);
        if (spin) { lam_seed(m, 2, rg.z, rg.sth, mf, mfe, &mant, &ex); lam_init(&sm, mant, ex); }
        real sq[2] = {0, 0}, aq[2] = {0, 0}, su[2] = {0, 0}, au[2] = {0, 0}; /* sym / antisym parts */
        for (int l = l0; l <= lmax; ++l) {
            int64_t id = alm_idx(lmax, l, m);
            int odd = (l + m) & 1;
            if (!spin) {
                if (sp.scale == 0) {
                    real er = (real)almE[2 * id], ei = (real)almE[2 * id + 1];
                    if (!odd) { sq[0] += er * sp.cur; sq[1] += ei * sp.cur; }
                    else      { aq[0] += er * sp.cur; aq[1] += ei * sp.cur; }
                }
            } else {
                real lp = sp.scale == 0 ? sp.cur : 0, lm = sm.scale == 0 ? sm.cur : 0;
                real f1 = (real)0.5 * (lp + lm), f2 = (real)0.5 * (lp - lm);
                real er = (real)almE[2 * id], ei = (real)almE[2 * id + 1];
                real br = (real)almB[2 * id], bi = (real)almB[2 * id + 1];
                /* F1 terms have parity (-1)^(l+m), F2 terms the opposite one */
                real *q1 = odd ? aq : sq, *q2 = odd ? sq : aq, *u1 = odd ? au : su, *u2 = odd ? su : au;
                q1[0] -= er * f1; q1[1] -= ei * f1; u1[0] -= br * f1; u1[1] -= bi * f1;
                q2[0] += bi * f2; q2[1] -= br * f2; u2[0] -= ei * f2; u2[1] += er * f2;
            }
            if (l < lmax) {
                lam_step(&sp, &rc[l - l0], x);
                if (spin) { rec_t c = rc[l - l0]; c.B = -c.B; lam_step(&sm, &c, x); }
            }
        }
        Fqr[ir - 1] = sq[0] + aq[0]; Fqi[ir - 1] = sq[1] + aq[1];
        if (spin) { Fur[ir - 1] = su[0] + au[0]; Fui[ir - 1] = su[1] + au[1]; }
        if (is != ir) {
            Fqr[is - 1] = sq[0] - aq[0]; Fqi[is - 1] = sq[1] - aq[1];
            if (spin) { Fur[is - 1] = su[0] - au[0]; Fui[is - 1] = su[1] - au[1]; }
        }
    }
    (void)nring;
    free(rc);
}

/* adjoint of leg_synth_m: given ring spectra G (any weights already applied) accumulate
 * alm[l,m] = sum_rings conj-structure of the synthesis (exact transpose in the real layout). */
static void leg_anal_m(int nside, int lmax, int spin, int m, double *almE, double *almB,
                       const real *Gqr, const real *Gqi, const real *Gur, const real *Gui)
{
    int l0 = m > spin ? m : spin;
    for (int l = m; l <= lmax; ++l) {
        int64_t id = alm_idx(lmax, l, m);
        almE[2 * id] = almE[2 * id + 1] = 0.0;
        if (spin) almB[2 * id] = almB[2 * id + 1] = 0.0;
    }
    if (l0 > lmax) return;
    int ncoef = lmax - l0 + 1;
    rec_t *rc = (rec_t *)malloc(sizeof(rec_t) * (size_t)ncoef);
    real *acc = (real *)calloc((size_t)ncoef * 4, sizeof(real));
    rec_coef(lmax, m, spin ? -2 : 0, rc);
    int mfe; long double mf = mfac_scaled(m, &mfe);
    int npair = 2 * nside;
    for (int ir = 1; ir <= npair; ++ir) {
        ring_t rg; ring_geom(nside, ir, &rg);
        int is = 4 * nside - ir;
        real x = (real)rg.z;
        real mant; int ex;
        lam_state sp, sm;
        lam_seed(m, spin ? -2 : 0, rg.z, rg.sth, mf, mfe, &mant, &ex); lam_init(&sp, mant, ex);
        if (spin) { lam_seed(m, 2, rg.z, rg.sth, mf, mfe, &mant, &ex); lam_init(&sm, mant, ex); }
        real nq[2] = {Gqr[ir - 1], Gqi[ir - 1]}, nu[2] = {0, 0}, zq[2] = {0, 0}, zu[2] = {0, 0};
        if (spin) { nu[0] = Gur[ir - 1]; nu[1] = Gui[ir - 1]; }
        if (is != ir) { zq[0] = Gqr[is - 1]; zq[1] = Gqi[is - 1]; if (spin) { zu[0] = Gur[is - 1]; zu[1] = Gui[is - 1]; } }
        real sq[2] = {nq[0] + zq[0], nq[1] + zq[1]}, aq[2] = {nq[0] - zq[0], nq[1] - zq[1]};
        real su[2] = {nu[0] + zu[0], nu[1] + zu[1]}, au[2] = {nu[0] - zu[0], nu[1] - zu[1]};
        for (int l = l0; l <= lmax; ++l) {
            int odd = (l + m) & 1;
            real *a = &acc[(size_t)(l - l0) * 4];
            if (!spin) {
                if (sp.scale == 0) { const real *q = odd ? aq : sq; a[0] += q[0] * sp.cur; a[1] += q[1] * sp.cur; }
            } else {
                real lp = sp.scale == 0 ? sp.cur : 0, lm = sm.scale == 0 ? sm.cur : 0;
                real f1 = (real)0.5 * (lp + lm), f2 = (real)0.5 * (lp - lm);
                const real *q1 = odd ? aq : sq, *q2 = odd ? sq : aq, *u1 = odd ? au : su, *u2 = odd ? su : au;
                /* transpose of: q1 -= E f1; u1 -= B f1; q2.re += Bi f2; q2.im -= Br f2; u2.re -= Ei f2; u2.im += Er f2 */
                a[0] += -q1[0] * f1 + u2[1] * f2; /* E re */
                a[1] += -q1[1] * f1 - u2[0] * f2; /* E im */
                a[2] += -u1[0] * f1 - q2[1] * f2; /* B re */
                a[3] += -u1[1] * f1 + q2[0] * f2; /* B im */
            }
            if (l < lmax) {
                lam_step(&sp, &rc[l - l0], x);
                if (spin) { rec_t c = rc[l - l0]; c.B = -c.B; lam_step(&sm, &c, x); }
            }
        }
    }
    for (int l = l0; l <= lmax; ++l) {
        int64_t id = alm_idx(lmax, l, m);
        const real *a = &acc[(size_t)(l - l0) * 4];
        almE[2 * id] = (double)a[0]; almE[2 * id + 1] = (double)a[1];
        if (spin) { almB[2 * id] = (double)a[2]; almB[2 * id + 1] = (double)a[3]; }
    }
    free(rc); free(acc);
}

/* alm (healpy m-major complex, interleaved re/im) -> RING maps.
 * spin 0: almE -> mapQ (almB/mapU ignored). spin 2: (almE, almB) -> (mapQ, mapU). */
int orc_alm2map(int nside, int lmax, int spin, const double *almE, const double *almB,
                double *mapQ, double *mapU)
{
    if (nside < 1 || lmax < 0 || (spin != 0 && spin != 2)) return -1;
    int nring = 4 * nside - 1, nm = lmax + 1;
    int ncomp = spin ? 2 : 1;
    size_t fsz = (size_t)nring * nm;
    real *F = (real *)calloc(fsz * 2 * ncomp, sizeof(real)); /* [comp][re/im][m][ring] */
    if (!F) return -2;
#pragma omp parallel for schedule(dynamic, 1)
    for (int m = 0; m <= lmax; ++m) {
        real *Fqr = F + (size_t)m * nring, *Fqi = F + fsz + (size_t)m * nring;
        real *Fur = spin ? F + 2 * fsz + (size_t)m * nring : NULL, *Fui = spin ? F + 3 * fsz + (size_t)m * nring : NULL;
        leg_synth_m(nside, lmax, spin, m, almE, almB, Fqr, Fqi, Fur, Fui, nring);
    }
#pragma omp parallel for schedule(dynamic, 4)
    for (int ir = 1; ir <= nring; ++ir) {
        ring_t rg; ring_geom(nside, ir, &rg);
        int n = rg.nphi;
        real *buf = (real *)malloc(sizeof(real) * 6 * (size_t)n);
        real *zr = buf, *zi = buf + n, *outr = buf + 2 * n, *outi = buf + 3 * n, *twr = buf + 4 * n, *twi = buf + 5 * n;
        make_twiddles(n, +1, twr, twi);
        for (int c = 0; c < ncomp; ++c) {
            const real *Fr = F + (size_t)(2 * c) * fsz, *Fi = F + (size_t)(2 * c + 1) * fsz;
            for (int k = 0; k < n; ++k) { zr[k] = 0; zi[k] = 0; }
            for (int m = 0; m <= lmax; ++m) {
                long double ph = fmodl((long double)m * rg.phi0, 2.0L * PI_L);
                real cr = (real)cosl(ph), ci = (real)sinl(ph);
                real fr = Fr[(size_t)m * nring + ir - 1], fi = Fi[(size_t)m * nring + ir - 1];
                real w = m ? 2 : 1;
                int k = m % n;
                zr[k] += w * (fr * cr - fi * ci);
                zi[k] += w * (fr * ci + fi * cr);
            }
            fft_rec(n, zr, zi, 1, outr, outi, twr, twi, n);
            double *mp = c ? mapU : mapQ;
            for (int j = 0; j < n; ++j) mp[rg.start + j] = (double)outr[j];
        }
        free(buf);
    }
    free(F);
    return 0;
}

/* RING maps -> alm = weight * sum_p conj(Y_lm(p)) f(p)   (weight = 4 pi / Npix is
 * hp.map2alm(iter=0, use_weights=False); weight = 1 is the plain adjoint A^T of the
 * reference's "adjoint synthesis", utils.py:79-111 / config.py:72). */
int orc_map2alm(int nside, int lmax, int spin, const double *mapQ, const double *mapU,
                double *almE, double *almB, double weight)
{
    if (nside < 1 || lmax < 0 || (spin != 0 && spin != 2)) return -1;
    int nring = 4 * nside - 1, nm = lmax + 1;
    int ncomp = spin ? 2 : 1;
    size_t fsz = (size_t)nring * nm;
    real *G = (real *)calloc(fsz * 2 * ncomp, sizeof(real));
    if (!G) return -2;
#pragma omp parallel for schedule(dynamic, 4)
    for (int ir = 1; ir <= nring; ++ir) {
        ring_t rg; ring_geom(nside, ir, &rg);
        int n = rg.nphi;
        real *buf = (real *)malloc(sizeof(real) * 6 * (size_t)n);
        real *zr = buf, *zi = buf + n, *outr = buf + 2 * n, *outi = buf + 3 * n, *twr = buf + 4 * n, *twi = buf + 5 * n;
        make_twiddles(n, -1, twr, twi);
        for (int c = 0; c < ncomp; ++c) {
            const double *mp = c ? mapU : mapQ;
            real *Gr = G + (size_t)(2 * c) * fsz, *Gi = G + (size_t)(2 * c + 1) * fsz;
            for (int j = 0; j < n; ++j) { zr[j] = (real)mp[rg.start + j]; zi[j] = 0; }
            fft_rec(n, zr, zi, 1, outr, outi, twr, twi, n);
            for (int m = 0; m <= lmax; ++m) {
                long double ph = fmodl((long double)m * rg.phi0, 2.0L * PI_L);
                real cr = (real)cosl(ph), ci = (real)(-sinl(ph));
                int k = m % n;
                Gr[(size_t)m * nring + ir - 1] = (real)weight * (outr[k] * cr - outi[k] * ci);
                Gi[(size_t)m * nring + ir - 1] = (real)weight * (outr[k] * ci + outi[k] * cr);
            }
        }
        free(buf);
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int m = 0; m <= lmax; ++m) {
        const real *Gqr = G + (size_t)m * nring, *Gqi = G + fsz + (size_t)m * nring;
        const real *Gur = spin ? G + 2 * fsz + (size_t)m * nring : NULL, *Gui = spin ? G + 3 * fsz + (size_t)m * nring : NULL;
        leg_anal_m(nside, lmax, spin, m, almE, almB, Gqr, Gqi, Gur, Gui);
    }
    free(G);
    return 0;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
int orc_real_bits(void) { return (int)(sizeof(real) * 8); }
