/*
 * oracle/sht_fast.c -- TEST / BENCH INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Vectorised CPU implementation of the spin-2 HEALPix transform pair that the reference's hot path
 * executes through healpy (libsharp2 + pocketfft; call sites CenteredGibbs.py:298-299, 505-513,
 * NonCenteredGibbs.py:350-351, utils.py:87-106, and qcinv's opfilt_pp.fwd_op every PCG iteration).
 * It restates the same mathematics as oracle/sht_oracle.c (which stays the accuracy checker) with the
 * organisation libsharp uses on a CPU, so that `bench.py --impl reference` and the `cpu_baseline` leg
 * time a credible stand-in for "healpy on the box's host cores" instead of a scalar loop:
 *   - SIMD lanes = ring pairs (8 doubles per vector, GCC vector extensions; target_clones picks the
 *     AVX-512 / AVX2 / baseline build of the hot loops at load time, so the .so travels between hosts),
 *     scalar loop over l, OpenMP over m, north/south rings share one recurrence through parity;
 *   - two-FMA normalised three-term recurrence for lambda^{+-}_{lm} = sqrt((2l+1)/4pi) d^l_{m,-+2}
 *     (tables in long double), integer-scaled range extension, libsharp's m_lim ring pruning;
 *   - lambda^{+-} basis: 4 recurrence FMAs + 8 accumulation FMAs per (ring pair, l, m);
 *   - ring FFTs: power-of-two Stockham passes, Bluestein for the 4i-pixel cap rings, 4 sequences
 *     (Q, U) x (north, south) per vector.
 * PARITY: checked against oracle/sht_oracle.c (long double) in tests/test_oracle_fast.py; healpy itself
 * is not installable here, so healpy parity is UNPINNED exactly as for sht_oracle.c.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SC_LO (-900)
#define SC_K 256
#define VL 8

typedef double v8d __attribute__((vector_size(64), aligned(64)));
typedef long long v8l __attribute__((vector_size(64), aligned(64)));
typedef double v4d __attribute__((vector_size(32), aligned(32)));

static const long double PI_L = 3.14159265358979323846264338327950288L;

static void *amalloc(size_t bytes)
{
    void *p = NULL;
    if (posix_memalign(&p, 64, bytes ? bytes : 64)) return NULL;
    return p;
}

/* ------------------------------------------------------------------ plan */
typedef struct {
    int n, M;            /* ring length, power-of-two convolution length (0: n is a power of two) */
    double *cr, *ci;     /* chirp c_j = exp(-i pi j^2 / n), j < n */
    double *br, *bi;     /* FFT_M of the wrapped conjugate chirp, divided by M */
} blue_t;

typedef struct {
    int nside, lmax, npair, npad, nring;
    double *cth;                 /* [npad] cos(theta) of the north ring of each pair (pole -> equator) */
    int *mlim;                   /* [npair] last m that matters on the pair (libsharp's rule, spin 2) */
    int *pmin;                   /* [lmax+1] first pair that reaches m */
    double *ra, *rb, *alpha;     /* [nalm] two-FMA recurrence (a_l, b_l) and normalisation alpha_l, idx(l,m) */
    double *seedp, *seedm;       /* [lmax+1][npad] mu^+ / mu^- at l0 = max(m,2), relative to 2^seede */
    int *seede;                  /* [lmax+1][npad] */
    int *nphi; int64_t *startN, *startS; int *phq, *phden; /* [npair] */
    double *twr, *twi; int twn;  /* exp(-2 pi i k / twn) */
    blue_t *blue;                /* [npair] Bluestein tables of the pair's ring length (M = 0: none) */
    double *belt_cr, *belt_ci;   /* exp(i pi m / (4 nside)), m <= lmax: phase of the shifted belt rings */
    double *F;                   /* [npad][lmax+1][8] ring spectra records (see sht_fast_hot.h) */
} fplan;

static fplan *g_plan = NULL;
static double g_t[4];   /* seconds of the last call: Legendre, ring stage (synthesis: 0,1; analysis: 2,3) */
static double now(void)
{
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

static int mlim_of(int lmax, int spin, long double sth, long double cth)
{   /* libsharp's sharp_get_mlim: lambda_lm is evanescent for m > l sin(theta) */
    long double ofs = lmax * 0.01L;
    if (ofs < 100.0L) ofs = 100.0L;
    long double b = -2.0L * spin * fabsl(cth);
    long double t1 = lmax * sth + ofs;
    long double c = (long double)spin * spin - t1 * t1;
    long double discr = b * b - 4 * c;
    if (discr <= 0) return lmax;
    long double res = (-b + sqrtl(discr)) / 2.0L;
    if (res > lmax) res = lmax;
    return (int)(res + 0.5L);
}

typedef struct { double a, b, c[8]; } coef_t;   /* per l: recurrence (a, b) + 8 broadcast coefficients */
#define CHK 16

#if defined(__x86_64__)
#pragma GCC push_options
#pragma GCC target("avx512f,avx512dq,avx512vl,avx2,fma")
#define SUF _avx512
#include "sht_fast_hot.h"
#undef SUF
#pragma GCC pop_options
#pragma GCC push_options
#pragma GCC target("avx2,fma")
#define SUF _avx2
#include "sht_fast_hot.h"
#undef SUF
#pragma GCC pop_options
#endif
#define SUF _base
#include "sht_fast_hot.h"
#undef SUF

typedef void (*fft_fn)(int, v4d *, v4d *, v4d *, v4d *, const double *, const double *, int, v4d **, v4d **);
typedef void (*rfft_fn)(const fplan *, const blue_t *, v4d *, v4d *, v4d *, v4d *, v4d **, v4d **);
typedef void (*rsyn_fn)(const fplan *, int, const double *, const double *, v4d *, size_t, double *, double *);
typedef void (*rana_fn)(const fplan *, int, const double *, const double *, v4d *, size_t, const double *, const double *, const double *, double);
typedef void (*synth_fn)(const fplan *, int, const coef_t *);
typedef void (*anal_fn)(const fplan *, int, const double *, v8d *, v8d *, int *);
static fft_fn fft_pow2 = NULL;
static rfft_fn ring_fft = NULL;
static synth_fn leg_synth_m = NULL;
static rsyn_fn ring_synth_pair = NULL;
static rana_fn ring_anal_pair = NULL;
static anal_fn leg_anal_m = NULL;
static double (*fma_peak_loop)(long, double, double) = NULL;
static int g_simd_bits = 0;

static void dispatch_init(void)
{
    if (g_simd_bits) return;
    const char *force = getenv("ORF_SIMD");   /* 512 / 256 / 128: testing aid */
    int want = force ? atoi(force) : 512;
#if defined(__x86_64__)
    __builtin_cpu_init();
    if (want >= 512 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl")) {
        fft_pow2 = fft_pow2_avx512; ring_fft = ring_fft_avx512; leg_synth_m = leg_synth_m_avx512; leg_anal_m = leg_anal_m_avx512;
        fma_peak_loop = fma_peak_loop_avx512; ring_synth_pair = ring_synth_pair_avx512; ring_anal_pair = ring_anal_pair_avx512;
        g_simd_bits = 512;
        return;
    }
    if (want >= 256 && __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) {
        fft_pow2 = fft_pow2_avx2; ring_fft = ring_fft_avx2; leg_synth_m = leg_synth_m_avx2; leg_anal_m = leg_anal_m_avx2;
        fma_peak_loop = fma_peak_loop_avx2; ring_synth_pair = ring_synth_pair_avx2; ring_anal_pair = ring_anal_pair_avx2;
        g_simd_bits = 256;
        return;
    }
#endif
    (void)want;
    fft_pow2 = fft_pow2_base; ring_fft = ring_fft_base; leg_synth_m = leg_synth_m_base; leg_anal_m = leg_anal_m_base;
        fma_peak_loop = fma_peak_loop_base; ring_synth_pair = ring_synth_pair_base; ring_anal_pair = ring_anal_pair_base;
    g_simd_bits = 128;
}

static void plan_free(fplan *p)
{
    if (!p) return;
    free(p->cth); free(p->mlim); free(p->pmin); free(p->ra); free(p->rb); free(p->alpha);
    free(p->seedp); free(p->seedm); free(p->seede); free(p->nphi); free(p->startN); free(p->startS);
    free(p->phq); free(p->phden); free(p->twr); free(p->twi); free(p->belt_cr); free(p->belt_ci); free(p->F);
    if (p->blue) {
        for (int q = 0; q < p->npair; ++q) { free(p->blue[q].cr); free(p->blue[q].ci); free(p->blue[q].br); free(p->blue[q].bi); }
        free(p->blue);
    }
    free(p);
}

static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

static fplan *plan_get(int nside, int lmax)
{
    if (g_plan && g_plan->nside == nside && g_plan->lmax == lmax) return g_plan;
    dispatch_init();
    plan_free(g_plan);
    g_plan = NULL;
    fplan *p = (fplan *)calloc(1, sizeof(fplan));
    const int L = lmax, npair = 2 * nside, npad = (npair + 15) & ~15;
    const int64_t ns = nside, npix = 12 * ns * ns, ncap = 2 * ns * (ns - 1);
    p->nside = nside; p->lmax = L; p->npair = npair; p->npad = npad; p->nring = 4 * nside - 1;
    p->cth = (double *)amalloc(sizeof(double) * npad);
    p->mlim = (int *)malloc(sizeof(int) * npair);
    p->pmin = (int *)malloc(sizeof(int) * (L + 1));
    p->nphi = (int *)malloc(sizeof(int) * npair);
    p->phq = (int *)malloc(sizeof(int) * npair);
    p->phden = (int *)malloc(sizeof(int) * npair);
    p->startN = (int64_t *)malloc(sizeof(int64_t) * npair);
    p->startS = (int64_t *)malloc(sizeof(int64_t) * npair);
    long double *sthl = (long double *)malloc(sizeof(long double) * npair);
    long double *t2l = (long double *)malloc(sizeof(long double) * npair);   /* tan^2(theta/2) */
    long double *c2l = (long double *)malloc(sizeof(long double) * npair), *s2l = (long double *)malloc(sizeof(long double) * npair);
    for (int q = 0; q < npad; ++q) p->cth[q] = 0.0;
    for (int i = 1; i <= npair; ++i) {
        long double z, omz;
        if (i < nside) {
            omz = (long double)i * i / (3.0L * ns * ns);
            z = 1.0L - omz;
            p->nphi[i - 1] = 4 * i; p->phq[i - 1] = 1; p->phden[i - 1] = 4 * i;
            p->startN[i - 1] = 2 * (int64_t)i * (i - 1);
        } else {
            z = 4.0L / 3.0L - 2.0L * i / (3.0L * ns);
            omz = 1.0L - z;
            p->nphi[i - 1] = 4 * nside; p->phq[i - 1] = (i - nside + 1) & 1; p->phden[i - 1] = 4 * nside;
            p->startN[i - 1] = ncap + (int64_t)(i - nside) * 4 * ns;
        }
        p->startS[i - 1] = npix - p->startN[i - 1] - p->nphi[i - 1];
        long double st = sqrtl(omz * (1.0L + z));
        sthl[i - 1] = st;
        s2l[i - 1] = 0.5L * omz; c2l[i - 1] = 0.5L * (1.0L + z);
        t2l[i - 1] = s2l[i - 1] / c2l[i - 1];
        p->cth[i - 1] = (double)z;
        p->mlim[i - 1] = mlim_of(L, 2, st, z);
    }
    for (int m = 0; m <= L; ++m) {
        p->pmin[m] = npair;
        for (int q = 0; q < npair; ++q) if (p->mlim[q] >= m) { p->pmin[m] = q; break; }
    }
    /* recurrence tables: lam_{l+1} = A_l (x - B_l) lam_l - C_l lam_{l-1}; lam_l = alpha_l mu_l with
     * alpha_{l+1} = C_l alpha_{l-1} gives mu_{l+1} = (a_l x -+ b_l) mu_l - mu_{l-1} (m' = +-2). */
    const int64_t nalm = (int64_t)(L + 1) * (L + 2) / 2;
    p->ra = (double *)calloc(nalm, sizeof(double));
    p->rb = (double *)calloc(nalm, sizeof(double));
    p->alpha = (double *)calloc(nalm, sizeof(double));
#pragma omp parallel for schedule(dynamic, 8)
    for (int m = 0; m <= L; ++m) {
        int l0 = m > 2 ? m : 2;
        if (l0 > L) continue;
        int64_t base = (int64_t)m * (2 * L + 1 - m) / 2;
        long double a_prev = 1.0L, a_cur = 1.0L;
        for (int l = l0; l <= L; ++l) {
            p->alpha[base + l] = (double)a_cur;
            if (l == L) break;
            long double ll = l, L1 = l + 1.0L, mm = m;
            long double den = sqrtl((L1 * L1 - mm * mm) * (L1 * L1 - 4.0L));
            long double f = L1 * (2.0L * ll + 1.0L) / den;
            long double A = sqrtl((2.0L * ll + 3.0L) / (2.0L * ll + 1.0L)) * f;
            long double B = mm * 2.0L / (ll * L1);
            long double a_next = 1.0L;
            if (l != l0) {
                long double Cc = sqrtl((2.0L * ll + 3.0L) / (2.0L * ll - 1.0L)) * f *
                                 sqrtl((ll * ll - mm * mm) * (ll * ll - 4.0L)) / (ll * (2.0L * ll + 1.0L));
                a_next = Cc * a_prev;
            }
            long double a = A * a_cur / a_next;
            p->ra[base + l] = (double)a;
            p->rb[base + l] = (double)(a * B);
            a_prev = a_cur; a_cur = a_next;
        }
    }
    /* seeds at l0 = max(m, 2):
     *   m >= 2: lam^{+-}_{mm} = (-1)^m sqrt((2m+1)/4pi) prod_{k<=m} sqrt((2k-1)/2k) sqrt(m(m-1)/((m+1)(m+2))) sin^m(t) tan^{+-2}(t/2)
     *   m < 2 : closed forms of d^2_{m,-+2} (sht_oracle.c:lam_seed) */
    p->seedp = (double *)amalloc(sizeof(double) * (size_t)(L + 1) * npad);
    p->seedm = (double *)amalloc(sizeof(double) * (size_t)(L + 1) * npad);
    p->seede = (int *)amalloc(sizeof(int) * (size_t)(L + 1) * npad);
    memset(p->seedp, 0, sizeof(double) * (size_t)(L + 1) * npad);
    memset(p->seedm, 0, sizeof(double) * (size_t)(L + 1) * npad);
    memset(p->seede, 0, sizeof(int) * (size_t)(L + 1) * npad);
    {
        long double *mf = (long double *)malloc(sizeof(long double) * (L + 1));
        int *mfe = (int *)malloc(sizeof(int) * (L + 1));
        long double v = 1.0L; int e = 0, t;
        for (int m = 0; m <= L; ++m) {
            if (m > 0) { v *= sqrtl((2.0L * m - 1.0L) / (2.0L * m)); v = frexpl(v, &t); e += t; }
            mf[m] = v; mfe[m] = e;
        }
#pragma omp parallel for schedule(static)
        for (int q = 0; q < npair; ++q) {
            long double sp = 1.0L; int se = 0, tt;   /* sin^m as sp * 2^se */
            for (int m = 0; m <= L; ++m) {
                if (m > 0) { sp *= sthl[q]; sp = frexpl(sp, &tt); se += tt; }
                long double vp, vm; int ex;
                if (m >= 2) {
                    long double base = mf[m] * sp * sqrtl((long double)m * (m - 1) / ((long double)(m + 1) * (m + 2))) *
                                       sqrtl((2.0L * m + 1.0L) / (4.0L * PI_L));
                    if (m & 1) base = -base;
                    ex = mfe[m] + se;
                    vp = base * t2l[q];      /* m' = -2 */
                    vm = base / t2l[q];      /* m' = +2 */
                } else {
                    long double c = sqrtl(c2l[q]), s = sqrtl(s2l[q]);
                    long double binom = (m == 0) ? 6.0L : 4.0L;
                    long double nrm = sqrtl(5.0L / (4.0L * PI_L)) * sqrtl(binom);
                    vm = nrm * powl(c, 2 + m) * powl(s, 2 - m);                                /* m' = +2 */
                    vp = ((m & 1) ? -1.0L : 1.0L) * nrm * powl(c, 2 - m) * powl(s, 2 + m);      /* m' = -2 */
                    ex = 0;
                }
                /* common exponent from the larger of the two */
                long double big = fabsl(vm) > fabsl(vp) ? fabsl(vm) : fabsl(vp);
                int eb = 0;
                if (big > 0) (void)frexpl(big, &eb);
                p->seedp[(size_t)m * npad + q] = (double)ldexpl(vp, -eb);
                p->seedm[(size_t)m * npad + q] = (double)ldexpl(vm, -eb);
                p->seede[(size_t)m * npad + q] = ex + eb;
            }
        }
        free(mf); free(mfe);
    }
    /* FFT tables */
    int maxM = 1;
    for (int q = 0; q < npair; ++q) {
        int n = p->nphi[q], M = n;
        if (!is_pow2(n)) { M = 1; while (M < 2 * n - 1) M *= 2; }
        if (M > maxM) maxM = M;
    }
    p->twn = maxM;
    p->twr = (double *)amalloc(sizeof(double) * maxM);
    p->twi = (double *)amalloc(sizeof(double) * maxM);
    for (int k = 0; k < maxM; ++k) {
        long double a = 2.0L * PI_L * k / maxM;
        p->twr[k] = (double)cosl(a); p->twi[k] = (double)(-sinl(a));
    }
    p->blue = (blue_t *)calloc(npair, sizeof(blue_t));
#pragma omp parallel for schedule(dynamic, 4)
    for (int q = 0; q < npair; ++q) {
        int n = p->nphi[q];
        blue_t *b = &p->blue[q];
        b->n = n; b->M = 0;
        if (is_pow2(n)) continue;
        int M = 1; while (M < 2 * n - 1) M *= 2;
        b->M = M;
        b->cr = (double *)amalloc(sizeof(double) * n); b->ci = (double *)amalloc(sizeof(double) * n);
        b->br = (double *)amalloc(sizeof(double) * M); b->bi = (double *)amalloc(sizeof(double) * M);
        for (int j = 0; j < n; ++j) {
            int64_t r = ((int64_t)j * j) % (2 * n);
            long double a = PI_L * r / n;
            b->cr[j] = (double)cosl(a); b->ci[j] = (double)(-sinl(a));
        }
        v4d *xr = (v4d *)amalloc(sizeof(v4d) * M * 4), *xi = xr + M, *yr = xi + M, *yi = yr + M;
        for (int j = 0; j < M; ++j) { xr[j] = (v4d){0, 0, 0, 0}; xi[j] = (v4d){0, 0, 0, 0}; }
        for (int j = 0; j < n; ++j) {   /* conj(c)[|j|] wrapped to length M */
            v4d vr = {b->cr[j], 0, 0, 0}, vi = {-b->ci[j], 0, 0, 0};
            xr[j] = vr; xi[j] = vi;
            if (j) { xr[M - j] = vr; xi[M - j] = vi; }
        }
        v4d *orr, *oii;
        fft_pow2(M, xr, xi, yr, yi, p->twr, p->twi, p->twn, &orr, &oii);
        for (int j = 0; j < M; ++j) { b->br[j] = orr[j][0] / M; b->bi[j] = oii[j][0] / M; }
        free(xr);
    }
    p->belt_cr = (double *)malloc(sizeof(double) * (L + 1));
    p->belt_ci = (double *)malloc(sizeof(double) * (L + 1));
    for (int m = 0; m <= L; ++m) {
        long double a = PI_L * (m % (8 * nside)) / (4.0L * nside);
        p->belt_cr[m] = (double)cosl(a); p->belt_ci[m] = (double)sinl(a);
    }
    p->F = (double *)amalloc(sizeof(double) * 8 * (size_t)(L + 1) * npad);
    memset(p->F, 0, sizeof(double) * 8 * (size_t)(L + 1) * npad);
    free(sthl); free(t2l); free(c2l); free(s2l);
    g_plan = p;
    return p;
}

/* phase exp(i m phi0) of pair q for m = 0..lmax into (cr, ci) */
static void ring_phases(const fplan *p, int q, double *cr, double *ci)
{
    const int L = p->lmax;
    if (p->nphi[q] == 4 * p->nside && p->phden[q] == 4 * p->nside) {
        if (p->phq[q] == 0) { for (int m = 0; m <= L; ++m) { cr[m] = 1.0; ci[m] = 0.0; } }
        else { memcpy(cr, p->belt_cr, sizeof(double) * (L + 1)); memcpy(ci, p->belt_ci, sizeof(double) * (L + 1)); }
        return;
    }
    const int den = p->phden[q], pq = p->phq[q];
    const double phi0 = (double)(PI_L * pq / den);
    const double sr = cos(phi0), si = sin(phi0);
    for (int m = 0; m <= L; ++m) {
        if ((m & 31) == 0) {   /* re-anchor the rotation every 32 steps */
            long double a = PI_L * (long double)(((int64_t)m * pq) % (2 * den)) / den;
            cr[m] = (double)cosl(a); ci[m] = (double)sinl(a);
        } else {
            cr[m] = cr[m - 1] * sr - ci[m - 1] * si;
            ci[m] = cr[m - 1] * si + ci[m - 1] * sr;
        }
    }
}

/* ------------------------------------------------------------------ transforms */
static inline __attribute__((always_inline)) double hsum(v8d v) { double s = 0; for (int i = 0; i < VL; ++i) s += v[i]; return s; }

/* (E, B) -> RING maps (Q, U): hp.alm2map([0, E, B], pol=True).  real_layout = 0: healpy complex alms (interleaved
 * re/im); 1: the reference's real layout (utils.py:49-76), i.e. real_to_complex is done while the coefficients are
 * staged.  flE / flB (nullable): per-l factors applied to the input (hp.almxfl). */
static int synth_impl(int nside, int lmax, const double *almE, const double *almB, int real_layout, const double *flE,
                      const double *flB, double *mapQ, double *mapU)
{
    if (nside < 1 || lmax < 2) return -1;
    fplan *p = plan_get(nside, lmax);
    if (!p) return -2;
    const int L = lmax, npad = p->npad;
    const double t_begin = now();
#pragma omp parallel
    {
        coef_t *cf = (coef_t *)amalloc(sizeof(coef_t) * (L + 1));
#pragma omp for schedule(dynamic, 1)
        for (int m = 0; m <= L; ++m) {
            const int l0 = m > 2 ? m : 2;
            if (l0 > L) { for (int q = 0; q < npad; ++q) memset(p->F + ((size_t)q * (L + 1) + m) * 8, 0, sizeof(double) * 8); continue; }
            const int64_t base = (int64_t)m * (2 * L + 1 - m) / 2;
            for (int l = l0; l <= L; ++l) {
                coef_t *c = &cf[l - l0];
                const double k = -0.5 * p->alpha[base + l], s = ((l + m) & 1) ? -1.0 : 1.0;
                double er, ei, br, bi;
                if (!real_layout) {
                    er = almE[2 * (base + l)]; ei = almE[2 * (base + l) + 1];
                    br = almB[2 * (base + l)]; bi = almB[2 * (base + l) + 1];
                } else if (m == 0) {
                    er = almE[l]; ei = 0.0; br = almB[l]; bi = 0.0;
                } else {
                    const int64_t o = 2 * (base + l) - (L + 1);
                    er = almE[o] * M_SQRT1_2; ei = almE[o + 1] * M_SQRT1_2; br = almB[o] * M_SQRT1_2; bi = almB[o + 1] * M_SQRT1_2;
                }
                if (flE) { er *= flE[l]; ei *= flE[l]; }
                if (flB) { br *= flB[l]; bi *= flB[l]; }
                c->a = p->ra[base + l]; c->b = p->rb[base + l];
                const double cpr = k * (er - bi), cpi = k * (ei + br);   /* c+ = E + iB */
                const double cmr = k * (er + bi), cmi = k * (ei - br);   /* c- = E - iB */
                c->c[0] = cpr; c->c[1] = cpi; c->c[2] = cmr; c->c[3] = cmi;
                c->c[4] = s * cpr; c->c[5] = s * cpi; c->c[6] = s * cmr; c->c[7] = s * cmi;
            }
            leg_synth_m(p, m, cf);
        }
        free(cf);
#pragma omp master
        g_t[0] = now() - t_begin;
        /* ring stage: blocks of 8 pairs share the cache lines of the F rows */
        size_t wlen = 16;
        for (int q = 0; q < p->npair; ++q) { size_t w = p->blue[q].M ? (size_t)p->blue[q].M : (size_t)p->nphi[q]; if (w > wlen) wlen = w; }
        v4d *w = (v4d *)amalloc(sizeof(v4d) * 4 * wlen);
        double *cr = (double *)malloc(sizeof(double) * 2 * (L + 1)), *ci = cr + L + 1;
#pragma omp for schedule(dynamic, 1)
        for (int blk = 0; blk < (p->npair + 7) / 8; ++blk) {
            for (int q = blk * 8; q < p->npair && q < blk * 8 + 8; ++q) {
                ring_phases(p, q, cr, ci);
                ring_synth_pair(p, q, cr, ci, w, wlen, mapQ, mapU);
            }
        }
        free(w); free(cr);
    }
    g_t[1] = now() - t_begin - g_t[0];
    return 0;
}

int orf_alm2map_spin2(int nside, int lmax, const double *almE, const double *almB, double *mapQ, double *mapU)
{
    return synth_impl(nside, lmax, almE, almB, 0, NULL, NULL, mapQ, mapU);
}

/* real-layout coefficients times per-l factors -> maps: alm2map(almxfl(real_to_complex(s), fl)) in one call */
int orf_synth_real(int nside, int lmax, const double *sE, const double *sB, const double *flE, const double *flB,
                   double *mapQ, double *mapU)
{
    return synth_impl(nside, lmax, sE, sB, 1, flE, flB, mapQ, mapU);
}

/* RING maps (Q, U) -> (E, B) = weight * sum_p conj(Y)(p) f(p): weight = 4 pi / Npix is hp.map2alm(iter=0,
 * use_weights=False); weight = 1 is the plain transpose A^T of the synthesis (utils.py:79-111 / config.py:72) */
static int anal_impl(int nside, int lmax, const double *mapQ, const double *mapU, const double *pixw, int real_layout,
                     const double *flE, const double *flB, double *almE, double *almB, double weight)
{
    if (nside < 1 || lmax < 2) return -1;
    fplan *p = plan_get(nside, lmax);
    if (!p) return -2;
    const int L = lmax, npad = p->npad;
    const int64_t nalm = (int64_t)(L + 1) * (L + 2) / 2;
    const size_t nout = real_layout ? (size_t)(L + 1) * (L + 1) : 2 * (size_t)nalm;
    memset(almE, 0, sizeof(double) * nout);
    memset(almB, 0, sizeof(double) * nout);
    const double t_begin = now();
#pragma omp parallel
    {
        size_t wlen = 16;
        for (int q = 0; q < p->npair; ++q) { size_t w = p->blue[q].M ? (size_t)p->blue[q].M : (size_t)p->nphi[q]; if (w > wlen) wlen = w; }
        v4d *w = (v4d *)amalloc(sizeof(v4d) * 4 * wlen);
        double *cr = (double *)malloc(sizeof(double) * 2 * (L + 1)), *ci = cr + L + 1;
#pragma omp for schedule(dynamic, 1)
        for (int blk = 0; blk < (p->npad + 7) / 8; ++blk) {
            for (int q = blk * 8; q < blk * 8 + 8; ++q) {
                if (q >= p->npair) {   /* padding lanes of the last vector */
                    memset(p->F + (size_t)q * (L + 1) * 8, 0, sizeof(double) * 8 * (size_t)(L + 1));
                    continue;
                }
                ring_phases(p, q, cr, ci);
                ring_anal_pair(p, q, cr, ci, w, wlen, mapQ, mapU, pixw, weight);
            }
        }
        free(w); free(cr);
#pragma omp master
        g_t[3] = now() - t_begin;
        v8d *acc = (v8d *)amalloc(sizeof(v8d) * 4 * (size_t)(L + 2));
        double *rab = (double *)amalloc(sizeof(double) * 2 * (size_t)(L + 2));
        v8d *st = (v8d *)amalloc(sizeof(v8d) * 12 * (size_t)(npad / VL));
        int *lst = (int *)malloc(sizeof(int) * (size_t)(npad / VL));
#pragma omp for schedule(dynamic, 1)
        for (int m = 0; m <= L; ++m) {
            const int l0 = m > 2 ? m : 2;
            if (l0 > L) continue;
            const int64_t base = (int64_t)m * (2 * L + 1 - m) / 2;
            for (int l = l0; l <= L; ++l) { rab[2 * (l - l0)] = p->ra[base + l]; rab[2 * (l - l0) + 1] = p->rb[base + l]; }
            rab[2 * (L - l0 + 1)] = rab[2 * (L - l0 + 1) + 1] = 0.0;
            leg_anal_m(p, m, rab, acc, st, lst);
            for (int l = l0; l <= L; ++l) {
                const v8d *a = acc + 4 * (l - l0);
                const double xr = hsum(a[0]), xi = hsum(a[1]), yr = hsum(a[2]), yi = hsum(a[3]);
                const double k = -0.5 * p->alpha[base + l];
                /* E = k (X + Y), B = -i k (X - Y) */
                double er = k * (xr + yr), ei = k * (xi + yi), br = k * (xi - yi), bi = -k * (xr - yr);
                if (flE) { er *= flE[l]; ei *= flE[l]; }
                if (flB) { br *= flB[l]; bi *= flB[l]; }
                if (!real_layout) {
                    almE[2 * (base + l)] = er; almE[2 * (base + l) + 1] = ei;
                    almB[2 * (base + l)] = br; almB[2 * (base + l) + 1] = bi;
                } else if (m == 0) {
                    almE[l] = er; almB[l] = br;
                } else {
                    const int64_t o = 2 * (base + l) - (L + 1);
                    almE[o] = er * M_SQRT2; almE[o + 1] = ei * M_SQRT2; almB[o] = br * M_SQRT2; almB[o + 1] = bi * M_SQRT2;
                }
            }
        }
        free(acc); free(rab); free(st); free(lst);
    }
    g_t[2] = now() - t_begin - g_t[3];
    return 0;
}

int orf_map2alm_spin2(int nside, int lmax, const double *mapQ, const double *mapU, double *almE, double *almB, double weight)
{
    return anal_impl(nside, lmax, mapQ, mapU, NULL, 0, NULL, NULL, almE, almB, weight);
}

/* maps (times the per-pixel weights pixw, nullable) -> real-layout coefficients times per-l factors and `weight`:
 * complex_to_real(almxfl(map2alm(map * pixw), fl)) * weight / (4 pi / Npix) in one call (weight = 1: A^T) */
int orf_adjoint_real(int nside, int lmax, const double *mapQ, const double *mapU, const double *pixw, const double *flE,
                     const double *flB, double weight, double *outE, double *outB)
{
    return anal_impl(nside, lmax, mapQ, mapU, pixw, 1, flE, flB, outE, outB, weight);
}

/* sum_p w_p [(dQ_p - q_p)^2 + (dU_p - u_p)^2]: the pixel-space chi^2 of NonCenteredGibbs.py:353-355 in one pass */
double orf_chi2(const double *dQ, const double *dU, const double *q, const double *u, const double *w, int64_t npix)
{
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < npix; ++i) {
        const double a = dQ[i] - q[i], b = dU[i] - u[i];
        s += w[i] * (a * a + b * b);
    }
    return s;
}

void orf_last_times(double *t4) { for (int i = 0; i < 4; ++i) t4[i] = g_t[i]; }

/* test hook: forward DFT of one complex sequence of arbitrary length n through the ring-FFT code path */
int orf_test_fft(int n, const double *inr, const double *ini, double *outr, double *outi)
{
    dispatch_init();
    fplan tmp; memset(&tmp, 0, sizeof(tmp));
    int M = n;
    if (!is_pow2(n)) { M = 1; while (M < 2 * n - 1) M *= 2; }
    tmp.twn = M;
    tmp.twr = (double *)amalloc(sizeof(double) * M); tmp.twi = (double *)amalloc(sizeof(double) * M);
    for (int k = 0; k < M; ++k) { long double a = 2.0L * PI_L * k / M; tmp.twr[k] = (double)cosl(a); tmp.twi[k] = (double)(-sinl(a)); }
    blue_t b; memset(&b, 0, sizeof(b)); b.n = n;
    v4d *w = (v4d *)amalloc(sizeof(v4d) * 4 * M);
    if (!is_pow2(n)) {
        b.M = M;
        b.cr = (double *)amalloc(sizeof(double) * n); b.ci = (double *)amalloc(sizeof(double) * n);
        b.br = (double *)amalloc(sizeof(double) * M); b.bi = (double *)amalloc(sizeof(double) * M);
        for (int j = 0; j < n; ++j) { int64_t r = ((int64_t)j * j) % (2 * n); long double a = PI_L * r / n; b.cr[j] = (double)cosl(a); b.ci[j] = (double)(-sinl(a)); }
        v4d *xr = w, *xi = w + M, *yr = w + 2 * M, *yi = w + 3 * M;
        for (int j = 0; j < M; ++j) { xr[j] = (v4d){0, 0, 0, 0}; xi[j] = (v4d){0, 0, 0, 0}; }
        for (int j = 0; j < n; ++j) { v4d vr = {b.cr[j], 0, 0, 0}, vi = {-b.ci[j], 0, 0, 0}; xr[j] = vr; xi[j] = vi; if (j) { xr[M - j] = vr; xi[M - j] = vi; } }
        v4d *orr, *oii;
        fft_pow2(M, xr, xi, yr, yi, tmp.twr, tmp.twi, tmp.twn, &orr, &oii);
        for (int j = 0; j < M; ++j) { b.br[j] = orr[j][0] / M; b.bi[j] = oii[j][0] / M; }
    }
    v4d *xr = w, *xi = w + M, *yr = w + 2 * M, *yi = w + 3 * M;
    for (int j = 0; j < n; ++j) { xr[j] = (v4d){inr[j], 0, 0, 0}; xi[j] = (v4d){ini[j], 0, 0, 0}; }
    v4d *orr, *oii;
    ring_fft(&tmp, &b, xr, xi, yr, yi, &orr, &oii);
    for (int k = 0; k < n; ++k) { outr[k] = orr[k][0]; outi[k] = oii[k][0]; }
    free(w); free(tmp.twr); free(tmp.twi); free(b.cr); free(b.ci); free(b.br); free(b.bi);
    return 0;
}

int orf_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* widest vector ISA the hot loops were dispatched to on this host: 512, 256 or 128 (bits) */
int orf_simd_bits(void)
{
    dispatch_init();
    return g_simd_bits;
}

/* measured DFMA peak of the host with the dispatched vector ISA: GFLOP/s per core while all `threads` cores run the
 * loop together (the clock a vector-heavy job really gets), for about `seconds` of work */
double orf_fma_peak_gflops_per_core(double seconds)
{
    dispatch_init();
    volatile double sink = 0;
    long iters = 1000000;
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        double t0 = now();
#pragma omp parallel
        {
            double r = fma_peak_loop(iters, 0.999999, 1e-9);
#pragma omp atomic
            sink += r;
        }
        double dt = now() - t0;
        double g = (double)iters * 16 * VL * 2 / dt * 1e-9;
        if (g > best) best = g;
        if (dt < seconds / 6 && rep < 2) iters = (long)(iters * (seconds / 6) / (dt > 1e-6 ? dt : 1e-6));
    }
    return best;
}

void orf_free_plan(void) { plan_free(g_plan); g_plan = NULL; }
