/*
 * gibbs_b200.h -- C ABI of libgibbs_b200.so: the B200 (sm_100a) implementation of the
 * constrained-realization + C_l-sampling Gibbs step of Gabriel-Ducrocq/GibbsSampler.
 *
 * The reference has no C ABI at this boundary: its samplers call healpy, a
 * forked qcinv, scipy.stats and the Cython module variance_expension.pyx directly from
 * Python.  Each entry point below names the reference call it replaces (file:line in the
 * reference tree).  Conventions:
 *   - every pointer argument is a CALLER-OWNED DEVICE pointer unless the name ends in _host;
 *   - sizes are passed explicitly or fixed by the plan (nside, lmax);
 *   - every call takes the CUDA stream to run on as `void* stream` (a cudaStream_t, NULL =
 *     default stream) and is asynchronous with respect to the host unless stated otherwise;
 *   - the return value is a status (GS_OK = 0, negative = error); no exception crosses the
 *     ABI; gs_last_error_string() describes the last failure on the calling thread;
 *   - a plan is used by one host thread at a time; distinct plans / streams are independent.
 *
 * Layouts (bit-exact integer conventions of the reference):
 *   GS_ALM_COMPLEX  healpy m-major complex128, idx(l,m) = m(2L+1-m)/2 + l
 *                   (variance_expension.pyx:19, utils.py:123); (L+1)(L+2)/2 complex numbers.
 *   GS_ALM_REAL     the reference's "real convention" (utils.py:49-76,
 *                   variance_expension.pyx:65-100): r[i] = Re a_i for i <= L (m = 0) and
 *                   r[2i-(L+1)] = sqrt2 Re a_i, r[2i-(L+1)+1] = sqrt2 Im a_i for i > L;
 *                   (L+1)^2 doubles.
 *   maps            HEALPix RING order, float64, Npix = 12 nside^2.
 */
#ifndef GIBBS_B200_H
#define GIBBS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_OK 0
#define GS_E_BADARG (-1)
#define GS_E_CUDA (-2)
#define GS_E_NOTCONVERGED (-3)
#define GS_E_NCCL (-4)
#define GS_E_NOMEM (-5)

#define GS_ALM_COMPLEX 0
#define GS_ALM_REAL 1

typedef struct gs_plan gs_plan;

/* ---- library / plan ------------------------------------------------------------------ */
const char* gs_last_error_string(void);
int gs_version(void);

/* Builds ring geometry, FP64 recurrence tables and ring-FFT tables for (nside, lmax) on CUDA
 * device `device` (-1 = current).  Replaces the implicit healpy/libsharp geometry + plan that
 * hp.alm2map / hp.map2alm rebuild on every call.  Synchronous. */
int gs_plan_create(gs_plan** plan, int nside, int lmax, int device);
int gs_plan_destroy(gs_plan* plan);
int gs_plan_nside(const gs_plan* plan);
int gs_plan_lmax(const gs_plan* plan);
int64_t gs_plan_npix(const gs_plan* plan);
int64_t gs_plan_nalm(const gs_plan* plan);   /* (L+1)(L+2)/2 */
int64_t gs_plan_nreal(const gs_plan* plan);  /* (L+1)^2      */

/* ---- spherical-harmonic transforms ----------------------------------------------------- */
/* hp.alm2map(alm, nside, lmax)  (NonCenteredGibbs.py:179,204; ClsSampler.py:108 via
 * synthesis_hp, variance_expension.pyx:114-123).  `fl` (nullable, L+1 doubles) multiplies a_lm
 * by fl[l] first, i.e. a fused hp.almxfl(alm, fl). */
int gs_alm2map_spin0(gs_plan* plan, const double* alm, int layout, const double* fl, double* map,
                     void* stream);

/* (Q, U) of hp.alm2map([0, almE, almB], pol=True)  (CenteredGibbs.py:505-508, 698-699,
 * 751-753; NonCenteredGibbs.py:350-351; qcinv opfilt_pp.fwd_op).  HEALPix convention
 * Q +- iU = sum -(E +- iB) (+-2)Y_lm. */
int gs_alm2map_spin2(gs_plan* plan, const double* almE, const double* almB, int layout,
                     const double* fl, double* mapQ, double* mapU, void* stream);

/* hp.map2alm(map, lmax, iter=iter, use_weights=False) (utils.py:104; CenteredGibbs.py:209).
 * adjoint != 0 computes the plain transpose A^T f = (Npix/4pi) map2alm_iter0(f) instead
 * (utils.adjoint_synthesis_hp with iter = 0; config.py:72) and ignores `iter`.
 * `pixw` (nullable, Npix doubles) multiplies the map first (fused N^-1); `fl` (nullable)
 * multiplies the result by fl[l].  iter > 0 uses plan-owned scratch. */
int gs_map2alm_spin0(gs_plan* plan, const double* map, const double* pixw, int iter, int adjoint,
                     const double* fl, double* alm, int layout, void* stream);

/* (E, B) of hp.map2alm([0, Q, U], lmax, pol=True, iter=iter)  (CenteredGibbs.py:298-299, 513,
 * 717-719; utils.py:89 with iter=3; NonCenteredGibbs.py:155). */
int gs_map2alm_spin2(gs_plan* plan, const double* mapQ, const double* mapU, const double* pixw,
                     int iter, int adjoint, const double* fl, double* almE, double* almB,
                     int layout, void* stream);

/* ---- harmonic-space array utilities (all device pointers) ------------------------------ */
/* utils.real_to_complex (utils.py:49-60) / variance_expension.real_to_complex (.pyx:84-100) */
int gs_real_to_complex(const double* real_alm, double* complex_alm, int lmax, void* stream);
/* utils.complex_to_real (utils.py:63-76) / variance_expension.complex_to_real (.pyx:65-81) */
int gs_complex_to_real(const double* complex_alm, double* real_alm, int lmax, void* stream);
/* Expansion of a per-l array x[0..L] to the real alm layout ((L+1)^2 doubles).
 *   mode 0: x_l                     GibbsSampler.compute_bl_map (GibbsSampler.py:64-74), config.py:75-84
 *   mode 1: C_l = D_l 2pi/(l(l+1))  utils.generate_var_cl (utils.py:114-147), .pyx:8-33 (l = 0 copied)
 *   mode 2: 1/C_l where C_l != 0    CenteredGibbs.py:463-466
 *   mode 3: sqrt(C_l)               ASIS.py:198-203
 *   mode 4: sqrt(1/C_l) where != 0  CenteredGibbs.py:478-479, NonCenteredGibbs.py:192-194 */
int gs_expand_per_l(const double* x, int lmax, int mode, double* out, void* stream);
/* utils.unfold_bins (utils.py:150-162): np.repeat(binned, diff(bins)); bins has nbins + 1
 * ascending int32 edges, out has nout = bins[nbins] - bins[0] entries. */
int gs_unfold_bins(const double* binned, const int* bins, int nbins, double* out, int nout,
                   void* stream);
/* hp.almxfl(alm, fl) (CenteredGibbs.py:304-305, 514-515); out may alias alm. */
int gs_almxfl(const double* alm, int layout, int lmax, const double* fl, double* out, void* stream);
/* hp.alm2cl(alm, lmax) (CenteredGibbs.py:30, 61): (|a_l0|^2 + 2 sum_m |a_lm|^2) / (2l+1). */
int gs_alm2cl(const double* alm, int layout, int lmax, double* cl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GIBBS_B200_H */
