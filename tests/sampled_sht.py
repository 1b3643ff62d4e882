"""Direct-sum references for SHT checks at sizes where a full CPU transform is too slow (TEST INFRASTRUCTURE).

A synthesis restricted to a few m is  f(theta, phi) = sum_m (2 - delta_m0) Re[F_m(theta) e^{i m phi}]  with
F_m = sum_l a_lm lambda_lm(theta); the lambda come from the oracle's long-double recurrence (oracle.sht.lam), one
(ring, m) at a time, so any ring of any nside can be checked in O(lmax).  Conventions as hp.alm2map / hp.map2alm
(HEALPix: Q +- iU = sum -(E +- iB) (+-2)Y_lm); pinned to the oracle's full transforms by tests/test_oracle_sht.py.
"""
import numpy as np

from oracle import sht as O


def sampled_m(lmax):
    return sorted({m for m in (0, 1, 2, 3, 17, lmax // 4 - 1, lmax // 2 - 1, lmax // 2, 3 * lmax // 4 + 5, lmax - 96, lmax - 1, lmax)
                   if 0 <= m <= lmax})


def sampled_alm(lmax, ms, rng):
    """(E, B, T) complex alm that vanish outside the m of `ms`, and their per-m coefficient slices."""
    e, b, t = (np.zeros(O.nalm(lmax), complex) for _ in range(3))
    coef = {}
    for m in ms:
        i0 = O.alm_index(lmax, m, m)
        n = lmax - m + 1
        for arr, lmin in ((e, 2), (b, 2), (t, 0)):
            v = rng.standard_normal(n) + 1j * rng.standard_normal(n)
            if m == 0:
                v = v.real + 0j
            v[:max(0, lmin - m)] = 0
            arr[i0:i0 + n] = v
        coef[m] = (e[i0:i0 + n], b[i0:i0 + n], t[i0:i0 + n])
    return e, b, t, coef


def _basis(lmax, m, z, sth):
    """(F1, F2, lambda^0)_{l >= m} of one ring.  cos(theta) and sin(theta) enter as the doubles of the ring table (as in libsharp
    and in the CUDA plan): at ring 1 of nside 2048 the rounding of cos(theta) to a double alone moves lambda_lm by up to
    l ulp / sin(theta) ~ 5e-10 at l = 4096, so the FP64 transform under test and its reference must share it."""
    lp, lm_, l0 = O.lam_zs(lmax, m, -2, z, sth)[m:], O.lam_zs(lmax, m, 2, z, sth)[m:], O.lam_zs(lmax, m, 0, z, sth)[m:]
    return 0.5 * (lp + lm_), 0.5 * (lp - lm_), l0


def synthesis_error(nside, lmax, ms, coef, rings, q, u, t, rng, npick=48):
    """max |map - direct sum| / max |map| over sampled pixels of `rings` for the maps (q, u) = alm2map_spin2, t = alm2map."""
    worst = 0.0
    scale, scale_t = max(np.abs(q).max(), np.abs(u).max()), np.abs(t).max()
    for ring in rings:
        z, sth, phi0, nphi, start = O.ring_info(nside, ring)
        j = np.unique(rng.integers(0, nphi, size=min(npick, nphi)))
        phi = phi0 + 2 * np.pi * j / nphi
        rq, ru, rt = np.zeros(len(j)), np.zeros(len(j)), np.zeros(len(j))
        for m in ms:
            ce, cb, ct = coef[m]
            f1, f2, l0 = _basis(lmax, m, z, sth)
            fq = np.sum(-ce * f1 - 1j * cb * f2)
            fu = np.sum(-cb * f1 + 1j * ce * f2)
            ft = np.sum(ct * l0)
            w = 1.0 if m == 0 else 2.0
            ph = np.exp(1j * m * phi)
            rq += w * (fq * ph).real
            ru += w * (fu * ph).real
            rt += w * (ft * ph).real
        worst = max(worst, np.abs(q[start + j] - rq).max() / scale, np.abs(u[start + j] - ru).max() / scale,
                    np.abs(t[start + j] - rt).max() / scale_t)
    return worst


def ring_supported_maps(nside, trings, rng):
    fq, fu = np.zeros(12 * nside ** 2), np.zeros(12 * nside ** 2)
    geo = {}
    for ring in trings:
        z, sth, phi0, nphi, start = O.ring_info(nside, ring)
        fq[start:start + nphi] = rng.standard_normal(nphi)
        fu[start:start + nphi] = rng.standard_normal(nphi)
        geo[ring] = (z, sth, phi0 + 2 * np.pi * np.arange(nphi) / nphi, start, nphi)
    return fq, fu, geo


def analysis_error(nside, lmax, ms, trings, fq, fu, geo, ge, gb, gt):
    """max error of (ge, gb) = A^T_spin2 (fq, fu) and gt = A^T_spin0 fq (complex healpy layout, weight 1) on the m of `ms`."""
    sc_e, sc_t = max(np.abs(ge).max(), np.abs(gb).max()), np.abs(gt).max()
    worst = 0.0
    for m in ms:
        n = lmax - m + 1
        re_, rb_, rt = np.zeros(n, complex), np.zeros(n, complex), np.zeros(n, complex)
        for ring in trings:
            z, sth, phi, start, nphi = geo[ring]
            ph = np.exp(-1j * m * phi)
            gq, gu = np.sum(fq[start:start + nphi] * ph), np.sum(fu[start:start + nphi] * ph)
            f1, f2, l0 = _basis(lmax, m, z, sth)
            # conjugate transpose of the synthesis: E = sum -f1 G^Q - i f2 G^U,  B = sum -f1 G^U + i f2 G^Q
            re_ += -f1 * gq - 1j * f2 * gu
            rb_ += -f1 * gu + 1j * f2 * gq
            rt += l0 * gq
        if m == 0:
            re_, rb_, rt = re_.real + 0j, rb_.real + 0j, rt.real + 0j
        i0 = O.alm_index(lmax, m, m)
        worst = max(worst, np.abs(ge[i0:i0 + n] - re_).max() / sc_e, np.abs(gb[i0:i0 + n] - rb_).max() / sc_e,
                    np.abs(gt[i0:i0 + n] - rt).max() / sc_t)
    return worst
