"""Independent closed-form references used to pin the oracle (tests only).

Wigner small-d by Wigner's explicit factorial sum (exact integer factorials), and the
spin-weighted harmonics  sY_lm = (-1)^s sqrt((2l+1)/4pi) d^l_{m,-s}(theta) e^{i m phi}
(Goldberg et al. 1967), independent of any recurrence.
"""
from math import factorial, sqrt, cos, sin, pi

import numpy as np


def wigner_d(j, mp, m, beta):
    """d^j_{mp,m}(beta), Wikipedia/Sakurai convention <j mp| exp(-i beta J_y) |j m>."""
    c, s = cos(beta / 2.0), sin(beta / 2.0)
    pref = sqrt(factorial(j + mp) * factorial(j - mp) * factorial(j + m) * factorial(j - m))
    tot = 0.0
    for k in range(max(0, m - mp), min(j + m, j - mp) + 1):
        den = factorial(j + m - k) * factorial(k) * factorial(mp - m + k) * factorial(j - mp - k)
        tot += (-1) ** (mp - m + k) * c ** (2 * j + m - mp - 2 * k) * s ** (mp - m + 2 * k) / den
    return pref * tot


def lam_ref(l, m, mp, theta):
    """sqrt((2l+1)/4pi) d^l_{m,mp}(theta)"""
    if l < max(abs(m), abs(mp)):
        return 0.0
    return sqrt((2 * l + 1) / (4 * pi)) * wigner_d(l, m, mp, theta)


def direct_alm2map_spin0(alm, lmax, theta, phi, idx):
    out = np.zeros(len(theta))
    for p in range(len(theta)):
        acc = 0.0
        for m in range(lmax + 1):
            w = 1.0 if m == 0 else 2.0
            for l in range(m, lmax + 1):
                acc += w * (alm[idx(lmax, l, m)] * lam_ref(l, m, 0, theta[p]) * np.exp(1j * m * phi[p])).real
        out[p] = acc
    return out


def direct_alm2map_spin2(almE, almB, lmax, theta, phi, idx):
    """Q + iU = sum_{l, m=-l..l} -(E_lm + i B_lm) 2Y_lm ; negative m by E_{l,-m} = (-1)^m conj(E_lm)."""
    Q = np.zeros(len(theta))
    U = np.zeros(len(theta))
    for p in range(len(theta)):
        acc = 0.0 + 0.0j
        for l in range(2, lmax + 1):
            for m in range(-l, l + 1):
                if m >= 0:
                    e, b = almE[idx(lmax, l, m)], almB[idx(lmax, l, m)]
                else:
                    e = (-1) ** m * np.conj(almE[idx(lmax, l, -m)])
                    b = (-1) ** m * np.conj(almB[idx(lmax, l, -m)])
                y2 = lam_ref(l, m, -2, theta[p]) * np.exp(1j * m * phi[p])
                acc_term = -(e + 1j * b) * y2
                acc += acc_term
        Q[p], U[p] = acc.real, acc.imag
        acc = 0.0
    return Q, U
