"""ctypes front-end of oracle/sht_oracle.c (TEST INFRASTRUCTURE ONLY).

Restates the healpy calls of the reference's hot path: hp.alm2map,
hp.map2alm(iter=k, use_weights=False), hp.alm2cl, hp.almxfl, hp.gauss_beam, and the
HEALPix RING geometry.  Parity against healpy itself is UNPINNED (healpy not installable
here); pinned by analytic answers / scipy / sympy in tests/test_oracle_sht.py.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build():
    subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)


def _lib(kind="ld"):
    if kind not in _LIBS:
        path = os.path.join(_HERE, "_build", "liboracle_%s.so" % kind)
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        lib.orc_alm2map.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp]
        lib.orc_map2alm.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, C.c_double]
        lib.orc_lambda.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, dp]
        lib.orc_lambda_zs.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, dp]
        lib.orc_ring_info.argtypes = [C.c_int, C.c_int, dp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int64)]
        _LIBS[kind] = lib
    return _LIBS[kind]


def _fast():
    """oracle/sht_fast.c: the vectorised spin-2 pair (libsharp-style stand-in for healpy on the host cores)."""
    if "fast" not in _LIBS:
        path = os.path.join(_HERE, "_build", "liboracle_fast.so")
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        lib.orf_alm2map_spin2.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp]
        lib.orf_map2alm_spin2.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, C.c_double]
        lib.orf_test_fft.argtypes = [C.c_int, dp, dp, dp, dp]
        lib.orf_chi2.argtypes = [dp, dp, dp, dp, dp, C.c_int64]
        lib.orf_chi2.restype = C.c_double
        lib.orf_synth_real.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, dp]
        lib.orf_adjoint_real.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, C.c_double, dp, dp]
        _LIBS["fast"] = lib
    return _LIBS["fast"]


def num_threads(kind="f64"):
    return _fast().orf_num_threads() if kind == "fast" else _lib(kind).orc_num_threads()


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def nalm(lmax):
    return (lmax + 1) * (lmax + 2) // 2


def alm_index(lmax, l, m):
    """healpy m-major index (variance_expension.pyx:19, utils.py:123)."""
    return m * (2 * lmax + 1 - m) // 2 + l


def ring_info(nside, ring):
    z, s, p = C.c_double(), C.c_double(), C.c_double()
    n, st = C.c_int(), C.c_int64()
    rc = _lib().orc_ring_info(nside, ring, C.byref(z), C.byref(s), C.byref(p), C.byref(n), C.byref(st))
    if rc:
        raise ValueError("bad ring")
    return z.value, s.value, p.value, n.value, st.value


def pix_angles(nside):
    """(theta, phi) of every RING pixel."""
    npix = 12 * nside * nside
    th, ph = np.empty(npix), np.empty(npix)
    for r in range(1, 4 * nside):
        z, s, p0, n, st = ring_info(nside, r)
        th[st:st + n] = np.arctan2(s, z)
        ph[st:st + n] = p0 + 2 * np.pi * np.arange(n) / n
    return th, ph


def lam(lmax, m, mp, z):
    out = np.zeros(lmax + 1)
    if _lib().orc_lambda(lmax, m, mp, float(z), _p(out)):
        raise ValueError("bad args")
    return out


def lam_zs(lmax, m, mp, z, sth):
    """lambda^{mp}_{lm} from cos(theta) AND sin(theta) as doubles (libsharp's inputs; see orc_lambda_zs)."""
    out = np.zeros(lmax + 1)
    if _lib().orc_lambda_zs(lmax, m, mp, float(z), float(sth), _p(out)):
        raise ValueError("bad args")
    return out


def alm2map(alm, nside, lmax, kind="ld"):
    """hp.alm2map(alm, nside, lmax) for one spin-0 field."""
    a = np.ascontiguousarray(alm, dtype=np.complex128)
    assert a.shape == (nalm(lmax),)
    out = np.empty(12 * nside * nside)
    rc = _lib(kind).orc_alm2map(nside, lmax, 0, _p(a.view(np.float64)), None, _p(out), None)
    assert rc == 0
    return out


def alm2map_spin2(almE, almB, nside, lmax, kind="ld"):
    """(Q, U) of hp.alm2map([0, E, B], pol=True)."""
    e = np.ascontiguousarray(almE, dtype=np.complex128)
    b = np.ascontiguousarray(almB, dtype=np.complex128)
    q, u = np.empty(12 * nside * nside), np.empty(12 * nside * nside)
    if kind == "fast":
        rc = _fast().orf_alm2map_spin2(nside, lmax, _p(e.view(np.float64)), _p(b.view(np.float64)), _p(q), _p(u))
        assert rc == 0
        return q, u
    rc = _lib(kind).orc_alm2map(nside, lmax, 2, _p(e.view(np.float64)), _p(b.view(np.float64)), _p(q), _p(u))
    assert rc == 0
    return q, u


def _map2alm0(m, nside, lmax, weight, kind):
    f = np.ascontiguousarray(m, dtype=np.float64)
    a = np.zeros(nalm(lmax), dtype=np.complex128)
    rc = _lib(kind).orc_map2alm(nside, lmax, 0, _p(f), None, _p(a.view(np.float64)), None, weight)
    assert rc == 0
    return a


def _map2alm2(q, u, nside, lmax, weight, kind):
    q = np.ascontiguousarray(q, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    e = np.zeros(nalm(lmax), dtype=np.complex128)
    b = np.zeros(nalm(lmax), dtype=np.complex128)
    if kind == "fast":
        rc = _fast().orf_map2alm_spin2(nside, lmax, _p(q), _p(u), _p(e.view(np.float64)), _p(b.view(np.float64)), weight)
        assert rc == 0
        return e, b
    rc = _lib(kind).orc_map2alm(nside, lmax, 2, _p(q), _p(u), _p(e.view(np.float64)), _p(b.view(np.float64)), weight)
    assert rc == 0
    return e, b


def map2alm(m, nside, lmax, iter=0, adjoint=False, kind="ld"):
    """hp.map2alm(m, lmax, iter=iter, use_weights=False); adjoint=True gives A^T (weight 1)."""
    w = 1.0 if adjoint else 4 * np.pi / (12 * nside * nside)
    a = _map2alm0(m, nside, lmax, w, kind)
    for _ in range(0 if adjoint else iter):
        a = a + _map2alm0(np.asarray(m) - alm2map(a, nside, lmax, kind), nside, lmax, w, kind)
    return a


def map2alm_spin2(q, u, nside, lmax, iter=0, adjoint=False, kind="ld"):
    """(E, B) of hp.map2alm([0, Q, U], pol=True, iter=iter, use_weights=False)."""
    w = 1.0 if adjoint else 4 * np.pi / (12 * nside * nside)
    e, b = _map2alm2(q, u, nside, lmax, w, kind)
    for _ in range(0 if adjoint else iter):
        q2, u2 = alm2map_spin2(e, b, nside, lmax, kind)
        de, db = _map2alm2(np.asarray(q) - q2, np.asarray(u) - u2, nside, lmax, w, kind)
        e, b = e + de, b + db
    return e, b


def synth_real_fast(sE, sB, flE, flB, nside, lmax):
    """alm2map_spin2(almxfl(real_to_complex(s), fl)) in one call of the vectorised port (real layout in, per-l factors)."""
    sE, sB = np.ascontiguousarray(sE, dtype=np.float64), np.ascontiguousarray(sB, dtype=np.float64)
    flE = None if flE is None else np.ascontiguousarray(flE, dtype=np.float64)
    flB = None if flB is None else np.ascontiguousarray(flB, dtype=np.float64)
    q, u = np.empty(12 * nside * nside), np.empty(12 * nside * nside)
    rc = _fast().orf_synth_real(nside, lmax, _p(sE), _p(sB), _p(flE), _p(flB), _p(q), _p(u))
    assert rc == 0
    return q, u


def adjoint_real_fast(q, u, pixw, flE, flB, nside, lmax, weight=1.0):
    """complex_to_real(almxfl(A^T (map * pixw), fl)) * weight in one call of the vectorised port (real layout out)."""
    q, u = np.ascontiguousarray(q, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
    pixw = None if pixw is None else np.ascontiguousarray(pixw, dtype=np.float64)
    flE = None if flE is None else np.ascontiguousarray(flE, dtype=np.float64)
    flB = None if flB is None else np.ascontiguousarray(flB, dtype=np.float64)
    e, b = np.empty((lmax + 1) ** 2), np.empty((lmax + 1) ** 2)
    rc = _fast().orf_adjoint_real(nside, lmax, _p(q), _p(u), _p(pixw), _p(flE), _p(flB), weight, _p(e), _p(b))
    assert rc == 0
    return e, b


def chi2_fast(dQ, dU, q, u, w):
    """sum_p w_p [(dQ - q)^2 + (dU - u)^2] in one pass (the two np.sum of NonCenteredGibbs.py:353-355)."""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (dQ, dU, q, u, w)]
    return float(_fast().orf_chi2(*[_p(a) for a in arrs], arrs[0].size))


def alm2cl(alm, lmax):
    """hp.alm2cl: (|a_l0|^2 + 2 sum_{m>0} |a_lm|^2) / (2l+1)."""
    cl = np.zeros(lmax + 1)
    for m in range(lmax + 1):
        seg = alm[alm_index(lmax, m, m):alm_index(lmax, lmax, m) + 1]
        cl[m:] += (1.0 if m == 0 else 2.0) * (seg.real ** 2 + seg.imag ** 2)
    return cl / (2 * np.arange(lmax + 1) + 1)


def almxfl(alm, fl, lmax):
    """hp.almxfl: a_lm * f_l."""
    out = np.array(alm, dtype=np.complex128)
    for m in range(lmax + 1):
        i0 = alm_index(lmax, m, m)
        out[i0:i0 + lmax - m + 1] *= fl[m:lmax + 1]
    return out


def gauss_beam(fwhm, lmax):
    """hp.gauss_beam(fwhm, lmax) (temperature beam): exp(-l(l+1) sigma^2 / 2)."""
    sigma = fwhm / np.sqrt(8.0 * np.log(2.0))
    ell = np.arange(lmax + 1)
    return np.exp(-0.5 * ell * (ell + 1) * sigma ** 2)
