"""Pins the CPU oracle's SHT (oracle/sht_oracle.c) with references that share no code with it:
analytic answers, scipy.special.sph_harm_y, Wigner's closed-form d-matrix,
sympy's Rotation.d, plus adjointness / round-trip properties."""
import numpy as np
import pytest
from scipy.special import sph_harm_y

from oracle import sht
from tests import wigner_ref as W


def rand_alm(lmax, rng, lmin=0):
    a = rng.standard_normal(sht.nalm(lmax)) + 1j * rng.standard_normal(sht.nalm(lmax))
    a[:lmax + 1] = a[:lmax + 1].real
    for m in range(lmax + 1):
        for l in range(m, min(lmin, lmax + 1)):
            a[sht.alm_index(lmax, l, m)] = 0
    return a


def test_ring_geometry_matches_healpix_definition():
    for nside in (1, 2, 4, 16):
        npix = 12 * nside * nside
        tot, starts = 0, []
        for r in range(1, 4 * nside):
            z, s, p0, n, st = sht.ring_info(nside, r)
            assert st == tot
            tot += n
            assert abs(z * z + s * s - 1) < 1e-15
            i = min(r, 4 * nside - r)
            if i < nside:
                assert n == 4 * i and abs(abs(z) - (1 - i * i / (3.0 * nside ** 2))) < 1e-15
                assert abs(p0 - np.pi / (4 * i)) < 1e-15
            else:
                assert n == 4 * nside and abs(abs(z) - abs(4 / 3 - 2 * i / (3.0 * nside))) < 1e-15
        assert tot == npix


def test_wigner_closed_form_agrees_with_sympy():
    from sympy.physics.quantum.spin import Rotation
    from sympy import N
    for (j, mp, m, b) in [(2, 0, 2, 0.7), (2, 1, -2, 1.9), (3, 2, 2, 0.4), (4, 3, -2, 2.5), (5, 0, 0, 1.1)]:
        assert abs(W.wigner_d(j, mp, m, b) - complex(N(Rotation.d(j, mp, m, b).doit())).real) < 1e-13


@pytest.mark.parametrize("mp", [0, 2, -2])
def test_lambda_recurrence_vs_closed_form(mp):
    lmax = 12
    for z in (0.97, 0.3, 0.0, -0.55, -0.99):
        th = np.arccos(z)
        for m in range(0, lmax + 1):
            got = sht.lam(lmax, m, mp, z)
            ref = np.array([W.lam_ref(l, m, mp, th) for l in range(lmax + 1)])
            assert np.abs(got - ref).max() < 2e-13, (mp, z, m)


def test_analytic_known_answers():
    nside, lmax = 4, 8
    th, ph = sht.pix_angles(nside)
    a = np.zeros(sht.nalm(lmax), complex)
    a[0] = np.sqrt(4 * np.pi)
    assert np.abs(sht.alm2map(a, nside, lmax) - 1).max() < 1e-15
    a[:] = 0
    a[1] = 1
    assert np.abs(sht.alm2map(a, nside, lmax) - np.sqrt(3 / 4 / np.pi) * np.cos(th)).max() < 1e-15
    a[:] = 0
    a[sht.alm_index(lmax, 1, 1)] = 0.3 + 0.7j
    ref = -2 * np.sqrt(3 / 8 / np.pi) * np.sin(th) * (0.3 * np.cos(ph) - 0.7 * np.sin(ph))
    assert np.abs(sht.alm2map(a, nside, lmax) - ref).max() < 1e-15
    e = np.zeros(sht.nalm(lmax), complex)
    b = e.copy()
    e[2] = 1
    q, u = sht.alm2map_spin2(e, b, nside, lmax)
    assert np.abs(q + 0.25 * np.sqrt(15 / 2 / np.pi) * np.sin(th) ** 2).max() < 1e-15
    assert np.abs(u).max() < 1e-15


def test_spin0_vs_scipy_direct_sum():
    nside, lmax = 4, 11
    rng = np.random.default_rng(1)
    a = rand_alm(lmax, rng)
    th, ph = sht.pix_angles(nside)
    ref = np.zeros(len(th))
    for m in range(lmax + 1):
        for l in range(m, lmax + 1):
            ref += (1 if m == 0 else 2) * (a[sht.alm_index(lmax, l, m)] * sph_harm_y(l, m, th, ph)).real
    got = sht.alm2map(a, nside, lmax)
    assert np.abs(got - ref).max() < 1e-12 * np.abs(ref).max()


def test_spin2_vs_closed_form_direct_sum():
    nside, lmax = 2, 6
    rng = np.random.default_rng(2)
    e, b = rand_alm(lmax, rng, 2), rand_alm(lmax, rng, 2)
    th, ph = sht.pix_angles(nside)
    qr, ur = W.direct_alm2map_spin2(e, b, lmax, th, ph, sht.alm_index)
    q, u = sht.alm2map_spin2(e, b, nside, lmax)
    assert np.abs(q - qr).max() < 1e-12 and np.abs(u - ur).max() < 1e-12


def real_dot(a, b, lmax):
    """<a,b> in the reference's real layout = sum_m (2 - delta_m0) Re(conj(a) b)"""
    w = np.full(len(a), 2.0)
    w[:lmax + 1] = 1.0
    return float(np.sum(w * (a.conj() * b).real))


@pytest.mark.parametrize("nside,lmax", [(4, 8), (8, 20), (16, 32)])
def test_adjointness_and_roundtrip(nside, lmax):
    rng = np.random.default_rng(3)
    npix = 12 * nside * nside
    a = rand_alm(lmax, rng)
    f = rng.standard_normal(npix)
    lhs = float(np.dot(sht.alm2map(a, nside, lmax), f))
    rhs = real_dot(a, sht.map2alm(f, nside, lmax, adjoint=True), lmax)
    assert abs(lhs - rhs) < 1e-12 * abs(lhs)
    e, b = rand_alm(lmax, rng, 2), rand_alm(lmax, rng, 2)
    q, u = rng.standard_normal(npix), rng.standard_normal(npix)
    mq, mu = sht.alm2map_spin2(e, b, nside, lmax)
    te, tb = sht.map2alm_spin2(q, u, nside, lmax, adjoint=True)
    lhs = float(np.dot(mq, q) + np.dot(mu, u))
    rhs = real_dot(e, te, lmax) + real_dot(b, tb, lmax)
    assert abs(lhs - rhs) < 1e-12 * abs(lhs)
    # Jacobi-refined analysis converges back to band-limited input (l <= 2 nside); pure E stays pure E
    if lmax <= 2 * nside:
        a2 = sht.map2alm(sht.alm2map(a, nside, lmax), nside, lmax, iter=8)
        assert np.abs(a2 - a).max() < 1e-3
        e2, b2 = sht.map2alm_spin2(*sht.alm2map_spin2(e, 0 * b, nside, lmax), nside, lmax, iter=8)
        assert np.abs(e2 - e).max() < 1e-3 and np.abs(b2).max() < 1e-3


def test_double_build_matches_long_double_build():
    nside, lmax = 32, 64
    rng = np.random.default_rng(4)
    e, b = rand_alm(lmax, rng, 2), rand_alm(lmax, rng, 2)
    q1, u1 = sht.alm2map_spin2(e, b, nside, lmax, kind="ld")
    q2, u2 = sht.alm2map_spin2(e, b, nside, lmax, kind="f64")
    assert np.abs(q1 - q2).max() < 1e-11 * np.abs(q1).max()
    e1, b1 = sht.map2alm_spin2(q1, u1, nside, lmax, kind="ld")
    e2, b2 = sht.map2alm_spin2(q1, u1, nside, lmax, kind="f64")
    assert np.abs(e1 - e2).max() < 1e-11 * np.abs(e1).max()


def test_alm_helpers():
    lmax = 6
    rng = np.random.default_rng(5)
    a = rand_alm(lmax, rng)
    cl = sht.alm2cl(a, lmax)
    ref = np.zeros(lmax + 1)
    for l in range(lmax + 1):
        for m in range(l + 1):
            ref[l] += (1 if m == 0 else 2) * abs(a[sht.alm_index(lmax, l, m)]) ** 2
        ref[l] /= 2 * l + 1
    assert np.allclose(cl, ref, rtol=1e-14)
    fl = rng.standard_normal(lmax + 1)
    x = sht.almxfl(a, fl, lmax)
    for l in range(lmax + 1):
        for m in range(l + 1):
            i = sht.alm_index(lmax, l, m)
            assert x[i] == a[i] * fl[l]
    bl = sht.gauss_beam(np.radians(0.5), 512)
    assert bl[0] == 1.0 and abs(bl[512] - np.exp(-0.5 * 512 * 513 * (np.radians(0.5) / 2.3548200450309493) ** 2)) < 1e-15


def test_pixel_centres_known_answers():
    """Pixel positions pinned to healpy facts that do not depend on this repo: the 12 base pixels (hp.pix2ang(1, p)) and
    the HEALPix rule that belt rings alternate between a half-pixel shift (rings nside, nside + 2, ...) and a start at
    phi = 0 (pix2ang_ring: phi = (j - fodd) pi / (2 nside)).  tests/test_healpix_io.py adds the NESTED-hierarchy check."""
    from oracle import sht as O
    th, ph = O.pix_angles(1)
    assert np.allclose(np.cos(th), [2 / 3] * 4 + [0] * 4 + [-2 / 3] * 4, atol=1e-15)
    assert np.allclose(ph, [np.pi / 4 + k * np.pi / 2 for k in range(4)] + [k * np.pi / 2 for k in range(4)]
                       + [np.pi / 4 + k * np.pi / 2 for k in range(4)], atol=1e-15)
    th, ph = O.pix_angles(2)
    # ring 1 (cap, 4 pixels), ring 2 = nside (belt, shifted), ring 3 (belt, starts at 0), ring 4 (equator, shifted)
    assert np.isclose(ph[0], np.pi / 4) and np.isclose(np.cos(th[0]), 1 - 1 / 12)
    assert np.isclose(ph[4], np.pi / 8) and np.isclose(np.cos(th[4]), 2 / 3)
    assert np.isclose(ph[12], 0.0) and np.isclose(np.cos(th[12]), 1 / 3)
    assert np.isclose(ph[20], np.pi / 8) and abs(np.cos(th[20])) < 1e-15
    assert np.isclose(ph[28], 0.0) and np.isclose(np.cos(th[28]), -1 / 3)
    for nside in (4, 8):
        th, ph = O.pix_angles(nside)
        ncap = 2 * nside * (nside - 1)
        first = ph[ncap::4 * nside][:2 * nside + 1]           # first pixel of every belt ring
        want = np.where(np.arange(2 * nside + 1) % 2 == 0, np.pi / (4 * nside), 0.0)
        assert np.allclose(first, want, atol=1e-15)


@pytest.mark.parametrize("nside,lmax", [(8, 16), (16, 40)])
def test_sampled_direct_sum_helper_matches_full_transforms(nside, lmax):
    """tests/sampled_sht.py (the reference of the nside 1024 / 2048 GPU checks) against the oracle's own full transforms."""
    from tests import sampled_sht as S
    rng = np.random.default_rng(77 + nside)
    ms = S.sampled_m(lmax)
    e, b, t, coef = S.sampled_alm(lmax, ms, rng)
    q, u = sht.alm2map_spin2(e, b, nside, lmax)
    tm = sht.alm2map(t, nside, lmax)
    rings = list(range(1, 4 * nside))
    assert S.synthesis_error(nside, lmax, ms, coef, rings, q, u, tm, rng) < 1e-13
    # a deliberately wrong map must be seen
    assert S.synthesis_error(nside, lmax, ms, coef, rings, q, -u, tm, rng) > 1e-3
    trings = [1, 2, nside - 1, nside, 2 * nside, 3 * nside + 1, 4 * nside - 1]
    fq, fu, geo = S.ring_supported_maps(nside, trings, rng)
    ge, gb = sht.map2alm_spin2(fq, fu, nside, lmax, adjoint=True)
    gt = sht.map2alm(fq, nside, lmax, adjoint=True)
    assert S.analysis_error(nside, lmax, ms, trings, fq, fu, geo, ge, gb, gt) < 1e-13
    assert S.analysis_error(nside, lmax, ms, trings, fq, fu, geo, gb, ge, gt) > 1e-3
