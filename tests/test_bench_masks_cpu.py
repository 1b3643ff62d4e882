"""The synthetic sky masks of bench.py (both arms, config #4 and the scripts use them) and the CPU model of the transform-free ring
stage: on a ring whose pixel weights are all equal, DFT^H diag(w) DFT of the alias-folded spectrum is n w times that spectrum
(csrc/ringfft.cu, ring_apply_kernel's constant-weight branch), and the weighted sum of squares the Metropolis sweep forms on such a
ring is invariant under the unitary DFT of Q + iU (spectral ring storage)."""
import numpy as np
import pytest

import bench
from oracle import sht as O


def _ring_slices(nside):
    n = np.r_[4 * np.arange(1, nside), np.full(2 * nside + 1, 4 * nside), 4 * np.arange(nside - 1, 0, -1)]
    s = np.r_[0, np.cumsum(n)]
    return [slice(s[i], s[i + 1]) for i in range(4 * nside - 1)]


def test_pixel_geometry_matches_the_oracle():
    for nside in (1, 2, 8, 16):
        th, ph = O.pix_angles(nside)
        assert np.abs(np.cos(th) - bench.pixel_z(nside)).max() < 1e-14
        assert np.abs(np.angle(np.exp(1j * (ph - bench.pixel_phi(nside))))).max() < 1e-13


@pytest.mark.parametrize("kind", bench.MASK_KINDS)
@pytest.mark.parametrize("nside", [16, 64, 256])
def test_mask_sky_fraction_and_range(kind, nside):
    m = bench.make_mask(nside, 0.8, kind)
    assert m.shape == (12 * nside * nside,) and m.min() == 0.0 and m.max() == 1.0
    assert abs(m.mean() - 0.8) < (0.01 if nside == 16 else 2e-3)
    assert np.array_equal(m, bench.make_mask(nside, 0.8, kind))          # deterministic: both arms build the same mask
    z = bench.pixel_z(nside)
    assert m[np.abs(z) > 0.62].min() == 1.0                             # polar caps and high latitudes are never masked
    assert m[np.abs(z) < 0.02].max() == 0.0                             # the plane always is


def test_band_mask_cuts_no_ring_and_galplane_mask_does():
    nside = 64
    sl = _ring_slices(nside)
    band, gal = bench.make_mask(nside, 0.8, "band"), bench.make_mask(nside, 0.8, "galplane")
    assert all(np.ptp(band[s]) == 0.0 for s in sl)
    cut = sum(np.ptp(gal[s]) > 0.0 for s in sl)
    idle_b = sum(not band[s].any() for s in sl)
    idle_g = sum(not gal[s].any() for s in sl)
    assert 0.2 * len(sl) < cut < 0.5 * len(sl)          # the rings near the plane are cut, caps and high-latitude belt are whole
    assert 0 < idle_g < idle_b                          # and fewer rings lie wholly inside the mask than under the band


@pytest.mark.parametrize("n,mtop", [(8, 3), (8, 4), (12, 11), (20, 33), (64, 32), (36, 100)])
def test_constant_weight_ring_needs_no_transform(n, mtop):
    """numpy model of both branches of the constant-weight path against the literal synthesis -> weights -> analysis of a ring."""
    rng = np.random.default_rng(n * 1000 + mtop)
    phi0, w = 0.37, 2.5
    F = rng.standard_normal(mtop + 1) + 1j * rng.standard_normal(mtop + 1)
    F[0] = F[0].real
    phi = phi0 + 2 * np.pi * np.arange(n) / n
    wm = np.where(np.arange(mtop + 1) == 0, 1.0, 2.0)
    x = np.real((wm * F)[None, :] * np.exp(1j * np.outer(phi, np.arange(mtop + 1)))).sum(axis=1)          # pixels of the ring
    ref = (w * x)[None, :] @ np.exp(-1j * np.outer(phi, np.arange(mtop + 1)))                            # F'_m = sum_j w x_j e^{-i m phi_j}
    ref = ref.ravel()
    # alias fold: G_k = sum_{m = k mod n} w_m/2 F_m e^{i m phi0}, X_k = G_k + conj G_{n-k}
    G = np.zeros(n, complex)
    for m in range(mtop + 1):
        G[m % n] += 0.5 * wm[m] * F[m] * np.exp(1j * m * phi0)
    X = G + np.conj(G[(-np.arange(n)) % n])
    got = n * w * X[np.arange(mtop + 1) % n] * np.exp(-1j * np.arange(mtop + 1) * phi0)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    if 2 * mtop <= n:   # the streaming branch: F'_m = n w F_m, real part at m = 0, mirror term at m = n/2
        direct = n * w * F.copy()
        direct[0] = n * w * F[0].real
        if 2 * mtop == n:
            direct[mtop] = n * w * (F[mtop] + np.conj(F[mtop] * np.exp(2j * mtop * phi0)))
        assert np.abs(direct - ref).max() <= 1e-12 * np.abs(ref).max()


def test_unitary_ring_dft_keeps_the_weighted_sum_of_squares():
    rng = np.random.default_rng(4)
    for n in (4, 12, 60, 128):
        q, u, gq, gu = rng.standard_normal((4, n))
        w = 1.7
        chi2 = w * ((q - gq) ** 2 + (u - gu) ** 2).sum()
        zt = np.fft.fft(q + 1j * u) / np.sqrt(n)
        gt = np.fft.fft(gq + 1j * gu) / np.sqrt(n)
        chi2_t = w * ((zt.real - gt.real) ** 2 + (zt.imag - gt.imag) ** 2).sum()
        assert abs(chi2 - chi2_t) <= 1e-13 * chi2
