"""Fused ring stage of the PCG mat-vec (ring_apply_kernel: F_m -> ring pixels in shared memory -> N^-1 -> F'_m) against
the two-kernel path (ring synthesis to maps + weighted ring analysis), and both against the oracle operator.
Covers power-of-two belt rings, Bluestein cap rings, rings shorter than lmax (staged alias fold) and longer ones
(direct fold), spin 2 and spin 0."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _apply(nside, lmax, spin, fused, seed=0):
    from gibbssampler_b200 import _dev, _lib
    from gibbssampler_b200._dev import f64, ptr, stream
    from gibbssampler_b200.sht import Plan
    L = _lib.lib()
    rng = np.random.default_rng(seed)
    nre, npix = (lmax + 1) ** 2, 12 * nside ** 2
    ell = np.arange(lmax + 1)
    dl = f64(np.where(ell >= 2, 1.0 + 0.1 * ell, 0.0))
    bl = f64(np.exp(-1e-4 * ell * (ell + 1.0)))
    invn = f64(rng.uniform(0.0, 2.0, npix) * (rng.uniform(size=npix) > 0.2))
    xe, xb = f64(rng.standard_normal(nre)), f64(rng.standard_normal(nre))
    ye, yb = torch.empty_like(xe), torch.empty_like(xb)
    plan = Plan.get(nside, lmax)
    old = L.gs_set_ring_fused(1 if fused else 0)
    try:
        if spin == 2:
            _lib.check(L.gs_cr_apply_q_pol(plan._h, ptr(dl), ptr(dl), ptr(bl), ptr(invn), ptr(xe), ptr(xb), ptr(ye), ptr(yb), stream()))
        else:
            _lib.check(L.gs_cr_apply_q_tt(plan._h, ptr(dl), ptr(bl), ptr(invn), ptr(xe), ptr(ye), stream()))
            yb.zero_()
        torch.cuda.synchronize()
    finally:
        L.gs_set_ring_fused(old)
    return ye.cpu().numpy(), yb.cpu().numpy()


@pytest.mark.parametrize("nside,lmax", [(4, 8), (8, 23), (16, 32), (16, 47), (32, 64), (64, 128), (128, 300), (256, 512)])
@pytest.mark.parametrize("spin", [2, 0])
def test_fused_ring_stage_equals_two_kernel_path(nside, lmax, spin):
    a = _apply(nside, lmax, spin, True)
    b = _apply(nside, lmax, spin, False)
    for u, v in zip(a, b):
        assert np.abs(u - v).max() <= 1e-12 * max(1.0, np.abs(v).max())


def test_fused_operator_vs_oracle():
    from oracle import reference_logic as R
    from oracle import sht as O
    nside, lmax = 16, 40
    rng = np.random.default_rng(0)
    nre, npix = (lmax + 1) ** 2, 12 * nside ** 2
    ell = np.arange(lmax + 1)
    dl = np.where(ell >= 2, 1.0 + 0.1 * ell, 0.0)
    bl = np.exp(-1e-4 * ell * (ell + 1.0))
    invn = rng.uniform(0.0, 2.0, npix) * (rng.uniform(size=npix) > 0.2)
    xe, xb = rng.standard_normal(nre), rng.standard_normal(nre)
    ye, yb = _apply(nside, lmax, 2, True)
    blm = R.expand_per_l(bl)
    q, u = R.synth_pol(xe * blm, xb * blm, nside, lmax, "ld")
    ae, ab = R.adjoint_pol(q * invn, u * invn, nside, lmax, 0, "ld")
    ic = R.safe_inv(R.generate_var_cl(dl))
    re, rb = ic * xe + blm * ae, ic * xb + blm * ab
    assert np.abs(ye - re).max() <= 1e-10 * np.abs(re).max()
    assert np.abs(yb - rb).max() <= 1e-10 * np.abs(rb).max()
