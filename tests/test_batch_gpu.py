"""Chain batches (BASELINE config #5 "batched over chains", north_star (a)): two right-hand sides per launch share one
Legendre recurrence.  A batch must give what separate calls give -- the transforms against the CPU oracle and against the
single-chain kernels, the batched PCG against two single solves, and the batched constrained-realization draw against
sample_mask on the same injected Gaussians (reference call sites: CenteredGibbs.py:448-491)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import sht as O

pytestmark = pytest.mark.gpu


def dev(x):
    return torch.as_tensor(x, device="cuda")


def rand_alm(lmax, rng, lmin=0):
    a = rng.standard_normal(O.nalm(lmax)) + 1j * rng.standard_normal(O.nalm(lmax))
    a[:lmax + 1] = a[:lmax + 1].real
    ell = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])
    a[ell < lmin] = 0
    return a


def to_real(a, lmax):
    r = np.empty((lmax + 1) ** 2)
    r[:lmax + 1] = a[:lmax + 1].real
    r[lmax + 1::2] = a[lmax + 1:].real * np.sqrt(2)
    r[lmax + 2::2] = a[lmax + 1:].imag * np.sqrt(2)
    return r


def relerr(got, ref):
    return float(np.abs(got - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize("nside,lmax,k", [(4, 8, 2), (16, 47, 3), (32, 64, 2), (64, 128, 5), (128, 256, 2)])
def test_batched_transforms_vs_oracle_and_single_calls(nside, lmax, k):
    from gibbssampler_b200.sht import Plan
    plan = Plan(nside, lmax)   # own plan: the batch re-sizes the plan's workspaces
    rng = np.random.default_rng(nside + k)
    es = [rand_alm(lmax, rng, 2) for _ in range(k)]
    bs = [rand_alm(lmax, rng, 2) for _ in range(k)]
    ts = [rand_alm(lmax, rng) for _ in range(k)]
    # single-chain results first (before the plan switches to batch-sized buffers), then the batch, then single again
    single = [plan.alm2map_spin2(dev(e), dev(b)) for e, b in zip(es, bs)]
    single = [(q.cpu().numpy(), u.cpu().numpy()) for q, u in single]
    q, u = plan.alm2map_spin2_batch(dev(np.stack(es)), dev(np.stack(bs)))
    qr, ur = plan.alm2map_spin2_batch(dev(np.stack([to_real(e, lmax) for e in es])), dev(np.stack([to_real(b, lmax) for b in bs])))
    t = plan.alm2map_batch(dev(np.stack(ts)))
    for c in range(k):
        rq, ru = O.alm2map_spin2(es[c], bs[c], nside, lmax)
        assert relerr(q[c].cpu().numpy(), rq) < 1e-10 and relerr(u[c].cpu().numpy(), ru) < 1e-10
        assert relerr(qr[c].cpu().numpy(), rq) < 1e-10 and relerr(ur[c].cpu().numpy(), ru) < 1e-10
        assert relerr(q[c].cpu().numpy(), single[c][0]) < 1e-13 and relerr(u[c].cpu().numpy(), single[c][1]) < 1e-13
        assert relerr(t[c].cpu().numpy(), O.alm2map(ts[c], nside, lmax)) < 1e-10
    again = plan.alm2map_spin2(dev(es[0]), dev(bs[0]))
    assert relerr(again[0].cpu().numpy(), single[0][0]) < 1e-14
    # analysis
    npix = 12 * nside ** 2
    fq, fu = rng.standard_normal((k, npix)), rng.standard_normal((k, npix))
    w = rng.uniform(0.5, 1.5, npix)
    bl = O.gauss_beam(np.radians(2.0), lmax)
    ge, gb = plan.map2alm_spin2_batch(dev(fq), dev(fu))
    ae, ab = plan.map2alm_spin2_batch(dev(fq), dev(fu), adjoint=True, pixw=dev(w), fl=dev(bl), real_layout=True)
    g0 = plan.map2alm_batch(dev(fq), adjoint=True, real_layout=True)
    for c in range(k):
        re_, rb_ = O.map2alm_spin2(fq[c], fu[c], nside, lmax)
        assert relerr(ge[c].cpu().numpy(), re_) < 1e-10 and relerr(gb[c].cpu().numpy(), rb_) < 1e-10
        se, sb = plan.map2alm_spin2(dev(fq[c]), dev(fu[c]), adjoint=True, pixw=dev(w), fl=dev(bl), real_layout=True)
        assert relerr(ae[c].cpu().numpy(), se.cpu().numpy()) < 1e-12 and relerr(ab[c].cpu().numpy(), sb.cpu().numpy()) < 1e-12
        assert relerr(g0[c].cpu().numpy(), to_real(O.map2alm(fq[c], nside, lmax, adjoint=True), lmax)) < 1e-10


def _problem(nside, lmax, rng):
    from gibbssampler_b200 import utils
    npix = 12 * nside * nside
    th, _ = O.pix_angles(nside)
    mask = (np.abs(np.cos(th)) > 0.3).astype(float)
    ell = np.arange(lmax + 1)
    dls = [{"EE": np.where(ell >= 2, 1.0 + 0.03 * ell, 0.0), "BB": np.where(ell >= 2, 0.4 + 0.01 * ell, 0.0)},
           {"EE": np.where(ell >= 2, 0.6 + 0.05 * ell, 0.0), "BB": np.where(ell >= 2, 0.2 + 0.02 * ell, 0.0)}]
    fwhm = 60.0 / nside * 4
    bl_map = utils.expand_per_l(O.gauss_beam(np.radians(fwhm), lmax))
    dQ, dU = rng.standard_normal(npix) * mask, rng.standard_normal(npix) * mask
    return npix, mask, dls, fwhm, bl_map, dQ, dU


@pytest.mark.parametrize("nside,lmax", [(8, 16), (32, 64), (64, 128)])
def test_batched_constrained_realization_equals_two_single_draws(nside, lmax):
    from gibbssampler_b200.CenteredGibbs import PolarizedCenteredConstrainedRealization, sample_mask_batch
    from gibbssampler_b200.sht import Plan
    rng = np.random.default_rng(5 + nside)
    npix, mask, dls, fwhm, bl_map, dQ, dU = _problem(nside, lmax, rng)
    nre = (lmax + 1) ** 2
    plan = Plan(nside, lmax)
    noise = np.full(npix, 0.05)
    crs = [PolarizedCenteredConstrainedRealization({"Q": dQ, "U": dU}, noise * 1e4, noise, bl_map, lmax, npix, fwhm, mask=mask,
                                                   plan=plan, seed=11 + k) for k in range(2)]
    crs[1].inv_noise_pol = crs[0].inv_noise_pol
    xis = [tuple(rng.standard_normal(n) for n in (npix, npix, nre, nre)) for _ in range(2)]
    ref = []
    for cr, d, xi in zip(crs, dls, xis):
        sol, _ = cr.sample_mask(d, xi)
        ref.append((np.asarray(sol["EE"]).copy(), np.asarray(sol["BB"]).copy(), cr.last_pcg_iterations, cr.last_pcg_residual))
    out = sample_mask_batch(crs, dls, xis)
    for k in range(2):
        sol = out[k][0]
        assert abs(crs[k].last_pcg_iterations - ref[k][2]) <= 1, (crs[k].last_pcg_iterations, ref[k][2])
        assert crs[k].last_pcg_residual <= crs[k].pcg_accuracy
        scale = max(np.abs(ref[k][0]).max(), np.abs(ref[k][1]).max())
        # two converged PCG runs of the same system: equal to the solver tolerance (1e-5 on the residual)
        assert np.abs(np.asarray(sol["EE"]) - ref[k][0]).max() < 2e-4 * scale
        assert np.abs(np.asarray(sol["BB"]) - ref[k][1]).max() < 2e-4 * scale
    # a single solve on the same plan still works after the batch and reproduces itself bit for bit
    sol, _ = crs[0].sample_mask(dls[0], xis[0])
    assert np.array_equal(np.asarray(sol["EE"]), ref[0][0])


def test_batched_pcg_first_iterations_match_single_chain_exactly_in_count():
    """itermax cut: after n iterations the batched solver has done n iterations on both chains (n_iter_out) and the iterates
    agree with the single-chain solver's to rounding (same alpha / beta recurrences, mat-vec equal to ~1e-13)."""
    from gibbssampler_b200 import _lib
    from gibbssampler_b200._dev import f64, ptr, stream
    from gibbssampler_b200.CenteredGibbs import PolarizedCenteredConstrainedRealization
    from gibbssampler_b200.sht import Plan
    nside, lmax = 32, 64
    rng = np.random.default_rng(77)
    npix, mask, dls, fwhm, bl_map, dQ, dU = _problem(nside, lmax, rng)
    nre = (lmax + 1) ** 2
    plan = Plan(nside, lmax)
    noise = np.full(npix, 0.05)
    cr = PolarizedCenteredConstrainedRealization({"Q": dQ, "U": dU}, noise * 1e4, noise, bl_map, lmax, npix, fwhm, mask=mask, plan=plan, seed=3)
    L = _lib.lib()
    rhs = f64(rng.standard_normal((2, 2, nre)))
    dl = f64(np.stack([np.stack([dls[0]["EE"], dls[1]["EE"]]), np.stack([dls[0]["BB"], dls[1]["BB"]])]))
    nit_cut = 7
    single = []
    for k in range(2):
        xe, xb = torch.empty(nre, dtype=torch.float64, device="cuda"), torch.empty(nre, dtype=torch.float64, device="cuda")
        nit, res = C.c_int(0), C.c_double(0)
        rc = L.gs_cr_pcg_pol(plan._h, ptr(dl[0, k]), ptr(dl[1, k]), ptr(cr.bl_gauss_d), ptr(cr.inv_noise_pol), cr.ninv_sum_over_4pi,
                             ptr(rhs[k, 0]), ptr(rhs[k, 1]), ptr(xe), ptr(xb), 0, 1e-12, nit_cut, 8, C.byref(nit), C.byref(res), stream())
        assert rc == -3 and nit.value == nit_cut
        single.append((xe.cpu().numpy(), xb.cpu().numpy(), res.value))
    x = torch.empty_like(rhs)
    nit2, res2 = (C.c_int * 2)(), (C.c_double * 2)()
    rc = L.gs_cr_pcg_pol_batch(plan._h, 2, ptr(dl[0]), ptr(dl[1]), ptr(cr.bl_gauss_d), ptr(cr.inv_noise_pol), cr.ninv_sum_over_4pi,
                               ptr(rhs[0, 0]), ptr(rhs[0, 1]), ptr(x[0, 0]), ptr(x[0, 1]), 2 * nre, 1e-12, nit_cut, 8, nit2, res2, stream())
    assert rc == -3 and list(nit2) == [nit_cut, nit_cut]
    for k in range(2):
        assert relerr(x[k, 0].cpu().numpy(), single[k][0]) < 1e-9 and relerr(x[k, 1].cpu().numpy(), single[k][1]) < 1e-9
        assert abs(res2[k] - single[k][2]) < 1e-8 * single[k][2]


def test_batched_transforms_on_split_ring_path(monkeypatch):
    """Rings too long for one CTA (nside >= 1024 in production) go through a shared scratch buffer: a chain batch shares the
    Legendre recurrence and runs that ring stage one chain after the other.  GS_RING_MCAP forces the path at nside 32."""
    from gibbssampler_b200.sht import Plan
    nside, lmax, k = 32, 64, 3
    monkeypatch.setenv("GS_RING_MCAP", "64")
    plan = Plan(nside, lmax)
    monkeypatch.delenv("GS_RING_MCAP")
    rng = np.random.default_rng(8)
    es = [rand_alm(lmax, rng, 2) for _ in range(k)]
    bs = [rand_alm(lmax, rng, 2) for _ in range(k)]
    q, u = plan.alm2map_spin2_batch(dev(np.stack(es)), dev(np.stack(bs)))
    npix = 12 * nside ** 2
    fq, fu = rng.standard_normal((k, npix)), rng.standard_normal((k, npix))
    ge, gb = plan.map2alm_spin2_batch(dev(fq), dev(fu))
    for c in range(k):
        rq, ru = O.alm2map_spin2(es[c], bs[c], nside, lmax)
        assert relerr(q[c].cpu().numpy(), rq) < 1e-10 and relerr(u[c].cpu().numpy(), ru) < 1e-10
        re_, rb_ = O.map2alm_spin2(fq[c], fu[c], nside, lmax)
        assert relerr(ge[c].cpu().numpy(), re_) < 1e-10 and relerr(gb[c].cpu().numpy(), rb_) < 1e-10


def test_run_chains_reproduces_each_chain_own_run():
    """multichain.run_chains (two chains per GPU, batched constrained realizations) against the chains' own run(): same Philox
    seeds -> the same right-hand sides; the batched PCG stops at the same tolerance, so the D_l histories agree to the solver
    tolerance (and exactly in the accept flags of the Metropolis sweep)."""
    from gibbssampler_b200.multichain import run_chains
    from gibbssampler_b200.PNCP import PNCPGibbs
    nside, lmax, l_cut, n_iter = 8, 16, 4, 4
    rng = np.random.default_rng(21)
    npix, mask, dls, fwhm, bl_map, dQ, dU = _problem(nside, lmax, rng)
    ell = np.arange(lmax + 1)
    bins = {"EE": np.arange(0, lmax + 2), "BB": np.arange(0, lmax + 2)}
    blocks = {"EE": [4, 8, 12, lmax + 1], "BB": [4, 8, 12, lmax + 1]}
    pv = {"EE": np.full(lmax - 1, 0.05), "BB": np.full(lmax - 1, 0.05)}
    nt, npol = np.full(npix, 1.0), np.full(npix, 0.05)
    init = [{"EE": np.where(ell >= 2, 1.0, 0.0), "BB": np.where(ell >= 2, 0.5, 0.0)},
            {"EE": np.where(ell >= 2, 0.7, 0.0), "BB": np.where(ell >= 2, 0.3, 0.0)}]

    def make(seed):
        g = PNCPGibbs({"Q": dQ, "U": dU}, nt, fwhm, nside, lmax, npix, pv, l_cut, metropolis_blocks=blocks, polarization=True, bins=bins,
                      n_iter=n_iter, noise_Q=npol, mask=mask, seed=seed)
        g.constrained_sampler.pcg_accuracy = 1e-12
        return g
    ref = [make(41 + k).run(init[k]) for k in range(2)]
    got = run_chains([make(41 + k) for k in range(2)], init)
    for k in range(2):
        for pol in ("EE", "BB"):
            assert np.array_equal(np.asarray(got[k][1][pol]), np.asarray(ref[k][1][pol])), (k, pol)
            a, b = np.asarray(got[k][0][pol]), np.asarray(ref[k][0][pol])
            assert a.shape == b.shape == (n_iter + 1, lmax + 1)
            assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max(), (k, pol, np.abs(a - b).max())
