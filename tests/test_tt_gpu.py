"""Temperature-only samplers (gibbssampler_b200/Temperature.py) against the numpy restatement of the reference's TT
classes (oracle.reference_logic.TTProblem: CenteredGibbs.py:95-235, NonCenteredGibbs.py:17-101, ClsSampler.py:94-108,
recovered PNCP) on the same injected draws; FP64 tolerance 1e-10 where no iterative solver is involved."""
import numpy as np
import pytest
import torch

from oracle import reference_logic as R

pytestmark = pytest.mark.gpu

NSIDE, LMAX, FWHM = 8, 16, 5.0
NPIX, NRE = 12 * NSIDE ** 2, (LMAX + 1) ** 2


def setup(masked, seed=0):
    rng = np.random.default_rng(seed)
    ell = np.arange(LMAX + 1)
    dl = np.where(ell >= 2, 300.0 / (ell + 5.0), 0.0)
    noise = np.full(NPIX, 4.0)
    mask = None
    if masked:
        mask = np.ones(NPIX)
        mask[NPIX // 3: NPIX // 3 + NPIX // 8] = 0
        mask[NPIX // 2: NPIX // 2 + 20] = 0.5
    d = rng.standard_normal(NPIX) * 3
    if masked:
        d = d * mask
    inv_noise = (1 / noise) * (mask if masked else 1)
    prob = R.TTProblem(NSIDE, LMAX, d, inv_noise, FWHM)
    var = R.generate_var_cl(dl)
    cl = dl * np.array([2 * np.pi / (l * (l + 1)) if l else 0 for l in range(LMAX + 1)])
    xi = (rng.standard_normal(NRE), rng.standard_normal(NPIX))
    return rng, dl, cl, var, noise, mask, d, prob, xi


def rel(a, b):
    a = a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.abs(a - b).max() / np.abs(b).max())


def test_tt_direct_draws_centred_noncentred_pncp():
    from gibbssampler_b200 import utils
    from gibbssampler_b200.Temperature import (CenteredConstrainedRealization, NonCenteredConstrainedRealization,
                                               PNCPConstrainedRealizationTT)
    rng, dl, cl, var, noise, mask, d, prob, xi = setup(False)
    bl_map = utils.expand_per_l(prob.bl_gauss, 0)
    c = CenteredConstrainedRealization(d, noise, bl_map, FWHM, LMAX, NPIX, seed=1)
    s, acc = c.sample_no_mask(var, xi=xi)
    assert acc == 1 and rel(s, prob.sample_no_mask(var, *xi)) < 1e-10
    s2, _ = c.sample(cl, var, None)          # dispatcher: no mask -> diagonal solve (CenteredGibbs.py:232-233)
    assert np.shape(s2) == (NRE,)
    nc = NonCenteredConstrainedRealization(d, noise, bl_map, FWHM, LMAX, NPIX, seed=1)
    s, _ = nc.sample_no_mask(cl, var, xi=xi)
    assert rel(s, prob.sample_no_mask_nc(var, *xi)) < 1e-10
    pn = PNCPConstrainedRealizationTT(d, noise, bl_map, FWHM, LMAX, NPIX, seed=1, l_cut=5)
    s, t, err = pn.sample(var, xi=xi)
    assert err == 0 and rel(s, prob.sample_pncp(var, *xi, 5)) < 1e-10
    assert np.all(np.asarray(s)[[0, 1, LMAX + 1, LMAX + 2]] == 0)


def test_tt_masked_pcg_rjpo_and_aux_variable():
    from gibbssampler_b200 import utils
    from gibbssampler_b200.Temperature import CenteredConstrainedRealization, NonCenteredConstrainedRealization
    rng, dl, cl, var, noise, mask, d, prob, xi = setup(True)
    bl_map = utils.expand_per_l(prob.bl_gauss, 0)
    c = CenteredConstrainedRealization(d, noise, bl_map, FWHM, LMAX, NPIX, mask=mask, seed=1)
    c.pcg_accuracy = 1e-10
    s, acc = c.sample_mask(cl, var, None, xi=xi)
    b = prob.rhs(var, *xi)
    assert rel(c.last_rhs, b) < 1e-10
    x, it = prob.pcg(var, b, eps=1e-10)
    assert abs(it - c.last_pcg_iterations) <= 2
    assert rel(s, x) < 1e-7
    # the solution solves Q x = b
    q = c.apply_Q(c._dl(var), torch.as_tensor(s, device="cuda"))
    assert rel(q, b) < 1e-8
    assert rel(q, prob.apply_Q(var, np.asarray(s))) < 1e-10
    # non-centred twin: C^-1/2 s, monopole / dipole zero
    nc = NonCenteredConstrainedRealization(d, noise, bl_map, FWHM, LMAX, NPIX, mask=mask, seed=1)
    nc.pcg_accuracy = 1e-10
    snc, _ = nc.sample_mask(cl, var, None, xi=xi)
    assert rel(snc, np.sqrt(R.safe_inv(var)) * x) < 1e-7
    # RJPO from the previous map: accepted with u small, and the accepted map solves the system to eps
    s_old = np.asarray(s) + 0.01 * rng.standard_normal(NRE)
    s_rj, acc = c.sample_mask(cl, var, s_old, metropolis_step=True, xi=xi, u=1e-300)
    assert acc == 1 and rel(s_rj, x) < 1e-6
    # auxiliary-variable step keeps shapes / finite values and moves the map
    s_aux, acc = c.sample_gibbs_change_variable(var, np.asarray(s))
    assert acc == 1 and np.all(np.isfinite(s_aux)) and rel(s_aux, np.asarray(s)) > 1e-6


def test_tt_noncentred_likelihood_and_mwg_sweep():
    from gibbssampler_b200 import utils
    from gibbssampler_b200.Temperature import NonCenteredClsSampler
    rng, dl, cl, var, noise, mask, d, prob, xi = setup(True)
    bl_map = utils.expand_per_l(prob.bl_gauss, 0)
    bins = np.arange(LMAX + 2)
    blocks = [2, 6, 10, LMAX + 1]
    pv = np.full(LMAX + 1 - 2, 4.0)
    smp = NonCenteredClsSampler(d, LMAX, NSIDE, bins, bl_map, noise, blocks, pv, mask=mask, rng="numpy")
    s_nc = rng.standard_normal(NRE)
    s_nc[[0, 1, LMAX + 1, LMAX + 2]] = 0
    assert abs(smp.compute_log_likelihood(var, s_nc) - prob.loglik(var, s_nc)) < 1e-9 * abs(prob.loglik(var, s_nc))
    # one sweep with numpy's stream reproduces a python restatement of NonCenteredGibbs.py:212-248
    from scipy.stats import truncnorm
    np.random.seed(5)
    out, var_new, acc = smp.sample(s_nc, dl.copy(), var)
    np.random.seed(5)
    old = dl.copy()
    clip = -old[2:] / np.sqrt(pv)
    prop = np.concatenate([np.zeros(2), truncnorm.rvs(a=clip, b=np.inf, loc=old[2:], scale=np.sqrt(pv))])

    def logq(frm, to):
        return np.concatenate([np.zeros(2), truncnorm.logpdf(to[2:], a=-frm[2:] / np.sqrt(pv), b=np.inf, loc=frm[2:], scale=np.sqrt(pv))])
    lr = logq(prop, old) - logq(old, prop)
    old_lik = prob.loglik(R.generate_var_cl(old), s_nc)
    us = [np.random.uniform() for _ in range(len(blocks) - 1)]
    ref_acc = []
    for i in range(len(blocks) - 1):
        new = old.copy()
        new[blocks[i]:blocks[i + 1]] = prop[blocks[i]:blocks[i + 1]]
        nl = prob.loglik(R.generate_var_cl(new), s_nc)
        if np.log(us[i]) < np.sum(lr[blocks[i]:blocks[i + 1]]) + nl - old_lik:
            old, old_lik = new, nl
            ref_acc.append(1)
        else:
            ref_acc.append(0)
    assert acc == ref_acc
    assert np.allclose(out, old, rtol=1e-9, atol=1e-12)
    assert np.allclose(var_new, R.generate_var_cl(old), rtol=1e-9, atol=1e-15)


@pytest.mark.parametrize("kind", ["centered", "noncentered", "asis", "pncp"])
def test_tt_gibbs_loops_run_and_return_reference_shapes(kind):
    from gibbssampler_b200.ASIS import ASIS
    from gibbssampler_b200.CenteredGibbs import CenteredGibbs
    from gibbssampler_b200.NonCenteredGibbs import NonCenteredGibbs
    from gibbssampler_b200.PNCP import PNCPGibbs
    rng, dl, cl, var, noise, mask, d, prob, xi = setup(kind != "pncp", seed=3)
    bins = np.arange(LMAX + 2)
    pv = np.full(LMAX + 1 - 2, 2.0)
    n_iter = 4
    if kind == "centered":
        g = CenteredGibbs(d, noise, None, FWHM, NSIDE, LMAX, NPIX, mask=mask, polarization=False, bins=bins, n_iter=n_iter, seed=2)
        h, acc, t = g.run(dl)
        assert h.shape == (n_iter + 1, LMAX + 1) and len(acc) == n_iter and len(t) == n_iter
    elif kind == "noncentered":
        g = NonCenteredGibbs(d, noise, None, FWHM, NSIDE, LMAX, NPIX, pv, metropolis_blocks=[2, 8, LMAX + 1], polarization=False, bins=bins,
                             n_iter=n_iter, mask=mask, seed=2)
        h, acc, t = g.run(dl)
        assert h.shape == (n_iter, LMAX + 1) and acc.shape == (n_iter, 2)
    elif kind == "asis":
        g = ASIS(d, noise, None, FWHM, NSIDE, LMAX, NPIX, pv, metropolis_blocks=[2, 8, LMAX + 1], polarization=False, bins=bins,
                 n_iter=n_iter, mask=mask, seed=2)
        h, acc, acr, t = g.run(dl)
        assert h.shape == (n_iter + 1, LMAX + 1) and acc.shape == (n_iter, 2) and len(acr) == n_iter
    else:
        g = PNCPGibbs(d, noise, FWHM, NSIDE, LMAX, NPIX, pv, 5, metropolis_blocks=[5, 9, LMAX + 1], polarization=False, bins=bins,
                      n_iter=n_iter, seed=2)
        h, acc, t = g.run(dl)
        assert h.shape == (n_iter + 1, LMAX + 1) and acc.shape == (n_iter, 2)
    assert np.all(np.isfinite(h)) and np.all(h[:, :2] == 0) and np.all(h[1:, 2:] > 0)
