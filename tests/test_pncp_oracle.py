"""Deterministic parity of the polarised masked-sky PNCP sampler (BASELINE config #3).

The reference has no implementation of this sampler (PNCP exists only as TT / full-sky bytecode), so it is DEFINED
by oracle/reference_logic.py:PNCPPol as a composition of reference pieces.  CPU part: the pieces of that composition
agree with the restatements that ARE pinned by the reference's own modules (tests/golden/reference_nside4.npz).
GPU part: gibbssampler_b200.PNCP.PNCPGibbs on numpy's legacy random stream reproduces the oracle chain draw for
draw (accept flags exact)."""
import os

import numpy as np
import pytest

from oracle import reference_logic as R
from oracle import sht as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_nside4.npz"))


def small_problem(nside=8, lmax=16, l_cut=5, seed=3):
    npix, n = 12 * nside ** 2, (lmax + 1) ** 2
    rng = np.random.default_rng(seed)
    ell = np.arange(lmax + 1)
    dl_true = np.where(ell >= 2, 1.0 + 0.05 * ell, 0.0)
    fwhm, noise0 = 6.0, 0.05
    th, _ = O.pix_angles(nside)
    mask = np.clip((np.abs(np.cos(th)) - 0.2) / 0.15, 0.0, 1.0)          # fractional edge like a ud_graded mask
    bl_map = R.expand_per_l(O.gauss_beam(np.radians(fwhm), lmax))
    sE = rng.standard_normal(n) * np.sqrt(R.generate_var_cl(dl_true))
    sB = rng.standard_normal(n) * np.sqrt(R.generate_var_cl(0.3 * dl_true))
    q, u = R.synth_pol(sE * bl_map, sB * bl_map, nside, lmax)
    dQ = (q + rng.standard_normal(npix) * np.sqrt(noise0)) * mask
    dU = (u + rng.standard_normal(npix) * np.sqrt(noise0)) * mask
    bins = {"EE": np.arange(0, lmax + 2), "BB": np.unique(list(range(0, lmax // 2 + 1)) + [lmax // 2 + 2, lmax - 2, lmax + 1])}
    first = {p: int(np.searchsorted(bins[p], l_cut, side="left")) for p in bins}
    nb = {p: len(bins[p]) - 1 for p in bins}
    blocks = {"EE": [first["EE"], (first["EE"] + nb["EE"]) // 2, nb["EE"]],
              "BB": [first["BB"], first["BB"] + 1] + list(range(first["BB"] + 2, nb["BB"] + 1))}
    pv = {p: np.full(len(bins[p]) - 3, 0.2) for p in bins}
    init = {"EE": np.array([dl_true[bins["EE"][i]:bins["EE"][i + 1]].mean() for i in range(len(bins["EE"]) - 1)]),
            "BB": np.array([0.3 * dl_true[bins["BB"][i]:bins["BB"][i + 1]].mean() for i in range(len(bins["BB"]) - 1)])}
    return dict(nside=nside, lmax=lmax, npix=npix, l_cut=l_cut, fwhm=fwhm, noise0=noise0, mask=mask, dQ=dQ, dU=dU, bins=bins,
                blocks=blocks, pv=pv, init=init)


def test_mwg_sweep_restatement_matches_reference_golden():
    """mwg_propose / mwg_log_proposal / mwg_sweep (the pieces PNCPPol composes) against the proposals, log-proposal
    densities, likelihood values and the whole blocked sweep (accept flags, new D_l) recorded from the reference's own
    NonCenteredGibbs module."""
    nside, lmax = 4, 8
    bins = {"EE": G["bins_EE"], "BB": G["bins_BB"]}
    blocks = {"EE": list(G["blocks_EE"]), "BB": list(G["blocks_BB"])}
    pvar = {"EE": G["prop_var_EE"], "BB": G["prop_var_BB"]}
    old = {"EE": G["binned_old_EE"], "BB": G["binned_old_BB"]}
    np.random.seed(int(G["propose_seed"]))
    for pol in ("EE", "BB"):
        new = R.mwg_propose(old[pol], pvar[pol])
        assert np.array_equal(new, G["propose_" + pol])
        assert np.allclose(R.mwg_log_proposal(new, old[pol], pvar[pol]), G["logprop_" + pol], rtol=1e-13)
    prob = R.PolProblem(nside, lmax, G["dQ"], G["dU"], G["mask"] / G["noise_pol"], float(G["fwhm"]))
    s_nc = {"EE": G["s_nc_E"], "BB": G["s_nc_B"]}
    # l_cut = 0 is the reference's likelihood; the vectorised twins give the same value
    assert abs(R.nc_loglik(old, bins, s_nc, prob, 0) - float(G["loglik_old"])) < 1e-10 * abs(float(G["loglik_old"]))
    pv = R.PolProblem(nside, lmax, G["dQ"], G["dU"], G["mask"] / G["noise_pol"], float(G["fwhm"]), vectorised=True)
    assert R.nc_loglik(old, bins, s_nc, pv, 0) == R.nc_loglik(old, bins, s_nc, prob, 0)
    np.random.seed(int(G["mwg_seed"]))
    new, accept = R.mwg_sweep(prob, bins, blocks, pvar, s_nc, old)
    assert accept["EE"] == list(G["mwg_accept_EE"]) and accept["BB"] == list(G["mwg_accept_BB"])
    assert np.allclose(new["EE"], G["mwg_EE"], rtol=1e-12) and np.allclose(new["BB"], G["mwg_BB"], rtol=1e-12)


def test_pncp_mixed_variable_is_a_bijection_and_l_cut_zero_is_noncentred():
    P = small_problem()
    prob = R.PolProblem(P["nside"], P["lmax"], P["dQ"], P["dU"], P["mask"] / P["noise0"], P["fwhm"])
    pn = R.PNCPPol(prob, P["bins"], P["blocks"], P["pv"], P["l_cut"])
    dl = R.unfold_bins(P["init"]["EE"], P["bins"]["EE"])
    f, g = pn.factor(dl, True), pn.factor(dl, False)
    ell = R.ell_index(P["lmax"])
    assert np.all(f[ell < P["l_cut"]] == 1.0) and np.all(g[ell < P["l_cut"]] == 1.0)
    assert np.allclose((f * g)[ell >= P["l_cut"]], 1.0, rtol=1e-15)
    # likelihood in the mixed variable == likelihood of the centred map it encodes (the change of variable is exact)
    rng = np.random.default_rng(0)
    s = {k: rng.standard_normal((P["lmax"] + 1) ** 2) for k in ("EE", "BB")}
    for k in s:
        s[k][[0, 1, P["lmax"] + 1, P["lmax"] + 2]] = 0
    dls = {k: R.unfold_bins(P["init"][k], P["bins"][k]) for k in s}
    mixed = {k: s[k] * pn.factor(dls[k], True) for k in s}
    q, u = R.synth_pol(s["EE"] * prob.bl_map, s["BB"] * prob.bl_map, P["nside"], P["lmax"])
    want = -0.5 * (np.sum((P["dQ"] - q) ** 2 * prob.inv_noise) + np.sum((P["dU"] - u) ** 2 * prob.inv_noise))
    assert abs(R.nc_loglik(P["init"], P["bins"], mixed, prob, P["l_cut"]) - want) < 1e-11 * abs(want)


def test_pncp_iteration_runs_and_is_reproducible():
    P = small_problem(nside=4, lmax=8, l_cut=4)
    prob = R.PolProblem(P["nside"], P["lmax"], P["dQ"], P["dU"], P["mask"] / P["noise0"], P["fwhm"])
    pn = R.PNCPPol(prob, P["bins"], P["blocks"], P["pv"], P["l_cut"])
    np.random.seed(5)
    a, acc_a, _ = pn.iteration(P["init"])
    np.random.seed(5)
    b, acc_b, _ = pn.iteration(P["init"])
    assert acc_a == acc_b and all(np.array_equal(a[k], b[k]) for k in a)
    for pol in ("EE", "BB"):
        k = pn.low_bins[pol]
        assert np.all(a[pol][:2] == 0) and np.all(a[pol][2:k] > 0)          # low bins redrawn (inverse-gamma), monopole/dipole 0
        assert len(acc_a[pol]) == len(P["blocks"][pol]) - 1


@pytest.mark.gpu
def test_pncp_gibbs_reproduces_oracle_chain():
    """3 iterations of PNCPGibbs.run (polarised, masked, fractional mask edge) on numpy's random stream against
    oracle.reference_logic.PNCPPol: accept flags exact, binned D_l to 1e-7 after three PCG solves."""
    from gibbssampler_b200.PNCP import PNCPGibbs
    P = small_problem()
    n_iter = 3
    prob = R.PolProblem(P["nside"], P["lmax"], P["dQ"], P["dU"], P["mask"] / P["noise0"], P["fwhm"])
    pn = R.PNCPPol(prob, P["bins"], P["blocks"], P["pv"], P["l_cut"], eps=1e-13)
    np.random.seed(77)
    cur, want, want_acc = P["init"], {"EE": [P["init"]["EE"]], "BB": [P["init"]["BB"]]}, {"EE": [], "BB": []}
    for _ in range(n_iter):
        cur, acc, _ = pn.iteration(cur)
        for k in cur:
            want[k].append(cur[k])
            want_acc[k].append(acc[k])
    npix = P["npix"]
    for batched in (True, False):
        g = PNCPGibbs({"Q": P["dQ"], "U": P["dU"]}, np.full(npix, 1600.0), P["fwhm"], P["nside"], P["lmax"], npix, P["pv"], P["l_cut"],
                      metropolis_blocks=P["blocks"], polarization=True, bins=P["bins"], n_iter=n_iter, noise_Q=np.full(npix, P["noise0"]),
                      mask=P["mask"], rng="numpy")
        g.constrained_sampler.pcg_accuracy = 1e-13
        g.cls_sampler.batched_blocks = batched
        np.random.seed(77)
        h, acc, _, _ = g.run(P["init"])
        for k in ("EE", "BB"):
            assert np.array_equal(np.asarray(acc[k]), np.asarray(want_acc[k])), (batched, k)
            got, ref = np.asarray(h[k]), np.asarray(want[k])
            assert got.shape == ref.shape
            assert np.abs(got - ref).max() <= 1e-7 * np.abs(ref).max(), (batched, k, np.abs(got - ref).max())
