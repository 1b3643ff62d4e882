"""GPU: the drop-in classes of gibbssampler_b200 against golden vectors produced by the REFERENCE'S OWN
modules (tests/golden/make_golden.py) on the same numpy random stream (rng="numpy" injects numpy's
legacy global draws in the reference's order).  Index work bit-exact; FP64 work within 1e-10 relative."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_nside4.npz"))
NSIDE, LMAX = 4, 8
NPIX, NRE = 12 * NSIDE ** 2, (LMAX + 1) ** 2
RTOL = 1e-10


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() <= rtol * max(np.abs(b).max(), 1e-300)


def common():
    pix_map = {"Q": G["dQ"], "U": G["dU"], "EE": G["dE"], "BB": G["dB"]}
    bins = {"EE": G["bins_EE"], "BB": G["bins_BB"]}
    blocks = {"EE": list(G["blocks_EE"]), "BB": list(G["blocks_BB"])}
    pv = {"EE": G["prop_var_EE"], "BB": G["prop_var_BB"]}
    dls = {"EE": G["dls_EE"], "BB": G["dls_BB"]}
    return pix_map, bins, blocks, pv, dls, np.full(NPIX, 1600.0), G["noise_pol"], float(G["fwhm"])


def test_layout_helpers_bit_exact():
    from gibbssampler_b200 import utils
    for L in (2, 3, 4, 7, 8):
        assert np.array_equal(utils.real_to_complex(G["r2c_in_%d" % L]), G["r2c_out_%d" % L])
        assert np.array_equal(utils.complex_to_real(G["r2c_out_%d" % L]), G["c2r_out_%d" % L])
        assert close(utils.generate_var_cl(G["varcl_in_%d" % L]), G["varcl_out_%d" % L], 1e-15)
        # index pattern of the expansion is bit-exact: expanding the identity l -> l reproduces the reference's l map
        ell = np.arange(L + 1, dtype=float)
        ref_l = np.concatenate([ell, np.array([c for m in range(1, L + 1) for c in ell[m:] for _ in range(2)])])
        assert np.array_equal(utils.expand_per_l(ell, 0), ref_l)
    assert np.array_equal(utils.unfold_bins(G["unfold_in"], G["bins_BB"]), G["unfold_out"])


def test_compute_bl_map_and_second_part_grad():
    from gibbssampler_b200.CenteredGibbs import CenteredGibbs
    pix_map, bins, blocks, pv, dls, nt, npol, fwhm = common()
    g = CenteredGibbs(pix_map, nt, npol, fwhm, NSIDE, LMAX, NPIX, mask=G["mask"], polarization=True, bins=bins, n_iter=1, rng="numpy")
    assert close(g.bl_map.cpu().numpy(), G["bl_map"], 1e-15)
    assert close(g.constrained_sampler.second_part_grad_E.cpu().numpy(), G["second_part_grad_E"])
    assert close(g.constrained_sampler.second_part_grad_B.cpu().numpy(), G["second_part_grad_B"])


def test_sample_mask_and_sample_no_mask():
    from gibbssampler_b200.CenteredGibbs import PolarizedCenteredConstrainedRealization
    pix_map, bins, blocks, pv, dls, nt, npol, fwhm = common()
    cr = PolarizedCenteredConstrainedRealization(pix_map, nt, npol, G["bl_map"], LMAX, NPIX, fwhm, mask=G["mask"], rng="numpy")
    cr.pcg_accuracy = 1e-13
    np.random.seed(int(G["sample_mask_seed"]))
    sol, acc = cr.sample_mask(dls)
    assert acc == 1
    assert close(cr.last_rhs[0].cpu().numpy(), G["sample_mask_rhs_E"]) and close(cr.last_rhs[1].cpu().numpy(), G["sample_mask_rhs_B"])
    assert close(sol["EE"], G["sample_mask_E"], 1e-8) and close(sol["BB"], G["sample_mask_B"], 1e-8)
    np.random.seed(int(G["sample_no_mask_seed"]))
    sol, acc = cr.sample_no_mask(dls)
    assert close(sol["EE"], G["sample_no_mask_E"], 1e-12) and close(sol["BB"], G["sample_no_mask_B"], 1e-12)


def test_centered_cls_sampler():
    from gibbssampler_b200.CenteredGibbs import PolarizedCenteredClsSampler
    pix_map, bins, blocks, pv, dls, nt, npol, fwhm = common()
    cs = PolarizedCenteredClsSampler(pix_map, LMAX, NSIDE, bins, G["bl_map"], nt, mask=G["mask"], rng="numpy")
    np.random.seed(int(G["cls_sample_seed"]))
    d = cs.sample({"EE": G["sample_mask_E"], "BB": G["sample_mask_B"]})
    assert close(d["EE"], G["cls_sample_EE"], 1e-12) and close(d["BB"], G["cls_sample_BB"], 1e-12)


def test_noncentred_sampler_pieces_and_sweep():
    from gibbssampler_b200.NonCenteredGibbs import PolarizationNonCenteredClsSampler, PolarizedNonCenteredConstrainedRealization
    pix_map, bins, blocks, pv, dls, nt, npol, fwhm = common()
    nc = PolarizationNonCenteredClsSampler(pix_map, LMAX, NSIDE, bins, G["bl_map"], nt, npol, blocks, pv, n_iter=1, mask=G["mask"],
                                           rng="numpy")
    old = {"EE": G["binned_old_EE"], "BB": G["binned_old_BB"]}
    s_nc = {"EE": G["s_nc_E"], "BB": G["s_nc_B"]}
    assert abs(nc.compute_log_likelihood(old, s_nc) - float(G["loglik_old"])) < RTOL * abs(float(G["loglik_old"]))
    prop = {"EE": G["propose_EE"], "BB": G["propose_BB"]}
    assert abs(nc.compute_log_likelihood(prop, s_nc) - float(G["loglik_prop"])) < RTOL * abs(float(G["loglik_prop"]))
    np.random.seed(int(G["propose_seed"]))
    p = nc.propose_dl(old)
    assert np.array_equal(p["EE"].cpu().numpy(), G["propose_EE"]) and np.array_equal(p["BB"].cpu().numpy(), G["propose_BB"])
    lp = nc.compute_log_proposal(old, prop)
    assert close(lp["EE"].cpu().numpy(), G["logprop_EE"], 1e-11) and close(lp["BB"].cpu().numpy(), G["logprop_BB"], 1e-11)
    np.random.seed(int(G["mwg_seed"]))
    new, accept = nc.sample(s_nc, old)
    assert accept["EE"] == list(G["mwg_accept_EE"]) and accept["BB"] == list(G["mwg_accept_BB"])
    assert close(new["EE"], G["mwg_EE"], 1e-12) and close(new["BB"], G["mwg_BB"], 1e-12)
    ncr = PolarizedNonCenteredConstrainedRealization(pix_map, nt, npol, G["bl_map"], LMAX, NPIX, fwhm, mask=G["mask"], rng="numpy")
    ncr.pol_centered_constraint_realizer.pcg_accuracy = 1e-13
    np.random.seed(int(G["nc_sample_mask_seed"]))
    sol, _ = ncr.sample_mask(dls)
    assert close(sol["EE"], G["nc_sample_mask_E"], 1e-8) and close(sol["BB"], G["nc_sample_mask_B"], 1e-8)


def test_full_gibbs_loops_reproduce_reference_chains():
    """3 iterations of CenteredGibbs.run, ASIS.run and NonCenteredGibbs.run on the reference's random stream."""
    from gibbssampler_b200.ASIS import ASIS
    from gibbssampler_b200.CenteredGibbs import CenteredGibbs
    from gibbssampler_b200.NonCenteredGibbs import NonCenteredGibbs
    pix_map, bins, blocks, pv, dls, nt, npol, fwhm = common()
    init = {"EE": G["binned_init_EE"], "BB": G["binned_init_BB"]}
    cg = CenteredGibbs(pix_map, nt, npol, fwhm, NSIDE, LMAX, NPIX, mask=G["mask"], polarization=True, bins=bins, n_iter=3, rng="numpy")
    cg.constrained_sampler.pcg_accuracy = 1e-13
    np.random.seed(int(G["centered_run_seed"]))
    h, acc, t1, t2 = cg.run(init)
    assert h["EE"].shape == G["centered_run_EE"].shape
    assert close(h["EE"], G["centered_run_EE"], 1e-7) and close(h["BB"], G["centered_run_BB"], 1e-7)

    asis = ASIS(pix_map, nt, npol, fwhm, NSIDE, LMAX, NPIX, pv, metropolis_blocks=blocks, polarization=True, bins=bins, n_iter=3,
                mask=G["mask"], rng="numpy")
    asis.constrained_sampler.pcg_accuracy = 1e-13
    np.random.seed(int(G["asis_run_seed"]))
    res = asis.run(init)
    assert np.array_equal(res[1]["EE"], G["asis_accept_EE"]) and np.array_equal(res[1]["BB"], G["asis_accept_BB"])
    assert close(res[0]["EE"], G["asis_run_EE"], 1e-7) and close(res[0]["BB"], G["asis_run_BB"], 1e-7)

    ncg = NonCenteredGibbs(pix_map, nt, npol, fwhm, NSIDE, LMAX, NPIX, pv, metropolis_blocks=blocks, polarization=True, bins=bins, n_iter=3,
                           mask=G["mask"], rng="numpy")
    ncg.constrained_sampler.pol_centered_constraint_realizer.pcg_accuracy = 1e-13
    np.random.seed(int(G["nc_run_seed"]))
    res = ncg.run(init)
    assert np.array_equal(res[1]["EE"], G["nc_accept_EE"]) and np.array_equal(res[1]["BB"], G["nc_accept_BB"])
    assert close(res[0]["EE"], G["nc_run_EE"], 1e-7) and close(res[0]["BB"], G["nc_run_BB"], 1e-7)


def test_alternative_cr_kernels_match_reference():
    """Auxiliary-variable Gibbs, over-relaxation, MALA and RJPO (CenteredGibbs.py:494-825) on the reference's stream."""
    from gibbssampler_b200 import cr_extra
    from gibbssampler_b200.CenteredGibbs import PolarizedCenteredConstrainedRealization
    pix_map, bins, blocks, pv, dls, nt, npol, fwhm = common()
    cr = PolarizedCenteredConstrainedRealization(pix_map, nt, npol, G["bl_map"], LMAX, NPIX, fwhm, mask=G["mask"], rng="numpy", n_gibbs=2)
    s_old = {"EE": G["sample_mask_E"].copy(), "BB": G["sample_mask_B"].copy()}
    np.random.seed(int(G["aux_seed"]))
    sol, acc = cr_extra.sample_gibbs_change_variable(cr, dls, s_old)
    assert close(sol["EE"], G["aux_E"]) and close(sol["BB"], G["aux_B"])
    np.random.seed(int(G["overrelax_seed"]))
    sol, acc = cr_extra.overrelaxation_sampler(cr, dls, s_old)
    assert close(sol["EE"], G["overrelax_E"], 1e-8) and close(sol["BB"], G["overrelax_B"], 1e-8)
    for seed in (112, 113, 114):
        np.random.seed(seed)
        sol, acc = cr_extra.sample_mala(cr, dls, s_old)
        assert acc == int(G["mala_acc_%d" % seed])
        assert close(sol["EE"], G["mala_E_%d" % seed]) and close(sol["BB"], G["mala_B_%d" % seed])
    cr.pcg_accuracy = 1e-3
    np.random.seed(115)
    sol, acc = cr.sample_mask_rj(dls, s_old)
    assert acc == int(G["rj_acc"])
    assert close(sol["EE"], G["rj_E"], 1e-6) and close(sol["BB"], G["rj_B"], 1e-6)
    # dispatcher: gibbs_cr + overrelaxation / ula combinations route like CenteredGibbs.py:828-850
    cr.gibbs_cr, cr.overrelaxation, cr.ula = True, True, False
    np.random.seed(int(G["overrelax_seed"]))
    sol, _ = cr.sample(dls, s_old)
    assert close(sol["EE"], G["overrelax_E"], 1e-8)


def test_all_sph_sampler_matches_reference():
    """all_sph branch (full sky, isotropic noise, harmonic data: NonCenteredGibbs.py:357-377, 385-393, 414-419) against the
    reference's own module (tests/golden/make_golden_allsph.py): likelihood values and the whole blocked sweep."""
    from gibbssampler_b200.NonCenteredGibbs import PolarizationNonCenteredClsSampler
    A = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_allsph_nside4.npz"))
    bins = {"EE": A["bins_EE"], "BB": A["bins_BB"]}
    blocks = {"EE": list(A["blocks_EE"]), "BB": list(A["blocks_BB"])}
    pv = {"EE": A["prop_var_EE"], "BB": A["prop_var_BB"]}
    pix_map = {"EE": A["dE"], "BB": A["dB"]}
    nc = PolarizationNonCenteredClsSampler(pix_map, LMAX, NSIDE, bins, A["bl_map"], np.full(NPIX, 1600.0), A["noise_pol"], blocks, pv,
                                           n_iter=2, all_sph=True, rng="numpy")
    old = {"EE": A["binned_old_EE"], "BB": A["binned_old_BB"]}
    s_nc = {"EE": A["s_nc_E"], "BB": A["s_nc_B"]}
    assert abs(nc.compute_log_likelihood_all_sph(old, s_nc) - float(A["loglik_old"])) < 1e-12 * abs(float(A["loglik_old"]))
    prop = {"EE": A["propose_EE"], "BB": A["propose_BB"]}
    assert abs(nc.compute_log_likelihood_all_sph(prop, s_nc) - float(A["loglik_prop"])) < 1e-12 * abs(float(A["loglik_prop"]))
    np.random.seed(int(A["mwg_seed"]))
    new, accept = nc.sample(s_nc, old)
    assert accept["EE"] == list(A["mwg_accept_EE"]) and accept["BB"] == list(A["mwg_accept_BB"])
    assert close(new["EE"], A["mwg_EE"], 1e-12) and close(new["BB"], A["mwg_BB"], 1e-12)
    with pytest.raises(ValueError):
        PolarizationNonCenteredClsSampler(pix_map, LMAX, NSIDE, bins, A["bl_map"], np.full(NPIX, 1600.0), A["noise_pol"], blocks, pv,
                                          all_sph=True, mask=np.ones(NPIX), rng="numpy")
