"""Generates tests/golden/*.npz by running the REFERENCE'S OWN Python modules (imported from
/root/reference, unmodified) on seeded inputs, with the absent third-party packages replaced by
tests/golden/ref_stubs.py (healpy/qcinv arithmetic restated by the CPU oracle).

Run in the build container only:   python tests/golden/make_golden.py
The GPU box never needs /root/reference: tests read the committed .npz files."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import ref_stubs  # noqa: E402

NSIDE, LMAX = 4, 8
NPIX, NRE = 12 * NSIDE ** 2, (LMAX + 1) ** 2


def main():
    rng = np.random.default_rng(2024)
    bins = {"EE": np.arange(0, LMAX + 2), "BB": np.array([0, 1, 2, 3, 4, 6, LMAX + 1])}
    blocks = {"EE": [2, len(bins["EE"]) - 1], "BB": [2, 4, 5, 6]}
    mask = np.clip(rng.uniform(-0.3, 1.5, NPIX), 0, 1)
    mask_path = os.path.join(HERE, "mask_nside4.npy")
    np.save(mask_path, mask)
    qc = ref_stubs.install(NSIDE, LMAX, mask_path=mask_path, bins=bins, blocks=blocks)
    import utils as ref_utils
    import CenteredGibbs as ref_C
    import NonCenteredGibbs as ref_NC
    import ASIS as ref_ASIS
    from GibbsSampler import GibbsSampler as RefGibbs

    out = {}
    # ---- layout helpers (utils.py:49-76, 114-162) at several lmax
    for L in (2, 3, 4, 7, 8):
        sys.modules["config"].L_MAX_SCALARS = L
        r = rng.standard_normal((L + 1) ** 2)
        c = ref_utils.real_to_complex(r)
        out["r2c_in_%d" % L], out["r2c_out_%d" % L] = r, c
        out["c2r_out_%d" % L] = ref_utils.complex_to_real(c)
        dl = rng.uniform(0.1, 2, L + 1)
        out["varcl_in_%d" % L], out["varcl_out_%d" % L] = dl, ref_utils.generate_var_cl(dl)
    sys.modules["config"].L_MAX_SCALARS = LMAX
    out["unfold_in"] = rng.uniform(0, 1, len(bins["BB"]) - 1)
    out["unfold_out"] = ref_utils.unfold_bins(out["unfold_in"], bins["BB"])

    # ---- problem
    fwhm = 10.0
    noise_pol = np.full(NPIX, 0.3)
    noise_temp = np.full(NPIX, 1600.0)
    dQ, dU = rng.standard_normal(NPIX) * mask, rng.standard_normal(NPIX) * mask
    dE, dB = rng.standard_normal(NRE), rng.standard_normal(NRE)
    pix_map = {"Q": dQ, "U": dU, "EE": dE, "BB": dB}
    ell = np.arange(LMAX + 1)
    dls = {"EE": np.where(ell >= 2, 1.0 + 0.1 * ell, 0.0), "BB": np.where(ell >= 2, 0.3 + 0.02 * ell, 0.0)}
    out.update(mask=mask, noise_pol=noise_pol, dQ=dQ, dU=dU, dE=dE, dB=dB, dls_EE=dls["EE"], dls_BB=dls["BB"], fwhm=fwhm,
               bins_EE=bins["EE"], bins_BB=bins["BB"], blocks_EE=np.array(blocks["EE"]), blocks_BB=np.array(blocks["BB"]))

    g = RefGibbs(pix_map, noise_temp, fwhm, NSIDE, LMAX, polarization=True, bins=bins, n_iter=1)
    out["bl_map"] = g.bl_map                                   # GibbsSampler.compute_bl_map (GibbsSampler.py:64-74)

    # ---- PolarizedCenteredConstrainedRealization (CenteredGibbs.py:239-850)
    cr = ref_C.PolarizedCenteredConstrainedRealization(pix_map, noise_temp, noise_pol, g.bl_map, LMAX, NPIX, fwhm, mask_path=mask_path)
    out["second_part_grad_E"], out["second_part_grad_B"] = cr.second_part_grad_E, cr.second_part_grad_B
    cr.chain_descr[0][5] = 1e-13                               # tight PCG so that the solution itself is a golden vector
    np.random.seed(101)
    sol, acc = cr.sample_mask({k: v.copy() for k, v in dls.items()})
    out["sample_mask_seed"] = 101
    out["sample_mask_E"], out["sample_mask_B"] = sol["EE"], sol["BB"]
    out["sample_mask_rhs_E"], out["sample_mask_rhs_B"] = qc.multigrid.multigrid_chain.log[-1]["bE"], qc.multigrid.multigrid_chain.log[-1]["bB"]
    # full-sky object for the diagonal solve (the mask only enters sample_no_mask through inv_noise_pol[0])
    cr_full = ref_C.PolarizedCenteredConstrainedRealization(pix_map, noise_temp, noise_pol, g.bl_map, LMAX, NPIX, fwhm, mask_path=None)
    np.random.seed(102)
    sol, acc = cr_full.sample_no_mask({k: v.copy() for k, v in dls.items()})
    out["sample_no_mask_seed"] = 102
    out["sample_no_mask_E"], out["sample_no_mask_B"] = sol["EE"], sol["BB"]

    # ---- alternative CR kernels (CenteredGibbs.py:494-825) from a fixed starting map
    s_old = {"EE": out["sample_mask_E"].copy(), "BB": out["sample_mask_B"].copy()}
    cr.n_gibbs = 2
    np.random.seed(110)
    sol, _ = cr.sample_gibbs_change_variable({k: v.copy() for k, v in dls.items()}, {k: v.copy() for k, v in s_old.items()})
    out["aux_seed"] = 110
    out["aux_E"], out["aux_B"] = sol["EE"], sol["BB"]
    np.random.seed(111)
    sol, _ = cr.overrelaxation_sampler({k: v.copy() for k, v in dls.items()}, {k: v.copy() for k, v in s_old.items()})
    out["overrelax_seed"] = 111
    out["overrelax_E"], out["overrelax_B"] = sol["EE"], sol["BB"]
    for seed in (112, 113, 114):
        np.random.seed(seed)
        sol, acc = cr.sample_mala({k: v.copy() for k, v in dls.items()}, {k: v.copy() for k, v in s_old.items()})
        out["mala_E_%d" % seed], out["mala_B_%d" % seed], out["mala_acc_%d" % seed] = sol["EE"], sol["BB"], acc
    cr.chain_descr[0][5] = 1e-3                                # truncated PCG: the RJPO accept test is non-trivial
    np.random.seed(115)
    sol, acc = cr.sample_mask_rj({k: v.copy() for k, v in dls.items()}, {k: v.copy() for k, v in s_old.items()})
    out["rj_E"], out["rj_B"], out["rj_acc"] = sol["EE"], sol["BB"], acc
    cr.chain_descr[0][5] = 1e-13

    # ---- PolarizedCenteredClsSampler (CenteredGibbs.py:51-93)
    cs = ref_C.PolarizedCenteredClsSampler(pix_map, LMAX, NSIDE, bins, g.bl_map, noise_temp, mask_path=mask_path)
    np.random.seed(103)
    d = cs.sample({"EE": out["sample_mask_E"], "BB": out["sample_mask_B"]})
    out["cls_sample_seed"] = 103
    out["cls_sample_EE"], out["cls_sample_BB"] = d["EE"], d["BB"]

    # ---- non-centred pieces (NonCenteredGibbs.py:105-445)
    pv = {"EE": np.full(len(bins["EE"]) - 3, 0.05), "BB": np.full(len(bins["BB"]) - 3, 0.02)}
    out["prop_var_EE"], out["prop_var_BB"] = pv["EE"], pv["BB"]
    nc = ref_NC.PolarizationNonCenteredClsSampler(pix_map, LMAX, NSIDE, bins, g.bl_map, noise_temp, noise_pol, blocks, pv,
                                                  n_iter=1, mask_path=mask_path)
    binned_old = {"EE": dls["EE"][bins["EE"][:-1]].copy(), "BB": np.array([0, 0, 0.34, 0.36, 0.4, 0.44])}
    out["binned_old_EE"], out["binned_old_BB"] = binned_old["EE"], binned_old["BB"]
    s_nc = {"EE": rng.standard_normal(NRE), "BB": rng.standard_normal(NRE)}
    out["s_nc_E"], out["s_nc_B"] = s_nc["EE"], s_nc["BB"]
    np.random.seed(104)
    prop = nc.propose_dl(binned_old)
    out["propose_seed"] = 104
    out["propose_EE"], out["propose_BB"] = prop["EE"], prop["BB"]
    lp = nc.compute_log_proposal(binned_old, prop)
    out["logprop_EE"], out["logprop_BB"] = lp["EE"], lp["BB"]
    out["loglik_old"] = nc.compute_log_likelihood(binned_old, s_nc)
    out["loglik_prop"] = nc.compute_log_likelihood(prop, s_nc)
    np.random.seed(105)
    new, accept = nc.sample(s_nc, {k: v.copy() for k, v in binned_old.items()})
    out["mwg_seed"] = 105
    out["mwg_EE"], out["mwg_BB"] = new["EE"], new["BB"]
    out["mwg_accept_EE"], out["mwg_accept_BB"] = np.array(accept["EE"]), np.array(accept["BB"])

    ncr = ref_NC.PolarizedNonCenteredConstrainedRealization(pix_map, noise_temp, noise_pol, g.bl_map, LMAX, NPIX, fwhm, mask_path=mask_path)
    ncr.pol_centered_constraint_realizer.chain_descr[0][5] = 1e-13
    np.random.seed(106)
    sol, _ = ncr.sample_mask({k: v.copy() for k, v in dls.items()})
    out["nc_sample_mask_seed"] = 106
    out["nc_sample_mask_E"], out["nc_sample_mask_B"] = sol["EE"], sol["BB"]

    # ---- full loops: CenteredGibbs.run (GibbsSampler.py:118-180) and ASIS.run (ASIS.py:134-226), 3 iterations
    binned_init = {"EE": dls["EE"][bins["EE"][:-1]].copy(), "BB": np.array([0, 0, 0.34, 0.36, 0.4, 0.44])}
    out["binned_init_EE"], out["binned_init_BB"] = binned_init["EE"], binned_init["BB"]
    cg = ref_C.CenteredGibbs(pix_map, noise_temp, noise_pol, fwhm, NSIDE, LMAX, NPIX, mask_path=mask_path, polarization=True, bins=bins,
                             n_iter=3)
    cg.constrained_sampler.chain_descr[0][5] = 1e-13
    cg.constrained_sampler.ula = False     # CenteredGibbs forwards ula=False (CenteredGibbs.py:875) -> PCG branch of the dispatcher
    np.random.seed(107)
    h_dls, h_acc, _, _ = cg.run({k: v.copy() for k, v in binned_init.items()})
    out["centered_run_seed"] = 107
    out["centered_run_EE"], out["centered_run_BB"] = h_dls["EE"], h_dls["BB"]

    asis = ref_ASIS.ASIS(pix_map, noise_temp, noise_pol, fwhm, NSIDE, LMAX, NPIX, pv, metropolis_blocks=blocks, polarization=True, bins=bins,
                         n_iter=3, mask_path=mask_path)
    asis.constrained_sampler.chain_descr[0][5] = 1e-13
    asis.constrained_sampler.ula = False
    np.random.seed(108)
    res = asis.run({k: v.copy() for k, v in binned_init.items()})
    out["asis_run_seed"] = 108
    out["asis_run_EE"], out["asis_run_BB"] = res[0]["EE"], res[0]["BB"]
    out["asis_accept_EE"], out["asis_accept_BB"] = res[1]["EE"], res[1]["BB"]

    ncg = ref_NC.NonCenteredGibbs(pix_map, noise_temp, noise_pol, fwhm, NSIDE, LMAX, NPIX, pv, metropolis_blocks=blocks, polarization=True,
                                  bins=bins, n_iter=3, mask_path=mask_path)
    ncg.constrained_sampler.pol_centered_constraint_realizer.chain_descr[0][5] = 1e-13
    np.random.seed(109)
    res = ncg.run({k: v.copy() for k, v in binned_init.items()})
    out["nc_run_seed"] = 109
    out["nc_run_EE"], out["nc_run_BB"] = res[0]["EE"], res[0]["BB"]
    out["nc_accept_EE"], out["nc_accept_BB"] = res[1]["EE"], res[1]["BB"]

    np.savez_compressed(os.path.join(HERE, "reference_nside4.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_nside4.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
