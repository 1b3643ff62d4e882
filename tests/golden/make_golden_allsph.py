"""Golden vectors for the all_sph branch of the reference's non-centred C_l sampler (full sky, isotropic noise, data in
harmonic space: NonCenteredGibbs.py:357-377, 385-393, 414-419), produced by the REFERENCE'S OWN module imported from
/root/reference with the absent third-party packages stubbed (tests/golden/ref_stubs.py).

Run in the build container only:   python tests/golden/make_golden_allsph.py
The GPU box never needs /root/reference: tests read the committed reference_allsph_nside4.npz."""
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
    rng = np.random.default_rng(4048)
    bins = {"EE": np.arange(0, LMAX + 2), "BB": np.array([0, 1, 2, 3, 4, 6, LMAX + 1])}
    blocks = {"EE": [2, 5, len(bins["EE"]) - 1], "BB": [2, 4, 5, 6]}
    ref_stubs.install(NSIDE, LMAX, mask_path=None, bins=bins, blocks=blocks)
    import GibbsSampler as ref_G
    import NonCenteredGibbs as ref_NC

    fwhm = 10.0
    noise_pol = np.full(NPIX, 0.3)
    noise_temp = np.full(NPIX, 1600.0)
    dE, dB = rng.standard_normal(NRE), rng.standard_normal(NRE)
    pix_map = {"Q": np.zeros(NPIX), "U": np.zeros(NPIX), "EE": dE, "BB": dB}
    g = ref_G.GibbsSampler(pix_map, noise_pol, fwhm, NSIDE, LMAX, polarization=True, bins=bins, n_iter=1)
    pv = {"EE": np.full(len(bins["EE"]) - 3, 0.05), "BB": np.full(len(bins["BB"]) - 3, 0.02)}
    nc = ref_NC.PolarizationNonCenteredClsSampler(pix_map, LMAX, NSIDE, bins, g.bl_map, noise_temp, noise_pol, blocks, pv, n_iter=2,
                                                  mask_path=None, all_sph=True)
    ell = np.arange(LMAX + 1)
    dls = {"EE": np.where(ell >= 2, 1.0 + 0.1 * ell, 0.0), "BB": np.where(ell >= 2, 0.3 + 0.02 * ell, 0.0)}
    binned_old = {"EE": dls["EE"][bins["EE"][:-1]].copy(), "BB": np.array([0, 0, 0.34, 0.36, 0.4, 0.44])}
    s_nc = {"EE": rng.standard_normal(NRE), "BB": rng.standard_normal(NRE)}
    out = dict(fwhm=fwhm, noise_pol=noise_pol, dE=dE, dB=dB, bins_EE=bins["EE"], bins_BB=bins["BB"], blocks_EE=np.array(blocks["EE"]),
               blocks_BB=np.array(blocks["BB"]), prop_var_EE=pv["EE"], prop_var_BB=pv["BB"], binned_old_EE=binned_old["EE"],
               binned_old_BB=binned_old["BB"], s_nc_E=s_nc["EE"], s_nc_B=s_nc["BB"], bl_map=g.bl_map)
    out["loglik_old"] = nc.compute_log_likelihood_all_sph(binned_old, s_nc)
    np.random.seed(311)
    prop = nc.propose_dl(binned_old)
    out["propose_seed"] = 311
    out["propose_EE"], out["propose_BB"] = prop["EE"], prop["BB"]
    out["loglik_prop"] = nc.compute_log_likelihood_all_sph(prop, s_nc)
    np.random.seed(312)
    new, accept = nc.sample(s_nc, {k: v.copy() for k, v in binned_old.items()})
    out["mwg_seed"] = 312
    out["mwg_EE"], out["mwg_BB"] = new["EE"], new["BB"]
    out["mwg_accept_EE"], out["mwg_accept_BB"] = np.array(accept["EE"]), np.array(accept["BB"])
    path = os.path.join(HERE, "reference_allsph_nside4.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays; loglik", out["loglik_old"], out["loglik_prop"], "accept", accept)


if __name__ == "__main__":
    main()
