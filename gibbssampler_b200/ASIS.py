"""Ancillarity-sufficiency interweaving (mirror of ASIS.py:16-232), polarised EE/BB."""
import time

import numpy as np
import torch

from . import _dev, _lib, utils
from ._dev import f64, ptr, stream
from ._lib import check
from .CenteredGibbs import PolarizedCenteredClsSampler, PolarizedCenteredConstrainedRealization
from .GibbsSampler import GibbsSampler
from .NonCenteredGibbs import PolarizationNonCenteredClsSampler


class ASIS(GibbsSampler):
    def __init__(self, pix_map, noise, noise_Q, beam, nside, lmax, Npix, proposal_variances, metropolis_blocks=None,
                 polarization=False, bins=None, n_iter=10000, n_iter_metropolis=1, mask_path=None, gibbs_cr=False,
                 rj_step=False, all_sph=False, n_gibbs=20, overrelaxation=False, *, mask=None, rng="philox", seed=None,
                 verbose=False):
        """Mirror of ASIS.__init__ (ASIS.py:18-65)."""
        super().__init__(pix_map, noise, beam, nside, lmax, polarization=polarization, bins=bins, n_iter=n_iter, gibbs_cr=gibbs_cr,
                         rj_step=rj_step, verbose=verbose)
        shared = _dev.Rng(rng, seed)
        if not polarization:  # ASIS.py:46-54
            from .CenteredGibbs import CenteredClsSampler
            from .Temperature import CenteredConstrainedRealization, NonCenteredClsSampler
            self.constrained_sampler = CenteredConstrainedRealization(pix_map, noise, self.bl_map, beam, lmax, Npix, mask_path=mask_path,
                                                                      mask=mask, rng=shared)
            self.non_centered_cls_sampler = NonCenteredClsSampler(pix_map, lmax, nside, self.bins, self.bl_map, noise, metropolis_blocks,
                                                                  proposal_variances, n_iter=n_iter_metropolis, mask_path=mask_path,
                                                                  mask=mask, rng=shared)
            self.centered_cls_sampler = CenteredClsSampler(pix_map, lmax, nside, self.bins, self.bl_map, noise, rng=shared)
            return
        self.non_centered_cls_sampler = PolarizationNonCenteredClsSampler(pix_map, lmax, nside, self.bins, self.bl_map, noise, noise_Q,
                                                                          metropolis_blocks, proposal_variances, n_iter=n_iter_metropolis,
                                                                          mask_path=mask_path, all_sph=all_sph, mask=mask, rng=shared)
        self.centered_cls_sampler = PolarizedCenteredClsSampler(pix_map, lmax, nside, self.bins, self.bl_map, noise, rng=shared)
        self.constrained_sampler = PolarizedCenteredConstrainedRealization(pix_map, noise, noise_Q, self.bl_map, lmax, Npix, beam,
                                                                           mask_path=mask_path, gibbs_cr=gibbs_cr, n_gibbs=n_gibbs,
                                                                           overrelaxation=overrelaxation, mask=mask, rng=shared,
                                                                           ula=False)

    def _scale(self, skymap, dls_unbinned, mode):
        out = {}
        for pol in ("EE", "BB"):
            f = utils.expand_per_l(dls_unbinned[pol], mode)
            o = torch.empty_like(skymap[pol])
            check(_lib.lib().gs_mul(ptr(skymap[pol]), ptr(f), ptr(o), o.numel(), stream()))
            out[pol] = o
        return out

    def run_temperature(self, dls_init):
        """Mirror of ASIS.run_temperature (ASIS.py:68-132): CR, centred C_l draw, non-centred MwG, re-centring.
        Returns (h_dls, h_accept, h_accept_cr, h_time_seconds)."""
        h_accept_cr, h_accept, h_dls, h_time_seconds = [], [], [], []
        binned_dls = f64(dls_init)
        cls, var_cls_full = self._tt_state(binned_dls)
        h_dls.append(_dev.to_host(binned_dls))
        skymap, _ = self.constrained_sampler.sample(cls, var_cls_full, None, metropolis_step=False)
        for i in range(self.n_iter):
            if self.verbose:
                print("Interweaving, iteration:", i)
            t0 = time.perf_counter()
            skymap, accept_cr = self.constrained_sampler.sample(cls, var_cls_full, skymap, use_gibbs=self.gibbs_cr)
            h_accept_cr.append(accept_cr)
            binned_dls_temp = self.centered_cls_sampler.sample(skymap)
            dls_temp = utils.unfold_bins(binned_dls_temp, self.bins)
            s_nonCentered = f64(skymap) * utils.expand_per_l(dls_temp, 4)          # sqrt(1/C) s (ASIS.py:104-106)
            binned_dls, var_cls_full, accept = self.non_centered_cls_sampler.sample(s_nonCentered, binned_dls_temp,
                                                                                    utils.generate_var_cl(dls_temp))
            dls = utils.unfold_bins(binned_dls, self.bins)
            cls = dls * f64(self.dls_to_cls_array)
            skymap = torch.sqrt(f64(var_cls_full)) * s_nonCentered                   # re-centre (ASIS.py:116)
            torch.cuda.synchronize()
            h_time_seconds.append(time.perf_counter() - t0)
            h_accept.append(accept)
            h_dls.append(_dev.to_host(binned_dls))
        return np.array(h_dls), np.array(h_accept), np.array(h_accept_cr), np.array(h_time_seconds)

    def run_polarization(self, dls_init):
        """Mirror of ASIS.run_polarization (ASIS.py:134-226); same 7-tuple.  Deviation: the re-centring at
        ASIS.py:203 multiplies the CENTRED map by sqrt(C_new); the interweaving step is s = sqrt(C_new) s_nc
        (as the TT twin does at ASIS.py:116), which is what is computed here."""
        h_duration_cr, h_duration_cls_nc_sampling, h_duration_cls_sampling, h_iteration_duration = [], [], [], []
        accept = {"EE": [], "BB": []}
        accept_cr = []
        h_dls = {"EE": [], "BB": []}
        binned_dls = {k: f64(v) for k, v in dls_init.items()}
        all_dls = {"EE": self._unfold(binned_dls, "EE"), "BB": self._unfold(binned_dls, "BB")}
        if self.rj_step == True or self.gibbs_cr == True:
            skymap, _ = self.constrained_sampler.sample(all_dls)
        for i in range(self.n_iter):
            if self.verbose:
                print("Interweaving, iteration: " + str(i))
            t_it = t0 = time.perf_counter()
            if self.rj_step is False and self.gibbs_cr is False:
                skymap, _ = self.constrained_sampler.sample(all_dls)
            elif self.rj_step is True:
                skymap, acc = self.constrained_sampler.sample_mask_rj(all_dls, skymap)
                accept_cr.append(acc)
            else:
                skymap, acc = self.constrained_sampler.sample(all_dls, skymap)
                accept_cr.append(acc)
            torch.cuda.synchronize()
            h_duration_cr.append(time.perf_counter() - t0)

            t0 = time.perf_counter()
            binned_dls_temp = self.centered_cls_sampler.sample(skymap)
            torch.cuda.synchronize()
            h_duration_cls_sampling.append(time.perf_counter() - t0)
            dls_temp = {"EE": self._unfold(binned_dls_temp, "EE"), "BB": self._unfold(binned_dls_temp, "BB")}
            s_nonCentered = self._scale(skymap, dls_temp, 4)           # sqrt(1/C) s (ASIS.py:181-189)
            t0 = time.perf_counter()
            binned_dls, acception = self.non_centered_cls_sampler.sample(s_nonCentered, binned_dls_temp)
            h_duration_cls_nc_sampling.append(time.perf_counter() - t0)
            accept["EE"].append(acception["EE"])
            accept["BB"].append(acception["BB"])
            all_dls = {"EE": self._unfold(binned_dls, "EE"), "BB": self._unfold(binned_dls, "BB")}
            skymap = self._scale(s_nonCentered, all_dls, 3)            # s = sqrt(C_new) s_nc
            h_dls["EE"].append(_dev.to_host(binned_dls["EE"]))
            h_dls["BB"].append(_dev.to_host(binned_dls["BB"]))
            h_iteration_duration.append(time.perf_counter() - t_it)
        total_accept = {"EE": np.array(accept["EE"]), "BB": np.array(accept["BB"])}
        h_dls["EE"] = np.array(h_dls["EE"])
        h_dls["BB"] = np.array(h_dls["BB"])
        acr = np.array(accept_cr) if self.rj_step else None
        return (h_dls, total_accept, acr, np.array(h_iteration_duration), np.array(h_duration_cr),
                np.array(h_duration_cls_sampling), np.array(h_duration_cls_nc_sampling))
