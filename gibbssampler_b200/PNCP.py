"""Partially non-centred parametrisation (PNCP).

The reference ships PNCP only as bytecode (__pycache__/PNCP.cpython-38.pyc) and only for TT, full sky,
isotropic noise.  BASELINE.json config #3 asks for the polarised masked-sky variant,
which has no reference implementation; it is DEFINED here by analogy:

  * multipoles l < l_cut stay centred:   D_l | s  ~ inverse-gamma        (CenteredGibbs.py:54-79)
  * multipoles l >= l_cut are non-centred: D_l | s_nc by blocked Metropolis-within-Gibbs with the
    pixel-space likelihood of NonCenteredGibbs.py:333-355, where the synthesised field is
    A B (s_l for l < l_cut ; sqrt(C_l) s_nc,l for l >= l_cut)
  * the CR step is the centred PCG draw (CenteredGibbs.py:448-491) followed by s_nc = C^-1/2 s on
    l >= l_cut (generalising NonCenteredGibbs.py:192-194 and the recovered
    PNCPConstrainedRealization.sample).

Both conditionals leave the joint posterior invariant (each is a valid Gibbs / Metropolis update of D
under a bijective reparametrisation of s), so the chain targets the same distribution as CenteredGibbs;
tests/test_statistics_gpu.py checks that numerically.  The constructor keeps the recovered signature
(pix_map, noise, beam, nside, lmax, Npix, proposal_variances, l_cut, metropolis_blocks=None,
polarization=False, bins=None, n_iter=10000, n_iter_metropolis=1) with noise_Q / mask added."""
import numpy as np
import torch

from . import _dev, _lib, utils
from ._dev import f64, ptr, stream
from ._lib import check
from .CenteredGibbs import PolarizedCenteredClsSampler, PolarizedCenteredConstrainedRealization
from .GibbsSampler import GibbsSampler
from .NonCenteredGibbs import PolarizationNonCenteredClsSampler


class PNCPConstrainedRealization(PolarizedCenteredConstrainedRealization):
    """Centred PCG draw; `to_mixed` / `from_mixed` move l >= l_cut to / from the non-centred variable."""

    def __init__(self, *args, l_cut=5, **kw):
        super().__init__(*args, **kw)
        self.l_cut = int(l_cut)

    def _mixed(self, skymap, all_dls, mode):
        out = {}
        for pol in ("EE", "BB"):
            dl = f64(all_dls[pol])
            fl = torch.empty(self.lmax + 1, dtype=torch.float64, device=self.dev)
            check(_lib.lib().gs_pncp_factor(ptr(dl), self.lmax, self.l_cut, mode, ptr(fl), stream()))
            f = utils.expand_per_l(fl, 0)
            o = torch.empty_like(skymap[pol])
            check(_lib.lib().gs_mul(ptr(skymap[pol]), ptr(f), ptr(o), o.numel(), stream()))
            out[pol] = o
        return out

    def to_mixed(self, skymap, all_dls):
        return self._mixed(skymap, all_dls, 0)

    def from_mixed(self, mixed, all_dls):
        return self._mixed(mixed, all_dls, 1)


class PNCPClsSampler(PolarizationNonCenteredClsSampler):
    """low-l centred inverse-gamma draw + high-l non-centred Metropolis-within-Gibbs."""

    def __init__(self, pix_map, lmax, nside, bins, bl_map, noise_I, noise_Q, metropolis_blocks, proposal_variances, l_cut,
                 n_iter=1, mask_path=None, *, mask=None, rng="philox", seed=None):
        # blocks must only touch bins made of multipoles >= l_cut
        for pol in ("EE", "BB"):
            first = int(metropolis_blocks[pol][0])
            if int(np.asarray(bins[pol])[first]) < l_cut:
                raise ValueError("metropolis_blocks[%s] starts below l_cut" % pol)
        super().__init__(pix_map, lmax, nside, bins, bl_map, noise_I, noise_Q, metropolis_blocks, proposal_variances, n_iter=n_iter,
                         mask_path=mask_path, mask=mask, rng=rng, seed=seed, l_cut=l_cut)
        self.centered = PolarizedCenteredClsSampler(pix_map, lmax, nside, bins, bl_map, noise_I, rng=self.rng)
        self.low_bins = {p: int(np.searchsorted(np.asarray(bins[p]), l_cut, side="left")) for p in ("EE", "BB")}

    def sample_low_l(self, skymap, binned_dls):
        """Inverse-gamma draw of the bins below l_cut given the centred map (recovered PNCPClsSampler.sample_low_l)."""
        draw = self.centered.sample(skymap)
        out = {}
        for pol in ("EE", "BB"):
            o = f64(binned_dls[pol]).clone()
            k = self.low_bins[pol]
            o[:k] = draw[pol][:k]
            out[pol] = o
        return out

    def sample_high_l(self, mixed, binned_dls):
        """Blocked MwG on the bins >= l_cut (recovered PNCPClsSampler.sample_high_l)."""
        return self.sample(mixed, binned_dls)


class PNCPGibbs(GibbsSampler):
    def __init__(self, pix_map, noise, beam, nside, lmax, Npix, proposal_variances, l_cut, metropolis_blocks=None,
                 polarization=False, bins=None, n_iter=10000, n_iter_metropolis=1, *, noise_Q=None, mask_path=None, mask=None,
                 rng="philox", seed=None, verbose=False):
        super().__init__(pix_map, noise, beam, nside, lmax, polarization=polarization, bins=bins, n_iter=n_iter, verbose=verbose)
        self.l_cut = int(l_cut)
        if not polarization:  # the recovered TT class: full sky, isotropic noise
            from .Temperature import PNCPClsSamplerTT, PNCPConstrainedRealizationTT
            shared = _dev.Rng(rng, seed)
            self.constrained_sampler = PNCPConstrainedRealizationTT(pix_map, noise, self.bl_map, beam, lmax, Npix, isotropic=True,
                                                                    rng=shared, l_cut=l_cut)
            self.cls_sampler = PNCPClsSamplerTT(pix_map, lmax, nside, self.bins, self.bl_map, noise, metropolis_blocks,
                                                proposal_variances, l_cut, n_iter=n_iter_metropolis, rng=shared)
            return
        if noise_Q is None:
            raise ValueError("noise_Q (polarisation noise variance per pixel) is required")
        self.l_cut = int(l_cut)
        shared = _dev.Rng(rng, seed)
        self.constrained_sampler = PNCPConstrainedRealization(pix_map, noise, noise_Q, self.bl_map, lmax, Npix, beam, mask_path=mask_path,
                                                              mask=mask, rng=shared, ula=False, l_cut=l_cut)
        self.cls_sampler = PNCPClsSampler(pix_map, lmax, nside, self.bins, self.bl_map, noise, noise_Q, metropolis_blocks,
                                          proposal_variances, l_cut, n_iter=n_iter_metropolis, mask_path=mask_path, mask=mask, rng=shared)

    def run_temperature(self, dls_init):
        """Recovered PNCPGibbs.run loop for TT: CR -> low-l centred draw -> high-l non-centred MwG.
        Returns (h_dls, h_accept, h_time_cr)."""
        h_dls, h_accept, h_time = [], [], []
        binned = f64(dls_init)
        h_dls.append(_dev.to_host(binned))
        for i in range(self.n_iter):
            _, var_cls = self._tt_state(binned)
            mixed, t_cr, _ = self.constrained_sampler.sample(var_cls)
            binned = self.cls_sampler.sample_low_l(mixed, binned)
            binned, acc = self.cls_sampler.sample_high_l(mixed, binned)
            h_time.append(t_cr)
            h_accept.append(acc)
            h_dls.append(_dev.to_host(binned))
        return np.array(h_dls), np.array(h_accept), np.array(h_time)

    def run_polarization(self, dls_init):
        """CR -> low-l centred draw -> high-l non-centred MwG (recovered PNCPGibbs.run loop).
        Returns (h_dls, accept, h_duration_cr, h_duration_cls)."""
        import time
        h_dls = {"EE": [], "BB": []}
        accept = {"EE": [], "BB": []}
        t_cr, t_cls = [], []
        binned = {k: f64(v) for k, v in dls_init.items()}
        h_dls["EE"].append(_dev.to_host(binned["EE"]))
        h_dls["BB"].append(_dev.to_host(binned["BB"]))
        for i in range(self.n_iter):
            t0 = time.perf_counter()
            all_dls = {"EE": self._unfold(binned, "EE"), "BB": self._unfold(binned, "BB")}
            skymap, _ = self.constrained_sampler.sample_mask(all_dls)
            torch.cuda.synchronize()
            t_cr.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            binned = self.cls_sampler.sample_low_l(skymap, binned)
            all_dls = {"EE": self._unfold(binned, "EE"), "BB": self._unfold(binned, "BB")}
            mixed = self.constrained_sampler.to_mixed(skymap, all_dls)
            binned, acc = self.cls_sampler.sample_high_l(mixed, binned)
            torch.cuda.synchronize()
            t_cls.append(time.perf_counter() - t0)
            accept["EE"].append(acc["EE"])
            accept["BB"].append(acc["BB"])
            h_dls["EE"].append(_dev.to_host(binned["EE"]))
            h_dls["BB"].append(_dev.to_host(binned["BB"]))
        h_dls["EE"], h_dls["BB"] = np.array(h_dls["EE"]), np.array(h_dls["BB"])
        return h_dls, {k: np.array(v) for k, v in accept.items()}, np.array(t_cr), np.array(t_cls)
