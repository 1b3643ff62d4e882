"""Non-centred parametrisation (mirror of NonCenteredGibbs.py): CR step = centred draw times C^-1/2,
C_l step = blocked Metropolis-within-Gibbs whose likelihood costs one spin-2 synthesis per block.
The whole Metropolis sweep (filters -> synthesis -> chi^2 reduction -> accept/commit) runs on the
device without host synchronisation; only the accept flags are read back at the end."""
import numpy as np
import torch

from . import _dev, _lib, utils
from ._dev import f64, ptr, stream
from ._lib import check, GS_ALM_REAL
from .CenteredGibbs import PolarizedCenteredConstrainedRealization
from .ClsSampler import MHClsSampler
from .ConstrainedRealization import ConstrainedRealization
from .GibbsSampler import GibbsSampler


class PolarizedNonCenteredConstrainedRealization(ConstrainedRealization):
    def __init__(self, pix_map, noise_temp, noise_pol, bl_map, lmax, Npix, bl_fwhm, mask_path=None, all_sph=False, *,
                 mask=None, rng="philox", seed=None):
        """Mirror of NonCenteredGibbs.py:106-135."""
        super().__init__(pix_map, noise_temp, bl_map, bl_fwhm, lmax, Npix, mask=None, rng=rng, seed=seed)
        self.noise_temp = noise_temp
        self.noise_pol = noise_pol
        self.mask_path = mask_path
        self.masked = mask is not None or mask_path is not None
        self.all_sph = all_sph
        self.bl_fwhm = bl_fwhm
        self.pol_centered_constraint_realizer = PolarizedCenteredConstrainedRealization(
            pix_map, noise_temp, noise_pol, bl_map, lmax, Npix, bl_fwhm, mask_path=mask_path, mask=mask, rng=self.rng)
        self.noise_pol0 = self.pol_centered_constraint_realizer.noise_pol0

    def _data_alm(self):
        """Harmonic-space data term of the no-mask draw, as d with  r = sqrt(C) b (Npix N^-1[0] / 4 pi) d :
        all_sph (NonCenteredGibbs.py:162-168): d = pix_map["EE"/"BB"];
        otherwise (NonCenteredGibbs.py:155-160): (Npix/4pi) map2alm([0, Q N^-1, U N^-1]) with healpy's default iter = 3, i.e.
        d = map2alm_iter3(Q N^-1, U N^-1) / N^-1[0] in the real layout (computed once, the data do not change)."""
        c = self.pol_centered_constraint_realizer
        if self.all_sph or c.d_Q is None:
            if c.d_E is None:
                raise _lib.GibbsB200Error("sample_no_mask needs pix_map['EE'] / ['BB'] (all_sph) or the pixel maps 'Q' / 'U'")
            return c.d_E, c.d_B
        if getattr(self, "_d_pix_alm", None) is None:
            inv0 = float(c.inv_noise_pol[0].item())
            e, b = self.plan.map2alm_spin2(c.d_Q * c.inv_noise_pol, c.d_U * c.inv_noise_pol, iter=3, real_layout=True)
            self._d_pix_alm = (e / inv0, b / inv0)
        return self._d_pix_alm

    def sample_no_mask(self, all_dls):
        """Full sky, isotropic noise (NonCenteredGibbs.py:138-176): harmonic-space data when all_sph, else the pixel maps."""
        c = self.pol_centered_constraint_realizer
        d_E, d_B = self._data_alm()
        dle, dlb = c._dls(all_dls)
        w = self.Npix / (self.noise_pol0 * 4 * np.pi)
        n = self.dimension_alm
        xe, xb = self.rng.normal(n), self.rng.normal(n)
        oe, ob = torch.empty_like(xe), torch.empty_like(xb)
        L = _lib.lib()
        check(L.gs_cr_direct(ptr(dle), ptr(self.bl_gauss_d), ptr(d_E), ptr(xe), w, self.lmax, 1, ptr(oe), stream()))
        check(L.gs_cr_direct(ptr(dlb), ptr(self.bl_gauss_d), ptr(d_B), ptr(xb), w, self.lmax, 1, ptr(ob), stream()))
        return c._ret({"EE": oe, "BB": ob}, all_dls["EE"]), 0

    def sample_mask(self, all_dls):
        """Centred PCG draw, then s_nc = C^-1/2 s (NonCenteredGibbs.py:178-195)."""
        c = self.pol_centered_constraint_realizer
        dle, dlb = c._dls(all_dls)
        alms, _ = c.sample_mask({"EE": dle, "BB": dlb})
        for pol, dl in (("EE", dle), ("BB", dlb)):
            sq = utils.expand_per_l(dl, 4)
            check(_lib.lib().gs_mul(ptr(alms[pol]), ptr(sq), ptr(alms[pol]), alms[pol].numel(), stream()))
        return c._ret(alms, all_dls["EE"]), 1

    def sample(self, all_dls):
        if not self.masked:
            return self.sample_no_mask(all_dls)
        return self.sample_mask(all_dls)


class PolarizationNonCenteredClsSampler(MHClsSampler):
    def __init__(self, pix_map, lmax, nside, bins, bl_map, noise_I, noise_Q, metropolis_blocks, proposal_variances, n_iter=1,
                 mask_path=None, polarization=True, all_sph=False, *, mask=None, rng="philox", seed=None, l_cut=0,
                 batched_blocks=True, workspace_bytes=0):
        """Mirror of NonCenteredGibbs.py:256-289.  l_cut > 0 keeps l < l_cut centred (PNCP).
        batched_blocks: run the sweep through gs_mwg_sweep_blocks (one Legendre pass for all blocks; NSIDE <= 1024)
        instead of one synthesis per block; both give the same chain."""
        super().__init__(pix_map, lmax, nside, bins, bl_map, noise_I, metropolis_blocks, proposal_variances, n_iter=n_iter,
                         polarization=polarization, mask=None, rng=rng, seed=seed)
        self.noise_temp = noise_I
        self.noise_pol = noise_Q
        self.mask_path = mask_path
        m = _dev.load_mask(mask_path, self.nside, mask)
        self.inv_noise_pol = 1.0 / f64(noise_Q)
        if self.inv_noise_pol.numel() == 1:
            self.inv_noise_pol = self.inv_noise_pol.expand(12 * self.nside ** 2).contiguous()
        if m is not None:
            self.inv_noise_pol = self.inv_noise_pol * f64(m)
        self.sigma = 0.8
        self.Npix = 12 * nside ** 2
        self.all_sph = bool(all_sph)
        self.l_cut = int(l_cut)
        from .sht import Plan
        self.plan = Plan.get(self.nside, self.lmax)
        if self.all_sph:
            # full sky, isotropic noise, data in harmonic space (NonCenteredGibbs.py:273-274, 357-377): the likelihood is a sum
            # over the real alm layout weighted by N^-1[0] Npix / 4 pi
            if m is not None:
                raise ValueError("all_sph=True needs a full sky (no mask), NonCenteredGibbs.py:273-274")
            self.d_E, self.d_B = f64(pix_map["EE"]), f64(pix_map["BB"])
            self.allsph_weight = float(self.inv_noise_pol[0].item()) * self.Npix / (4.0 * np.pi)
        if "Q" in pix_map and "U" in pix_map:
            self.d_Q, self.d_U = f64(pix_map["Q"]), f64(pix_map["U"])
        elif not self.all_sph:
            raise KeyError("pix_map needs the pixel maps 'Q' and 'U'")
        fw = getattr(self, "bl_gauss", None)
        # b_l from the expanded bl_map: entries 0..lmax are the m = 0 column
        self.bl_gauss_d = f64(bl_map)[: self.lmax + 1].contiguous()
        self.bins_d = {p: _dev.i32(np.asarray(self.bins[p])) for p in ("EE", "BB")}
        self.nb = {p: len(self.bins[p]) - 1 for p in ("EE", "BB")}
        self.pv_d = {p: f64(self.proposal_variances[p]) for p in ("EE", "BB")}
        self._mq = torch.empty(self.Npix, dtype=torch.float64, device=self.dev)
        self._mu = torch.empty_like(self._mq)
        self._flE = torch.empty(self.lmax + 1, dtype=torch.float64, device=self.dev)
        self._flB = torch.empty_like(self._flE)
        self._scratch = torch.empty(592, dtype=torch.float64, device=self.dev)
        self.batched_blocks = bool(batched_blocks) and self.nside <= 1024 and not self.all_sph
        self.workspace_bytes = int(workspace_bytes)
        self._bins_h = {p: np.ascontiguousarray(self.bins[p], dtype=np.int32) for p in ("EE", "BB")}
        self._blocks_h = {p: np.ascontiguousarray(self.metropolis_blocks[p], dtype=np.int32) for p in ("EE", "BB")}

    # ---- reference-named helpers -----------------------------------------------------------------
    def propose_dl(self, dls_old):
        """NonCenteredGibbs.py:292-309 (EE drawn first, then BB)."""
        ee = self._propose(f64(dls_old["EE"]), self.pv_d["EE"])
        bb = self._propose(f64(dls_old["BB"]), self.pv_d["BB"])
        return {"EE": ee, "BB": bb}

    def compute_log_proposal(self, dl_old, dl_new):
        """NonCenteredGibbs.py:313-330: log q(dl_new | dl_old) per bin."""
        return {p: self._log_proposal(f64(dl_new[p]), f64(dl_old[p]), self.pv_d[p]) for p in ("EE", "BB")}

    def _loglik_device(self, cur, prop, pol, b0, b1, s_nc, out):
        """log-likelihood of (cur with block [b0,b1) of `pol` replaced by prop) -> out[0] on the device."""
        L = _lib.lib()
        pe = prop["EE"] if prop is not None else None
        pb = prop["BB"] if prop is not None else None
        check(L.gs_mwg_filters(ptr(cur["EE"]), ptr(cur["BB"]), ptr(pe), ptr(pb), ptr(self.bins_d["EE"]), self.nb["EE"],
                               ptr(self.bins_d["BB"]), self.nb["BB"], pol, b0, b1, ptr(self.bl_gauss_d), self.lmax, self.l_cut,
                               ptr(self._flE), ptr(self._flB), stream()))
        if self.all_sph:   # compute_log_MH_ratio's all_sph branch (NonCenteredGibbs.py:385-393)
            check(L.gs_loglik_alm(ptr(self.d_E), ptr(self.d_B), ptr(s_nc["EE"]), ptr(s_nc["BB"]), ptr(self._flE), ptr(self._flB),
                                  self.lmax, self.allsph_weight, ptr(self._scratch), ptr(out), stream()))
            return
        check(L.gs_alm2map_spin2_fl2(self.plan._h, ptr(s_nc["EE"]), ptr(s_nc["BB"]), GS_ALM_REAL, ptr(self._flE), ptr(self._flB),
                                     ptr(self._mq), ptr(self._mu), stream()))
        check(L.gs_loglik_pix(ptr(self.d_Q), ptr(self.d_U), ptr(self._mq), ptr(self._mu), ptr(self.inv_noise_pol), self.Npix,
                              ptr(self._scratch), ptr(out), stream()))

    def compute_log_likelihood(self, dls, s_nonCentered):
        """NonCenteredGibbs.py:333-355 (pixel domain; with all_sph=True the harmonic form of :357-377) -> python float."""
        cur = {p: f64(dls[p]) for p in ("EE", "BB")}
        s = {p: f64(s_nonCentered[p]) for p in ("EE", "BB")}
        out = torch.empty(1, dtype=torch.float64, device=self.dev)
        self._loglik_device(cur, None, -1, 0, 0, s, out)
        return float(out.item())

    def compute_log_likelihood_all_sph(self, dls, s_nonCentered):
        """NonCenteredGibbs.py:357-377 -> python float (needs all_sph=True at construction: harmonic data, no mask)."""
        if not self.all_sph:
            raise ValueError("compute_log_likelihood_all_sph needs a sampler built with all_sph=True")
        return self.compute_log_likelihood(dls, s_nonCentered)

    def sample(self, s_nonCentered, binned_dls_old):
        """Blocked Metropolis-within-Gibbs sweep (NonCenteredGibbs.py:401-445); same return values."""
        host = not isinstance(binned_dls_old["EE"], torch.Tensor)
        L = _lib.lib()
        s = {p: f64(s_nonCentered[p]) for p in ("EE", "BB")}
        cur = {p: f64(binned_dls_old[p]).clone() for p in ("EE", "BB")}
        prop = self.propose_dl(cur)
        num = self.compute_log_proposal(prop, cur)
        den = self.compute_log_proposal(cur, prop)
        logr = {p: (num[p] - den[p]).contiguous() for p in ("EE", "BB")}
        nblk = {p: len(self.metropolis_blocks[p]) - 1 for p in ("EE", "BB")}
        ntot = (nblk["EE"] + nblk["BB"]) * self.n_iter
        if self.rng.mode == "numpy":
            u = f64(np.array([np.random.uniform() for _ in range(ntot)]))
        else:
            u = self.rng.uniform(max(ntot, 2))
        acc = torch.zeros(max(ntot, 1), dtype=torch.int32, device=self.dev)
        if self.batched_blocks and ntot > 0:
            bh, kh = self._bins_h, self._blocks_h
            check(L.gs_mwg_sweep_blocks(self.plan._h, ptr(s["EE"]), ptr(s["BB"]), ptr(cur["EE"]), ptr(cur["BB"]), ptr(prop["EE"]),
                                        ptr(prop["BB"]), ptr(logr["EE"]), ptr(logr["BB"]), bh["EE"].ctypes.data, self.nb["EE"],
                                        bh["BB"].ctypes.data, self.nb["BB"], kh["EE"].ctypes.data, nblk["EE"], kh["BB"].ctypes.data,
                                        nblk["BB"], self.n_iter, ptr(self.bl_gauss_d), self.l_cut, ptr(self.d_Q), ptr(self.d_U),
                                        ptr(self.inv_noise_pol), ptr(u), ptr(acc), None, self.workspace_bytes, stream()))
        else:
            old_lik = torch.empty(1, dtype=torch.float64, device=self.dev)
            new_lik = torch.empty(1, dtype=torch.float64, device=self.dev)
            self._loglik_device(cur, None, -1, 0, 0, s, old_lik)
            k = 0
            for ip, pol in enumerate(("EE", "BB")):
                blocks = self.metropolis_blocks[pol]
                for i in range(nblk[pol]):
                    b0, b1 = int(blocks[i]), int(blocks[i + 1])
                    for _ in range(self.n_iter):
                        self._loglik_device(cur, prop, ip, b0, b1, s, new_lik)
                        check(L.gs_mwg_accept(ptr(cur[pol]), ptr(prop[pol]), ptr(logr[pol]), b0, b1, ptr(new_lik), ptr(old_lik),
                                              ptr(u[k:]), ptr(acc[k:]), stream()))
                        k += 1
        a = acc.cpu().numpy()
        ne = nblk["EE"] * self.n_iter
        accept = {"EE": [int(x) for x in a[:ne]], "BB": [int(x) for x in a[ne:ntot]]}
        if host:
            return {p: cur[p].cpu().numpy() for p in cur}, accept
        return cur, accept


class NonCenteredGibbs(GibbsSampler):
    def __init__(self, pix_map, noise_I, noise_Q, beam, nside, lmax, Npix, proposal_variances, metropolis_blocks=None,
                 polarization=False, bins=None, n_iter=10000, n_iter_metropolis=1, mask_path=None, all_sph=False, *,
                 mask=None, rng="philox", seed=None, verbose=False):
        """Mirror of NonCenteredGibbs.__init__ (NonCenteredGibbs.py:450-486)."""
        super().__init__(pix_map, noise_I, beam, nside, lmax, polarization=polarization, bins=bins, n_iter=n_iter, verbose=verbose)
        shared = _dev.Rng(rng, seed)
        if not polarization:  # NonCenteredGibbs.py:471-477
            from .Temperature import NonCenteredClsSampler, NonCenteredConstrainedRealization
            self.constrained_sampler = NonCenteredConstrainedRealization(pix_map, noise_I, self.bl_map, beam, lmax, Npix, isotropic=True,
                                                                         mask_path=mask_path, mask=mask, rng=shared)
            self.cls_sampler = NonCenteredClsSampler(pix_map, lmax, nside, self.bins, self.bl_map, noise_I, metropolis_blocks,
                                                     proposal_variances, n_iter=n_iter_metropolis, mask_path=mask_path, mask=mask,
                                                     rng=shared)
            return
        self.constrained_sampler = PolarizedNonCenteredConstrainedRealization(pix_map, noise_I, noise_Q, self.bl_map, lmax, Npix, beam,
                                                                              mask_path=mask_path, all_sph=all_sph, mask=mask, rng=shared)
        self.cls_sampler = PolarizationNonCenteredClsSampler(pix_map, lmax, nside, self.bins, self.bl_map, noise_I, noise_Q,
                                                             metropolis_blocks, proposal_variances, n_iter=n_iter_metropolis,
                                                             mask_path=mask_path, all_sph=all_sph, mask=mask, rng=shared)

    def run_temperature(self, dl_init):
        """Mirror of NonCenteredGibbs.run_temperature (NonCenteredGibbs.py:488-527); same return tuple."""
        import time
        h_time_seconds, total_accept, h_dl = [], [], []
        binned_dls = f64(dl_init)
        cls, var_cls = self._tt_state(binned_dls)
        for i in range(self.n_iter):
            if self.verbose and i % 100 == 0:
                print("Non centered gibbs")
                print(i)
            start_time = time.perf_counter()
            s_nonCentered, _ = self.constrained_sampler.sample(cls, var_cls, None, False)
            binned_dls, var_cls, accept = self.cls_sampler.sample(s_nonCentered, binned_dls, var_cls)
            cls = utils.unfold_bins(binned_dls, self.bins) * f64(self.dls_to_cls_array)
            total_accept.append(accept)
            h_dl.append(_dev.to_host(binned_dls))
            h_time_seconds.append(time.perf_counter() - start_time)
        return np.array(h_dl), np.array(total_accept), np.array(h_time_seconds)

    def run_polarization(self, dls_init):
        """Mirror of NonCenteredGibbs.run_polarization (NonCenteredGibbs.py:529-571); same return tuple."""
        h_duration_cr, h_duration_cls_sampling = [], []
        total_accept = {"EE": [], "BB": []}
        h_dls = {"EE": [], "BB": []}
        binned_dls = {k: f64(v) for k, v in dls_init.items()}
        dls_unbinned = {"EE": self._unfold(binned_dls, "EE"), "BB": self._unfold(binned_dls, "BB")}
        h_dls["EE"].append(_dev.to_host(binned_dls["EE"]))
        h_dls["BB"].append(_dev.to_host(binned_dls["BB"]))
        for i in range(self.n_iter):
            if self.verbose:
                print("Non centered gibbs")
                print(i)
            s_nonCentered, _ = self.constrained_sampler.sample(dls_unbinned)
            binned_dls, accept = self.cls_sampler.sample(s_nonCentered, binned_dls)
            dls_unbinned = {"EE": self._unfold(binned_dls, "EE"), "BB": self._unfold(binned_dls, "BB")}
            total_accept["EE"].append(accept["EE"])
            total_accept["BB"].append(accept["BB"])
            h_dls["EE"].append(_dev.to_host(binned_dls["EE"]))
            h_dls["BB"].append(_dev.to_host(binned_dls["BB"]))
        total_accept = {"EE": np.array(total_accept["EE"]), "BB": np.array(total_accept["BB"])}
        h_dls["EE"] = np.array(h_dls["EE"])
        h_dls["BB"] = np.array(h_dls["BB"])
        return h_dls, total_accept, np.array(h_duration_cr), np.array(h_duration_cls_sampling)
