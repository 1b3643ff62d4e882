"""Joint TT/TE/EE/BB Gibbs sampling with per-multipole 3x3 covariances.

The reference never finished this path: the per-l 3x3 helpers survive only as bytecode / object files
(utils.compute_inverse_and_cholesky, utils.matrix_product, linear_algebra.pyx) and a buggy Cython expansion
(variance_expension.pyx:36-61); the intended C_l conditional is the inverse-Wishart of
.ipynb_checkpoints/main-checkpoint.py:39-44,333-346.  Built here for the case those helpers were written for
-- full sky, isotropic noise, data in harmonic space -- where the constrained realization is a per-coefficient
3x3 solve:
    Sigma_l = (C_l^-1 + diag(b_l^2 w_X))^-1,  s = Sigma_l (b_l w_X d_X)_X + chol(Sigma_l) xi,  w_X = Npix / (4 pi sigma^2_X)
and  (C^TT, C^TE, C^EE)_l | s ~ IW(2l - 2, (2l+1) Chat_l),  C^BB_l | s ~ inverse-gamma as in CenteredGibbs.py:54-79.
Every step is a batched one-thread-per-l (or per-coefficient) kernel of libgibbs_b200.so."""
import numpy as np
import torch

from . import _dev, _lib
from ._dev import f64, ptr, stream
from ._lib import check, GS_ALM_REAL


class JointConstrainedRealization():
    def __init__(self, pix_map, noise_temp, noise_pol, bl_gauss, lmax, Npix, *, rng="philox", seed=None):
        """pix_map: {"TT","EE","BB"} data alms in the real layout; noise_*: per-pixel noise variances (scalars)."""
        self.lmax, self.Npix = int(lmax), int(Npix)
        self.dev = _dev.device()
        self.n = (self.lmax + 1) ** 2
        self.bl = f64(bl_gauss)
        self.w = torch.tensor([self.Npix / (4 * np.pi * float(noise_temp)), self.Npix / (4 * np.pi * float(noise_pol)),
                               self.Npix / (4 * np.pi * float(noise_pol))], dtype=torch.float64, device=self.dev)
        self.rng = rng if isinstance(rng, _dev.Rng) else _dev.Rng(rng, seed)
        d = torch.stack([f64(pix_map["TT"]), f64(pix_map["EE"]), f64(pix_map["BB"])], dim=1).contiguous()   # (n, 3)
        from . import utils
        blx = utils.expand_per_l(self.bl, 0)
        self.b_w_d = (d * blx[:, None] * self.w[None, :]).contiguous()               # B N^-1 d per coefficient
        self.pix_part = (self.bl[:, None] ** 2 * self.w[None, :]).contiguous()      # (L+1, 3): b_l^2 w_X

    def covariances(self, all_dls):
        """D_l dict {"TT","EE","BB","TE"} (unbinned, L+1) -> per-l C_l matrices (L+1,3,3)."""
        ell = torch.arange(self.lmax + 1, dtype=torch.float64, device=self.dev)
        f = torch.where(ell > 0, 2 * np.pi / (ell * (ell + 1)).clamp(min=1), torch.ones_like(ell))
        c = torch.zeros(self.lmax + 1, 3, 3, dtype=torch.float64, device=self.dev)
        c[:, 0, 0] = f64(all_dls["TT"]) * f
        c[:, 1, 1] = f64(all_dls["EE"]) * f
        c[:, 2, 2] = f64(all_dls["BB"]) * f
        c[:, 0, 1] = c[:, 1, 0] = f64(all_dls["TE"]) * f
        return c.contiguous()

    def sample(self, all_dls, xi=None):
        L = _lib.lib()
        cl = self.covariances(all_dls)
        sig = torch.empty_like(cl)
        cho = torch.empty_like(cl)
        check(L.gs_inv_chol_3x3(ptr(cl), ptr(self.pix_part), self.lmax, ptr(sig), ptr(cho), stream()))
        if xi is None:
            xi = self.rng.normal(3 * self.n)
        xi = f64(xi).reshape(self.n, 3).contiguous()
        out = torch.empty(self.n, 3, dtype=torch.float64, device=self.dev)
        check(L.gs_matvec_3x3(ptr(sig), ptr(self.b_w_d), None, self.lmax, ptr(out), stream()))     # mean
        check(L.gs_matvec_3x3(ptr(cho), ptr(xi), ptr(out), self.lmax, ptr(out), stream()))          # + fluctuation
        return {"TT": out[:, 0].contiguous(), "EE": out[:, 1].contiguous(), "BB": out[:, 2].contiguous()}, 1


class JointClsSampler():
    def __init__(self, lmax, *, rng="philox", seed=None):
        self.lmax = int(lmax)
        self.dev = _dev.device()
        self.rng = rng if isinstance(rng, _dev.Rng) else _dev.Rng(rng, seed)
        self._call = 0
        self.bins1 = _dev.i32(np.arange(self.lmax + 2))

    def empirical(self, alms):
        L = _lib.lib()
        out = {}
        dev_alms = {k: f64(alms[k]) for k in ("TT", "EE", "BB")}   # held until the launches below are queued
        for key, (a, b) in {"TT": ("TT", "TT"), "EE": ("EE", "EE"), "BB": ("BB", "BB"), "TE": ("TT", "EE")}.items():
            cl = torch.empty(self.lmax + 1, dtype=torch.float64, device=self.dev)
            check(L.gs_alm2cl_cross(ptr(dev_alms[a]), ptr(dev_alms[b]), self.lmax, ptr(cl), stream()))
            out[key] = cl
        return out

    def sample(self, alms, inject=None, gamma_inject=None):
        """-> unbinned D_l dict {"TT","EE","BB","TE"}.  inject: (L+1,3) (chi2_df, chi2_{df-1}, normal) and gamma_inject
        (L+1) Gamma variates for parity runs; default Philox."""
        L = _lib.lib()
        ch = self.empirical(alms)
        self.rng.counter += 2          # shared generator counter (see ClsSampler._invgamma_draw): one index per draw kernel
        self._call = self.rng.counter - 1
        tt, te, ee = (torch.empty(self.lmax + 1, dtype=torch.float64, device=self.dev) for _ in range(3))
        inj = f64(inject).contiguous() if inject is not None else None
        check(L.gs_cls_invwishart(ptr(ch["TT"]), ptr(ch["TE"]), ptr(ch["EE"]), self.lmax, ptr(inj), self.rng.seed, self._call,
                                  ptr(tt), ptr(te), ptr(ee), stream()))
        ell = torch.arange(self.lmax + 1, dtype=torch.float64, device=self.dev)
        c2d = ell * (ell + 1) / (2 * np.pi)
        bb = torch.empty(self.lmax + 1, dtype=torch.float64, device=self.dev)
        gi = f64(gamma_inject).contiguous() if gamma_inject is not None else None
        check(L.gs_cls_invgamma(ptr(ch["BB"]), ptr(self.bins1), self.lmax + 1, ptr(gi), self.rng.seed, self._call + 1,
                                ptr(bb), None, None, stream()))
        return {"TT": tt * c2d, "TE": te * c2d, "EE": ee * c2d, "BB": bb}


class JointGibbs():
    """CR <-> C_l Gibbs loop for (T, E, B) with TE correlation; full sky, isotropic noise."""

    def __init__(self, pix_map, noise_temp, noise_pol, beam_fwhm_deg, nside, lmax, n_iter=1000, *, rng="philox", seed=None):
        self.lmax, self.nside, self.n_iter = int(lmax), int(nside), int(n_iter)
        bl = _dev.gauss_beam(np.radians(beam_fwhm_deg), self.lmax)
        shared = _dev.Rng(rng, seed)
        self.constrained_sampler = JointConstrainedRealization(pix_map, noise_temp, noise_pol, bl, lmax, 12 * nside ** 2, rng=shared)
        self.cls_sampler = JointClsSampler(lmax, rng=shared)

    def run(self, dls_init):
        keys = ("TT", "EE", "BB", "TE")
        h = {k: [_dev.to_host(f64(dls_init[k]))] for k in keys}
        dls = {k: f64(dls_init[k]) for k in keys}
        for _ in range(self.n_iter):
            sky, _ = self.constrained_sampler.sample(dls)
            dls = self.cls_sampler.sample(sky)
            for k in keys:
                h[k].append(_dev.to_host(dls[k]))
        return {k: np.array(v) for k, v in h.items()}
