"""Base class of the constrained-realization step (mirror of ConstrainedRealization.py:5-53)."""
import math

import numpy as np
import torch

from . import _dev, _lib
from ._dev import f64, ptr, stream
from ._lib import check
from .sht import Plan


class ConstrainedRealization():
    """Same constructor as the reference (ConstrainedRealization.py:8) plus two keyword-only
    extensions: ``mask`` (a RING-ordered array at the map's nside, instead of a FITS path: healpy is
    not a dependency) and ``rng`` ("philox": device draws; "numpy": numpy's global stream, injected
    in the reference's order).  The qcinv objects of the reference (n_inv_filt, chain_descr) are
    replaced by the device-resident N^-1, b_l and PCG settings below."""

    def __init__(self, pix_map, noise, bl_map, fwhm_deg, lmax, Npix, mask_path=None, isotropic=True,
                 *, mask=None, rng="philox", seed=None, plan=None):
        self.pix_map = pix_map
        self.isotropic = isotropic
        self.noise = noise
        self.dev = _dev.device()
        self.lmax = int(lmax)
        self.Npix = int(Npix)
        self.nside = int(round(math.sqrt(self.Npix / 12)))
        if 12 * self.nside ** 2 != self.Npix:
            raise ValueError("Npix = %d is not 12 nside^2" % self.Npix)
        # `plan` = a gibbssampler_b200.sharded.ShardedPlan runs the single chain m-sharded over the ranks of
        # its process group: full-sky inputs are cut to the local ring shard, alms are local m shards
        self.plan = plan if plan is not None else Plan.get(self.nside, self.lmax)
        self.dimension_alm = self.plan.nreal     # (lmax+1)^2 on one GPU
        self.npix_local = self.plan.npix         # Npix on one GPU
        self.fwhm_radians = (np.pi / 180) * fwhm_deg
        self.bl_gauss = _dev.gauss_beam(self.fwhm_radians, self.lmax)  # hp.gauss_beam (ConstrainedRealization.py:31)
        self.bl_gauss_d = f64(self.bl_gauss)
        self.bl_map = bl_map
        self.bl_map_d = f64(bl_map)
        self.mask_path = mask_path
        self._mask_arr = _dev.load_mask(mask_path, self.nside, mask)
        self.masked = self._mask_arr is not None
        self.inv_noise = self._inv_noise_from(noise)
        self.rng = rng if isinstance(rng, _dev.Rng) else _dev.Rng(rng, seed)
        # PCG settings of the reference's qcinv chain (ConstrainedRealization.py:41): diag_cl, 4000 its, 1e-6
        self.pcg_itermax = 4000
        self.pcg_accuracy = 1.0e-6
        self.mu = float(self.inv_noise.max().item()) + 0.0000001  # ConstrainedRealization.py:44
        self.last_pcg_iterations = 0
        self.last_pcg_residual = 0.0

    def _inv_noise_from(self, noise):
        n = f64(noise)
        if n.numel() == 1:
            n = n.expand(self.Npix).contiguous()
        inv = torch.empty_like(n)
        # 1/noise (* mask): ConstrainedRealization.py:26,37 -- a setup-time elementwise op
        inv.copy_(1.0 / n)
        if self._mask_arr is not None:
            inv.mul_(f64(self._mask_arr))
        return self.plan.local_map(inv)

    def sample(self, cls, var_cls):
        return None
