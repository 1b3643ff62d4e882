"""Device-resident spherical-harmonic transforms: thin torch-tensor front-end of the C ABI.

The methods mirror the healpy calls of the reference but take and return CUDA
float64 / complex128 tensors that stay in HBM; nothing here computes on the CPU.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import GS_ALM_COMPLEX, GS_ALM_REAL, check


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Plan:
    """Geometry + recurrence tables + workspace for one (nside, lmax) on one GPU."""

    _cache = {}

    def __init__(self, nside, lmax, device=None):
        if not torch.cuda.is_available():
            raise _lib.GibbsB200Error("gibbssampler_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.nside, self.lmax = int(nside), int(lmax)
        self._h = C.c_void_p()
        check(_lib.lib().gs_plan_create(C.byref(self._h), self.nside, self.lmax, self.device.index))
        self.npix = 12 * self.nside ** 2
        self.nalm = (self.lmax + 1) * (self.lmax + 2) // 2
        self.nreal = (self.lmax + 1) ** 2
        self.npix_global, self.nreal_global = self.npix, self.nreal

    @classmethod
    def get(cls, nside, lmax, device=None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        key = (int(nside), int(lmax), dev.index)
        if key not in cls._cache:
            cls._cache[key] = cls(nside, lmax, dev)
        return cls._cache[key]

    def __del__(self):
        try:
            if self._h:
                _lib.lib().gs_plan_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ single-GPU versions of the shard helpers
    # (gibbssampler_b200.sharded.ShardedPlan overrides them; the sampler classes are written against these)
    world, rank = 1, 0

    def local_map(self, full):
        return full

    def local_alm(self, full_real):
        return full_real

    def gather_alm(self, local):
        return local

    def gather_map(self, local):
        return local

    def allreduce_sum(self, value):
        return float(value)

    def expand_per_l(self, x, mode=0):
        from . import utils
        return utils.expand_per_l(torch.as_tensor(x, dtype=torch.float64, device=self.device), mode)

    def alm2cl(self, alm_real):
        cl = torch.empty(self.lmax + 1, dtype=torch.float64, device=self.device)
        check(_lib.lib().gs_alm2cl(_ptr(alm_real.contiguous()), GS_ALM_REAL, self.lmax, _ptr(cl), _stream()))
        return cl

    # ------------------------------------------------------------------ helpers
    def _alm_in(self, a):
        """-> (contiguous tensor, layout flag)"""
        if a.dtype == torch.complex128:
            assert a.numel() == self.nalm, "complex alm must have (L+1)(L+2)/2 entries"
            return a.contiguous(), GS_ALM_COMPLEX
        assert a.dtype == torch.float64 and a.numel() == self.nreal, "real-layout alm must have (L+1)^2 float64 entries"
        return a.contiguous(), GS_ALM_REAL

    def _alm_out(self, layout):
        if layout == GS_ALM_COMPLEX:
            return torch.empty(self.nalm, dtype=torch.complex128, device=self.device)
        return torch.empty(self.nreal, dtype=torch.float64, device=self.device)

    def _fl(self, fl):
        if fl is None:
            return None
        fl = torch.as_tensor(fl, dtype=torch.float64, device=self.device).contiguous()
        assert fl.numel() == self.lmax + 1
        return fl

    # ------------------------------------------------------------------ transforms
    def alm2map(self, alm, fl=None, out=None):
        """hp.alm2map(alm, nside, lmax) of one spin-0 field (complex or real-layout alm)."""
        a, lay = self._alm_in(alm)
        fl = self._fl(fl)
        m = torch.empty(self.npix, dtype=torch.float64, device=self.device) if out is None else out
        check(_lib.lib().gs_alm2map_spin0(self._h, _ptr(a), lay, _ptr(fl), _ptr(m), _stream()))
        return m

    def alm2map_spin2(self, almE, almB, fl=None, out=None):
        """(Q, U) of hp.alm2map([0, E, B], pol=True)."""
        e, lay = self._alm_in(almE)
        b, lay2 = self._alm_in(almB)
        assert lay == lay2
        fl = self._fl(fl)
        if out is None:
            q = torch.empty(self.npix, dtype=torch.float64, device=self.device)
            u = torch.empty(self.npix, dtype=torch.float64, device=self.device)
        else:
            q, u = out
        check(_lib.lib().gs_alm2map_spin2(self._h, _ptr(e), _ptr(b), lay, _ptr(fl), _ptr(q), _ptr(u), _stream()))
        return q, u

    def map2alm(self, m, iter=0, adjoint=False, pixw=None, fl=None, real_layout=False):
        """hp.map2alm(m, lmax, iter=iter, use_weights=False); adjoint=True -> A^T m."""
        m = m.contiguous()
        assert m.dtype == torch.float64 and m.numel() == self.npix
        lay = GS_ALM_REAL if real_layout else GS_ALM_COMPLEX
        a = self._alm_out(lay)
        fl = self._fl(fl)
        check(_lib.lib().gs_map2alm_spin0(self._h, _ptr(m), _ptr(pixw), int(iter), int(bool(adjoint)), _ptr(fl),
                                          _ptr(a), lay, _stream()))
        return a

    def map2alm_spin2(self, q, u, iter=0, adjoint=False, pixw=None, fl=None, real_layout=False):
        """(E, B) of hp.map2alm([0, Q, U], lmax, pol=True, iter=iter, use_weights=False)."""
        q, u = q.contiguous(), u.contiguous()
        assert q.dtype == torch.float64 and q.numel() == self.npix and u.numel() == self.npix
        lay = GS_ALM_REAL if real_layout else GS_ALM_COMPLEX
        e, b = self._alm_out(lay), self._alm_out(lay)
        fl = self._fl(fl)
        check(_lib.lib().gs_map2alm_spin2(self._h, _ptr(q), _ptr(u), _ptr(pixw), int(iter), int(bool(adjoint)),
                                          _ptr(fl), _ptr(e), _ptr(b), lay, _stream()))
        return e, b

    # ------------------------------------------------------------------ chain batches (BASELINE config #5)
    def alm2map_spin2_batch(self, almE, almB, fl=None):
        """(Q, U) of n_chain right-hand sides at once: almE / almB are (n_chain, n) stacks (real layout, float64, or complex
        healpy layout); two chains per launch share one Legendre recurrence (gs_alm2map_batch).  Returns (n_chain, npix) maps."""
        e, b = almE.contiguous(), almB.contiguous()
        assert e.dim() == 2 and e.shape == b.shape and e.dtype == b.dtype
        k = e.shape[0]
        lay = GS_ALM_REAL if e.dtype == torch.float64 else GS_ALM_COMPLEX
        stride = e.shape[1] * (1 if lay == GS_ALM_REAL else 2)   # in doubles
        q = torch.empty((k, self.npix), dtype=torch.float64, device=self.device)
        u = torch.empty((k, self.npix), dtype=torch.float64, device=self.device)
        check(_lib.lib().gs_alm2map_batch(self._h, 2, k, _ptr(e), _ptr(b), stride, lay, _ptr(self._fl(fl)), _ptr(q), _ptr(u),
                                          self.npix, _stream()))
        return q, u

    def alm2map_batch(self, alm, fl=None):
        """spin-0 twin of alm2map_spin2_batch."""
        a = alm.contiguous()
        assert a.dim() == 2
        k = a.shape[0]
        lay = GS_ALM_REAL if a.dtype == torch.float64 else GS_ALM_COMPLEX
        stride = a.shape[1] * (1 if lay == GS_ALM_REAL else 2)
        m = torch.empty((k, self.npix), dtype=torch.float64, device=self.device)
        check(_lib.lib().gs_alm2map_batch(self._h, 0, k, _ptr(a), None, stride, lay, _ptr(self._fl(fl)), _ptr(m), None,
                                          self.npix, _stream()))
        return m

    def map2alm_spin2_batch(self, q, u, adjoint=False, pixw=None, fl=None, real_layout=False):
        """(E, B) of map2alm(iter=0) (adjoint=True: A^T) of (n_chain, npix) map stacks; (n_chain, n) alm stacks out."""
        q, u = q.contiguous(), u.contiguous()
        assert q.dim() == 2 and q.shape == u.shape and q.shape[1] == self.npix and q.dtype == torch.float64
        k = q.shape[0]
        lay = GS_ALM_REAL if real_layout else GS_ALM_COMPLEX
        n = self.nreal if real_layout else self.nalm
        dt = torch.float64 if real_layout else torch.complex128
        e = torch.empty((k, n), dtype=dt, device=self.device)
        b = torch.empty((k, n), dtype=dt, device=self.device)
        stride = n * (1 if real_layout else 2)
        check(_lib.lib().gs_map2alm_batch(self._h, 2, k, _ptr(q), _ptr(u), self.npix, _ptr(pixw), int(bool(adjoint)),
                                          _ptr(self._fl(fl)), _ptr(e), _ptr(b), stride, lay, _stream()))
        return e, b

    def map2alm_batch(self, m, adjoint=False, pixw=None, fl=None, real_layout=False):
        """spin-0 twin of map2alm_spin2_batch."""
        m = m.contiguous()
        assert m.dim() == 2 and m.shape[1] == self.npix and m.dtype == torch.float64
        k = m.shape[0]
        lay = GS_ALM_REAL if real_layout else GS_ALM_COMPLEX
        n = self.nreal if real_layout else self.nalm
        a = torch.empty((k, n), dtype=torch.float64 if real_layout else torch.complex128, device=self.device)
        stride = n * (1 if real_layout else 2)
        check(_lib.lib().gs_map2alm_batch(self._h, 0, k, _ptr(m), None, self.npix, _ptr(pixw), int(bool(adjoint)),
                                          _ptr(self._fl(fl)), _ptr(a), None, stride, lay, _stream()))
        return a
