"""Host-side plumbing shared by the sampler classes: tensor conversion, RNG front-end, masks.

PyTorch is used only for device memory and streams; every computation is a call into
libgibbs_b200.so through gibbssampler_b200._lib."""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import check


def device():
    if not torch.cuda.is_available():
        raise _lib.GibbsB200Error("gibbssampler_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def f64(x, dev=None):
    """numpy / list / tensor -> contiguous float64 CUDA tensor (no copy if already one)."""
    dev = dev or device()
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=dev)


def i32(x, dev=None):
    dev = dev or device()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x), dtype=np.int32), device=dev)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_host(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def lmax_from_real(n):
    l = int(math.isqrt(int(n))) - 1
    if (l + 1) ** 2 != n:
        raise ValueError("length %d is not (lmax+1)^2" % n)
    return l


def lmax_from_complex(n):
    l = int((-3 + math.isqrt(9 + 8 * (int(n) - 1))) // 2)
    if (l + 1) * (l + 2) // 2 != n:
        raise ValueError("length %d is not (lmax+1)(lmax+2)/2" % n)
    return l


def gauss_beam(fwhm_rad, lmax):
    """hp.gauss_beam(fwhm, lmax) (temperature window; ConstrainedRealization.py:31, GibbsSampler.py:72)."""
    sigma = fwhm_rad / math.sqrt(8.0 * math.log(2.0))
    ell = np.arange(lmax + 1, dtype=np.float64)
    return np.exp(-0.5 * ell * (ell + 1.0) * sigma * sigma)


def load_mask(mask_path, nside, mask=None):
    """The reference reads a FITS mask and ud_grades it (ConstrainedRealization.py:33-37, CenteredGibbs.py:266-268):
    hp.ud_grade(hp.read_map(mask_path), nside).  Same here, with the package's own FITS reader and ud_grade
    (healpix_io.py; healpy is not a dependency); a mask array (RING order, already at `nside`) or a .npy path is
    accepted as well."""
    if mask is not None:
        m = to_host(mask).astype(np.float64)
    elif mask_path is None:
        return None
    elif str(mask_path).endswith(".npy"):
        m = np.load(mask_path).astype(np.float64)
    else:
        from . import healpix_io
        m = healpix_io.ud_grade(healpix_io.read_map(mask_path), nside)
    if m.shape != (12 * nside * nside,):
        raise ValueError("mask must have 12 nside^2 = %d RING pixels, got %s" % (12 * nside * nside, m.shape))
    return m


class Rng:
    """Standard normals / uniforms for the samplers.

    mode "philox": generated on the device (gs_randn, Philox4x32-10), the production path.
    mode "numpy" : drawn from numpy's legacy global state in the reference's order and copied
                   to the device -- the injected-draw parity mode (SURVEY.md 7.2 "RNG parity")."""

    def __init__(self, mode="philox", seed=None):
        if mode not in ("philox", "numpy"):
            raise ValueError("rng mode must be 'philox' or 'numpy'")
        self.mode = mode
        self.seed = int(seed) if seed is not None else int.from_bytes(os.urandom(8), "little")
        self.counter = 0

    def normal(self, n):
        if self.mode == "numpy":
            return f64(np.random.normal(loc=0, scale=1, size=n))
        out = torch.empty(n, dtype=torch.float64, device=device())
        self.counter += 1
        check(_lib.lib().gs_randn(ptr(out), n, self.seed, self.counter, stream()))
        return out

    def uniform(self, n):
        if self.mode == "numpy":
            return f64(np.random.uniform(size=n))
        out = torch.empty(n, dtype=torch.float64, device=device())
        self.counter += 1
        check(_lib.lib().gs_randu(ptr(out), n, self.seed, self.counter, stream()))
        return out


def dsum(x):
    """Deterministic device sum (gs_sum) -> python float."""
    scratch = torch.empty(592, dtype=torch.float64, device=x.device)
    out = torch.empty(1, dtype=torch.float64, device=x.device)
    check(_lib.lib().gs_sum(ptr(x), x.numel(), ptr(scratch), ptr(out), stream()))
    return float(out.item())
