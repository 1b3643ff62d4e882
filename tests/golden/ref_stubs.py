"""Stub modules that let the reference's own Python files (/root/reference/*.py) be imported in this
container, where healpy, qcinv, classy and matplotlib are absent.  The stubs provide
exactly the third-party calls the hot path makes, implemented with the CPU oracle (oracle/sht.py,
oracle/reference_logic.py).  Used only by make_golden.py to generate tests/golden/*.npz; nothing in
the GPU tests, smoke() or bench.py imports this file or reads /root/reference."""
import sys
import time
import types

import numpy as np

from oracle import reference_logic as R
from oracle import sht as O


def install(nside, lmax, mask_path=None, bins=None, blocks=None):
    np.complex = complex            # removed in numpy 1.24 (CenteredGibbs.py:167,486,645)
    time.clock = time.perf_counter  # removed in Python 3.8 (GibbsSampler.py:151)

    # ---- healpy ------------------------------------------------------------------------
    hp = types.ModuleType("healpy")
    hp.npix2nside = lambda npix: int(round((npix / 12) ** 0.5))
    hp.gauss_beam = lambda fwhm, lmax=lmax, pol=False: O.gauss_beam(fwhm, lmax)
    hp.read_map = lambda path, field=0, **kw: np.load(path)
    hp.ud_grade = lambda m, nside_out, **kw: np.asarray(m)
    hp.alm2cl = lambda alm, lmax=lmax: O.alm2cl(np.asarray(alm), lmax)

    def almxfl(alm, fl, inplace=False):
        out = O.almxfl(np.asarray(alm), np.asarray(fl), lmax)
        if inplace:
            alm[:] = out
            return alm
        return out

    def alm2map(alms, nside=nside, lmax=lmax, pol=True, **kw):
        if isinstance(alms, (list, tuple)) or (isinstance(alms, np.ndarray) and alms.ndim == 2):
            t = O.alm2map(np.asarray(alms[0]), nside, lmax)
            q, u = O.alm2map_spin2(np.asarray(alms[1]), np.asarray(alms[2]), nside, lmax)
            return np.array([t, q, u])
        return O.alm2map(np.asarray(alms), nside, lmax)

    def map2alm(maps, lmax=lmax, iter=3, pol=True, **kw):
        if isinstance(maps, (list, tuple)) or (isinstance(maps, np.ndarray) and maps.ndim == 2):
            t = O.map2alm(np.asarray(maps[0]), nside, lmax, iter=iter)
            e, b = O.map2alm_spin2(np.asarray(maps[1]), np.asarray(maps[2]), nside, lmax, iter=iter)
            return np.array([t, e, b])
        return O.map2alm(np.asarray(maps), nside, lmax, iter=iter)

    hp.almxfl, hp.alm2map, hp.map2alm = almxfl, alm2map, map2alm
    sys.modules["healpy"] = hp

    # ---- qcinv (forked; only the calls at CenteredGibbs.py:281-282, 467, 486-488) --------------
    qc = types.ModuleType("qcinv")

    class eblm:
        def __init__(self, arr):
            self.elm, self.blm = arr[0], arr[1]

    class alm_filter_ninv:
        def __init__(self, n_inv, b_transf, marge_maps=None):
            self.n_inv = n_inv[0] if isinstance(n_inv, list) else n_inv
            self.b_transf = b_transf

    class multigrid_chain:
        log = []

        def __init__(self, opfilt, chain_descr, s_cls, n_inv_filt, debug_log_prefix=None):
            self.opfilt, self.descr, self.s_cls, self.f = qc.opfilt_pp, chain_descr[0], s_cls, n_inv_filt

        def sample(self, soltn, pix_map, fluctuations, pol=False):
            """b = calc_prep(d) + fluctuations; PCG with diag_cl preconditioner; solution written into soltn."""
            itermax, eps = self.descr[4], self.descr[5]
            prob = R.PolProblem.__new__(R.PolProblem)
            prob.nside, prob.lmax, prob.kind, prob.npix = nside, lmax, "ld", 12 * nside * nside
            prob.inv_noise, prob.bl_gauss = self.f.n_inv, self.f.b_transf
            prob.bl_map = R.expand_per_l(prob.bl_gauss)
            e, b = R.adjoint_pol(pix_map[0] * prob.inv_noise, pix_map[1] * prob.inv_noise, nside, lmax, 0)
            bE = e * prob.bl_map + R.complex_to_real(fluctuations["elm"])
            bB = b * prob.bl_map + R.complex_to_real(fluctuations["blm"])
            ell = np.arange(lmax + 1)
            with np.errstate(divide="ignore", invalid="ignore"):
                dlE = np.where(ell > 0, self.s_cls.clee * ell * (ell + 1) / (2 * np.pi), self.s_cls.clee)
                dlB = np.where(ell > 0, self.s_cls.clbb * ell * (ell + 1) / (2 * np.pi), self.s_cls.clbb)
            x0 = None
            if np.any(soltn.elm != 0) or np.any(soltn.blm != 0):
                x0 = (R.complex_to_real(soltn.elm), R.complex_to_real(soltn.blm))
            xE, xB, it, res = prob.pcg(dlE, dlB, bE, bB, eps=eps, itermax=itermax, x0=x0)
            soltn.elm[:] = R.real_to_complex(xE)
            soltn.blm[:] = R.real_to_complex(xB)
            multigrid_chain.log.append(dict(bE=bE, bB=bB, it=it, res=res))
            return eblm(np.array([R.real_to_complex(bE), R.real_to_complex(bB)]))

    class fwd_op:
        """opfilt_pp.fwd_op(s_cls, n_inv_filt)(alm) = Q alm (CenteredGibbs.py:629, 653)"""

        def __init__(self, s_cls, n_inv_filt):
            self.s_cls, self.f = s_cls, n_inv_filt

        def __call__(self, x):
            prob = R.PolProblem.__new__(R.PolProblem)
            prob.nside, prob.lmax, prob.kind, prob.npix = nside, lmax, "ld", 12 * nside * nside
            prob.inv_noise, prob.bl_gauss = self.f.n_inv, self.f.b_transf
            prob.bl_map = R.expand_per_l(prob.bl_gauss)
            ell = np.arange(lmax + 1)
            dlE = np.where(ell > 0, self.s_cls.clee * ell * (ell + 1) / (2 * np.pi), self.s_cls.clee)
            dlB = np.where(ell > 0, self.s_cls.clbb * ell * (ell + 1) / (2 * np.pi), self.s_cls.clbb)
            yE, yB = prob.apply_Q(dlE, dlB, R.complex_to_real(x.elm), R.complex_to_real(x.blm))
            return eblm(np.array([R.real_to_complex(yE), R.real_to_complex(yB)]))

    for name in ("opfilt_pp", "opfilt_tt", "cd_solve", "multigrid", "util_alm"):
        setattr(qc, name, types.ModuleType("qcinv." + name))
    qc.opfilt_pp.alm_filter_ninv = alm_filter_ninv
    qc.opfilt_tt.alm_filter_ninv = alm_filter_ninv
    qc.opfilt_pp.eblm = eblm
    qc.opfilt_pp.fwd_op = fwd_op
    qc.cd_solve.tr_cg = "tr_cg"
    qc.cd_solve.cache_mem = lambda: {}
    qc.multigrid.multigrid_chain = multigrid_chain
    qc.util_alm.lmax2nlm = lambda l: (l + 1) * (l + 2) // 2
    sys.modules["qcinv"] = qc

    # ---- classy / matplotlib: imported at module level by utils.py:4,7 and CenteredGibbs.py:12 ----
    classy = types.ModuleType("classy")
    classy.Class = lambda: None
    sys.modules["classy"] = classy
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = mpl.pyplot

    # ---- config (config.py has import-time side effects: env vars, FITS mask, healpy) -----------
    cfg = types.ModuleType("config")
    cfg.NSIDE, cfg.L_MAX_SCALARS, cfg.Npix = nside, lmax, 12 * nside * nside
    cfg.rescaling_map2alm = cfg.Npix / (4 * np.pi)          # config.py:72
    cfg.w = 4 * np.pi / cfg.Npix                            # config.py:73
    cfg.mask_path = mask_path
    cfg.bins = bins
    cfg.blocks = blocks
    sys.modules["config"] = cfg
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    return qc
