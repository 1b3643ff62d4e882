"""Per-multipole 3x3 TT/TE/EE/BB kernels against the numpy restatement of the
reference's recovered helpers (oracle/reference_logic.py) and against scipy's inverse-Wishart moments."""
import numpy as np
import pytest
import torch

from oracle import reference_logic as R

pytestmark = pytest.mark.gpu


def dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), device="cuda")


def spectra(lmax, rng):
    ell = np.arange(lmax + 1)
    tt = 1000.0 / (ell + 10.0) ** 2 * (1 + 0.1 * rng.random(lmax + 1))
    ee = 20.0 / (ell + 10.0) ** 2 * (1 + 0.1 * rng.random(lmax + 1))
    bb = 1.0 / (ell + 10.0) ** 2 * (1 + 0.1 * rng.random(lmax + 1))
    te = 0.6 * np.sqrt(tt * ee) * np.cos(ell / 7.0)
    cls = np.zeros((lmax + 1, 3, 3))
    cls[:, 0, 0], cls[:, 1, 1], cls[:, 2, 2], cls[:, 0, 1], cls[:, 1, 0] = tt, ee, bb, te, te
    return cls


@pytest.mark.parametrize("lmax", [2, 3, 7, 64, 300])
def test_expand_inverse_cholesky_and_matvec_vs_oracle(lmax):
    from gibbssampler_b200 import _lib
    from gibbssampler_b200._dev import ptr, stream
    L = _lib.lib()
    rng = np.random.default_rng(lmax)
    cls = spectra(lmax, rng)
    n = (lmax + 1) ** 2
    # expansion: index map bit-exact, values to rounding of the D_l -> C_l scaling
    out = torch.empty(n, 3, 3, dtype=torch.float64, device="cuda")
    cls_d = dev(cls)   # keep device inputs alive: a temporary's memory may be reused before the kernel runs
    _lib.check(L.gs_expand_var_cl_3x3(ptr(cls_d), lmax, ptr(out), stream()))
    ref = R.expand_var_cl_3x3(cls)
    assert np.array_equal(out.cpu().numpy() != 0, ref != 0)
    assert np.allclose(out.cpu().numpy(), ref, rtol=1e-15, atol=0)
    # an l-tagged input makes the index map check exact
    tag = np.zeros((lmax + 1, 3, 3))
    tag[:, 0, 0] = np.arange(lmax + 1)
    tag[0, 0, 0] = -1.0
    tag_d = dev(tag)
    _lib.check(L.gs_expand_var_cl_3x3(ptr(tag_d), lmax, ptr(out), stream()))
    ell = R.l_of_real_layout(lmax)
    got = out[:, 0, 0].cpu().numpy()
    assert np.array_equal(np.round(got[ell > 0] * ell[ell > 0] * (ell[ell > 0] + 1) / (2 * np.pi)).astype(int), ell[ell > 0])
    assert np.all(got[ell == 0] == -1.0)
    # Sigma_l and its Cholesky factor
    pix = rng.random((lmax + 1, 3)) * 50 + 1
    sig = torch.empty(lmax + 1, 3, 3, dtype=torch.float64, device="cuda")
    cho = torch.empty_like(sig)
    pix_d = dev(pix)
    _lib.check(L.gs_inv_chol_3x3(ptr(cls_d), ptr(pix_d), lmax, ptr(sig), ptr(cho), stream()))
    rs, rc = R.compute_inverse_and_cholesky(cls, pix)
    gs_, gc_ = sig.cpu().numpy(), cho.cpu().numpy()
    for l in range(2, lmax + 1):  # LAPACK's own inverse carries cond * eps: compare per l relative to the matrix scale
        assert np.abs(gs_[l] - rs[l]).max() < 1e-11 * np.abs(rs[l]).max()
        assert np.abs(gc_[l] - rc[l]).max() < 1e-11 * np.abs(rc[l]).max()
        assert np.abs(gc_[l] @ gc_[l].T - gs_[l]).max() < 1e-14 * np.abs(gs_[l]).max()   # L L^T = Sigma
    assert np.all(sig.cpu().numpy()[:2] == 0)
    # matrix_product
    v = rng.standard_normal((n, 3))
    o = torch.empty(n, 3, dtype=torch.float64, device="cuda")
    v_d = dev(v)
    _lib.check(L.gs_matvec_3x3(ptr(sig), ptr(v_d), None, lmax, ptr(o), stream()))
    assert np.abs(o.cpu().numpy() - R.matrix_product(rs, v)).max() < 1e-10 * np.abs(rs).max() * np.abs(v).max()


def test_cross_spectrum_and_inverse_wishart_injected_draws_vs_oracle():
    from gibbssampler_b200.JointSampler import JointClsSampler
    lmax = 40
    rng = np.random.default_rng(0)
    n = (lmax + 1) ** 2
    alms = {k: rng.standard_normal(n) for k in ("TT", "EE", "BB")}
    alms["EE"] = 0.5 * alms["TT"] + alms["EE"]
    s = JointClsSampler(lmax, seed=1)
    ch = s.empirical({k: dev(v) for k, v in alms.items()})
    ell = R.l_of_real_layout(lmax)
    for key, (a, b) in {"TT": ("TT", "TT"), "TE": ("TT", "EE"), "EE": ("EE", "EE")}.items():
        ref = np.bincount(ell, weights=alms[a] * alms[b], minlength=lmax + 1) / (2 * np.arange(lmax + 1) + 1)
        assert np.allclose(ch[key].cpu().numpy(), ref, rtol=1e-13)
    df = 2 * np.arange(lmax + 1) - 2.0
    draws = np.zeros((lmax + 1, 3))
    draws[2:, 0] = rng.chisquare(df[2:])
    draws[2:, 1] = rng.chisquare(np.maximum(df[2:] - 1, 1))
    draws[2:, 2] = rng.standard_normal(lmax - 1)
    gam = rng.gamma(np.maximum((2 * np.arange(lmax + 1) - 1) / 2, 0.5))
    out = s.sample({k: dev(v) for k, v in alms.items()}, inject=draws, gamma_inject=gam)
    ref = R.invwishart_bartlett(ch["TT"].cpu().numpy(), ch["TE"].cpu().numpy(), ch["EE"].cpu().numpy(), draws)
    c2d = np.arange(lmax + 1) * (np.arange(lmax + 1) + 1) / (2 * np.pi)
    assert np.allclose(out["TT"].cpu().numpy(), ref[:, 0, 0] * c2d, rtol=1e-11)
    assert np.allclose(out["TE"].cpu().numpy(), ref[:, 0, 1] * c2d, rtol=1e-11, atol=1e-18)
    assert np.allclose(out["EE"].cpu().numpy(), ref[:, 1, 1] * c2d, rtol=1e-11)
    assert np.all(out["TT"].cpu().numpy()[:2] == 0) and np.all(out["BB"].cpu().numpy()[:2] == 0)


def test_inverse_wishart_device_draws_have_the_right_moments():
    """E[X] = Psi / (df - 3) for IW(df, Psi), p = 2; 4000 device draws at l = 12 and l = 30."""
    from gibbssampler_b200 import _lib
    from gibbssampler_b200._dev import ptr, stream
    L = _lib.lib()
    lmax, ndraw = 30, 4000
    tt, te, ee = np.full(lmax + 1, 2.0), np.full(lmax + 1, 0.7), np.full(lmax + 1, 1.0)
    o = [torch.empty(lmax + 1, dtype=torch.float64, device="cuda") for _ in range(3)]
    acc = np.zeros((3, lmax + 1))
    tt_d, te_d, ee_d = dev(tt), dev(te), dev(ee)
    for k in range(ndraw):
        _lib.check(L.gs_cls_invwishart(ptr(tt_d), ptr(te_d), ptr(ee_d), lmax, None, 1234, k + 1, ptr(o[0]), ptr(o[1]), ptr(o[2]),
                                       stream()))
        acc += np.stack([x.cpu().numpy() for x in o])
    acc /= ndraw
    for l in (12, 30):
        df, f = 2 * l - 2, 2 * l + 1
        for i, v in enumerate((tt, te, ee)):
            mean = f * v[l] / (df - 3)
            # var of IW entries is O(mean^2 / df); 5 sigma of the Monte-Carlo mean
            assert abs(acc[i, l] - mean) < 5 * mean * np.sqrt(4.0 / df / ndraw) + 1e-3 * abs(mean), (l, i, acc[i, l], mean)


def test_joint_constrained_realization_moments():
    """s | C, d is Gaussian with mean Sigma_l B N^-1 d and covariance Sigma_l per coefficient: check with injected draws
    against the oracle, and the sample covariance of device draws at one multipole."""
    from gibbssampler_b200 import _dev
    from gibbssampler_b200.JointSampler import JointConstrainedRealization
    lmax, nside = 24, 16
    npix = 12 * nside ** 2
    rng = np.random.default_rng(4)
    cls = spectra(lmax, rng)
    ell = np.arange(lmax + 1)
    c2d = ell * (ell + 1) / (2 * np.pi)
    dls = {"TT": cls[:, 0, 0] * c2d, "EE": cls[:, 1, 1] * c2d, "BB": cls[:, 2, 2] * c2d, "TE": cls[:, 0, 1] * c2d}
    n = (lmax + 1) ** 2
    d = {k: rng.standard_normal(n) for k in ("TT", "EE", "BB")}
    bl = _dev.gauss_beam(np.radians(3.0), lmax)
    nt, npol = 4.0, 0.5
    cr = JointConstrainedRealization(d, nt, npol, bl, lmax, npix, seed=9)
    xi = rng.standard_normal((n, 3))
    s, _ = cr.sample(dls, xi=xi)
    w = np.array([npix / (4 * np.pi * nt), npix / (4 * np.pi * npol), npix / (4 * np.pi * npol)])
    clm = cls.copy()
    clm[0] = cls[0] * 0  # D_0 = 0 -> C_0 = 0 (l = 0 copied unscaled)
    clm[1] = cls[1]
    pix = bl[:, None] ** 2 * w[None, :]
    rs, rc = R.compute_inverse_and_cholesky(cls, pix)
    lof = R.l_of_real_layout(lmax)
    bwd = np.stack([d["TT"], d["EE"], d["BB"]], axis=1) * bl[lof][:, None] * w[None, :]
    ref = R.matrix_product(rs, bwd) + R.matrix_product(rc, xi)
    got = np.stack([s["TT"].cpu().numpy(), s["EE"].cpu().numpy(), s["BB"].cpu().numpy()], axis=1)
    sel = lof >= 2
    assert np.allclose(got[sel], ref[sel], rtol=1e-10, atol=1e-14)
    assert np.all(got[~sel] == 0)


def test_joint_gibbs_chain_runs_and_recovers_input_spectra():
    """Short TT/TE/EE/BB chain on a high-S/N full-sky simulation: posterior means of D_l track the realised spectra."""
    from gibbssampler_b200 import _dev
    from gibbssampler_b200.JointSampler import JointGibbs
    lmax, nside = 32, 16
    npix = 12 * nside ** 2
    rng = np.random.default_rng(8)
    cls = spectra(lmax, rng)
    lof = R.l_of_real_layout(lmax)
    chol = np.zeros_like(cls)
    chol[2:] = np.linalg.cholesky(cls[2:])
    s_true = np.einsum("iab,ib->ia", chol[lof], rng.standard_normal(((lmax + 1) ** 2, 3)))
    fwhm = 1.0
    bl = _dev.gauss_beam(np.radians(fwhm), lmax)
    nt, npol = 1e-6, 1e-8
    nl = np.array([nt, npol, npol]) * 4 * np.pi / npix
    d = s_true * bl[lof][:, None] + rng.standard_normal(s_true.shape) * np.sqrt(nl)[None, :]
    ell = np.arange(lmax + 1)
    c2d = ell * (ell + 1) / (2 * np.pi)
    init = {"TT": cls[:, 0, 0] * c2d, "EE": cls[:, 1, 1] * c2d, "BB": cls[:, 2, 2] * c2d, "TE": cls[:, 0, 1] * c2d}
    gs = JointGibbs({"TT": d[:, 0], "EE": d[:, 1], "BB": d[:, 2]}, nt, npol, fwhm, nside, lmax, n_iter=400, seed=3)
    h = gs.run(init)
    # noise is negligible: the chain samples C | s_true, whose mean is (2l+1) Chat / (2l - 5) (TT/TE/EE)
    for i, key in enumerate(("TT", "EE")):
        chat = np.bincount(lof, weights=s_true[:, i] ** 2, minlength=lmax + 1) / (2 * ell + 1) * c2d
        for l in (10, 20, 30):
            expect = (2 * l + 1) * chat[l] / (2 * l - 5)
            got = h[key][50:, l].mean()
            assert abs(got - expect) < 0.25 * expect, (key, l, got, expect)
    assert np.all(np.isfinite(h["TE"])) and np.all(h["BB"][1:, 2:] > 0)
