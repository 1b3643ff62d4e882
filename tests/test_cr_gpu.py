"""GPU parity of the constrained-realization and C_l kernels against the numpy restatement of the
reference logic (oracle/reference_logic.py) on the same injected draws."""
import numpy as np
import pytest
import torch

from oracle import reference_logic as R
from oracle import sht as O

pytestmark = pytest.mark.gpu


def make_problem(nside, lmax, seed=0, fsky=0.8, kind="ld"):
    rng = np.random.default_rng(seed)
    npix = 12 * nside * nside
    th, ph = O.pix_angles(nside)
    mask = np.clip((np.abs(np.cos(th)) - np.cos(np.radians(90 - 90 * (1 - fsky) * 0.5))) / 0.05 + 0.5, 0, 1)
    noise = np.full(npix, 0.04 * (npix / 786432.0))
    ell = np.arange(lmax + 1)
    dlE = np.where(ell >= 2, 1.0 * np.exp(-(ell / 1200.0) ** 2) + 0.02, 0.0)
    dlB = np.where(ell >= 2, 0.05 * (np.maximum(ell, 1) / 80.0) ** -0.5 * np.exp(-(ell / 1500.0) ** 2) + 1e-3, 0.0)
    fwhm = 0.5 * 512 / nside if nside < 512 else 0.5
    bl = O.gauss_beam(np.radians(fwhm), lmax)
    sE = rng.standard_normal((lmax + 1) ** 2) * np.sqrt(R.generate_var_cl(dlE))
    sB = rng.standard_normal((lmax + 1) ** 2) * np.sqrt(R.generate_var_cl(dlB))
    q, u = R.synth_pol(sE * R.expand_per_l(bl), sB * R.expand_per_l(bl), nside, lmax, kind)
    dQ = (q + rng.standard_normal(npix) * np.sqrt(noise)) * mask
    dU = (u + rng.standard_normal(npix) * np.sqrt(noise)) * mask
    return dict(nside=nside, lmax=lmax, npix=npix, mask=mask, noise=noise, dlE=dlE, dlB=dlB, fwhm=fwhm, dQ=dQ, dU=dU,
                rng=rng)


def gpu_cr(P, **kw):
    from gibbssampler_b200.CenteredGibbs import PolarizedCenteredConstrainedRealization
    from gibbssampler_b200 import utils
    bl_map = utils.expand_per_l(O.gauss_beam(np.radians(P["fwhm"]), P["lmax"]))
    return PolarizedCenteredConstrainedRealization({"Q": P["dQ"], "U": P["dU"]}, np.full(P["npix"], 1600.0), P["noise"], bl_map,
                                                   P["lmax"], P["npix"], P["fwhm"], mask=P["mask"], **kw)


def relerr(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def test_rhs_and_operator_vs_oracle():
    P = make_problem(8, 16)
    prob = R.PolProblem(8, 16, P["dQ"], P["dU"], P["mask"] / P["noise"], P["fwhm"])
    cr = gpu_cr(P)
    rng = P["rng"]
    xi = (rng.standard_normal(P["npix"]), rng.standard_normal(P["npix"]), rng.standard_normal(17 ** 2), rng.standard_normal(17 ** 2))
    dls = {"EE": P["dlE"], "BB": P["dlB"]}
    be, bb = cr.build_rhs(dls, xi)
    re_, rb_ = prob.rhs(P["dlE"], P["dlB"], *xi)
    assert relerr(be.cpu().numpy(), re_) < 1e-10 and relerr(bb.cpu().numpy(), rb_) < 1e-10
    assert relerr(cr.second_part_grad_E.cpu().numpy(), prob.bdata_E) < 1e-10
    x = {"EE": rng.standard_normal(17 ** 2), "BB": rng.standard_normal(17 ** 2)}
    for k in x:
        x[k][[0, 1, 17, 18]] = 0
    y = cr.apply_Q(dls, x)
    ye, yb = prob.apply_Q(P["dlE"], P["dlB"], x["EE"], x["BB"])
    assert relerr(y["EE"].cpu().numpy(), ye) < 1e-10 and relerr(y["BB"].cpu().numpy(), yb) < 1e-10


def test_pcg_equals_dense_solve():
    """equality with a dense solve of Q at nside 4 / lmax 8."""
    P = make_problem(4, 8)
    prob = R.PolProblem(4, 8, P["dQ"], P["dU"], P["mask"] / P["noise"], P["fwhm"])
    cr = gpu_cr(P)
    cr.pcg_accuracy = 1e-13
    rng = P["rng"]
    xi = (rng.standard_normal(P["npix"]), rng.standard_normal(P["npix"]), rng.standard_normal(81), rng.standard_normal(81))
    dls = {"EE": P["dlE"], "BB": P["dlB"]}
    sol, acc = cr.sample_mask(dls, xi)
    assert acc == 1
    bE, bB = prob.rhs(P["dlE"], P["dlB"], *xi)
    Q = prob.dense_Q(P["dlE"], P["dlB"])
    keep = np.ones(162, bool)
    keep[[0, 1, 9, 10, 81, 82, 90, 91]] = False  # l < 2: Q is zero there and so is b
    xs = np.zeros(162)
    xs[keep] = np.linalg.solve(Q[np.ix_(keep, keep)], np.concatenate([bE, bB])[keep])
    got = np.concatenate([sol["EE"], sol["BB"]])
    assert np.abs(got[~keep]).max() == 0
    assert relerr(got, xs) < 1e-9
    assert cr.last_pcg_residual <= 1e-13 * 1.01


@pytest.mark.parametrize("nside,lmax", [(16, 32), (32, 64)])
def test_pcg_matches_oracle_pcg_and_converges(nside, lmax):
    P = make_problem(nside, lmax, seed=3)
    prob = R.PolProblem(nside, lmax, P["dQ"], P["dU"], P["mask"] / P["noise"], P["fwhm"], kind="f64")
    cr = gpu_cr(P)
    rng = P["rng"]
    n = (lmax + 1) ** 2
    xi = (rng.standard_normal(P["npix"]), rng.standard_normal(P["npix"]), rng.standard_normal(n), rng.standard_normal(n))
    dls = {"EE": P["dlE"], "BB": P["dlB"]}
    sol, _ = cr.sample_mask(dls, xi)
    bE, bB = prob.rhs(P["dlE"], P["dlB"], *xi)
    xE, xB, it, res = prob.pcg(P["dlE"], P["dlB"], bE, bB, eps=1e-5)
    # same algorithm, same stopping rule: iteration counts agree (summation order may shift it by one)
    assert abs(cr.last_pcg_iterations - it) <= 1
    assert cr.last_pcg_residual <= 1e-5
    # true residual of the GPU solution, evaluated by the oracle
    yE, yB = prob.apply_Q(P["dlE"], P["dlB"], sol["EE"], sol["BB"])
    r = np.concatenate([bE - yE, bB - yB])
    assert np.linalg.norm(r) / np.linalg.norm(np.concatenate([bE, bB])) < 2e-5
    assert relerr(np.concatenate([sol["EE"], sol["BB"]]), np.concatenate([xE, xB])) < 1e-3


def test_direct_solve_and_cls_draw_vs_oracle():
    from gibbssampler_b200.CenteredGibbs import PolarizedCenteredConstrainedRealization, PolarizedCenteredClsSampler
    from gibbssampler_b200 import utils
    nside, lmax = 8, 16
    npix, n = 12 * nside ** 2, (lmax + 1) ** 2
    P = make_problem(nside, lmax, seed=5)
    rng = P["rng"]
    dE, dB = rng.standard_normal(n), rng.standard_normal(n)
    bl_map = utils.expand_per_l(O.gauss_beam(np.radians(P["fwhm"]), lmax))
    np.random.seed(11)
    cr = PolarizedCenteredConstrainedRealization({"EE": dE, "BB": dB}, np.full(npix, 1600.0), P["noise"], bl_map, lmax, npix,
                                                 P["fwhm"], rng="numpy")
    sol, acc = cr.sample({"EE": P["dlE"], "BB": P["dlB"]})
    np.random.seed(11)
    xiE, xiB = np.random.normal(size=n), np.random.normal(size=n)
    refE = R.sample_no_mask(P["dlE"], bl_map, dE, xiE, npix, P["noise"][0])
    refB = R.sample_no_mask(P["dlB"], bl_map, dB, xiB, npix, P["noise"][0])
    assert relerr(sol["EE"], refE) < 1e-12 and relerr(sol["BB"], refB) < 1e-12
    # inverse-gamma draw with the reference's own numpy/scipy stream
    bins = {"EE": np.arange(0, lmax + 2), "BB": np.concatenate([np.arange(0, 10), [10, 12, 15, lmax + 1]])}
    cs = PolarizedCenteredClsSampler({"EE": dE, "BB": dB}, lmax, nside, bins, bl_map, P["noise"], rng="numpy")
    np.random.seed(12)
    got = cs.sample({"EE": refE, "BB": refB})
    from scipy.stats import invgamma
    np.random.seed(12)
    for pol, s in (("EE", refE), ("BB", refB)):
        a, b = R.cls_alpha_beta(s, bins[pol], lmax)
        ref = b * invgamma.rvs(a=a)
        ref[:2] = 0
        assert relerr(got[pol], ref) < 1e-12


@pytest.mark.parametrize("eps,check_every", [(3e-1, 8), (1e-1, 8), (1e-2, 8), (1e-3, 8), (1e-5, 8), (1e-3, 3), (1e-4, 1), (1e-5, 16)])
def test_pcg_graph_replay_equals_plain_launches(eps, check_every):
    """The iterations between two polls of the convergence flag are replayed from a CUDA graph with two batches in flight
    (solver.cu, gs_set_pcg_graph, default on); the same solve with plain launches and one poll at a time runs the same kernels in the
    same order: identical iteration count and bit-identical solution, whether the solve stops inside the first batch, at a batch
    boundary or in the iterations left over before iter_max."""
    from gibbssampler_b200 import _lib
    L = _lib.lib()
    P = make_problem(16, 32, seed=11)
    ell = np.arange(P["lmax"] + 1)
    dls = {"EE": P["dlE"], "BB": P["dlB"]}
    rng = np.random.default_rng(4)
    nre = (P["lmax"] + 1) ** 2
    xi = (rng.standard_normal(P["npix"]), rng.standard_normal(P["npix"]), rng.standard_normal(nre), rng.standard_normal(nre))
    del ell

    def solve(graph, itermax):
        old = L.gs_set_pcg_graph(graph)
        try:
            cr = gpu_cr(P)
            cr.pcg_accuracy, cr.pcg_check_every, cr.pcg_itermax = eps, check_every, itermax
            sol, _ = cr.sample_mask(dls, xi)
            solve.residual = cr.last_pcg_residual
            return np.concatenate([np.asarray(sol["EE"]), np.asarray(sol["BB"])]), cr.last_pcg_iterations
        finally:
            L.gs_set_pcg_graph(old)

    xa, ia = solve(1, 4000)
    xb, ib = solve(0, 4000)
    assert ia == ib and ia >= 1
    assert np.array_equal(xa, xb)
    # iter_max exactly at the iteration that converges (the device stops there in both modes), one batch above, and not a multiple
    # of check_every: the left-over iterations run as plain launches after the last whole batch
    for itermax in sorted({ia, ia + check_every, ia + 1}):
        xc, ic = solve(1, itermax)
        assert ic == ia and np.array_equal(xc, xa)
    if ia > 1:   # one iteration short: both modes stop at iter_max without having converged (as qcinv does, the residual says so)
        (xd, idd), rd = solve(1, ia - 1), solve.residual
        (xe_, ie), re_ = solve(0, ia - 1), solve.residual
        assert idd == ie == ia - 1 and rd == re_ and rd > eps
        assert np.array_equal(xd, xe_)
