"""Block-batched Metropolis-within-Gibbs sweep (gs_mwg_sweep_blocks) against the per-block path
(one spin-2 synthesis per block, the literal NonCenteredGibbs.py:401-445) and against the numpy oracle likelihood."""
import numpy as np
import pytest
import torch

from oracle import reference_logic as R
from oracle import sht as O
from tests.test_cr_gpu import make_problem

pytestmark = pytest.mark.gpu


def build(P, bins, blocks, n_iter, l_cut, batched, workspace_bytes=0):
    from gibbssampler_b200 import utils
    from gibbssampler_b200.NonCenteredGibbs import PolarizationNonCenteredClsSampler
    lmax, nside = P["lmax"], P["nside"]
    bl_map = utils.expand_per_l(O.gauss_beam(np.radians(P["fwhm"]), lmax))
    pv = {p: np.full(len(bins[p]) - 1, 0.05) * np.concatenate([[0, 0], np.linspace(1, 0.1, len(bins[p]) - 3)]) for p in ("EE", "BB")}
    pv = {p: pv[p][2:] ** 2 + 1e-6 for p in pv}
    return PolarizationNonCenteredClsSampler({"Q": P["dQ"], "U": P["dU"]}, lmax, nside, bins, bl_map, np.full(P["npix"], 1600.0),
                                             P["noise"], blocks, pv, n_iter=n_iter, mask=P["mask"], rng="numpy", l_cut=l_cut,
                                             batched_blocks=batched, workspace_bytes=workspace_bytes)


def binned_start(P, bins):
    out = {}
    for p, dl in (("EE", P["dlE"]), ("BB", P["dlB"])):
        b = np.asarray(bins[p])
        out[p] = np.array([dl[b[i]:b[i + 1]].mean() for i in range(len(b) - 1)])
        out[p][:2] = 0.0
    return out


@pytest.mark.parametrize("nside,lmax,l_cut,n_iter,ws", [(16, 32, 0, 1, 0), (16, 32, 5, 2, 0), (32, 64, 0, 1, 1), (32, 80, 3, 1, 1200000)])
def test_batched_sweep_equals_per_block_sweep(nside, lmax, l_cut, n_iter, ws):
    P = make_problem(nside, lmax, seed=3)
    # EE: one multipole per bin; BB: coarser bins at the top (config.py:45-46 shape)
    ee = np.arange(0, lmax + 2)
    cut = (3 * lmax) // 4
    bb = np.concatenate([np.arange(0, cut), np.arange(cut, lmax + 1, 3)])
    bb[-1] = lmax + 1
    bins = {"EE": ee, "BB": bb}
    # blocks over the binned arrays (config.py:51-52 shape): one big block + single-bin blocks
    blocks = {"EE": [2, len(ee) // 2, len(ee) - 1], "BB": np.concatenate([[2, cut // 2], np.arange(cut // 2 + 1, len(bb))])}
    rng = np.random.default_rng(11)
    s_nc = {p: rng.standard_normal((lmax + 1) ** 2) for p in ("EE", "BB")}
    start = binned_start(P, bins)
    res = []
    for batched in (False, True):
        nc = build(P, bins, blocks, n_iter, l_cut, batched, ws)
        assert nc.batched_blocks == batched
        np.random.seed(77)
        dls, acc = nc.sample(s_nc, {k: v.copy() for k, v in start.items()})
        res.append((dls, acc, nc))
    (d0, a0, _), (d1, a1, nc) = res
    assert a0["EE"] == a1["EE"] and a0["BB"] == a1["BB"]
    assert 0 < sum(a0["BB"]) + sum(a0["EE"]) < len(a0["BB"]) + len(a0["EE"])   # both outcomes exercised
    for p in ("EE", "BB"):
        assert np.array_equal(d0[p], d1[p])
    # the final likelihood the batched sweep tracks incrementally equals a fresh evaluation at the final state
    lik = nc.compute_log_likelihood(d1, s_nc)
    from types import SimpleNamespace
    prob = SimpleNamespace(nside=nside, lmax=lmax, bl_map=R.expand_per_l(O.gauss_beam(np.radians(P["fwhm"]), lmax)), kind="ld",
                           inv_noise=P["mask"] / P["noise"], d_Q=P["dQ"], d_U=P["dU"])
    if l_cut == 0:
        ref = R.nc_loglik(d1, bins, s_nc, prob)
        assert abs(lik - ref) <= 1e-9 * abs(ref)


def test_batched_sweep_tracks_likelihood():
    """loglik_out of the sweep (r updated incrementally through accepted blocks) == likelihood recomputed from scratch."""
    from gibbssampler_b200 import _lib
    from gibbssampler_b200._dev import f64, ptr, stream
    nside, lmax = 32, 64
    P = make_problem(nside, lmax, seed=5)
    ee = np.arange(0, lmax + 2)
    bins = {"EE": ee, "BB": ee.copy()}
    blocks = {"EE": np.arange(2, lmax + 2), "BB": np.arange(2, lmax + 2)}
    nc = build(P, bins, blocks, 1, 0, True)
    rng = np.random.default_rng(2)
    s = {p: f64(rng.standard_normal((lmax + 1) ** 2)) for p in ("EE", "BB")}
    cur = {p: f64(binned_start(P, bins)[p]) for p in ("EE", "BB")}
    np.random.seed(5)
    prop = nc.propose_dl(cur)
    logr = {p: torch.zeros_like(cur[p]) for p in cur}
    nblk = lmax - 1
    ntot = 2 * nblk
    u = f64(np.random.uniform(size=ntot))
    acc = torch.zeros(ntot, dtype=torch.int32, device="cuda")
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    bh = {p: np.ascontiguousarray(bins[p], dtype=np.int32) for p in bins}
    kh = {p: np.ascontiguousarray(blocks[p], dtype=np.int32) for p in blocks}
    _lib.check(_lib.lib().gs_mwg_sweep_blocks(nc.plan._h, ptr(s["EE"]), ptr(s["BB"]), ptr(cur["EE"]), ptr(cur["BB"]), ptr(prop["EE"]),
                                              ptr(prop["BB"]), ptr(logr["EE"]), ptr(logr["BB"]), bh["EE"].ctypes.data, lmax + 1,
                                              bh["BB"].ctypes.data, lmax + 1, kh["EE"].ctypes.data, nblk, kh["BB"].ctypes.data, nblk,
                                              1, ptr(nc.bl_gauss_d), 0, ptr(nc.d_Q), ptr(nc.d_U), ptr(nc.inv_noise_pol), ptr(u),
                                              ptr(acc), ptr(out), 0, stream()))
    fresh = nc.compute_log_likelihood(cur, s)
    assert 10 < int(acc.sum()) < ntot - 10
    assert abs(float(out.item()) - fresh) <= 1e-10 * abs(fresh)
