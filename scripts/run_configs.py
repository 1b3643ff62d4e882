"""Run the BASELINE.json configurations that fit one GPU through the drop-in classes' run() and print one JSON line each
(wall clock around run(), device synchronised): #1 CenteredGibbs full-sky isotropic NSIDE 64 / lmax 128, 1000 iterations;
#2 NonCenteredGibbs and ASIS, polarised, NSIDE 256 / lmax 512, galactic mask, PCG.  (#3 is bench.py's default workload,
#4 is bench.py --mode sharded, #5 is scripts/sht_sweep.py.)   usage: python scripts/run_configs.py [iters_cfg2]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gibbssampler_b200 import _dev, utils  # noqa: E402
from gibbssampler_b200.ASIS import ASIS  # noqa: E402
from gibbssampler_b200.CenteredGibbs import CenteredGibbs  # noqa: E402
from gibbssampler_b200.NonCenteredGibbs import NonCenteredGibbs  # noqa: E402
from gibbssampler_b200.sht import Plan  # noqa: E402


def sky(nside, lmax, masked, fwhm):
    npix, nre = 12 * nside * nside, (lmax + 1) ** 2
    dlE, dlB = bench.fiducial(lmax)
    bl = _dev.gauss_beam(np.radians(fwhm), lmax)
    noise_var = 0.04 * npix / 786432.0
    plan = Plan.get(nside, lmax)
    rng = _dev.Rng("philox", seed=1234)
    sE = rng.normal(nre) * utils.expand_per_l(_dev.f64(dlE), 3)
    sB = rng.normal(nre) * utils.expand_per_l(_dev.f64(dlB), 3)
    q, u = plan.alm2map_spin2(sE, sB, fl=_dev.f64(bl))
    mask = bench.make_mask(nside) if masked else None
    m = _dev.f64(mask) if masked else 1.0
    dQ = (q + rng.normal(npix) * np.sqrt(noise_var)) * m
    dU = (u + rng.normal(npix) * np.sqrt(noise_var)) * m
    pix_map = {"Q": dQ, "U": dU}
    if not masked:  # harmonic data for the direct solve (main_polarization.py:44): (4 pi / Npix) A^T d
        e, b = plan.map2alm_spin2(dQ, dU, real_layout=True)
        pix_map["EE"], pix_map["BB"] = e, b
    return pix_map, mask, noise_var, dlE, dlB, bl


def binned(bins, dlE, dlB):
    return {p: np.array([dl[bins[p][i]:bins[p][i + 1]].mean() for i in range(len(bins[p]) - 1)]) for p, dl in (("EE", dlE), ("BB", dlB))}


def timed_run(g, init):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = g.run(init)
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


def main():
    n2 = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    # ---- config #1
    nside, lmax, fwhm, n_iter = 64, 128, 0.5 * 8, 1000
    pix_map, _, nv, dlE, dlB, bl = sky(nside, lmax, False, fwhm)
    npix = 12 * nside * nside
    bins = {"EE": np.arange(0, lmax + 2), "BB": np.arange(0, lmax + 2)}
    g = CenteredGibbs(pix_map, np.full(npix, nv * 1e4), np.full(npix, nv), fwhm, nside, lmax, npix, polarization=True, bins=bins,
                      n_iter=n_iter, seed=1)
    (h, *_), dt = timed_run(g, binned(bins, dlE, dlB))
    print(json.dumps({"config": 1, "workload": "CenteredGibbs full-sky isotropic noise (direct solve + inverse-gamma draw)", "nside": nside,
                      "lmax": lmax, "iterations": n_iter, "seconds": dt, "it_per_s": n_iter / dt,
                      "posterior_mean_DEE_l50": float(h["EE"][200:, 50].mean()), "input_DEE_l50": float(dlE[50])}), flush=True)
    # the same chain in ONE C call (gs_gibbs_run_centered_fullsky: the iteration as a replayed CUDA graph)
    g.run_fused(binned(bins, dlE, dlB))   # warm-up
    for use_graph in (True, False):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = g.run_fused(binned(bins, dlE, dlB), use_graph=use_graph)[0]
        dt = time.perf_counter() - t0
        print(json.dumps({"config": 1, "workload": "same chain through CenteredGibbs.run_fused (one C call, %s)" % ("CUDA graph replayed per iteration" if use_graph else "plain launches"),
                          "nside": nside, "lmax": lmax, "iterations": n_iter, "seconds": dt, "it_per_s": n_iter / dt, "us_per_iteration": 1e6 * dt / n_iter,
                          "posterior_mean_DEE_l50": float(h["EE"][200:, 50].mean()), "input_DEE_l50": float(dlE[50])}), flush=True)
    if len(sys.argv) > 2 and sys.argv[2] == "config1":
        return
    # ---- config #2
    nside, lmax, fwhm = 256, 512, 1.0
    pix_map, mask, nv, dlE, dlB, bl = sky(nside, lmax, True, fwhm)
    npix = 12 * nside * nside
    bins = bench.bins_for(lmax)
    blocks = bench.blocks_for(lmax, bins, 2)
    pv = bench.proposal_variances_for(lmax, bins, dlE, dlB, nv, npix, bl)
    init = binned(bins, dlE, dlB)
    for name, cls in (("NonCenteredGibbs", NonCenteredGibbs), ("ASIS", ASIS)):
        g = cls(pix_map, np.full(npix, nv * 1e4), np.full(npix, nv), fwhm, nside, lmax, npix, pv, metropolis_blocks=blocks,
                polarization=True, bins=bins, n_iter=1, mask=mask, seed=2)
        g.run(init)  # warm-up (plan tables, workspaces)
        g.n_iter = n2
        out, dt = timed_run(g, init)
        acc = out[1]
        rate = float(np.mean(np.concatenate([np.asarray(acc["EE"]).ravel(), np.asarray(acc["BB"]).ravel()]))) if isinstance(acc, dict) else None
        cr = g.constrained_sampler
        cr = getattr(cr, "pol_centered_constraint_realizer", cr)
        print(json.dumps({"config": 2, "workload": name + " polarised, galactic mask f_sky 0.8, PCG CR + blocked MwG (%d blocks)"
                          % (len(blocks["EE"]) + len(blocks["BB"]) - 2), "nside": nside, "lmax": lmax, "iterations": n2, "seconds": dt,
                          "it_per_s": n2 / dt, "last_pcg_iterations": int(cr.last_pcg_iterations), "mwg_accept_rate": rate}), flush=True)


if __name__ == "__main__":
    main()
