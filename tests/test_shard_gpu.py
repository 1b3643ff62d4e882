"""m-sharded SHT and sharded PCG (BASELINE config #4) against the single-GPU path.

Two harnesses over the SAME sharded kernels / index tables / solver code:
  * an in-process group (gs_local_group_create): `world` sharded plans on one device, one host thread
    and stream per rank, collectives = host barriers + device copies -- runs on a single-GPU box;
  * the production path: one process per GPU, NCCL all-to-all / all-reduce (tests/shard_worker.py under
    torch.distributed.run) -- runs when the box has >= 2 GPUs.
Tolerance: 1e-10 relative (FP64; BASELINE north_star) for transforms; the PCG solution to the solver's
own eps (the shards sum dot products in a different order)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-10


def relerr(got, ref):
    return float((got - ref).abs().max() / ref.abs().max())


def full_inputs(nside, lmax, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    nre, npix = (lmax + 1) ** 2, 12 * nside ** 2
    e = torch.randn(nre, generator=g, dtype=torch.float64).cuda()
    b = torch.randn(nre, generator=g, dtype=torch.float64).cuda()
    for a in (e, b):  # l < 2 carries no spin-2 signal
        a[[0, 1, lmax + 1, lmax + 2]] = 0
    q = torch.randn(npix, generator=g, dtype=torch.float64).cuda()
    u = torch.randn(npix, generator=g, dtype=torch.float64).cuda()
    w = torch.rand(npix, generator=g, dtype=torch.float64).cuda() + 0.5
    fl = torch.rand(lmax + 1, generator=g, dtype=torch.float64).cuda() + 0.5
    return e, b, q, u, w, fl


@pytest.mark.parametrize("nside,lmax,world,nb", [(8, 16, 2, 1), (16, 47, 3, 3), (32, 64, 4, 2), (64, 128, 8, 4), (128, 256, 2, 4)])
def test_local_group_transforms_match_single_gpu(nside, lmax, world, nb, monkeypatch):
    """nb: blocks of local m of the spectra layout (GS_SHARD_NB; production: 4 from 128 local m on), one all-to-all per block."""
    from gibbssampler_b200.sharded import ShardedPlan, run_local_group
    from gibbssampler_b200.sht import Plan
    monkeypatch.setenv("GS_SHARD_NB", str(nb))
    ref = Plan.get(nside, lmax)
    e, b, q, u, w, fl = full_inputs(nside, lmax, 5 + nside)
    rq, ru = ref.alm2map_spin2(e, b, fl=fl)
    re_, rb_ = ref.map2alm_spin2(q, u, adjoint=True, pixw=w, fl=fl, real_layout=True)
    ie, ib = ref.map2alm_spin2(q, u, iter=3, real_layout=True)
    r0 = ref.alm2map(e, fl=fl)
    a0 = ref.map2alm(q, adjoint=True, real_layout=True)
    plans = ShardedPlan.local_group(nside, lmax, world)

    def work(p):
        le, lb = p.local_alm(e), p.local_alm(b)
        lq, lu, lw = p.local_map(q), p.local_map(u), p.local_map(w)
        sq, su = p.alm2map_spin2(le, lb, fl=fl)
        ae, ab = p.map2alm_spin2(lq, lu, adjoint=True, pixw=lw, fl=fl, real_layout=True)
        je, jb = p.map2alm_spin2(lq, lu, iter=3, real_layout=True)
        s0 = p.alm2map(le, fl=fl)
        t0 = p.map2alm(lq, adjoint=True, real_layout=True)
        cl = p.alm2cl(le)
        return dict(sq=sq, su=su, ae=ae, ab=ab, je=je, jb=jb, s0=s0, t0=t0, cl=cl)

    res = run_local_group(plans, work)

    def join_map(key):
        out = torch.zeros(12 * nside ** 2, dtype=torch.float64, device="cuda")
        for p, r in zip(plans, res):
            out[p.pixel_index] = r[key]
        return out

    def join_alm(key):
        out = torch.zeros((lmax + 1) ** 2, dtype=torch.float64, device="cuda")
        for p, r in zip(plans, res):
            out[p.real_index] = r[key]
        return out

    assert relerr(join_map("sq"), rq) < RTOL and relerr(join_map("su"), ru) < RTOL
    assert relerr(join_alm("ae"), re_) < RTOL and relerr(join_alm("ab"), rb_) < RTOL
    assert relerr(join_alm("je"), ie) < RTOL and relerr(join_alm("jb"), ib) < RTOL
    assert relerr(join_map("s0"), r0) < RTOL
    assert relerr(join_alm("t0"), a0) < RTOL
    cl_ref = ref.alm2cl(e)
    for r in res:
        assert relerr(r["cl"], cl_ref) < 1e-13


def _cr_problem(nside, lmax, seed=3):
    """A small masked polarised CR system (same construction as bench.py)."""
    sys.path.insert(0, ROOT)
    import bench
    from gibbssampler_b200 import _dev
    rng = np.random.default_rng(seed)
    npix, nre = 12 * nside ** 2, (lmax + 1) ** 2
    dlE, dlB = bench.fiducial(lmax)
    dlE, dlB = dlE * (1200.0 / lmax) ** 0 + 0.0, dlB
    bl = _dev.gauss_beam(np.radians(0.5 * 512 / nside), lmax)
    noise_var = 0.04 * npix / 786432.0
    mask = bench.make_mask(nside)
    dQ = rng.standard_normal(npix) * mask
    dU = rng.standard_normal(npix) * mask
    xi = (rng.standard_normal(npix), rng.standard_normal(npix), rng.standard_normal(nre), rng.standard_normal(nre))
    return dlE, dlB, bl, noise_var, mask, dQ, dU, xi


@pytest.mark.parametrize("nside,lmax,world", [(16, 32, 2), (32, 64, 4)])
def test_local_group_cr_solve_matches_single_gpu(nside, lmax, world):
    from gibbssampler_b200 import _dev, _lib, utils
    from gibbssampler_b200.CenteredGibbs import PolarizedCenteredConstrainedRealization as CR
    from gibbssampler_b200.sharded import ShardedPlan, run_local_group
    dlE, dlB, bl, noise_var, mask, dQ, dU, xi = _cr_problem(nside, lmax)
    npix = 12 * nside ** 2
    bl_map = utils.expand_per_l(_dev.f64(bl), 0)
    fwhm = 0.5 * 512 / nside
    dls = {"EE": _dev.f64(dlE), "BB": _dev.f64(dlB)}
    xid = [_dev.f64(x) for x in xi]
    ref = CR({"Q": dQ, "U": dU}, 1.0, noise_var, bl_map, lmax, npix, fwhm, mask=mask, rng="philox", seed=1)
    ref.pcg_accuracy = 1e-9
    sol, _ = ref.sample_mask(dls, xi=xid)
    rhs_ref = ref.last_rhs
    plans = ShardedPlan.local_group(nside, lmax, world)

    def work(p):
        cr = CR({"Q": dQ, "U": dU}, 1.0, noise_var, bl_map, lmax, npix, fwhm, mask=mask, rng="philox", seed=1, plan=p)
        cr.pcg_accuracy = 1e-9
        s, _ = cr.sample_mask(dls, xi=xid)
        act, tot = C.c_int(-1), C.c_int(-1)
        _lib.check(_lib.lib().gs_active_ring_pairs(p._h, C.byref(act), C.byref(tot)))
        return dict(e=s["EE"], b=s["BB"], re=cr.last_rhs[0], rb=cr.last_rhs[1], it=cr.last_pcg_iterations,
                    ninv=cr.ninv_sum_over_4pi, act=(act.value, tot.value))

    res = run_local_group(plans, work)

    def join(key):
        out = torch.zeros((lmax + 1) ** 2, dtype=torch.float64, device="cuda")
        for p, r in zip(plans, res):
            out[p.real_index] = r[key]
        return out

    assert abs(res[0]["ninv"] - ref.ninv_sum_over_4pi) < 1e-12 * ref.ninv_sum_over_4pi
    # the ring pairs wholly inside the mask are left out of the sharded mat-vec too: every rank holds the same global list
    assert len({r["act"] for r in res}) == 1 and 0 < res[0]["act"][0] < res[0]["act"][1] == 2 * nside, [r["act"] for r in res]
    assert relerr(join("re"), rhs_ref[0]) < RTOL and relerr(join("rb"), rhs_ref[1]) < RTOL
    its = [r["it"] for r in res]
    assert len(set(its)) == 1, its                                     # every rank stops at the same iteration
    assert abs(its[0] - ref.last_pcg_iterations) <= max(2, ref.last_pcg_iterations // 50), (its, ref.last_pcg_iterations)
    assert relerr(join("e"), sol["EE"]) < 1e-7 and relerr(join("b"), sol["BB"]) < 1e-7


@pytest.mark.parametrize("world,nb", [(2, 1), (2, 4)])
def test_nccl_sharded_sht_and_gibbs(world, nb):
    """Production path: one process per GPU, NCCL all-to-all; needs >= `world` GPUs on the box.  nb = 4: the exchange runs block by
    block on the plan's communication stream, overlapped with the Legendre kernels of the neighbouring blocks."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29631 + nb), os.path.join(ROOT, "tests", "shard_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, GS_SHARD_NB=str(nb)))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SHARD_WORKER_OK" in out.stdout
