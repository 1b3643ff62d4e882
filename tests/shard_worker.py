"""Worker of tests/test_shard_gpu.py::test_nccl_sharded_sht_and_gibbs (run under torch.distributed.run, one
process per GPU): the NCCL m-sharded transforms, CR solve and a short CenteredGibbs chain must reproduce
the single-GPU path on the same inputs and numpy random stream."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def relerr(got, ref):
    return float((got - ref).abs().max() / ref.abs().max())


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import bench
    from gibbssampler_b200 import _dev, utils
    from gibbssampler_b200.CenteredGibbs import CenteredGibbs
    from gibbssampler_b200.sharded import ShardedPlan
    from gibbssampler_b200.sht import Plan

    for nside, lmax in ((16, 47), (128, 256)):
        g = torch.Generator(device="cpu").manual_seed(11 + nside)
        nre, npix = (lmax + 1) ** 2, 12 * nside ** 2
        e = torch.randn(nre, generator=g, dtype=torch.float64).cuda()
        b = torch.randn(nre, generator=g, dtype=torch.float64).cuda()
        for a in (e, b):
            a[[0, 1, lmax + 1, lmax + 2]] = 0
        q = torch.randn(npix, generator=g, dtype=torch.float64).cuda()
        u = torch.randn(npix, generator=g, dtype=torch.float64).cuda()
        w = torch.rand(npix, generator=g, dtype=torch.float64).cuda() + 0.5
        fl = torch.rand(lmax + 1, generator=g, dtype=torch.float64).cuda() + 0.5
        ref = Plan.get(nside, lmax)
        rq, ru = ref.alm2map_spin2(e, b, fl=fl)
        re_, rb_ = ref.map2alm_spin2(q, u, adjoint=True, pixw=w, fl=fl, real_layout=True)
        ie, ib = ref.map2alm_spin2(q, u, iter=3, real_layout=True)
        sp = ShardedPlan(nside, lmax)
        sq, su = sp.alm2map_spin2(sp.local_alm(e), sp.local_alm(b), fl=fl)
        ae, ab = sp.map2alm_spin2(sp.local_map(q), sp.local_map(u), adjoint=True, pixw=sp.local_map(w), fl=fl, real_layout=True)
        je, jb = sp.map2alm_spin2(sp.local_map(q), sp.local_map(u), iter=3, real_layout=True)
        errs = [relerr(sp.gather_map(sq), rq), relerr(sp.gather_map(su), ru), relerr(sp.gather_alm(ae), re_),
                relerr(sp.gather_alm(ab), rb_), relerr(sp.gather_alm(je), ie), relerr(sp.gather_alm(jb), ib),
                relerr(sp.alm2cl(sp.local_alm(e)), ref.alm2cl(e))]
        assert max(errs) < 1e-10, (nside, lmax, errs)
        if rank == 0:
            print("sharded SHT nside %d lmax %d world %d: max rel err %.2e" % (nside, lmax, world, max(errs)), flush=True)

    # ---- a short masked CenteredGibbs chain, sharded vs single GPU, same numpy stream on every rank
    nside, lmax, n_iter = 32, 64, 3
    npix = 12 * nside ** 2
    dlE, dlB = bench.fiducial(lmax)
    noise_var = 0.04 * npix / 786432.0
    mask = bench.make_mask(nside)
    rng = np.random.default_rng(5)
    dQ, dU = rng.standard_normal(npix) * mask, rng.standard_normal(npix) * mask
    bins = bench.bins_for(lmax)
    init = {p: np.array([d[bins[p][i]:bins[p][i + 1]].mean() for i in range(len(bins[p]) - 1)]) for p, d in (("EE", dlE), ("BB", dlB))}
    fwhm = 0.5 * 512 / nside
    hist = []
    for plan in (None, ShardedPlan(nside, lmax)):
        np.random.seed(77)
        gs = CenteredGibbs({"Q": dQ, "U": dU}, 1.0, noise_var, fwhm, nside, lmax, npix, mask=mask, polarization=True, bins=bins,
                           n_iter=n_iter, rng="numpy", plan=plan)
        gs.constrained_sampler.ula = False          # plain PCG constrained realization every iteration
        gs.ula = False
        gs.constrained_sampler.pcg_accuracy = 1e-9
        h, _, _, _ = gs.run(init)
        hist.append(h)
    for pol in ("EE", "BB"):
        d = np.abs(hist[1][pol] - hist[0][pol]).max() / np.abs(hist[0][pol]).max()
        assert d < 1e-6, (pol, d)
        # all ranks hold the same history
        t = torch.as_tensor(hist[1][pol], device="cuda")
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        assert all(torch.equal(parts[0], x) for x in parts)
    if rank == 0:
        print("sharded CenteredGibbs chain matches the single-GPU chain", flush=True)
        print("SHARD_WORKER_OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
