"""Does running K independent chains concurrently (one host thread + CUDA stream + Plan each) raise per-GPU
throughput of the PCG mat-vec?  Prints mat-vecs/s for K = 1, 2, 3, 4."""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gibbssampler_b200 import _dev, _lib  # noqa: E402
from gibbssampler_b200.sht import Plan  # noqa: E402

nside, lmax, nrep = 512, 1024, 60
npix, nre = 12 * nside ** 2, (lmax + 1) ** 2
dlE, dlB = bench.fiducial(lmax)
bl = _dev.f64(_dev.gauss_beam(np.radians(0.5), lmax))
invn = _dev.f64(bench.make_mask(nside) / (0.04 * npix / 786432.0))
L = _lib.lib()


def run(K):
    plans = [Plan(nside, lmax) for _ in range(K)]
    xs = [(torch.randn(nre, device="cuda", dtype=torch.float64), torch.randn(nre, device="cuda", dtype=torch.float64)) for _ in range(K)]
    ys = [(torch.empty(nre, device="cuda", dtype=torch.float64), torch.empty(nre, device="cuda", dtype=torch.float64)) for _ in range(K)]
    de, db = _dev.f64(dlE), _dev.f64(dlB)
    streams = [torch.cuda.Stream() for _ in range(K)]
    torch.cuda.synchronize()

    def work(i, n):
        with torch.cuda.stream(streams[i]):
            for _ in range(n):
                _lib.check(L.gs_cr_apply_q_pol(plans[i]._h, _dev.ptr(de), _dev.ptr(db), _dev.ptr(bl), _dev.ptr(invn), _dev.ptr(xs[i][0]),
                                               _dev.ptr(xs[i][1]), _dev.ptr(ys[i][0]), _dev.ptr(ys[i][1]), _dev.stream()))
            streams[i].synchronize()
    for n in (5, nrep):
        ts = [threading.Thread(target=work, args=(i, n)) for i in range(K)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print("K = %d chains: %.1f mat-vecs/s total, %.3f ms per mat-vec per chain" % (K, K * nrep / dt, 1e3 * dt / nrep), flush=True)


for K in (1, 2, 3, 4):
    run(K)
