"""BASELINE config #5: spin-0 / spin-2 alm2map + map2alm(iter=0) pairs/s, lmax 256..4096 (NSIDE = lmax/2), K independent
chains run concurrently (one plan + stream each), with the CPU oracle port on the host cores beside it.
usage: python scripts/sht_sweep.py [max_lmax] [cpu_max_lmax]   -> one JSON line per (lmax, spin, K)"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gibbssampler_b200 import _dev  # noqa: E402
from gibbssampler_b200.sht import Plan  # noqa: E402

max_l = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cpu_max = int(sys.argv[2]) if len(sys.argv) > 2 else 1024


def cpu_pair(nside, lmax, spin):
    from oracle import sht as O
    rng = np.random.default_rng(0)
    n = O.nalm(lmax)
    a = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    t0 = time.perf_counter()
    if spin == 2:
        q, u = O.alm2map_spin2(a, b, nside, lmax, kind="f64")
        O.map2alm_spin2(q, u, nside, lmax, kind="f64")
    else:
        m = O.alm2map(a, nside, lmax, kind="f64")
        O.map2alm(m, nside, lmax, kind="f64")
    return time.perf_counter() - t0, O._lib("f64").orc_num_threads()


def gpu_pairs(nside, lmax, spin, K, reps):
    plans = [Plan(nside, lmax) for _ in range(K)]
    streams = [torch.cuda.Stream() for _ in range(K)]
    nre = (lmax + 1) ** 2
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = [[torch.randn(nre, generator=g, device="cuda", dtype=torch.float64) for _ in range(2)] for _ in range(K)]
    torch.cuda.synchronize()

    def once():
        for k in range(K):
            with torch.cuda.stream(streams[k]):
                if spin == 2:
                    q, u = plans[k].alm2map_spin2(xs[k][0], xs[k][1])
                    plans[k].map2alm_spin2(q, u, real_layout=True)
                else:
                    m = plans[k].alm2map(xs[k][0])
                    plans[k].map2alm(m, real_layout=True)
    for _ in range(2):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for _ in range(reps):
        once()
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    del plans
    torch.cuda.empty_cache()
    return ms


lmax = 256
while lmax <= max_l:
    nside = lmax // 2
    for spin in (0, 2):
        cpu = None
        if lmax <= cpu_max:
            t, cores = cpu_pair(nside, lmax, spin)
            cpu = {"pairs_per_s": 1.0 / t, "cores": cores, "kind": "oracle port (C + OpenMP, FP64)"}
        for K in ((1, 2, 4, 8) if lmax <= 1024 else (1,)):
            reps = 20 if lmax <= 1024 else 5
            ms = gpu_pairs(nside, lmax, spin, K, reps)
            print(json.dumps({"config": 5, "lmax": lmax, "nside": nside, "spin": spin, "chains": K, "ms_per_round": ms,
                              "pairs_per_s": K * 1e3 / ms, "cpu": cpu if K == 1 else None}), flush=True)
    lmax *= 2
