"""BASELINE config #5: spin-0 / spin-2 alm2map + map2alm(iter=0) pairs/s, lmax 256..4096 (NSIDE = lmax/2), batched over K chains
with the chain-batched entry points (gs_alm2map_batch / gs_map2alm_batch: two right-hand sides per launch share one Legendre
recurrence), next to K separate single-chain calls on the same plan and to the CPU port on the host cores (spin 2: the vectorised
oracle/sht_fast.c; spin 0: the scalar FP64 OpenMP oracle).
usage: python scripts/sht_sweep.py [max_lmax] [cpu_max_lmax]   -> one JSON line per (lmax, spin, K)"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gibbssampler_b200.sht import Plan  # noqa: E402

max_l = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cpu_max = int(sys.argv[2]) if len(sys.argv) > 2 else 1024


def cpu_pair(nside, lmax, spin):
    from oracle import sht as O
    rng = np.random.default_rng(0)
    n = O.nalm(lmax)
    a = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    best = 1e30
    for _ in range(2):
        t0 = time.perf_counter()
        if spin == 2:
            q, u = O.alm2map_spin2(a, b, nside, lmax, kind="fast")
            O.map2alm_spin2(q, u, nside, lmax, kind="fast")
        else:
            m = O.alm2map(a, nside, lmax, kind="f64")
            O.map2alm(m, nside, lmax, kind="f64")
        best = min(best, time.perf_counter() - t0)
    return best, (O.num_threads("fast") if spin == 2 else O._lib("f64").orc_num_threads())


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def gpu_pairs(plan, lmax, spin, K, reps):
    nre = (lmax + 1) ** 2
    g = torch.Generator(device="cuda").manual_seed(1)
    xe = torch.randn((K, nre), generator=g, device="cuda", dtype=torch.float64)
    xb = torch.randn((K, nre), generator=g, device="cuda", dtype=torch.float64)

    def batched():
        if spin == 2:
            q, u = plan.alm2map_spin2_batch(xe, xb)
            plan.map2alm_spin2_batch(q, u, real_layout=True)
        else:
            m = plan.alm2map_batch(xe)
            plan.map2alm_batch(m, real_layout=True)

    def separate():
        for k in range(K):
            if spin == 2:
                q, u = plan.alm2map_spin2(xe[k], xb[k])
                plan.map2alm_spin2(q, u, real_layout=True)
            else:
                m = plan.alm2map(xe[k])
                plan.map2alm(m, real_layout=True)
    return timed(batched, reps), timed(separate, reps)


lmax = 256
while lmax <= max_l:
    nside = lmax // 2
    plan = Plan(nside, lmax)
    for spin in (0, 2):
        cpu = None
        if lmax <= cpu_max:
            t, cores = cpu_pair(nside, lmax, spin)
            cpu = {"pairs_per_s": 1.0 / t, "cores": cores,
                   "kind": "vectorised port oracle/sht_fast.c" if spin == 2 else "scalar oracle port (C + OpenMP, FP64)"}
        for K in ((1, 2, 4, 8) if lmax <= 1024 else (1, 2)):   # nside >= 1024: the long rings take the split path one chain at a time
            reps = 20 if lmax <= 1024 else 5
            ms_b, ms_s = gpu_pairs(plan, lmax, spin, K, reps)
            print(json.dumps({"config": 5, "lmax": lmax, "nside": nside, "spin": spin, "chains": K, "ms_per_round_batched": ms_b,
                              "pairs_per_s_batched": K * 1e3 / ms_b, "ms_per_round_separate_calls": ms_s,
                              "pairs_per_s_separate_calls": K * 1e3 / ms_s, "batch_gain": ms_s / ms_b, "cpu": cpu if K == 1 else None}), flush=True)
    del plan
    torch.cuda.empty_cache()
    lmax *= 2
