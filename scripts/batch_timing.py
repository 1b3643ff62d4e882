"""Stage timing of the PCG mat-vec with one and with two right-hand sides per launch (chain batch), NSIDE 512 / lmax 1024."""
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from gibbssampler_b200 import _dev, _lib
from gibbssampler_b200.sht import Plan
nside, lmax = 512, 1024
L = _lib.lib()
plan = Plan(nside, lmax)
nre, npix = (lmax + 1) ** 2, 12 * nside ** 2
g = torch.Generator(device='cuda').manual_seed(1)
x = torch.randn((2, 2, nre), generator=g, device='cuda', dtype=torch.float64)
y = torch.empty_like(x)
bl = _dev.f64(_dev.gauss_beam(np.radians(0.5), lmax))
for name, w in (("all rings", torch.rand(npix, generator=g, device='cuda', dtype=torch.float64) + 0.5),
                ("f_sky 0.8 band mask", _dev.f64(bench.make_mask(nside) / (0.04 * npix / 786432.0)))):
    for nc in (1, 2):
        ms = (C.c_float * 3)()
        for n in (3, 20):
            _lib.check(L.gs_profile_matvec_batch(plan._h, nc, _dev.ptr(x[0, 0]), _dev.ptr(x[0, 1]), 2 * nre, _dev.ptr(bl), _dev.ptr(w),
                                                 _dev.ptr(y[0, 0]), _dev.ptr(y[0, 1]), n, ms, _dev.stream()))
        print("%-20s chains/launch %d: leg_synth %.3f ring %.3f leg_anal %.3f total %.3f ms = %.3f ms per chain  chk %.6e"
              % (name, nc, ms[0], ms[1], ms[2], sum(ms), sum(ms) / nc, float(y[:nc].abs().sum().item())), flush=True)
